"""Shared helpers for the GPU parity tests and __graft_entry__.smoke()."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import glow_oracle as O  # noqa: E402  (checker only)
import pytorch_glow_b200 as G  # noqa: E402
from pytorch_glow_b200.hps import make_hps  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def sd_from(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith(prefix)}


def perms_from(z, prefix):
    out = {}
    for k in z.files:
        if k.startswith(prefix) and k.endswith("/indices"):
            i = int(k[len(prefix):].split("/")[0])
            out[i] = (np.array(z[k]), np.array(z[k[:-len("indices")] + "indices_inverse"]))
    return out


def adopt(model, sd, perms=None, strict=True):
    """Load a reference state_dict (+ the unsaved permutation indices) into a pytorch_glow_b200 module."""
    model.load_state_dict(sd, strict=strict)
    for m in model.modules():
        if isinstance(m, G.ActNorm):
            m.bias_inited = m.logs_inited = True
    if perms:
        flow = model.flow if hasattr(model, "flow") else model
        for i, (idx, _) in perms.items():
            flow.layers[i].perm_module.set_indices(idx)
    return model


def randomize_(sd, seed, coupling_std=0.05):
    """Same idea as tests/golden/make_golden.py: make zero-initialised tensors non-trivial."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if k == "h_top":
            continue
        if ".f.4." in k or k.startswith("f.4.") or "conv2d_zeros" in k:
            std = 0.1 if k.endswith("logs") else coupling_std
            v.copy_(torch.randn(v.shape, generator=g) * std)
        elif "actnorm" in k:
            std = 0.1 if k.endswith("logs") else 0.2
            v.copy_(torch.randn(v.shape, generator=g) * std)
        elif k.endswith("invconv.weight"):
            v.add_(0.05 * torch.randn(v.shape, generator=g))
    return sd


def tiny_glow_parity(device="cuda:0", conv_dtype="fp32"):
    """Glow (16x16x3, K=2, L=2, hidden 16) bits/dim + sampling vs the golden fixture. Returns rel errs."""
    z = load_golden("glow.npz")
    tag = "invconv_affine/"
    hps = make_hps((16, 16, 3), K=2, L=2, hidden_channels=16, coupling="affine", permutation="invconv", batch=4)
    np.random.seed(0)
    glow = G.Glow(hps)
    adopt(glow, sd_from(z, tag + "sd/"))
    glow = glow.to(device).eval()
    glow.flow.set_conv_dtype(conv_dtype)
    x = torch.from_numpy(z[tag + "x"]).to(device)
    noise = torch.from_numpy(z[tag + "noise"]).to(device)
    with torch.no_grad():
        zz, nll, _ = glow.normal_flow(x, None, noise=noise)
        top = torch.from_numpy(z[tag + "sample/eps/0"]).to(device)
        eps = [torch.from_numpy(z[tag + "sample/eps/1"]).to(device)]
        xs = glow.flow.decode(top, eps_list=eps)
    torch.cuda.synchronize()
    return (rel(zz, z[tag + "z"]), rel(nll, z[tag + "nll"]), rel(xs, z[tag + "sample/x"]))


def bf16_steps_parity(device="cuda:0"):
    """The tensor-core path the benchmark runs, on a small FlowModel (32x32x3, K=1, L=3) with the full 512-wide hidden
    layer: level 1 (C=12) and level 2 (C=24) run the fused coupling-net kernel with the in-kernel conv1 gather
    (csrc/cnet_fused_sm100.cu; level 2 in its deferred-GEMM3 layout), level 3 (C=48) three tcgen05 GEMMs
    (csrc/gemm_sm100.cu).  Forward and reverse against the fp32 oracle.
    Returns (rel err z, rel err logdet, rel err of the decoded x)."""
    np.random.seed(4)
    torch.manual_seed(4)
    fm = G.FlowModel((32, 32, 3), 512, K=1, L=3, permutation="invconv", coupling="affine")
    sd = randomize_({k: v.clone() for k, v in fm.state_dict().items()}, 5, coupling_std=0.01)
    adopt(fm, sd)
    fm.set_conv_dtype("bf16")
    fm = fm.to(device).eval()
    g = torch.Generator().manual_seed(6)
    x = torch.rand(4, 3, 32, 32, generator=g)
    z_ref, ld_ref = O.flow_encode(x, torch.zeros(4), sd, (32, 32, 3), 1, 3, "invconv", "affine", prefix="")
    top = torch.randn(4, 48, 4, 4, generator=g) * 0.7
    eps = [torch.randn(4, 12, 8, 8, generator=g) * 0.7, torch.randn(4, 6, 16, 16, generator=g) * 0.7]
    x_ref = O.flow_decode(top.clone(), sd, (32, 32, 3), 1, 3, "invconv", "affine", eps_list=[e.clone() for e in eps], prefix="")
    with torch.no_grad():
        z, ld = fm(x.to(device), torch.zeros(4, device=device))
        xs = fm.decode(top.to(device), eps_list=[e.to(device) for e in eps])
    torch.cuda.synchronize()
    return (rel(z, z_ref), rel(ld, ld_ref), rel(xs, x_ref))
