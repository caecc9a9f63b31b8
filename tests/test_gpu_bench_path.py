"""Parity of the configuration that bench.py measures: bf16 coupling convs on the fused tcgen05 kernels, full
depth (K = 32, L = 3, hidden 512, 64x64), CUDA graphs -- against the fp32 CPU oracle (oracle/glow_oracle.py, pinned to
the reference by tests/golden).  Reference: network/model.py:409-452 (bits/dim), 496-506 (loss), network/trainer.py
84-150 (iteration).  The measured errors are printed and written to GLOWK_PARITY_LOG (profiles/ keeps a copy)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import glow_oracle as O
import pytorch_glow_b200 as G
from pytorch_glow_b200.hps import make_hps
from pytorch_glow_b200.train import FusedTrainStep
from parity_util import adopt

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _log(record):
    path = os.environ.get("GLOWK_PARITY_LOG")
    print("parity:", json.dumps(record))
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(record) + "\n")


def _glow(K, L, hidden, batch, seed, coupling_std, logs_std=0.03):
    """A model in the state training leaves it in after step 0: data-dependent ActNorm initialisation on a batch
    (network/module.py:86-120, trainer.py:112-115; the engine's init pass is itself checked against the reference in
    test_gpu_model.py), then the zero-initialised convs randomised so that couplings and Split2d priors are
    non-trivial.  Returns (Glow on the device, its state_dict on the CPU for the oracle)."""
    hps = make_hps((64, 64, 3), K=K, L=L, hidden_channels=hidden, coupling="affine", permutation="invconv", batch=batch)
    np.random.seed(seed)
    torch.manual_seed(seed)
    glow = G.Glow(hps).to(DEV)
    glow.flow.set_conv_dtype("fp32")
    glow.train()
    with torch.no_grad():
        glow(x=torch.rand(batch, 3, 64, 64, generator=torch.Generator().manual_seed(seed + 100)).to(DEV))
    sd = {k: v.detach().cpu().clone() for k, v in glow.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    for k, v in sd.items():
        if ".f.4." in k or "conv2d_zeros" in k:
            v.copy_(torch.randn(v.shape, generator=g) * (logs_std if k.endswith("logs") else coupling_std))
    adopt(glow, sd)
    glow.flow.set_conv_dtype(None)
    return glow, sd


def test_bits_per_dim_bf16_full_depth_vs_oracle():
    """Glow 64x64 K=32 L=3 hidden 512 (profile/celeba.json:47-70), the zero-initialised convs randomised so that the
    couplings are non-trivial: bits/dim within 2e-3 relative, z within 1e-2 of its max -- the bf16 bound DESIGN.md
    states, measured here through all 96 flow steps (fp32 path on the same weights: 1e-4)."""
    glow, sd = _glow(32, 3, 512, 2, 21, 0.003)
    glow = glow.eval()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 64, 64, generator=g)
    noise = torch.rand(2, 3, 64, 64, generator=g) / 256
    z_ref, nll_ref = O.glow_nll(x, noise, sd, (64, 64, 3), 32, 3, "invconv", "affine")
    out = {}
    for mode in ("bf16", "fp32"):
        glow.flow.set_conv_dtype(mode)
        with torch.no_grad():
            z, nll, _ = glow.normal_flow(x.to(DEV), None, noise=noise.to(DEV))
        out[mode] = (rel_err(z, z_ref), rel_err(nll, nll_ref))
    _log({"test": "bits_per_dim_full_depth", "nll_ref": [float(v) for v in nll_ref],
          "bf16_rel_err_z": out["bf16"][0], "bf16_rel_err_nll": out["bf16"][1],
          "fp32_rel_err_z": out["fp32"][0], "fp32_rel_err_nll": out["fp32"][1]})
    assert out["fp32"][0] < 1e-4 and out["fp32"][1] < 1e-4
    assert out["bf16"][0] < 1e-2 and out["bf16"][1] < 2e-3


def test_sampling_bf16_full_depth_vs_oracle():
    """Reverse pass (network/model.py:454-471, eps_std 0.7) of the same model with the noise supplied: x within 1e-2
    of its max on the bf16 path (measured 4e-3), 1e-4 in fp32 (measured 3e-6)."""
    glow, sd = _glow(32, 3, 512, 2, 23, 0.003)
    glow = glow.eval()
    g = torch.Generator().manual_seed(6)
    top = torch.randn(2, 48, 8, 8, generator=g) * 0.7
    eps = [torch.randn(2, 12, 16, 16, generator=g) * 0.7, torch.randn(2, 6, 32, 32, generator=g) * 0.7]
    x_ref = O.flow_decode(top.clone(), sd, (64, 64, 3), 32, 3, "invconv", "affine", eps_list=[e.clone() for e in eps])
    errs = {}
    for mode in ("bf16", "fp32"):
        glow.flow.set_conv_dtype(mode)
        with torch.no_grad():
            x = glow.flow.decode(top.to(DEV), eps_list=[e.to(DEV) for e in eps])
        errs[mode] = rel_err(x, x_ref)
    _log({"test": "sampling_full_depth", "bf16_rel_err_x": errs["bf16"], "fp32_rel_err_x": errs["fp32"]})
    assert errs["fp32"] < 1e-4 and errs["bf16"] < 1e-2


def _oracle_train(sd0, x, noises, K, L, steps):
    p = {k: v.clone().requires_grad_(k != "h_top") for k, v in sd0.items()}
    names = [k for k in p if k != "h_top"]
    ms = {k: torch.zeros_like(p[k]) for k in names}
    vs = {k: torch.zeros_like(p[k]) for k in names}
    losses = []
    for t in range(steps):
        for k in names:
            p[k].grad = None
        _, nll = O.glow_nll(x, noises[t], p, (64, 64, 3), K, L, "invconv", "affine")
        loss = O.generative_loss(nll)
        loss.backward()
        losses.append(float(loss.detach()))
        O.clip_grads_([p[k].grad for k in names], 5.0, 100.0)
        with torch.no_grad():
            for k in names:
                O.adam_step_(p[k], p[k].grad, ms[k], vs[k], t + 1, O.noam_lr(1e-3, t, 4000, 1e-4))
    return losses, p


def test_twenty_train_iterations_bf16_graphs_track_fp32_oracle():
    """20 iterations (fwd + bwd + clip + Adam, Noam schedule) of a K=4 L=3 hidden-512 64x64 Glow on the bf16 fused
    kernels inside CUDA graphs against the same iterations through the fp32 oracle: per-step |loss - loss_ref| <= 2e-3
    bits/dim; the total parameter update after 20 steps has cosine >= 0.995 with the oracle's and differs by <= 10 % in
    L2 norm (Adam normalises every coordinate, so near-zero gradients take full-size steps of either sign and a
    per-tensor max-error bound is meaningless)."""
    steps, K, L, B = 20, 4, 3, 4
    glow, sd0 = _glow(K, L, 512, B, 31, 0.01)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(B, 3, 64, 64, generator=g)
    noises = [torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(70 + i)) / 256 for i in range(steps)]
    ref_losses, p_ref = _oracle_train(sd0, x, noises, K, L, steps)
    ts = FusedTrainStep(glow, use_graphs=True)
    static_noise = torch.empty(B, 3, 64, 64, device=DEV)
    orig = torch.nn.init.uniform_

    def fake_uniform(t, a=0., b=1.):
        t.copy_(static_noise)
        return t
    torch.nn.init.uniform_ = fake_uniform
    try:
        losses = []
        xd = x.to(DEV)
        for t in range(steps):
            static_noise.copy_(noises[t])
            losses.append(float(ts.step(xd)))
    finally:
        torch.nn.init.uniform_ = orig
    dl = [abs(a - b) for a, b in zip(losses, ref_losses)]
    got = glow.state_dict()
    names = [k for k in p_ref if k != "h_top"]
    upd = torch.cat([(got[k].detach().cpu() - sd0[k]).reshape(-1) for k in names]).double()
    upd_ref = torch.cat([(p_ref[k].detach() - sd0[k]).reshape(-1) for k in names]).double()
    cos = float(upd @ upd_ref / (upd.norm() * upd_ref.norm()))
    upd_rel = float((upd - upd_ref).norm() / upd_ref.norm())
    worst = max(rel_err(got[k], p_ref[k]) for k in names)
    _log({"test": "train_20_iterations_bf16_graphs", "loss_first": losses[0], "loss_last": losses[-1],
          "loss_ref_first": ref_losses[0], "loss_ref_last": ref_losses[-1], "max_abs_dloss": max(dl),
          "update_cosine": cos, "update_rel_l2_err": upd_rel, "worst_param_rel_err": worst})
    assert max(dl) <= 2e-3, (losses, ref_losses)
    assert losses[-1] < losses[0] - 0.2                           # the model actually trains
    assert cos >= 0.995 and upd_rel <= 0.1


def test_gradients_bf16_full_width_vs_oracle():
    """All parameter gradients of a K=2 L=3 hidden-512 64x64 Glow (every level runs its benchmark-path kernels) on
    the bf16 path against oracle autograd.  Measured on B200 (profiles/r2_parity_bench_path.jsonl): every tensor has
    cosine >= 0.9987 with its reference and no element is off by more than 0.11 of the tensor's max (the worst are the
    conv1 / conv2 weights of the 8x8 level: few pixels to average the bf16 rounding of the gradient signal and the
    ReLU-mask flips over); all gradients concatenated: cosine >= 0.9997.  Asserted: 0.15 / 0.997 per tensor,
    0.999 / 5 % L2 over all parameters."""
    K, L, B = 2, 3, 4
    glow, sd = _glow(K, L, 512, B, 41, 0.01)
    glow = glow.train()
    g = torch.Generator().manual_seed(8)
    x = torch.rand(B, 3, 64, 64, generator=g)
    noise = torch.rand(B, 3, 64, 64, generator=g) / 256
    p = {k: v.clone().requires_grad_(k != "h_top") for k, v in sd.items()}
    _, nll_ref = O.glow_nll(x, noise, p, (64, 64, 3), K, L, "invconv", "affine")
    O.generative_loss(nll_ref).backward()
    z, nll, _ = glow.normal_flow(x.to(DEV), None, noise=noise.to(DEV))
    G.Glow.generative_loss(nll).backward()
    rows, worst_rel, worst_cos = [], 0.0, 1.0
    gmax = max(float(p[k].grad.abs().max()) for k in p if p[k].grad is not None)
    for k, prm in glow.named_parameters():
        if k == "h_top" or p[k].grad is None:
            continue
        a, b = prm.grad.detach().double().cpu().reshape(-1), p[k].grad.double().reshape(-1)
        e = float((a - b).abs().max() / max(float(b.abs().max()), 1e-6 * gmax))
        cs = float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))
        rows.append((k, e, cs, float(b.abs().max())))
        if float(b.abs().max()) > 1e-4 * gmax:
            worst_rel, worst_cos = max(worst_rel, e), min(worst_cos, cs)
    rows.sort(key=lambda r: -r[1])
    names = [k for k, _ in glow.named_parameters() if k != "h_top" and p[k].grad is not None]
    ga = torch.cat([dict(glow.named_parameters())[k].grad.detach().double().cpu().reshape(-1) for k in names])
    gb = torch.cat([p[k].grad.double().reshape(-1) for k in names])
    cos_all = float(ga @ gb / (ga.norm() * gb.norm()))
    rel_all = float((ga - gb).norm() / gb.norm())
    _log({"test": "gradients_bf16_full_width", "tensors": len(rows), "worst_rel_err": worst_rel, "worst_cosine": worst_cos,
          "all_params_cosine": cos_all, "all_params_rel_l2_err": rel_all, "nll_rel_err": rel_err(nll, nll_ref),
          "top5": [(k, round(e, 5), round(c, 5)) for k, e, c, _ in rows[:5]]})
    assert rel_err(nll, nll_ref) < 2e-3
    assert worst_rel <= 0.15 and worst_cos >= 0.997, rows[:8]
    assert cos_all >= 0.999 and rel_all <= 0.05
