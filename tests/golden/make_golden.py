"""Generate golden fixtures by running the REAL reference (corenel/pytorch-glow).

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports network/module.py + network/model.py from /root/reference behind the
two import shims of tests/golden/_shims, runs seeded inputs through
the reference's own classes on CPU and writes tests/golden/*.npz.  The fixtures
pin oracle/glow_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA
path.  Gradients come from the reference with the out-of-place restatement of
FlowStep.normal_flow monkey-patched in (the in-place original cannot
run backward under torch>=2; forward is bit-identical).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GLOW_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)

from easydict import EasyDict  # noqa: E402  (shim)
from network import module as rm  # noqa: E402
from network import model as rmodel  # noqa: E402
from misc import ops as rops  # noqa: E402
import torch.nn.functional as F  # noqa: E402

torch.set_num_threads(4)


def randomize_(net, seed):
    """Make every zero-/identity-initialised tensor non-trivial."""
    g = torch.Generator().manual_seed(seed)
    for name, m in net.named_modules():
        cls = m.__class__.__name__
        if cls == "Conv2dZeros":
            m.weight.data.normal_(0, 0.05, generator=g)
            m.bias.data.normal_(0, 0.05, generator=g)
            m.logs.data.normal_(0, 0.1, generator=g)
        elif cls == "ActNorm":
            m.bias.data.normal_(0, 0.2, generator=g)
            m.logs.data.normal_(0, 0.1, generator=g)
            m.bias_inited = True
            m.logs_inited = True
        elif cls == "Invertible1x1Conv":
            m.weight.data.add_(0.05 * torch.randn(m.weight.shape, generator=g))
    return net


def sd_np(net, prefix="sd/"):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}


def rnd(gen, *shape, scale=1.0, shift=0.0):
    return torch.randn(*shape, generator=gen) * scale + shift


def save(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), len(d), "arrays")


class EpsTap:
    """Record what GaussianDiag.eps returns (the Split2d / top noise), in call order."""

    def __init__(self):
        self.rec = []
        self.orig = rm.GaussianDiag.eps

    def __enter__(self):
        orig, rec = self.orig, self.rec

        def tapped(shape_tensor, eps_std=None):
            e = orig(shape_tensor, eps_std)
            rec.append(e.clone())
            return e
        rm.GaussianDiag.eps = staticmethod(tapped)
        return self

    def __exit__(self, *a):
        rm.GaussianDiag.eps = staticmethod(self.orig)


def patched_normal_flow(self, x, logdet=None):
    """Out-of-place FlowStep.normal_flow (network/model.py:82-117), for backward only."""
    z, logdet = self.actnorm(x, logdet=logdet, reverse=False)
    if self.permutation == 'invconv':
        z, logdet = self.invconv(z, logdet, reverse=False)
    elif self.permutation == 'reverse':
        z = self.reverse(z, reverse=False)
    else:
        z = self.shuffle(z, reverse=False)
    z1, z2 = rops.split_channel(z, 'simple')
    if self.coupling == 'additive':
        z2 = z2 + self.f(z1)
    else:
        h = self.f(z1)
        shift, scale = rops.split_channel(h, 'cross')
        scale = torch.sigmoid(scale + 2.)
        z2 = (z2 + shift) * scale
        logdet = rops.reduce_sum(torch.log(scale), dim=[1, 2, 3]) + logdet
    return rops.cat_channel(z1, z2), logdet


# ------------------------------------------------------------------ layers
def gen_layers():
    out = {}
    g = torch.Generator().manual_seed(100)

    # ActNorm: data-dependent init + forward + reverse (module.py:86-149)
    x = rnd(g, 3, 4, 5, 6, scale=2.0, shift=1.0)
    an = rm.ActNorm(4, scale=1.3).train()
    y, ld = an(x.clone(), logdet=torch.zeros(3))
    out["actnorm/x"] = x
    out["actnorm/scale"] = np.float32(1.3)
    out["actnorm/bias"] = an.bias.data
    out["actnorm/logs"] = an.logs.data
    out["actnorm/y"] = y.detach()
    out["actnorm/logdet"] = ld.detach()
    xr, ldr = an(y.detach().clone(), logdet=ld.detach().clone(), reverse=True)
    out["actnorm/x_rev"] = xr.detach()
    out["actnorm/logdet_rev"] = ldr.detach()

    # Invertible1x1Conv (module.py:322-369)
    np.random.seed(1)
    ic = rm.Invertible1x1Conv(6)
    ic.weight.data.add_(0.1 * rnd(g, 6, 6))
    x = rnd(g, 2, 6, 4, 4)
    ld0 = rnd(g, 2)
    y, ld = ic(x, ld0.clone())
    xr, ldr = ic(y.detach(), ld.detach().clone(), reverse=True)
    out["invconv/weight"] = ic.weight.data
    out["invconv/x"] = x
    out["invconv/logdet_in"] = ld0
    out["invconv/y"] = y.detach()
    out["invconv/logdet"] = ld.detach()
    out["invconv/x_rev"] = xr.detach()
    out["invconv/logdet_rev"] = ldr.detach()
    np.random.seed(3)
    out["invconv/init_seed3_c5"] = rm.Invertible1x1Conv(5).weight.data

    # Permutation2d (module.py:372-397)
    for c, seed in ((6, 0), (12, 4)):
        np.random.seed(seed)
        pm = rm.Permutation2d(c, shuffle=True)
        x = rnd(g, 2, c, 3, 3)
        out["perm/c%d_seed%d/indices" % (c, seed)] = pm.indices
        out["perm/c%d_seed%d/indices_inverse" % (c, seed)] = pm.indices_inverse
        out["perm/c%d_seed%d/x" % (c, seed)] = x
        out["perm/c%d_seed%d/y" % (c, seed)] = pm(x)
        out["perm/c%d_seed%d/x_rev" % (c, seed)] = pm(pm(x), reverse=True)
    pr = rm.Permutation2d(6)
    out["perm/reverse6/indices"] = pr.indices

    # Squeeze2d (module.py:539-612)
    x = rnd(g, 2, 3, 4, 6)
    y = rm.Squeeze2d.squeeze(x, 2)
    out["squeeze/x"] = x
    out["squeeze/y"] = y
    out["squeeze/x_rev"] = rm.Squeeze2d.unsqueeze(y, 2)
    out["squeeze/arange16"] = rm.Squeeze2d.squeeze(torch.arange(16.).view(1, 1, 4, 4), 2)

    # GaussianDiag (module.py:400-483)
    mean, logs, x = rnd(g, 2, 3, 4, 4), rnd(g, 2, 3, 4, 4, scale=0.3), rnd(g, 2, 3, 4, 4)
    out["gauss/mean"], out["gauss/logs"], out["gauss/x"] = mean, logs, x
    out["gauss/logps"] = rm.GaussianDiag.logps(mean, logs, x)
    out["gauss/logp"] = rm.GaussianDiag.logp(mean, logs, x)
    torch.manual_seed(11)
    out["gauss/sample_seed11_std07"] = rm.GaussianDiag.sample(mean, logs, 0.7)

    # Conv2d(+ActNorm), Conv2dZeros, f() (module.py:188-319)
    torch.manual_seed(12)
    fnet = rm.f(4, 16, 8)
    randomize_(fnet, 13)
    x = rnd(g, 2, 4, 5, 5)
    out.update(sd_np(fnet, "f/sd/"))
    out["f/x"] = x
    out["f/conv1"] = fnet[0](x).detach()
    out["f/y"] = fnet(x).detach()

    # Split2d (module.py:486-536)
    sp = rm.Split2d(8)
    randomize_(sp, 14)
    x = rnd(g, 2, 8, 4, 4)
    ld0 = rnd(g, 2)
    z1, ld = sp(x, ld0.clone())
    out.update(sd_np(sp, "split/sd/"))
    out["split/x"], out["split/logdet_in"] = x, ld0
    out["split/z1"], out["split/logdet"] = z1.detach(), ld.detach()
    with EpsTap() as tap:
        torch.manual_seed(15)
        xr, _ = sp(z1.detach(), 0., reverse=True, eps_std=0.7)
    out["split/eps"] = tap.rec[0]
    out["split/x_rev"] = xr.detach()
    save("layers.npz", out)


# ------------------------------------------------------------------ flow steps
def gen_flowsteps():
    out = {}
    g = torch.Generator().manual_seed(200)
    for perm in rmodel.FlowStep.flow_permutation_list:
        for coup in rmodel.FlowStep.flow_coupling_list:
            tag = "%s_%s/" % (perm, coup)
            np.random.seed(21)
            torch.manual_seed(22)
            fs = rmodel.FlowStep(8, 16, permutation=perm, coupling=coup, actnorm_scale=1.0)
            randomize_(fs, 23)
            fs.eval()
            x = rnd(g, 2, 8, 4, 6)
            ld0 = rnd(g, 2)
            with torch.no_grad():
                z, ld = fs(x.clone(), ld0.clone(), reverse=False)
                xr, ldr = fs(z.clone(), ld.clone(), reverse=True)
            out.update(sd_np(fs, tag + "sd/"))
            if perm != "invconv":
                pm = getattr(fs, perm)
                out[tag + "indices"] = pm.indices
                out[tag + "indices_inverse"] = pm.indices_inverse
            out[tag + "x"], out[tag + "logdet_in"] = x, ld0
            out[tag + "z"], out[tag + "logdet"] = z, ld
            out[tag + "x_rev"], out[tag + "logdet_rev"] = xr, ldr
    save("flowstep.npz", out)


# ------------------------------------------------------------------ flow model
def gen_flowmodel():
    out = {}
    g = torch.Generator().manual_seed(300)
    for perm, coup in (("invconv", "affine"), ("shuffle", "additive"), ("reverse", "affine")):
        tag = "%s_%s/" % (perm, coup)
        np.random.seed(31)
        torch.manual_seed(32)
        fm = rmodel.FlowModel((16, 16, 3), 16, K=2, L=3, permutation=perm, coupling=coup)
        randomize_(fm, 33)
        fm.eval()
        x = torch.rand(2, 3, 16, 16, generator=g)
        ld0 = rnd(g, 2)
        with torch.no_grad():
            z, ld = fm(x.clone(), ld0.clone(), reverse=False)
            with EpsTap() as tap:
                torch.manual_seed(34)
                xr = fm(z.clone(), eps_std=0.7, reverse=True)
        out.update(sd_np(fm, tag + "sd/flow."))
        for i, layer in enumerate(fm.layers):
            if isinstance(layer, rmodel.FlowStep) and perm != "invconv":
                pm = getattr(layer, perm)
                out[tag + "perm/%d/indices" % i] = pm.indices
                out[tag + "perm/%d/indices_inverse" % i] = pm.indices_inverse
        out[tag + "x"], out[tag + "logdet_in"] = x, ld0
        out[tag + "z"], out[tag + "logdet"] = z, ld
        out[tag + "x_rev"] = xr
        for k, e in enumerate(tap.rec):
            out[tag + "eps/%d" % k] = e
        out[tag + "output_shapes"] = np.asarray(fm.output_shapes)
    save("flowmodel.npz", out)


# ------------------------------------------------------------------ Glow (bits/dim + grads)
def tiny_hps(coupling, perm, batch=4):
    return EasyDict({
        "model": {"image_shape": [16, 16, 3], "hidden_channels": 16, "K": 2, "L": 2,
                  "actnorm_scale": 1.0, "weight_y": 0.0, "n_bits_x": 8},
        "ablation": {"learn_top": False, "y_condition": False, "lu_decomposition": False,
                     "flow_permutation": perm, "flow_coupling": coupling, "seed": 0},
        "optim": {"num_batch_train": batch},
        "device": {"graph": ["cpu"]},
        "dataset": {"num_classes": 1},
    })


def gen_glow():
    out = {}
    g = torch.Generator().manual_seed(400)
    rmodel.FlowStep.normal_flow = patched_normal_flow  # F3 (forward bit-identical)
    for perm, coup in (("invconv", "affine"), ("reverse", "additive")):
        tag = "%s_%s/" % (perm, coup)
        np.random.seed(41)
        torch.manual_seed(42)
        glow = rmodel.Glow(tiny_hps(coup, perm))
        x = torch.rand(4, 3, 16, 16, generator=g)

        # (1) ActNorm data-dependent init through the whole graph (trainer.py:112-115)
        glow.train()
        torch.manual_seed(43)
        noise_init = torch.nn.init.uniform_(torch.empty(*x.shape), 0, 1. / 256)
        torch.manual_seed(43)
        with torch.no_grad():
            z0, nll0, _ = glow(x=x.clone())
        out.update(sd_np(glow, tag + "init/sd/"))
        out[tag + "init/noise"] = noise_init
        out[tag + "init/nll"] = nll0

        # (2) bits/dim + gradients with non-trivial weights
        randomize_(glow, 44)
        out.update(sd_np(glow, tag + "sd/"))
        torch.manual_seed(45)
        noise = torch.nn.init.uniform_(torch.empty(*x.shape), 0, 1. / 256)
        torch.manual_seed(45)
        z, nll, _ = glow(x=x.clone())
        loss = rmodel.Glow.generative_loss(nll)
        glow.zero_grad()
        loss.backward()
        out[tag + "x"], out[tag + "noise"] = x, noise
        out[tag + "z"], out[tag + "nll"], out[tag + "loss"] = z.detach(), nll.detach(), loss.detach()
        for k, v in glow.named_parameters():
            if v.grad is not None:
                out[tag + "grad/" + k] = v.grad.detach()
        for i, layer in enumerate(glow.flow.layers):
            if isinstance(layer, rmodel.FlowStep) and perm != "invconv":
                pm = getattr(layer, perm)
                out[tag + "perm/%d/indices" % i] = pm.indices
                out[tag + "perm/%d/indices_inverse" % i] = pm.indices_inverse

        # (3) sampling: z=None, eps_std=0.7 (BASELINE config 4)
        glow.eval()
        with EpsTap() as tap:
            torch.manual_seed(46)
            xs = glow(z=None, eps_std=0.7, reverse=True)
        out[tag + "sample/x"] = xs
        for k, e in enumerate(tap.rec):
            out[tag + "sample/eps/%d" % k] = e
    save("glow.npz", out)


if __name__ == "__main__":
    gen_layers()
    gen_flowsteps()
    gen_flowmodel()
    gen_glow()
