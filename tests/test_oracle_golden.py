"""Pin oracle/glow_oracle.py against the reference's own outputs (tests/golden/*.npz)
and the known-answer vectors of SURVEY.md section 8(c).  CPU only."""
import numpy as np
import torch

from conftest import assert_close
from oracle import glow_oracle as O

PERMS = ("invconv", "reverse", "shuffle")
COUPS = ("additive", "affine")


# ---------------------------------------------------------------- known answers
def test_known_squeeze_vectors():
    y = O.squeeze2d(torch.arange(16.).view(1, 1, 4, 4))
    assert tuple(y.shape) == (1, 4, 2, 2)
    assert y.flatten().tolist() == [0, 2, 8, 10, 1, 3, 9, 11, 4, 6, 12, 14, 5, 7, 13, 15]
    y = O.squeeze2d(torch.arange(8.).view(1, 2, 2, 2))
    assert tuple(y.shape) == (1, 8, 1, 1) and y.flatten().tolist() == list(range(8))


def test_known_permutation_vectors():
    idx, inv = O.permutation_indices(6)
    assert idx.tolist() == [5, 4, 3, 2, 1, 0] == inv.tolist() and idx.dtype == np.int64
    np.random.seed(0)
    idx, inv = O.permutation_indices(6, shuffle=True)
    assert idx.tolist() == [0, 3, 4, 2, 5, 1] and inv.tolist() == [0, 5, 3, 1, 2, 4]


def test_known_actnorm_init():
    x = torch.tensor([[[[1., 2.], [3., 4.]], [[0., 0.], [2., 2.]]]])
    bias, logs = O.actnorm_init(x)
    assert_close(bias.flatten(), [-2.5, -1.0], 0, 1e-7)
    assert_close(logs.flatten(), [-0.03719088435, -3.1789e-07], 1e-5, 1e-9)
    y, ld = O.actnorm(x, bias, logs, torch.zeros(1))
    assert_close(y.flatten(), [-1.34163964, -0.44721320, 0.44721320, 1.34163964,
                               -0.99999905, -0.99999905, 0.99999905, 0.99999905], 1e-6, 1e-7)
    assert_close(ld, [-0.44629443], 1e-6, 1e-7)


def test_known_gaussian_and_invconv():
    v = O.gaussian_logps(torch.zeros(1), torch.zeros(1), torch.ones(1))
    assert abs(float(v) - (-1.4189385175704956)) < 1e-6
    np.random.seed(0)
    w = O.invconv_init_weight(5)
    x = torch.randn(2, 5, 3, 3)
    y, ld = O.invconv(x, w, torch.zeros(2))
    assert float(ld.abs().max()) < 1e-4
    assert_close(y, torch.einsum("oi,nihw->nohw", w, x), 1e-5, 1e-6)


# ---------------------------------------------------------------- layers.npz
def test_actnorm(golden_layers):
    G = golden_layers
    x = G.t("actnorm/x")
    bias, logs = O.actnorm_init(x, scale=float(G.np("actnorm/scale")))
    assert_close(bias, G.t("actnorm/bias"), 1e-6, 1e-7, "bias")
    assert_close(logs, G.t("actnorm/logs"), 1e-6, 1e-7, "logs")
    y, ld = O.actnorm(x, G.t("actnorm/bias"), G.t("actnorm/logs"), torch.zeros(3))
    assert_close(y, G.t("actnorm/y"), 1e-6, 1e-7, "y")
    assert_close(ld, G.t("actnorm/logdet"), 1e-6, 1e-6, "logdet")
    xr, ldr = O.actnorm(y, G.t("actnorm/bias"), G.t("actnorm/logs"), ld, reverse=True)
    assert_close(xr, G.t("actnorm/x_rev"), 1e-6, 1e-6, "x_rev")
    assert_close(ldr, G.t("actnorm/logdet_rev"), 1e-6, 1e-6, "logdet_rev")


def test_invconv(golden_layers):
    G = golden_layers
    w, x = G.t("invconv/weight"), G.t("invconv/x")
    y, ld = O.invconv(x, w, G.t("invconv/logdet_in"))
    assert_close(y, G.t("invconv/y"), 1e-6, 1e-6, "y")
    assert_close(ld, G.t("invconv/logdet"), 1e-6, 1e-6, "logdet")
    xr, ldr = O.invconv(y, w, ld, reverse=True)
    assert_close(xr, G.t("invconv/x_rev"), 1e-5, 1e-6, "x_rev")
    assert_close(ldr, G.t("invconv/logdet_rev"), 1e-6, 1e-6, "logdet_rev")
    np.random.seed(3)
    assert torch.equal(O.invconv_init_weight(5), G.t("invconv/init_seed3_c5"))


def test_permutation_bit_exact(golden_layers):
    G = golden_layers
    for c, seed in ((6, 0), (12, 4)):
        tag = "perm/c%d_seed%d/" % (c, seed)
        np.random.seed(seed)
        idx, inv = O.permutation_indices(c, shuffle=True)
        assert np.array_equal(idx, G.np(tag + "indices"))
        assert np.array_equal(inv, G.np(tag + "indices_inverse"))
        x = G.t(tag + "x")
        assert torch.equal(O.permute(x, idx, inv), G.t(tag + "y"))
        assert torch.equal(O.permute(O.permute(x, idx, inv), idx, inv, reverse=True), G.t(tag + "x_rev"))
    assert np.array_equal(O.permutation_indices(6)[0], G.np("perm/reverse6/indices"))


def test_squeeze_bit_exact(golden_layers):
    G = golden_layers
    x = G.t("squeeze/x")
    assert torch.equal(O.squeeze2d(x), G.t("squeeze/y"))
    assert torch.equal(O.unsqueeze2d(G.t("squeeze/y")), G.t("squeeze/x_rev"))
    assert np.array_equal(O.squeeze2d_numpy(x.numpy()), G.np("squeeze/y"))
    assert torch.equal(O.squeeze2d(torch.arange(16.).view(1, 1, 4, 4)), G.t("squeeze/arange16"))


def test_gaussian(golden_layers):
    G = golden_layers
    m, l, x = G.t("gauss/mean"), G.t("gauss/logs"), G.t("gauss/x")
    assert_close(O.gaussian_logps(m, l, x), G.t("gauss/logps"), 1e-6, 1e-6)
    assert_close(O.gaussian_logp(m, l, x), G.t("gauss/logp"), 1e-6, 1e-5)
    torch.manual_seed(11)
    assert_close(O.gaussian_sample(m, l, 0.7), G.t("gauss/sample_seed11_std07"), 1e-6, 1e-6)


def test_f_net(golden_layers):
    G = golden_layers
    p = G.sd("f/sd/")
    x = G.t("f/x")
    c1 = O.conv2d_actnorm(x, p["0.weight"], p["0.actnorm.bias"], p["0.actnorm.logs"])
    assert_close(c1, G.t("f/conv1"), 1e-5, 1e-6, "conv1")
    assert_close(O.f_net(x, p, ""), G.t("f/y"), 1e-5, 1e-6, "f")


def test_split2d(golden_layers):
    G = golden_layers
    p = G.sd("split/sd/")
    x = G.t("split/x")
    z1, ld = O.split2d(x, G.t("split/logdet_in"), p, "")
    assert torch.equal(z1, G.t("split/z1"))
    assert_close(ld, G.t("split/logdet"), 1e-6, 1e-5, "logdet")
    xr, _ = O.split2d(z1, 0., p, "", reverse=True, eps=G.t("split/eps"))
    assert_close(xr, G.t("split/x_rev"), 1e-6, 1e-6, "x_rev")
    assert torch.equal(xr[:, :4], x[:, :4])          # test/test_module.py:93-94
    torch.manual_seed(15)
    xr2, _ = O.split2d(z1, 0., p, "", reverse=True, eps_std=0.7)  # same RNG consumption
    assert_close(xr2, G.t("split/x_rev"), 1e-6, 1e-6, "x_rev(rng)")


# ---------------------------------------------------------------- flowstep.npz
def test_flowstep_all_variants(golden_flowstep):
    G = golden_flowstep
    for perm in PERMS:
        for coup in COUPS:
            tag = "%s_%s/" % (perm, coup)
            p = G.sd(tag + "sd/")
            pm = None if perm == "invconv" else (G.np(tag + "indices"), G.np(tag + "indices_inverse"))
            z, ld = O.flowstep(G.t(tag + "x"), G.t(tag + "logdet_in"), p, "", perm, coup, pm)
            assert_close(z, G.t(tag + "z"), 1e-5, 1e-6, tag + "z")
            assert_close(ld, G.t(tag + "logdet"), 1e-5, 1e-5, tag + "logdet")
            xr, ldr = O.flowstep(z, ld, p, "", perm, coup, pm, reverse=True)
            assert_close(xr, G.t(tag + "x_rev"), 1e-5, 2e-6, tag + "x_rev")
            assert_close(ldr, G.t(tag + "logdet_rev"), 1e-5, 1e-5, tag + "logdet_rev")
            assert_close(xr, G.t(tag + "x"), 1e-4, 1e-5, tag + "roundtrip")  # test_model.py:32


# ---------------------------------------------------------------- flowmodel.npz
def test_flowmodel(golden_flowmodel):
    G = golden_flowmodel
    for perm, coup in (("invconv", "affine"), ("shuffle", "additive"), ("reverse", "affine")):
        tag = "%s_%s/" % (perm, coup)
        p = G.sd(tag + "sd/")
        perms = G.perms(tag + "perm/")
        _, shapes = O.flow_layout((16, 16, 3), 2, 3)
        assert np.array_equal(np.asarray(shapes), G.np(tag + "output_shapes"))
        z, ld = O.flow_encode(G.t(tag + "x"), G.t(tag + "logdet_in"), p, (16, 16, 3), 2, 3, perm, coup, perms)
        assert tuple(z.shape) == (2, 48, 2, 2)      # test/test_model.py:56
        assert_close(z, G.t(tag + "z"), 1e-5, 1e-5, tag + "z")
        assert_close(ld, G.t(tag + "logdet"), 1e-5, 1e-4, tag + "logdet")
        eps = [G.t(tag + "eps/%d" % k) for k in range(2)]
        xr = O.flow_decode(G.t(tag + "z"), p, (16, 16, 3), 2, 3, perm, coup, perms, eps_list=eps)
        assert_close(xr, G.t(tag + "x_rev"), 1e-4, 1e-5, tag + "x_rev")
        torch.manual_seed(34)
        xr = O.flow_decode(G.t(tag + "z"), p, (16, 16, 3), 2, 3, perm, coup, perms, eps_std=0.7)
        assert_close(xr, G.t(tag + "x_rev"), 1e-4, 1e-5, tag + "x_rev(rng)")


# ---------------------------------------------------------------- glow.npz
def _glow_cfg(perm, coup):
    return dict(in_shape=(16, 16, 3), K=2, L=2, permutation=perm, coupling=coup)


def test_glow_bits_per_dim_and_grads(golden_glow):
    G = golden_glow
    for perm, coup in (("invconv", "affine"), ("reverse", "additive")):
        tag = "%s_%s/" % (perm, coup)
        p = {k: v.clone().requires_grad_(k != "h_top") for k, v in G.sd(tag + "sd/").items()}
        perms = G.perms(tag + "perm/")
        z, nll = O.glow_nll(G.t(tag + "x"), G.t(tag + "noise"), p, perms=perms, **_glow_cfg(perm, coup))
        assert_close(z, G.t(tag + "z"), 1e-5, 1e-5, tag + "z")
        assert_close(nll, G.t(tag + "nll"), 1e-6, 1e-6, tag + "nll")
        loss = O.generative_loss(nll)
        assert_close(loss, G.t(tag + "loss"), 1e-6, 1e-6, tag + "loss")
        loss.backward()
        n = 0
        for k, v in p.items():
            gk = tag + "grad/" + k
            if G.has(gk):
                assert_close(v.grad, G.t(gk), 1e-4, 1e-7, gk)
                n += 1
        assert n > 20


def test_glow_actnorm_init_pass(golden_glow):
    """Layer-by-layer data-dependent init through the graph (trainer.py:112-115)."""
    G = golden_glow
    for perm, coup in (("invconv", "affine"), ("reverse", "additive")):
        tag = "%s_%s/" % (perm, coup)
        p = G.sd(tag + "init/sd/")
        perms = G.perms(tag + "perm/")
        x = G.t(tag + "x")
        _, nll = O.glow_nll(x, G.t(tag + "init/noise"), p, perms=perms, **_glow_cfg(perm, coup))
        assert_close(nll, G.t(tag + "init/nll"), 1e-5, 1e-5, tag + "init nll")


def test_glow_sample(golden_glow):
    G = golden_glow
    for perm, coup in (("invconv", "affine"), ("reverse", "additive")):
        tag = "%s_%s/" % (perm, coup)
        p = G.sd(tag + "sd/")
        perms = G.perms(tag + "perm/")
        top = G.t(tag + "sample/eps/0")               # h_top is zero => z_top = eps
        eps = [G.t(tag + "sample/eps/1")]
        x = O.glow_sample(top, p, perms=perms, eps_list=eps, **_glow_cfg(perm, coup))
        assert_close(x, G.t(tag + "sample/x"), 1e-4, 1e-5, tag + "sample")


# ---------------------------------------------------------------- LU path (unpinned by the reference, F2)
def test_lu_assemble_matches_dense_layer():
    torch.manual_seed(0)
    w = torch.randn(12, 12)
    pm, l, u, sign_s, log_s = O.lu_factor(w)
    w2, logabsdet = O.lu_assemble(pm, l, u, sign_s, log_s)
    assert_close(w2, w, 1e-4, 1e-5, "P L (U+diag s)")
    assert_close(logabsdet, torch.log(torch.abs(torch.det(w))), 1e-4, 1e-5, "sum log|s|")
    x = torch.randn(2, 12, 3, 3)
    y_lu, ld_lu = O.invconv(x, w2, torch.zeros(2))
    y, ld = O.invconv(x, w, torch.zeros(2))
    assert_close(y_lu, y, 1e-4, 1e-5)
    assert_close(ld_lu, ld, 1e-4, 1e-4)


def test_trainer_arithmetic():
    assert abs(O.noam_lr(1e-3, 0) - 1e-3 * 4000 ** 0.5 * 4000 ** -1.5) < 1e-12
    assert abs(O.noam_lr(1e-3, 3999) - 1e-3) < 1e-9
    assert O.noam_lr(1e-3, 10 ** 7, min_lr=1e-4) == 1e-4
    torch.manual_seed(1)
    ps = [torch.randn(5, 3), torch.randn(7)]
    gs = [torch.randn(5, 3) * 10, torch.randn(7) * 10]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    for r, g in zip(ref, gs):
        r.grad = g.clone()
    torch.nn.utils.clip_grad_value_(ref, 5)
    tn = torch.nn.utils.clip_grad_norm_(ref, 10)
    mine = [g.clone() for g in gs]
    tn2 = O.clip_grads_(mine, 5, 10)
    assert_close(tn2, tn, 1e-6, 1e-6)
    for r, g in zip(ref, mine):
        assert_close(g, r.grad, 1e-6, 1e-7)
    opt = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.9999), eps=1e-8)
    opt.step()
    for r, p0, g in zip(ref, ps, mine):
        m, v = torch.zeros_like(p0), torch.zeros_like(p0)
        O.adam_step_(p0, g, m, v, 1, 1e-3)
        assert_close(p0, r.data, 1e-5, 1e-7)


# ---------------------------------------------------------------- actnorm_init.npz (make_golden_actnorm.py)
def test_actnorm_init_variants_match_reference():
    """batch_variance=True and a first training-mode call with reverse=True (network/module.py:44-45, 62-63, 112-113,
    143-146), against the reference ActNorm's own parameters and outputs."""
    from conftest import Golden
    G = Golden("actnorm_init.npz")
    n = 0
    for c in (12, 48):
        for bv in (0, 1):
            for rev in (0, 1):
                tag = "c%d_bv%d_rev%d/" % (c, bv, rev)
                if not G.has(tag + "x"):
                    continue
                x = G.t(tag + "x")
                if rev:
                    bias, logs = O.actnorm_init_reverse(x, 1.3, 3.0, bool(bv))
                else:
                    bias, logs = O.actnorm_init(x, 1.3, 3.0, bool(bv))
                    logs = logs.expand(1, c, 1, 1)
                assert_close(bias, G.t(tag + "bias"), 1e-5, 1e-6, tag + "bias")
                assert_close(logs, G.t(tag + "logs"), 1e-5, 1e-6, tag + "logs")
                y, ld = O.actnorm(x, bias, logs, torch.zeros(x.shape[0]), reverse=bool(rev))
                assert_close(y, G.t(tag + "y"), 1e-5, 1e-5, tag + "y")
                assert_close(ld, G.t(tag + "logdet"), 1e-5, 1e-4, tag + "logdet")
                n += 1
    assert n == 6
