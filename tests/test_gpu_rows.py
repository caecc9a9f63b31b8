"""GPU parity of the pixel-major ("rows") flow kernels (csrc/flow_rows_kernels.cu, rows_path.py).

Each rows kernel is checked against its NCHW twin -- which test_gpu_kernels.py / test_gpu_backward.py pin
against the CPU oracle -- on the same seeded inputs: BIT-EXACT for everything computed per element
(squeeze / permutation / mix / coupling / sample / dz / du), to a tight tolerance for the reductions whose
summation order differs (logdet, parameter gradients).  The whole-model tests then compare FlowModel on
the rows path with the per-layer NCHW path and with the oracle (test_gpu_model.py runs on the rows path)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, rel_err
from oracle import glow_oracle as O
import pytorch_glow_b200 as G
from pytorch_glow_b200 import _C, config
from pytorch_glow_b200 import functional as K
from pytorch_glow_b200 import rows_path
from pytorch_glow_b200.functional import NCHW, ROWS, round_up
from pytorch_glow_b200.hps import make_hps
from parity_util import randomize_

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def g(seed):
    return torch.Generator().manual_seed(seed)


def cu(t):
    return t.to(DEV)


def to_rows(x):
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def to_nchw(r, n, h, w):
    return r.reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


# ---------------------------------------------------------------- squeeze / layout changes
@pytest.mark.parametrize("shape,f", [((2, 3, 8, 8), 2), ((3, 6, 4, 6), 2), ((2, 4, 4, 4), 1), ((1, 2, 6, 6), 3)])
def test_rows_squeeze_all_layouts(shape, f):
    n, c, h, w = shape
    x = torch.randn(*shape, generator=g(1))
    ref = O.squeeze2d(x, f) if f > 1 else x          # [n, c f f, h/f, w/f]
    cs, hs, ws_ = c * f * f, h // f, w // f
    xd = cu(x)
    # NCHW -> rows (squeeze)
    dst = torch.empty(n * hs * ws_, cs, device=DEV)
    K.rows_squeeze(xd, NCHW, c * h * w, dst, ROWS, cs, n, c, h, w, f, False)
    assert torch.equal(dst.cpu(), to_rows(ref))
    # rows (wider pitch: first c channels of 2c) -> rows
    wide = torch.full((n * h * w, 2 * c), 7.0, device=DEV)
    wide[:, :c] = cu(to_rows(x))
    dst2 = torch.empty_like(dst)
    K.rows_squeeze(wide, ROWS, 2 * c, dst2, ROWS, cs, n, c, h, w, f, False)
    assert torch.equal(dst2, dst)
    # rows -> NCHW (squeeze)
    dst3 = torch.empty(n, cs, hs, ws_, device=DEV)
    K.rows_squeeze(cu(to_rows(x)), ROWS, c, dst3, NCHW, cs * hs * ws_, n, c, h, w, f, False)
    assert torch.equal(dst3.cpu(), ref)
    # unsqueeze: rows -> NCHW, rows -> wider rows (only the first c channels are written)
    back = torch.empty(n, c, h, w, device=DEV)
    K.rows_squeeze(dst, ROWS, cs, back, NCHW, c * h * w, n, c, h, w, f, True)
    assert torch.equal(back.cpu(), x)
    back2 = torch.full((n * h * w, 2 * c), 7.0, device=DEV)
    K.rows_squeeze(dst, ROWS, cs, back2, ROWS, 2 * c, n, c, h, w, f, True)
    assert torch.equal(back2[:, :c].cpu(), to_rows(x)) and bool((back2[:, c:] == 7.0).all())
    # NCHW -> NCHW equals the reference kernel
    if f > 1:
        dst4 = torch.empty(n, cs, hs, ws_, device=DEV)
        K.rows_squeeze(xd, NCHW, c * h * w, dst4, NCHW, cs * hs * ws_, n, c, h, w, f, False)
        assert torch.equal(dst4, K.squeeze2d(xd, f))


def test_rows_squeeze_rejects_bad_shapes():
    x = torch.zeros(1, 3, 3, 4, device=DEV)
    with pytest.raises(ValueError):
        K.rows_squeeze(x, NCHW, 36, torch.empty(2, 12, device=DEV), ROWS, 12, 1, 3, 3, 4, 2, False)


# ---------------------------------------------------------------- ActNorm + mix / permutation
@pytest.mark.parametrize("shape", [(2, 12, 8, 8), (3, 24, 4, 4), (2, 48, 4, 4), (1, 4, 3, 5), (2, 96, 2, 2)])
@pytest.mark.parametrize("mode", ["mix", "perm"])
@pytest.mark.parametrize("actnorm", [True, False])
def test_rows_actnorm_mix_bit_exact(shape, mode, actnorm):
    n, c, h, w = shape
    x = cu(torch.randn(*shape, generator=g(2)))
    bias = cu(torch.randn(c, generator=g(3)) * 0.3) if actnorm else None
    logs = cu(torch.randn(c, generator=g(4)) * 0.2) if actnorm else None
    np.random.seed(5)
    wgt = cu(O.invconv_init_weight(c) + 0.1 * torch.randn(c, c, generator=g(6))) if mode == "mix" else None
    idx = cu(torch.randperm(c, generator=g(7))) if mode == "perm" else None
    for reverse in (False, True):
        ref = K.actnorm_mix(x, wgt, idx, bias, logs, 3.0, reverse)
        got = K.rows_actnorm_mix(to_rows(x), wgt, idx, bias, logs, 3.0, reverse)
        assert torch.equal(to_nchw(got, n, h, w), ref), "reverse=%s" % reverse


# ---------------------------------------------------------------- coupling (+ logdet)
def _coupling_inputs(n, c, h, w, affine, seed):
    cout = c if affine else c // 2
    n3p = round_up(9 * cout, 16)
    p3 = torch.randn(n * h * w, n3p, generator=g(seed)) * 0.1
    z = torch.randn(n, c, h, w, generator=g(seed + 1))
    bias3 = torch.randn(cout, generator=g(seed + 2)) * 0.1
    logs3 = torch.randn(cout, generator=g(seed + 3)) * 0.1
    return cu(p3), cu(z), cu(bias3), cu(logs3)


@pytest.mark.parametrize("shape", [(3, 12, 8, 8), (2, 24, 4, 4), (5, 48, 4, 4), (2, 4, 3, 5), (70, 12, 32, 32)])
@pytest.mark.parametrize("affine", [True, False])
@pytest.mark.parametrize("reverse", [False, True])
def test_rows_coupling_matches_nchw(shape, affine, reverse):
    n, c, h, w = shape
    p3, z, bias3, logs3 = _coupling_inputs(n, c, h, w, affine, 10)
    an_logs = cu(torch.randn(c, generator=g(20)) * 0.1)
    logabsdet = cu(torch.tensor([0.37]))
    ld_in = cu(torch.randn(n, generator=g(21)))
    sign = -1.0 if reverse else 1.0
    # NCHW twin
    z_ref = z.clone()
    partials, h_ref = K.coupling(p3, bias3, logs3, z_ref, affine, reverse, 3.0, save_h=True)
    ld_ref = K.logdet_finish(ld_in, n, h * w, logs=an_logs, logabsdet=logabsdet, partials=partials, sign=sign)
    # rows
    zr = to_rows(z)
    nblk = K.rows_coupling_nblk(h * w, c)
    tickets = torch.zeros(n, dtype=torch.int32, device=DEV)
    parts = torch.empty(n * nblk, device=DEV)
    for _ in range(2):                                     # twice: the tickets must come back to zero
        zz = zr.clone()
        ld, hs = K.rows_coupling(p3, bias3, logs3, zz, n, h, w, affine, reverse, 3.0, save_h=True, ld_in=ld_in,
                                 want_ld=True, an_logs=an_logs, an_f=3.0, logabsdet=logabsdet, sign=sign,
                                 partials=parts, tickets=tickets)
        assert torch.equal(to_nchw(zz, n, h, w), z_ref)
        assert torch.equal(hs, h_ref)
        assert_close(ld, ld_ref, 2e-6, 1e-4, "logdet")
        assert int(tickets.abs().sum()) == 0
    # without a logdet output nothing but z is touched
    zz = zr.clone()
    ld, hs = K.rows_coupling(p3, bias3, logs3, zz, n, h, w, affine, reverse)
    assert ld is None and hs is None and torch.equal(to_nchw(zz, n, h, w), z_ref)


@pytest.mark.parametrize("shape", [(3, 12, 8, 8), (2, 24, 4, 4), (4, 48, 4, 4), (2, 4, 3, 5), (40, 12, 32, 32)])
@pytest.mark.parametrize("affine", [True, False])
def test_rows_coupling_bwd_matches_nchw(shape, affine):
    n, c, h, w = shape
    cout = c if affine else c // 2
    y = cu(torch.randn(*shape, generator=g(30)))
    dy = cu(torch.randn(*shape, generator=g(31)))
    hrows = cu(torch.randn(n * h * w, cout, generator=g(32)) * 0.5)
    dld = cu(torch.randn(n, generator=g(33)))
    logs3 = cu(torch.randn(cout, generator=g(34)) * 0.1)
    dl_ref, db_ref = torch.zeros(cout, device=DEV), torch.zeros(cout, device=DEV)
    dz_ref, du_ref = K.coupling_bwd(y, hrows, dy, dld, logs3, affine, dl_ref, db_ref)
    dl, db = torch.zeros(cout, device=DEV), torch.zeros(cout, device=DEV)
    dz, du = K.rows_coupling_bwd(to_rows(y), hrows, to_rows(dy), dld, logs3, n, h * w, affine, dl, db)
    assert torch.equal(to_nchw(dz, n, h, w), dz_ref)
    assert torch.equal(du, du_ref)
    assert_close(dl, dl_ref, 1e-4, 1e-4, "dlogs3")
    assert_close(db, db_ref, 1e-4, 1e-4, "dbias3")
    # no logdet gradient
    dz0, du0 = K.rows_coupling_bwd(to_rows(y), hrows, to_rows(dy), None, logs3, n, h * w, affine,
                                   torch.zeros(cout, device=DEV), torch.zeros(cout, device=DEV))
    dz0_ref, du0_ref = K.coupling_bwd(y, hrows, dy, None, logs3, affine, torch.zeros(cout, device=DEV),
                                      torch.zeros(cout, device=DEV))
    assert torch.equal(to_nchw(dz0, n, h, w), dz0_ref) and torch.equal(du0, du0_ref)


# ---------------------------------------------------------------- ActNorm + mix backward (+ conv1 dgrad)
@pytest.mark.parametrize("shape", [(3, 12, 8, 8), (2, 24, 4, 4), (3, 48, 4, 4), (1, 4, 4, 4), (2, 96, 2, 2), (9, 12, 16, 16)])
@pytest.mark.parametrize("mode", ["mix", "perm"])
@pytest.mark.parametrize("with_da1", [True, False])
def test_rows_mix_bwd_matches_nchw(shape, mode, with_da1):
    n, c, h, w = shape
    cin = c // 2
    x = cu(torch.randn(*shape, generator=g(40)))
    dz = cu(torch.randn(*shape, generator=g(41)))
    bias = cu(torch.randn(c, generator=g(42)) * 0.3)
    logs = cu(torch.randn(c, generator=g(43)) * 0.2)
    np.random.seed(44)
    wgt = cu(O.invconv_init_weight(c) + 0.1 * torch.randn(c, c, generator=g(45))) if mode == "mix" else None
    idx = cu(torch.randperm(c, generator=g(46))) if mode == "perm" else None
    k1p = round_up(9 * cin, 64)
    da1 = cu(torch.randn(n * h * w, k1p, generator=g(47)) * 0.1) if with_da1 else None
    # NCHW twin: tap gather-sum into dz, then the mix adjoint
    dz_ref = dz.clone()
    if with_da1:
        K.tapsum_to_nchw(da1, dz_ref, 0, cin, flip=True, accumulate=True)
    dw_ref = torch.zeros(c * c, device=DEV) if mode == "mix" else None
    dl_ref, db_ref = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    dx_ref = K.actnorm_mix_bwd(x, dz_ref, wgt, idx, bias, logs, dw_ref, dl_ref, db_ref)
    dw = torch.zeros(c * c, device=DEV) if mode == "mix" else None
    dl, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    dx = K.rows_actnorm_mix_bwd(to_rows(x), to_rows(dz), n, h, w, da1=da1, cin=cin if with_da1 else 0, weight=wgt,
                                indices=idx, bias=bias, logs=logs, dw=dw, dlogs=dl, dbias=db)
    assert torch.equal(to_nchw(dx, n, h, w), dx_ref)
    assert_close(dl, dl_ref, 1e-4, 1e-4, "dlogs")
    assert_close(db, db_ref, 1e-4, 1e-4, "dbias")
    if mode == "mix":
        assert_close(dw, dw_ref, 1e-4, 1e-4, "dW")
    # with the gradient of the sample-independent logdet terms folded in (glowk_logdet_param_grad)
    dld = cu(torch.randn(n, generator=g(48)))
    winv = torch.linalg.inv(wgt.double()).float().contiguous() if mode == "mix" else None
    K.logdet_param_grad(dld, h * w, dl_ref, winv, dw_ref, 3.0)
    dw2 = torch.zeros(c * c, device=DEV) if mode == "mix" else None
    dl2, db2 = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    dx2 = K.rows_actnorm_mix_bwd(to_rows(x), to_rows(dz), n, h, w, da1=da1, cin=cin if with_da1 else 0, weight=wgt,
                                 indices=idx, bias=bias, logs=logs, dw=dw2, dlogs=dl2, dbias=db2, dld=dld, winv=winv)
    assert torch.equal(dx2, dx)
    assert_close(dl2, dl_ref, 1e-4, 2e-4, "dlogs + logdet term")
    if mode == "mix":
        assert_close(dw2, dw_ref, 1e-4, 2e-4, "dW + logdet term")
    # bf16 conv1-dgrad operand (the tcgen05 training path): taps widen exactly to fp32 and are summed in the same
    # order, so the result is bit-identical to feeding the same values as fp32
    if with_da1:
        da1_b = da1.bfloat16()
        outs = []
        for operand in (da1_b, da1_b.float()):
            dwb = torch.zeros(c * c, device=DEV) if mode == "mix" else None
            dlb, dbb = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
            dxb = K.rows_actnorm_mix_bwd(to_rows(x), to_rows(dz), n, h, w, da1=operand, cin=cin, weight=wgt, indices=idx,
                                         bias=bias, logs=logs, dw=dwb, dlogs=dlb, dbias=dbb)
            outs.append((dxb, dlb, dbb, dwb))
        assert torch.equal(outs[0][0], outs[1][0])
        # (the channel / weight reductions are atomic: their summation order differs from launch to launch)
        assert_close(outs[0][1], outs[1][1], 1e-4, 5e-4, "dlogs bf16 operand")
        assert_close(outs[0][2], outs[1][2], 1e-4, 5e-4, "dbias bf16 operand")
        if mode == "mix":
            assert_close(outs[0][3], outs[1][3], 1e-4, 5e-4, "dW bf16 operand")


# ---------------------------------------------------------------- Split2d pieces
@pytest.mark.parametrize("shape", [(3, 12, 8, 8), (2, 24, 4, 4), (2, 4, 3, 5)])
def test_rows_split2d_pieces(shape):
    n, c, h, w = shape
    ch = c // 2
    x = cu(torch.randn(*shape, generator=g(50)))
    hrows = cu(torch.randn(n * h * w, c, generator=g(51)) * 0.3)
    ld_in = cu(torch.randn(n, generator=g(52)))
    xr = to_rows(x)
    # log-prob of z2 under the learned prior, and the N(0, I) top prior
    assert_close(K.rows_gaussian_logp(hrows, xr, n, h * w, ch, ch, ld_in), K.gaussian_logp(hrows, x, ch, ch, ld_in),
                 2e-6, 1e-4, "logp")
    assert_close(K.rows_gaussian_logp(None, xr, n, h * w, 0, c, None), K.gaussian_logp(None, x, 0, c, None),
                 2e-6, 1e-4, "top prior")
    # reverse: sample z2
    z1 = x[:, :ch].contiguous()
    eps = cu(torch.randn(n, ch, h, w, generator=g(53)))
    ref = K.split2d_sample(hrows, z1, eps)
    got = K.rows_split2d_sample(hrows, to_rows(z1), ch, eps, n, ch, h * w)
    assert torch.equal(to_nchw(got, n, h, w), ref)
    got2 = K.rows_split2d_sample(hrows, xr, c, eps, n, ch, h * w)       # z1 read in place from wider rows
    assert torch.equal(got2, got)
    # backward
    dz1 = cu(torch.randn(n, ch, h, w, generator=g(54)))
    dld = cu(torch.randn(n, generator=g(55)))
    logs_p = cu(torch.randn(c, generator=g(56)) * 0.1)
    dl_ref, db_ref = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    dx_ref, du_ref = K.split2d_bwd(x, hrows, dz1, dld, logs_p, dl_ref, db_ref)
    dx = torch.zeros(n * h * w, c, device=DEV)
    dx[:, :ch] = to_rows(dz1)
    dl, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    du = K.rows_split2d_bwd(xr, hrows, dld, logs_p, dx, dl, db, n, h * w)
    assert torch.equal(to_nchw(dx, n, h, w), dx_ref) and torch.equal(du, du_ref)
    assert_close(dl, dl_ref, 1e-4, 1e-4, "dlogs_p")
    assert_close(db, db_ref, 1e-4, 1e-4, "dbias_p")
    # conv dgrad tap gather-sum
    kp = round_up(9 * ch, 64)
    da = cu(torch.randn(n * h * w, kp, generator=g(57)))
    ref = dx_ref.clone()
    K.tapsum_to_nchw(da, ref, 0, ch, flip=True, accumulate=True)
    K.rows_tapsum(da, dx, 0, ch, n, h, w, flip=True, accumulate=True)
    assert torch.equal(to_nchw(dx, n, h, w), ref)


# ---------------------------------------------------------------- im2col on rows == im2col on NCHW
@pytest.mark.parametrize("shape,c0,cin,ks,flip", [((2, 12, 8, 8), 0, 6, 3, False), ((2, 12, 8, 8), 0, 12, 3, True),
                                                  ((3, 24, 4, 4), 0, 24, 1, False), ((1, 6, 5, 7), 2, 3, 3, True),
                                                  ((2, 24, 4, 4), 0, 12, 3, False), ((2, 48, 4, 4), 0, 48, 3, True),
                                                  ((3, 4, 3, 5), 0, 2, 3, False), ((2, 8, 6, 6), 4, 4, 3, True),
                                                  ((2, 8, 6, 6), 2, 4, 3, False), ((40, 12, 32, 32), 0, 6, 3, False),
                                                  ((1, 96, 4, 4), 0, 96, 3, True), ((300, 24, 16, 16), 0, 12, 3, False)])
@pytest.mark.parametrize("dt", [_C.F32, _C.BF16])
def test_im2col_rows_equals_nchw(shape, c0, cin, ks, flip, dt):
    n, c, h, w = shape
    x = cu(torch.randn(*shape, generator=g(60)))
    ld = round_up(ks * ks * cin, 64)
    ref = K.im2col(x, c0, cin, ks, dt, ld, flip=flip)
    got = K.im2col_rows(to_rows(x), n, h, w, c0, cin, ks, dt, ld, flip=flip)
    assert torch.equal(got, ref)
    # and against plain torch unfold for the fp32, un-flipped case
    if dt == _C.F32 and not flip:
        cols = torch.nn.functional.unfold(x[:, c0:c0 + cin], ks, padding=(ks - 1) // 2)      # [n, cin*k*k, hw]
        cols = cols.reshape(n, cin, ks * ks, h * w).permute(0, 3, 2, 1).reshape(n * h * w, ks * ks * cin)
        assert torch.equal(got[:, :ks * ks * cin], cols) and float(got[:, ks * ks * cin:].abs().sum()) == 0.0


# ---------------------------------------------------------------- batched weight packing / unpacking
def test_batched_pack_and_unpack_match_single():
    np.random.seed(0); torch.manual_seed(0)
    flow = G.FlowModel((16, 16, 3), 64, K=2, L=2, permutation="invconv", coupling="affine")
    sd = flow.state_dict()
    randomize_(sd, 3)
    flow.load_state_dict(sd)
    flow = flow.to(DEV)
    flow.set_conv_dtype("bf16")
    rows_path.prepare_packs(flow, True, torch.device(DEV))
    steps, splits = rows_path._steps_and_splits(flow)
    for st in steps:
        net = st.f
        for key, prm, layout, rows, ld in net.pack_specs(_C.BF16, True):
            got = net._packs._d[(key, _C.BF16)][1]
            ref = K.pack_conv_weight(prm.detach(), layout, _C.BF16, rows, ld)
            assert torch.equal(got, ref), key
    for sp in splits:
        for conv, key, prm, layout, rows, ld in sp.pack_specs(_C.BF16, True):
            assert torch.equal(conv._packs._d[(key, _C.BF16)][1], K.pack_conv_weight(prm.detach(), layout, _C.BF16, rows, ld))
    # gradients: scratch slices -> .grad (accumulating)
    plan = rows_path.grad_plan(flow, torch.device(DEV))
    plan.begin()
    gen = torch.Generator().manual_seed(9)
    expect = {}
    for owner, tag, prm, layout, rows, ld in plan.items:
        v = plan.view(owner, tag)
        v.copy_(torch.randn(v.shape, generator=gen))
        prm.grad.fill_(0.5)
        ref = torch.full_like(prm, 0.5)
        K.unpack_weight_grad(v, prm.shape[0], prm.shape[1], prm.shape[2], layout, ref, accumulate=True)
        expect[id(prm)] = ref
    plan.finish()
    for owner, tag, prm, layout, rows, ld in plan.items:
        assert torch.equal(prm.grad, expect[id(prm)]), tag


# ---------------------------------------------------------------- whole model: rows path == per-layer NCHW path
def _flow(perm, coup, K_=3, L=3, hidden=32, shape=(16, 16, 3), seed=0):
    np.random.seed(seed); torch.manual_seed(seed)
    flow = G.FlowModel(shape, hidden, K=K_, L=L, permutation=perm, coupling=coup)
    sd = flow.state_dict()
    randomize_(sd, seed + 1)
    flow.load_state_dict(sd)
    for m in flow.modules():
        if isinstance(m, G.ActNorm):
            m.bias_inited = m.logs_inited = True
    return flow.to(DEV)


@pytest.mark.parametrize("perm,coup", [("invconv", "affine"), ("shuffle", "additive"), ("reverse", "affine")])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_flowmodel_rows_equals_nchw_path(perm, coup, dtype):
    hidden = 64 if dtype == "bf16" else 32
    flow = _flow(perm, coup, hidden=hidden).eval()
    flow.set_conv_dtype(dtype)
    x = cu(torch.rand(4, 3, 16, 16, generator=g(70)))
    ld0 = cu(torch.randn(4, generator=g(71)))
    assert rows_path.supported(flow, x)
    with torch.no_grad():
        z_r, ld_r = flow(x, logdet=ld0)
        eps = [cu(torch.randn(4, 12, 4, 4, generator=g(72))), cu(torch.randn(4, 6, 8, 8, generator=g(73)))]
        x_r = flow.decode(z_r.clone(), eps_list=eps)
        config.use_rows_path = False
        try:
            z_n, ld_n = flow(x, logdet=ld0)
            x_n = flow.decode(z_n.clone(), eps_list=eps)
        finally:
            config.use_rows_path = True
    assert torch.equal(z_r, z_n)
    assert_close(ld_r, ld_n, 2e-6, 1e-3, "logdet")
    assert torch.equal(x_r, x_n)


@pytest.mark.parametrize("perm,coup", [("invconv", "affine"), ("shuffle", "additive")])
def test_flowmodel_rows_gradients_equal_nchw_path(perm, coup):
    flow = _flow(perm, coup).train()
    flow.set_conv_dtype("fp32")
    x = cu(torch.rand(4, 3, 16, 16, generator=g(80))).requires_grad_(True)
    ld0 = cu(torch.randn(4, generator=g(81)))

    def run():
        for p in flow.parameters():
            p.grad = None
        x.grad = None
        z, ld = flow(x, logdet=ld0)
        loss = (ld.sum() * 1e-2 + (z * z).sum())
        loss.backward()
        return z.detach().clone(), ld.detach().clone(), {k: p.grad.clone() for k, p in flow.named_parameters()}, x.grad.clone()

    z_r, ld_r, g_r, dx_r = run()
    config.use_rows_path = False
    try:
        z_n, ld_n, g_n, dx_n = run()
    finally:
        config.use_rows_path = True
    assert torch.equal(z_r, z_n)
    assert_close(ld_r, ld_n, 2e-6, 1e-3, "logdet")
    assert_close(dx_r, dx_n, 1e-5, 1e-6, "dx")
    worst = 0.0
    for k in g_n:
        e = float((g_r[k] - g_n[k]).abs().max() / max(float(g_n[k].abs().max()), 1e-6))
        worst = max(worst, e)
        assert e < 2e-4, "%s: %.3e" % (k, e)
    print("rows vs NCHW parameter gradients: worst rel err %.2e over %d tensors" % (worst, len(g_n)))


@pytest.mark.parametrize("perm", ["invconv", "shuffle"])
def test_wide_levels_match_the_pure_nchw_path(perm):
    """L=5 (12 ... 192 channels).  invconv: the 192-channel level runs on the pixel-major path with ActNorm + 1x1
    conv as fp32 GEMMs.  shuffle: levels 1-4 run on the pixel-major kernels, level 5 on the per-layer NCHW kernels
    (hybrid, one autograd node).  Encode and every gradient (randomised weights) and decode with supplied noise
    (fresh weights: zero-initialised couplings keep the inverse of mismatched latents bounded) equal the all-NCHW path."""
    flow = _flow(perm, "affine", K_=2, L=5, hidden=32, shape=(32, 32, 3), seed=3).train()
    flow.set_conv_dtype("fp32")
    x = cu(torch.rand(3, 3, 32, 32, generator=g(80)))
    hd = rows_path.head(flow, x)
    if perm == "invconv":
        assert rows_path.supported(flow, x) and hd is None
    else:
        assert hd is not None and not rows_path.supported(flow, x)
        k = hd[1]
        assert isinstance(flow.layers[k - 1], G.Split2d) and flow.layers[k + 1].in_channels == 192
    np.random.seed(4); torch.manual_seed(4)
    fresh = G.FlowModel((32, 32, 3), 32, K=2, L=5, permutation=perm, coupling="affine")
    for m in fresh.modules():
        if isinstance(m, G.ActNorm):
            m.bias_inited = m.logs_inited = True
    fresh = fresh.to(DEV).eval()
    fresh.set_conv_dtype("fp32")
    eps, c, hh = [], 3, 32
    for lvl in range(4):
        c, hh = c * 4, hh // 2
        eps.append(cu(torch.randn(3, c // 2, hh, hh, generator=g(90 + lvl))))
        c //= 2
    eps = eps[::-1]                                                          # deepest split first
    res = {}
    for use_rows in (True, False):
        config.use_rows_path = use_rows
        try:
            flow.zero_grad(set_to_none=True)
            z, ld = flow(x, torch.zeros(3, device=DEV))
            (z.square().sum() + 0.01 * ld.sum()).backward()
            grads = {n_: p_.grad.clone() for n_, p_ in flow.named_parameters()}
            with torch.no_grad():
                z0, ld0 = flow(x, torch.zeros(3, device=DEV))
                zf, _ = fresh(x, torch.zeros(3, device=DEV))
                xr = fresh.decode(zf.clone(), eps_std=1.0, eps_list=eps)
            res[use_rows] = (z.detach(), ld.detach(), grads, z0, xr)
        finally:
            config.use_rows_path = True
    a, b = res[True], res[False]
    assert rel_err(a[0], b[0]) < 1e-5 and rel_err(a[1], b[1]) < 1e-5 and rel_err(a[3], b[3]) < 1e-5
    assert bool(torch.isfinite(b[4]).all()) and rel_err(a[4], b[4]) < 1e-5
    for n_ in a[2]:
        assert rel_err(a[2][n_], b[2][n_]) < 2e-4, n_


def test_rows_path_falls_back_for_wide_levels():
    flow = G.FlowModel((16, 16, 3), 32, K=1, L=4)          # last level has 3*4*8 = 96 ... 192 channels
    assert max(l.in_channels for l in flow.layers if isinstance(l, G.FlowStep)) == 96
    assert rows_path.supported(flow.to(DEV), torch.zeros(1, 3, 16, 16, device=DEV))
    flow5 = G.FlowModel((32, 32, 3), 32, K=1, L=5, permutation="reverse").to(DEV)     # 192-channel permutation
    assert not rows_path.supported(flow5, torch.zeros(1, 3, 32, 32, device=DEV))
    assert rows_path.supported(G.FlowModel((32, 32, 3), 32, K=1, L=5).to(DEV), torch.zeros(1, 3, 32, 32, device=DEV))
    with torch.no_grad():
        z, ld = flow5.eval()(torch.rand(2, 3, 32, 32, device=DEV), logdet=torch.zeros(2, device=DEV))
    assert z.shape == (2, 192, 1, 1) and ld.shape == (2,)


# ---------------------------------------------------------------- kernels kept behind A/B switches
def test_alternate_kernels_behind_ab_switches():
    """The direct-gather mix adjoint, the warp-per-pixel im2col and the shared-memory-window coupling kernel stay
    selectable through environment switches read at first use (DESIGN.md section 5 and the negative results): run their
    parity tests in a fresh process with the switches set so that they do not rot."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, GLOWK_MIXBWD_NOWIN="1", GLOWK_IM2COL_WARP="1", GLOWK_COUPLING_WIN="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_rows.py"), "-m", "gpu", "-q", "-x",
                        "-k", "rows_coupling_matches_nchw or rows_mix_bwd_matches_nchw or im2col_rows_equals_nchw or "
                              "flowmodel_rows_equals_nchw_path or flowmodel_rows_gradients_equal_nchw_path"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("c,h,w,affine,perm", [(12, 32, 32, True, False), (24, 16, 16, True, False), (48, 8, 8, True, False),
                                               (12, 10, 6, False, False), (24, 5, 7, True, True), (96, 4, 4, False, False)])
def test_rows_coupling_rev_mix_matches_two_launches(c, h, w, affine, perm):
    """One-launch reverse FlowStep tail (inverse coupling + W^-1 mix + ActNorm^-1, model.py:131-152) == glowk_rows_coupling
    (reverse) followed by glowk_rows_actnorm_mix (reverse), bit for bit."""
    n = 3
    g = torch.Generator().manual_seed(c * 100 + h)
    cout = c if affine else c // 2
    ldp = K.round_up(9 * cout, 16)
    x = torch.randn(n * h * w, c, generator=g).cuda()
    p3 = (torch.randn(n * h * w, ldp, generator=g) * 0.1).cuda()
    b3 = (torch.randn(cout, generator=g) * 0.1).cuda()
    l3 = (torch.randn(cout, generator=g) * 0.1).cuda()
    bias = (torch.randn(c, generator=g) * 0.2).cuda()
    logs = (torch.randn(c, generator=g) * 0.1).cuda()
    wm = idx = None
    if perm:
        idx = torch.randperm(c, generator=g).cuda()
    else:
        wm = (torch.eye(c) + 0.05 * torch.randn(c, c, generator=g)).cuda()
    ref = x.clone()
    K.rows_coupling(p3, b3, l3, ref, n, h, w, affine, True, 3.0)
    ref = K.rows_actnorm_mix(ref, wm, idx, bias, logs, 3.0, reverse=True)
    x0 = x.clone()
    got = K.rows_coupling_rev_mix(p3, b3, l3, x, n, h, w, affine, 3.0, wm, idx, bias, logs, 3.0)
    torch.cuda.synchronize()
    assert torch.equal(x, x0)                       # the input rows are not modified
    assert torch.equal(got, ref), (got - ref).abs().max().item()


# ---------------------------------------------------------------- a FlowStep called as a layer
@pytest.mark.parametrize("perm,coup,c,hw,hidden,lu", [("invconv", "affine", 12, 16, 32, False),
                                                      ("shuffle", "additive", 24, 8, 32, False),
                                                      ("invconv", "affine", 48, 4, 64, True)])
def test_standalone_flowstep_rows_route_matches_nchw_route_and_oracle(monkeypatch, perm, coup, c, hw, hidden, lu):
    """FlowStep.forward called as a layer (the reference's FlowModel driving this package's layers, INTEGRATION.md
    section 1) runs on the pixel-major kernels behind a layout change: forward, reverse (with the reference's
    logdet conventions: None / number / [N] tensor) and all gradients against the per-layer NCHW route
    (GLOWK_LAYER_ROWS=0) and, for the dense 1x1 conv, the oracle."""
    np.random.seed(c); torch.manual_seed(c)
    fs = G.FlowStep(c, hidden, permutation=perm, coupling=coup, lu_decomposition=lu)
    sd = randomize_({k: v.clone() for k, v in fs.state_dict().items()}, 31, coupling_std=0.02)
    fs.load_state_dict(sd)
    for m in fs.modules():
        if isinstance(m, G.ActNorm):
            m.bias_inited = m.logs_inited = True
    fs.conv_dtype = "fp32"
    fs = fs.to(DEV).train()
    n = 3
    x = torch.randn(n, c, hw, hw, generator=g(5))
    ld0 = torch.randn(n, generator=g(6))
    wz = cu(torch.randn(n, c, hw, hw, generator=g(7)))
    wl = cu(torch.randn(n, generator=g(8)))
    res = {}
    for route in ("1", "0"):
        monkeypatch.setenv("GLOWK_LAYER_ROWS", route)
        assert fs._rows_route(cu(x)) == (route == "1")
        with torch.no_grad():
            z, ld = fs(cu(x), cu(ld0))
            z_none, ld_none = fs(cu(x), None)
            z_s, ld_s = fs(cu(x), 0.5)
            xr, ldr = fs(z.clone(), ld.clone(), reverse=True)
            xr_none, _ = fs(z.clone(), None, reverse=True)
        assert ld_none is None and torch.equal(z_none, z) and torch.equal(z_s, z)
        assert ld_s.dim() == (1 if coup == "affine" else 0)           # reference broadcasting semantics
        fs.zero_grad(set_to_none=True)
        xg = cu(x).requires_grad_(True)
        zz, ll = fs(xg, cu(ld0))
        ((zz * wz).sum() + (ll * wl).sum()).backward()
        res[route] = dict(z=z, ld=ld, ld_s=ld_s.reshape(-1)[:1], xr=xr, ldr=ldr, xr_none=xr_none, dx=xg.grad.clone(),
                          grads={k: p.grad.clone() for k, p in fs.named_parameters() if p.grad is not None})
    a, b = res["1"], res["0"]
    for k in ("z", "ld", "ld_s", "xr", "ldr", "xr_none", "dx"):
        assert rel_err(a[k], b[k]) < 2e-5, k
    assert rel_err(a["xr"], x) < 1e-4 and rel_err(a["ldr"], ld0) < 1e-4      # round trip
    assert a["grads"].keys() == b["grads"].keys() and len(a["grads"]) >= 9
    for k in a["grads"]:
        ga, gb = a["grads"][k].double().cpu(), b["grads"][k].double().cpu()
        assert float((ga - gb).abs().max()) <= 2e-4 * max(float(gb.abs().max()), 1e-6), k
    if not lu and perm == "invconv":
        z_ref, ld_ref = O.flowstep(x, ld0.clone(), sd, "", perm, coup)
        assert rel_err(a["z"], z_ref) < 1e-4 and rel_err(a["ld"], ld_ref) < 1e-4
