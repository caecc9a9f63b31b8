"""GPU parity, kernel level: every glowk entry point (through the C ABI) vs the CPU oracle /
plain fp32 torch on the same seeded inputs.  Bit-exact where the op is an index map."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close, rel_err
from oracle import glow_oracle as O
from pytorch_glow_b200 import _C
from pytorch_glow_b200 import functional as K

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def g(seed):
    return torch.Generator().manual_seed(seed)


def cu(t):
    return t.to(DEV)


# ---------------------------------------------------------------- ActNorm
@pytest.mark.parametrize("shape", [(3, 4, 5, 6), (2, 12, 32, 32), (5, 48, 8, 8), (1, 6, 3, 5), (0, 4, 2, 2)])
def test_actnorm_fwd_rev(shape):
    x = torch.randn(*shape, generator=g(1)) * 2 + 0.5
    c = shape[1]
    bias = torch.randn(1, c, 1, 1, generator=g(2)) * 0.3
    logs = torch.randn(1, c, 1, 1, generator=g(3)) * 0.2
    y_ref, _ = O.actnorm(x, bias, logs)
    y = K.actnorm(cu(x), cu(bias).reshape(-1), cu(logs).reshape(-1), 3.0, False)
    assert_close(y, y_ref, 1e-6, 1e-6, "fwd")
    xr_ref, _ = O.actnorm(y_ref, bias, logs, reverse=True)
    xr = K.actnorm(cu(y_ref), cu(bias).reshape(-1), cu(logs).reshape(-1), 3.0, True)
    assert_close(xr, xr_ref, 1e-6, 1e-6, "rev")


@pytest.mark.parametrize("shape,scale", [((3, 4, 5, 6), 1.3), ((16, 12, 16, 16), 1.0), ((2, 48, 4, 4), 1.0)])
def test_actnorm_init(shape, scale):
    x = torch.randn(*shape, generator=g(4)) * 2 + 1
    b_ref, l_ref = O.actnorm_init(x, scale)
    b, l = K.actnorm_init_nchw(cu(x), scale)
    assert_close(b, b_ref.reshape(-1), 1e-5, 1e-6, "bias")
    assert_close(l, l_ref.reshape(-1), 1e-5, 1e-6, "logs")
    n, c, h, w = shape
    rows = x.permute(0, 2, 3, 1).reshape(-1, c)
    rows_p = torch.zeros(rows.shape[0], c + 3)
    rows_p[:, :c] = rows
    b2, l2 = K.actnorm_init_rows(cu(rows_p), c, scale)
    assert_close(b2, b_ref.reshape(-1), 1e-5, 1e-6, "bias(rows)")
    assert_close(l2, l_ref.reshape(-1), 1e-5, 1e-6, "logs(rows)")


def test_actnorm_init_variants_match_reference():
    """ActNorm(batch_variance=True) and the data-dependent init arriving with reverse=True (network/module.py:44-45,
    62-63, 112-113, 143-146): module outputs and initialised parameters against the reference's (actnorm_init.npz)."""
    from conftest import Golden
    G_ = Golden("actnorm_init.npz")
    import pytorch_glow_b200 as G
    n = 0
    for c in (12, 48):
        for bv in (0, 1):
            for rev in (0, 1):
                tag = "c%d_bv%d_rev%d/" % (c, bv, rev)
                if not G_.has(tag + "x"):
                    continue
                an = G.ActNorm(c, scale=1.3, logscale_factor=3., batch_variance=bool(bv)).to("cuda:0").train()
                with torch.no_grad():
                    y, ld = an(cu(G_.t(tag + "x")), torch.zeros(6, device="cuda:0"), reverse=bool(rev))
                assert an.bias_inited and an.logs_inited
                assert_close(an.bias, G_.t(tag + "bias"), 1e-5, 1e-6, tag + "bias")
                assert_close(an.logs, G_.t(tag + "logs"), 1e-5, 1e-6, tag + "logs")
                assert_close(y, G_.t(tag + "y"), 1e-5, 1e-5, tag + "y")
                assert_close(ld, G_.t(tag + "logdet"), 1e-5, 1e-4, tag + "logdet")
                if bv:                              # the rows entry point (Conv2d's ActNorm, FlowStep on the rows path)
                    x = G_.t(tag + "x")
                    rows = cu(x.permute(0, 2, 3, 1).reshape(-1, c).contiguous())
                    b2, l2 = K.actnorm_init_rows(rows, c, 1.3, 3.0, batch_variance=True, reverse=bool(rev))
                    assert_close(b2, G_.t(tag + "bias").reshape(-1), 1e-5, 1e-6, tag + "bias(rows)")
                    assert_close(l2, G_.t(tag + "logs").reshape(-1), 1e-5, 1e-6, tag + "logs(rows)")
                n += 1
    assert n == 6


def test_flowstep_reverse_direction_init_matches_oracle():
    """A training-mode FlowStep whose first call is reverse_flow: the step's ActNorm initialises from the un-mixed tensor
    (network/model.py:119-154 + module.py:143-146).  NCHW layer call and the rows path of FlowModel.decode."""
    import pytorch_glow_b200 as G
    from parity_util import randomize_
    torch.manual_seed(3); np.random.seed(3)
    c, hw, n = 12, 8, 4
    fs = G.FlowStep(c, 32, permutation="invconv", coupling="affine")
    sd = randomize_({k: v.clone() for k, v in fs.state_dict().items()}, 9, coupling_std=0.02)
    fs.load_state_dict(sd)
    for m in fs.modules():
        if isinstance(m, G.ActNorm):
            m.bias_inited = m.logs_inited = True
    fs.actnorm.bias_inited = fs.actnorm.logs_inited = False
    fs.conv_dtype = "fp32"
    fs = fs.to("cuda:0").train()
    z = torch.randn(n, c, hw, hw, generator=g(77)) * 1.5 + 0.3
    # oracle: coupling^-1, W^-1, then the reverse-direction init and ActNorm^-1
    z1, z2 = z[:, :c // 2], z[:, c // 2:]
    h = O.f_net(z1, sd, "f.")
    shift, sc = O.split_channel(h, "cross")
    sc = torch.sigmoid(sc + 2.)
    u = torch.cat([z1, z2 / sc - shift], 1)
    y, _ = O.invconv(u, sd["invconv.weight"], reverse=True)
    b, l = O.actnorm_init_reverse(y)
    x_ref, _ = O.actnorm(y, b, l, reverse=True)
    with torch.no_grad():
        x, _ = fs(cu(z).clone(), None, reverse=True)
    assert fs.actnorm.bias_inited
    assert_close(fs.actnorm.bias, b, 1e-4, 1e-5, "bias")
    assert_close(fs.actnorm.logs, l, 1e-4, 1e-5, "logs")
    assert_close(x, x_ref, 1e-4, 1e-4, "x")


# ---------------------------------------------------------------- 1x1 conv weight prep
@pytest.mark.parametrize("c", [2, 6, 12, 24, 48, 96, 130, 192, 384])
def test_invconv_prepare(c):
    np.random.seed(c)
    w = O.invconv_init_weight(c) + 0.1 * torch.randn(c, c, generator=g(c))
    ld, winv = K.invconv_prepare(cu(w), True)
    ref = torch.log(torch.abs(torch.det(w.double())))
    assert abs(float(ld) - float(ref)) < 1e-4 * max(1.0, abs(float(ref)))
    assert_close(winv, torch.linalg.inv(w.double()).float(), 1e-4, 1e-5, "inverse")
    ld2, none = K.invconv_prepare(cu(w), False)
    assert none is None and float(ld2) == float(ld)


@pytest.mark.parametrize("c", [4, 12, 48])
def test_lu_assemble(c):
    w = torch.randn(c, c, generator=g(10 + c))
    p, l, u, sign_s, log_s = O.lu_factor(w)
    w_ref, ld_ref = O.lu_assemble(p, l, u, sign_s, log_s)
    wk, winv, ld = K.invconv_lu_assemble(cu(p), cu(l), cu(u), cu(sign_s), cu(log_s), True)
    assert_close(wk, w_ref, 1e-5, 1e-6, "W")
    assert_close(ld, ld_ref.reshape(1), 1e-5, 1e-6, "logabsdet")
    assert_close(winv, torch.linalg.inv(w_ref.double()).float(), 2e-4, 1e-5, "W^-1")


# ---------------------------------------------------------------- fused actnorm + mix / permutation
MIX_SHAPES = [(2, 12, 32, 32), (3, 24, 16, 16), (2, 48, 8, 8), (5, 48, 4, 4), (2, 8, 4, 6), (1, 6, 3, 5),
              (2, 96, 4, 4), (1, 16, 2, 2), (1, 320, 2, 2)]


@pytest.mark.parametrize("shape", MIX_SHAPES)
def test_actnorm_mix(shape):
    n, c, h, w = shape
    x = torch.randn(*shape, generator=g(20))
    bias = torch.randn(1, c, 1, 1, generator=g(21)) * 0.3
    logs = torch.randn(1, c, 1, 1, generator=g(22)) * 0.2
    np.random.seed(5)
    wm = O.invconv_init_weight(c) + 0.05 * torch.randn(c, c, generator=g(23))
    a_ref, _ = O.actnorm(x, bias, logs)
    z_ref, _ = O.invconv(a_ref, wm)
    z = K.actnorm_mix(cu(x), weight=cu(wm), bias=cu(bias).reshape(-1), logs=cu(logs).reshape(-1))
    assert_close(z, z_ref, 1e-5, 1e-5, "fwd mix")
    # reverse: x = actnorm^-1(W^-1 z)
    winv = torch.linalg.inv(wm.double()).float().contiguous()
    zz, _ = O.invconv(z_ref, wm, reverse=True)
    xr_ref, _ = O.actnorm(zz, bias, logs, reverse=True)
    xr = K.actnorm_mix(cu(z_ref), weight=cu(winv), bias=cu(bias).reshape(-1), logs=cu(logs).reshape(-1), reverse=True)
    tol = 1e-4 if c <= 96 else 1e-3          # the oracle inverts W in fp32; wide W is worse conditioned
    assert_close(xr, xr_ref, tol, tol, "rev mix")
    # plain Invertible1x1Conv (no actnorm)
    z2 = K.actnorm_mix(cu(x), weight=cu(wm))
    assert_close(z2, O.invconv(x, wm)[0], 1e-5, 1e-5, "plain mix")


@pytest.mark.parametrize("shape", MIX_SHAPES)
def test_actnorm_permutation_bit_exact(shape):
    n, c, h, w = shape
    x = torch.randn(*shape, generator=g(30))
    np.random.seed(7)
    idx, inv = O.permutation_indices(c, shuffle=True)
    y = K.actnorm_mix(cu(x), indices=cu(torch.from_numpy(idx)))
    assert torch.equal(y.cpu(), O.permute(x, idx, inv))                       # bit-exact gather
    xr = K.actnorm_mix(y, indices=cu(torch.from_numpy(inv)))
    assert torch.equal(xr.cpu(), x)
    bias = torch.randn(1, c, 1, 1, generator=g(31)) * 0.3
    logs = torch.randn(1, c, 1, 1, generator=g(32)) * 0.2
    a_ref, _ = O.actnorm(x, bias, logs)
    yf = K.actnorm_mix(cu(x), indices=cu(torch.from_numpy(idx)), bias=cu(bias).reshape(-1), logs=cu(logs).reshape(-1))
    assert_close(yf, O.permute(a_ref, idx, inv), 1e-6, 1e-6, "actnorm+perm")


# ---------------------------------------------------------------- squeeze
@pytest.mark.parametrize("shape", [(2, 3, 4, 6), (4, 3, 64, 64), (3, 12, 16, 16), (1, 1, 2, 2), (0, 3, 4, 4)])
def test_squeeze_bit_exact(shape):
    x = torch.randn(*shape, generator=g(40))
    y = K.squeeze2d(cu(x), 2, reverse=False)
    assert torch.equal(y.cpu(), O.squeeze2d(x, 2))
    xr = K.squeeze2d(y, 2, reverse=True)
    assert torch.equal(xr.cpu(), x)


def test_squeeze_known_vectors_and_strided_batch():
    y = K.squeeze2d(cu(torch.arange(16.).view(1, 1, 4, 4)), 2)
    assert y.flatten().tolist() == [0, 2, 8, 10, 1, 3, 9, 11, 4, 6, 12, 14, 5, 7, 13, 15]
    x = torch.randn(3, 8, 4, 4, generator=g(41))
    view = cu(x)[:, :4]                                # what Split2d returns (module.py:111)
    assert not view.is_contiguous()
    assert torch.equal(K.squeeze2d(view, 2).cpu(), O.squeeze2d(x[:, :4].contiguous(), 2))
    with pytest.raises(ValueError):
        K.squeeze2d(cu(torch.zeros(1, 3, 3, 4)), 2)     # module.py:588


# ---------------------------------------------------------------- im2col / weight packing / tap sum
def _unfold_taps(x, ks):
    n, c, h, w = x.shape
    u = F.unfold(x, ks, padding=(ks - 1) // 2)                        # [N, C*k*k, HW], order (c, tap)
    u = u.view(n, c, ks * ks, h * w).permute(0, 3, 2, 1)              # [N, HW, tap, c]
    return u.reshape(n * h * w, ks * ks * c)


@pytest.mark.parametrize("dtype", [_C.F32, _C.BF16])
@pytest.mark.parametrize("ks", [1, 3])
def test_im2col(dtype, ks):
    x = torch.randn(2, 10, 5, 6, generator=g(50))
    c0, cin = 2, 6
    ld = 64
    rows = K.im2col(cu(x), c0, cin, ks, dtype, ld).float().cpu()
    ref = _unfold_taps(x[:, c0:c0 + cin], ks)
    if dtype == _C.BF16:
        ref = ref.bfloat16().float()
    assert torch.equal(rows[:, :ref.shape[1]], ref)
    assert float(rows[:, ref.shape[1]:].abs().max()) == 0.0
    # flipped taps (transposed conv)
    rows_f = K.im2col(cu(x), c0, cin, ks, _C.F32, ld, flip=True).cpu()
    ref_f = _unfold_taps(x[:, c0:c0 + cin], ks).view(-1, ks * ks, cin).flip(1).reshape(-1, ks * ks * cin)
    assert torch.equal(rows_f[:, :ref_f.shape[1]], ref_f)
    # pixel-major source
    src_rows = x.permute(0, 2, 3, 1).reshape(-1, 10).contiguous()
    rows_r = K.im2col_rows(cu(src_rows), 2, 5, 6, c0, cin, ks, _C.F32, ld).cpu()
    assert torch.equal(rows_r[:, :ref.shape[1]], _unfold_taps(x[:, c0:c0 + cin], ks))


def test_pack_conv_weight_layouts():
    w = torch.randn(5, 3, 3, 3, generator=g(51))
    o, i = 5, 3
    wt = w.view(o, i, 9)
    l0 = K.pack_conv_weight(cu(w), 0, _C.F32, 16, 64).cpu()
    assert torch.equal(l0[:o, :27], wt.permute(0, 2, 1).reshape(o, 27)) and float(l0[o:].abs().max()) == 0
    l1 = K.pack_conv_weight(cu(w), 1, _C.F32, 48, 8).cpu()
    assert torch.equal(l1[:45, :i], wt.permute(2, 0, 1).reshape(45, i)) and float(l1[:, i:].abs().max()) == 0
    l2 = K.pack_conv_weight(cu(w), 2, _C.F32, 32, 8).cpu()
    assert torch.equal(l2[:27, :o], wt.permute(2, 1, 0).reshape(27, o))
    l3 = K.pack_conv_weight(cu(w), 3, _C.F32, 8, 48).cpu()
    assert torch.equal(l3[:i, :45], wt.permute(1, 2, 0).reshape(i, 45))
    lb = K.pack_conv_weight(cu(w), 0, _C.BF16, 16, 64).float().cpu()
    assert torch.equal(lb[:o, :27], wt.permute(0, 2, 1).reshape(o, 27).bfloat16().float())


def test_tapsum_and_rows_to_nchw():
    n, c, h, w = 2, 5, 4, 6
    p = torch.randn(n * h * w, 48, generator=g(52))
    # reference: conv expressed as nine shifted adds
    pv = p[:, :45].view(n, h, w, 9, c)
    ref = torch.zeros(n, c, h, w)
    for t in range(9):
        dy, dx = t // 3 - 1, t % 3 - 1
        for y in range(h):
            for x in range(w):
                if 0 <= y + dy < h and 0 <= x + dx < w:
                    ref[:, :, y, x] += pv[:, y + dy, x + dx, t, :]
    dst = torch.zeros(n, 7, h, w, device=DEV)
    K.tapsum_to_nchw(cu(p), dst, 1, c)
    assert_close(dst[:, 1:6], ref, 1e-6, 1e-6, "tapsum")
    assert float(dst[:, 0].abs().max()) == 0 and float(dst[:, 6].abs().max()) == 0
    rows = torch.randn(n * h * w, 8, generator=g(53))
    out = K.rows_to_nchw(cu(rows), n, c, h, w).cpu()
    assert torch.equal(out, rows[:, :c].view(n, h, w, c).permute(0, 3, 1, 2))


# ---------------------------------------------------------------- GEMMs
def _gemm_case(m, n, k, dtype, epi, out_dtype, seed=60):
    a = torch.randn(m, k, generator=g(seed)) * 0.5
    b = torch.randn(n, k, generator=g(seed + 1)) * 0.2
    bias = torch.randn(n, generator=g(seed + 2)) * 0.3
    logs = torch.randn(n, generator=g(seed + 3)) * 0.1
    if dtype == _C.BF16:
        a, b = a.bfloat16(), b.bfloat16()
    acc = a.double() @ b.double().t()
    if epi == _C.EPI_STORE:
        ref = acc
    else:
        ref = (acc + bias.double()) * torch.exp(3.0 * logs.double())
        if epi == _C.EPI_ACTNORM_RELU:
            ref = ref.clamp_min(0)
    out = K.gemm(cu(a), cu(b), n, k, epi, cu(bias), cu(logs), 3.0, out_dtype=out_dtype,
                 ldo=(n + 7) // 8 * 8)
    torch.cuda.synchronize()
    return out[:, :n].float().cpu(), ref.float()


@pytest.mark.parametrize("m,n,k", [(64, 64, 16), (130, 50, 40), (1000, 112, 512), (257, 512, 64), (0, 8, 8)])
@pytest.mark.parametrize("epi", [_C.EPI_STORE, _C.EPI_ACTNORM_RELU, _C.EPI_ZEROS])
def test_gemm_fp32(m, n, k, epi):
    out, ref = _gemm_case(m, n, k, _C.F32, epi, _C.F32)
    assert_close(out, ref, 1e-5, 1e-5, "gemm fp32")


TC_SHAPES = [(128, 16, 64), (256, 512, 512), (1000, 112, 512), (130, 448, 512), (4096, 512, 64),
             (77, 64, 128), (512, 224, 192), (33000, 512, 512), (128, 48, 256)]


@pytest.mark.parametrize("m,n,k", TC_SHAPES)
@pytest.mark.parametrize("epi,out_dtype", [(_C.EPI_STORE, _C.F32), (_C.EPI_ACTNORM_RELU, _C.BF16),
                                           (_C.EPI_ZEROS, _C.F32), (_C.EPI_ACTNORM_RELU, _C.F32)])
def test_gemm_tcgen05(m, n, k, epi, out_dtype):
    """bf16 operands / fp32 accumulate: exact up to fp32 summation order (+ bf16 rounding of the output)."""
    assert _C.has_tcgen05(), "tcgen05 path must be available on the B200 box"
    out, ref = _gemm_case(m, n, k, _C.BF16, epi, out_dtype)
    if out_dtype == _C.BF16:
        assert_close(out, ref, 1e-2, 1e-2, "gemm tcgen05 (bf16 out)")      # 2^-8 output rounding
    else:
        assert_close(out, ref, 1e-4, 1e-4, "gemm tcgen05 (fp32 out)")


@pytest.mark.parametrize("dtype", [_C.F32, _C.BF16])
def test_gemm_relu_bwd(dtype):
    m, n, k = 700, 512 if dtype == _C.BF16 else 96, 128
    a = torch.randn(m, k, generator=g(70)) * 0.5
    b = torch.randn(n, k, generator=g(71)) * 0.2
    y = torch.randn(m, n, generator=g(72)).clamp_min(0)
    logs = torch.randn(n, generator=g(73)) * 0.1
    if dtype == _C.BF16:
        a, b, y = a.bfloat16(), b.bfloat16(), y.bfloat16()
    acc = a.double() @ b.double().t()
    gmask = torch.where(y.double() > 0, acc, torch.zeros_like(acc))
    s = torch.exp(3.0 * logs.double())
    ref_out = gmask * s
    ref_dlogs = 3.0 * (gmask * y.double()).sum(0)
    ref_dbias = s * gmask.sum(0)
    dlogs = torch.zeros(n, device=DEV)
    dbias = torch.zeros(n, device=DEV)
    out = K.gemm(cu(a), cu(b), n, k, _C.EPI_RELU_BWD, None, cu(logs), 3.0, y=cu(y), dlogs=dlogs, dbias=dbias,
                 out_dtype=dtype)
    tol = 1e-2 if dtype == _C.BF16 else 1e-5
    assert_close(out.float(), ref_out.float(), tol, tol, "relu_bwd out")
    assert_close(dlogs, ref_dlogs.float(), 1e-3, 1e-2, "dlogs")
    assert_close(dbias, ref_dbias.float(), 1e-3, 1e-2, "dbias")


def test_conv_actnorm_deferred_grads():
    """Conv2d backward on the bf16 path without epilogue column sums: the ActNorm bias gradient comes out of the
    weight-gradient GEMM through a ones column of the im2col operand (glowk_im2col_rows_ones), the logs gradient
    from <W, dW> + b*db (glowk_conv_actnorm_finish_batched).  Checked against the direct fp64 sums of the oracle
    formulas (module.py:34-84, 188-260)."""
    import numpy as np
    from pytorch_glow_b200 import rows_path
    n, h, w, cin, hid = 3, 6, 5, 6, 128
    kp = (9 * cin + 63) // 64 * 64
    ones_col = 9 * cin
    z = torch.randn(n * h * w, 2 * cin, generator=g(90))
    wgt = (torch.randn(hid, kp, generator=g(91)) * 0.2)
    wgt[:, 9 * cin:] = 0                                              # GEMM-layout weight: zero padding columns
    bias = torch.randn(hid, generator=g(92)) * 0.3
    logs = torch.randn(hid, generator=g(93)) * 0.1
    up = torch.randn(n * h * w, hid, generator=g(94)) * 0.5           # gradient wrt the ReLU output
    a_plain = K.im2col_rows(cu(z), n, h, w, 0, cin, 3, _C.BF16, kp)
    a1 = K.im2col_rows(cu(z), n, h, w, 0, cin, 3, _C.BF16, kp, ones_col=ones_col)
    assert torch.equal(a1[:, :ones_col], a_plain[:, :ones_col]) and torch.equal(a1[:, ones_col + 1:], a_plain[:, ones_col + 1:])
    assert float(a_plain[:, ones_col].float().abs().max()) == 0.0 and bool((a1[:, ones_col].float() == 1.0).all())
    wb = cu(wgt).bfloat16()
    s = torch.exp(3.0 * logs.double())
    # forward: the ones column meets a zero weight
    y = K.gemm(a1, wb, hid, kp, _C.EPI_ACTNORM_RELU, cu(bias), cu(logs), 3.0, out_dtype=_C.BF16)
    y_plain = K.gemm(a_plain, wb, hid, kp, _C.EPI_ACTNORM_RELU, cu(bias), cu(logs), 3.0, out_dtype=_C.BF16)
    assert torch.equal(y, y_plain)
    # backward through ReLU + ActNorm: "dgrad of the next layer" stands in as an identity GEMM on `up`
    eye = torch.eye(hid).bfloat16()
    upb = up.bfloat16()
    dl_direct = torch.zeros(hid, device=DEV); db_direct = torch.zeros(hid, device=DEV)
    v_direct = K.gemm(cu(upb), cu(eye), hid, hid, _C.EPI_RELU_BWD, None, cu(logs), 3.0, y=y, dlogs=dl_direct,
                      dbias=db_direct, out_dtype=_C.BF16)
    v = K.gemm(cu(upb), cu(eye), hid, hid, _C.EPI_RELU_BWD, None, cu(logs), 3.0, y=y, dlogs=None, dbias=None,
               out_dtype=_C.BF16)
    assert torch.equal(v, v_direct)
    dw = torch.zeros(hid, kp, device=DEV)
    K.gemm_wgrad(v, a1, hid, kp, dw)
    dbias = torch.full((hid,), 0.25, device=DEV); dlogs = torch.full((hid,), -0.5, device=DEV)   # accumulate into
    bias_d = cu(bias)
    job = np.array([(wb.data_ptr(), dw.data_ptr(), bias_d.data_ptr(), dw.data_ptr() + 4 * ones_col, dbias.data_ptr(),
                     dlogs.data_ptr(), hid, kp, kp, kp, 3.0, kp)], dtype=rows_path._FJOB)
    K.conv_actnorm_finish_batched(torch.from_numpy(job.view(np.uint8).copy()).to(DEV), 1, hid)
    gm = torch.where(y.double().cpu() > 0, upb.double(), torch.zeros(1, dtype=torch.float64))
    ref_db = (gm * s).sum(0)
    ref_dl = 3.0 * (gm * y.double().cpu()).sum(0)
    # the deferred sums run over the bf16-ROUNDED gradient v (and bf16 W, a1): bf16 tolerances
    assert_close(dbias - 0.25, ref_db.float(), 1e-2, 5e-2, "deferred dbias")
    assert_close(dlogs + 0.5, ref_dl.float(), 3e-2, 1e-1, "deferred dlogs")
    assert_close(dbias - 0.25, db_direct, 1e-2, 5e-2, "deferred vs epilogue dbias")
    assert_close(dlogs + 0.5, dl_direct, 3e-2, 1e-1, "deferred vs epilogue dlogs")


@pytest.mark.parametrize("dtype,p,mo,no", [(_C.F32, 1000, 70, 50), (_C.BF16, 1000, 70, 50), (_C.BF16, 5000, 512, 512),
                                           (_C.BF16, 300, 128, 64), (_C.BF16, 20000, 448, 512),
                                           (_C.BF16, 64 * 1024, 512, 256)])
def test_gemm_wgrad(dtype, p, mo, no):
    a = torch.randn(p, mo, generator=g(80)) * 0.3
    b = torch.randn(p, no, generator=g(81)) * 0.3
    if dtype == _C.BF16:
        a, b = a.bfloat16(), b.bfloat16()
    ref = (a.double().t() @ b.double()).float()
    dw = torch.zeros(mo, no, device=DEV)
    K.gemm_wgrad(cu(a), cu(b), mo, no, dw)
    assert_close(dw, ref, 2e-4, 2e-3 * (p ** 0.5) / 30, "wgrad")
    K.gemm_wgrad(cu(a), cu(b), mo, no, dw)                       # accumulates
    assert_close(dw, 2 * ref, 2e-4, 4e-3 * (p ** 0.5) / 30, "wgrad accumulate")


# ---------------------------------------------------------------- coupling / logdet / prior
def _coupling_ref(p3, bias3, logs3, z, affine, reverse, n, c, h, w):
    cout = c if affine else c // 2
    pv = p3[:, :9 * cout].view(n, h, w, 9, cout)
    u = torch.zeros(n, cout, h, w)
    for t in range(9):
        dy, dx = t // 3 - 1, t % 3 - 1
        ys, ye = max(0, -dy), min(h, h - dy)
        xs, xe = max(0, -dx), min(w, w - dx)
        u[:, :, ys:ye, xs:xe] += pv[:, ys + dy:ye + dy, xs + dx:xe + dx, t, :].permute(0, 3, 1, 2)
    hh = (u + bias3.view(1, -1, 1, 1)) * torch.exp(3.0 * logs3.view(1, -1, 1, 1))
    z1, z2 = z[:, :c // 2], z[:, c // 2:]
    ld = torch.zeros(n)
    if affine:
        shift, scale = hh[:, 0::2], torch.sigmoid(hh[:, 1::2] + 2.0)
        if not reverse:
            z2 = (z2 + shift) * scale
            ld = O.reduce_sum(torch.log(scale), [1, 2, 3])
        else:
            z2 = z2 / scale - shift
            ld = -O.reduce_sum(torch.log(scale), [1, 2, 3])
    else:
        z2 = z2 - hh if reverse else z2 + hh
    return torch.cat((z1, z2), 1), ld, hh


@pytest.mark.parametrize("shape", [(2, 12, 32, 32), (3, 24, 16, 16), (2, 48, 8, 8), (4, 8, 4, 6), (2, 4, 2, 2), (3, 6, 20, 20)])
@pytest.mark.parametrize("affine", [True, False])
@pytest.mark.parametrize("reverse", [False, True])
def test_coupling(shape, affine, reverse):
    n, c, h, w = shape
    cout = c if affine else c // 2
    ldp = (9 * cout + 15) // 16 * 16
    p3 = torch.randn(n * h * w, ldp, generator=g(90)) * 0.3
    bias3 = torch.randn(cout, generator=g(91)) * 0.2
    logs3 = torch.randn(cout, generator=g(92)) * 0.1
    z = torch.randn(*shape, generator=g(93))
    z_ref, ld_ref, h_ref = _coupling_ref(p3, bias3, logs3, z, affine, reverse, n, c, h, w)
    zc = cu(z).clone()
    partials, hs = K.coupling(cu(p3), cu(bias3), cu(logs3), zc, affine, reverse, save_h=True)
    assert_close(zc, z_ref, 1e-5, 1e-5, "z")
    assert torch.equal(zc[:, :c // 2].cpu(), z[:, :c // 2])
    hs_ref = h_ref.permute(0, 2, 3, 1).reshape(-1, cout)
    if affine:   # kernel saves (shift, pre-sigmoid scale) interleaved
        assert_close(hs, hs_ref, 1e-5, 1e-5, "h")
        ld0 = torch.randn(n, generator=g(94))
        out = K.logdet_finish(cu(ld0), n, h * w, partials=partials)
        assert_close(out, ld0 + ld_ref, 1e-5, 1e-4, "logdet")
    else:
        assert partials is None
        assert_close(hs, hs_ref, 1e-5, 1e-5, "h")


def test_logdet_finish_terms():
    n, c, hw = 5, 12, 1024
    ld0 = torch.randn(n, generator=g(95)) * 100
    logs = torch.randn(c, generator=g(96)) * 0.1
    lad = torch.tensor([0.37])
    for sign in (1.0, -1.0):
        ref = ld0 + sign * (torch.sum(logs * 3.0) * hw) + sign * (lad * hw)
        out = K.logdet_finish(cu(ld0), n, hw, logs=cu(logs), logabsdet=cu(lad), sign=sign)
        assert_close(out, ref, 1e-6, 1e-3, "logdet terms")
    out = K.logdet_finish(None, 3, 16, logs=cu(logs), device=DEV)
    assert_close(out, (torch.sum(logs * 3.0) * 16).expand(3), 1e-6, 1e-5)


@pytest.mark.parametrize("shape", [(2, 12, 32, 32), (3, 48, 8, 8), (2, 8, 4, 6)])
def test_gaussian_logp_and_split_sample(shape):
    n, c, h, w = shape
    ch = c // 2
    x = torch.randn(*shape, generator=g(97))
    hrows = torch.randn(n * h * w, c, generator=g(98)) * 0.3
    hn = hrows.view(n, h, w, c).permute(0, 3, 1, 2)
    mean, logs = hn[:, 0::2], hn[:, 1::2]
    ld0 = torch.randn(n, generator=g(99))
    ref = O.gaussian_logp(mean, logs, x[:, ch:]) + ld0
    out = K.gaussian_logp(cu(hrows), cu(x), ch, ch, cu(ld0))
    assert_close(out, ref, 1e-5, 1e-3, "split logp")
    ref0 = O.gaussian_logp(torch.zeros_like(x), torch.zeros_like(x), x)
    out0 = K.gaussian_logp(None, cu(x), 0, c)
    assert_close(out0, ref0, 1e-5, 1e-3, "top prior logp")
    eps = torch.randn(n, ch, h, w, generator=g(100)) * 0.7
    z1 = x[:, :ch].contiguous()
    ref_s = torch.cat((z1, O.gaussian_sample(mean, logs, eps=eps)), 1)
    out_s = K.split2d_sample(cu(hrows), cu(z1), cu(eps))
    assert_close(out_s, ref_s, 1e-6, 1e-6, "split sample")
    assert torch.equal(out_s[:, :ch].cpu(), z1)


def test_errors_are_python_exceptions():
    with pytest.raises(_C.GlowkError):
        K.actnorm(torch.zeros(1, 2, 2, 2), torch.zeros(2), torch.zeros(2))        # CPU tensor: no fallback
    with pytest.raises(ValueError):
        _C.call("glowk_coupling", 1, 8, 1, 1, 3.0, 1, 1, None, 1, 3, 2, 2, 1, 0)     # odd channel count
