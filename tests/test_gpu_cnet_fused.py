"""The fused coupling-network kernels (glowk_cnet_forward / glowk_cnet_backward, csrc/cnet_fused_sm100.cu) against
(i) the three-GEMM path they replace -- bit for bit on the same bf16 operands -- and (ii) a plain fp32 torch
restatement of network/module.py:300-319 in its GEMM form, with the bf16 tolerance written out."""
import pytest
import torch

from pytorch_glow_b200 import _C
from pytorch_glow_b200 import functional as K

pytestmark = pytest.mark.gpu

HID = 512


def _mk(m, k1, n3, seed):
    g = torch.Generator().manual_seed(seed)
    dev = "cuda"
    a1 = (torch.randn(m, k1, generator=g) * 0.5).to(dev).bfloat16()
    w1 = (torch.randn(HID, k1, generator=g) * 0.05).to(dev).bfloat16()
    w2 = (torch.randn(HID, HID, generator=g) * 0.05).to(dev).bfloat16()
    w3 = (torch.randn(n3, HID, generator=g) * 0.05).to(dev).bfloat16()
    vec = lambda s: (torch.randn(HID, generator=g) * s).to(dev)
    return a1, w1, w2, w3, vec(0.1), vec(0.1), vec(0.1), vec(0.1)


def _three_gemms(a1, w1, w2, w3, b1, l1, b2, l2, n3, k1):
    h1 = K.gemm(a1, w1, HID, k1, _C.EPI_ACTNORM_RELU, b1, l1, 3.0, out_dtype=_C.BF16)
    h2 = K.gemm(h1, w2, HID, HID, _C.EPI_ACTNORM_RELU, b2, l2, 3.0, out_dtype=_C.BF16)
    p3 = K.gemm(h2, w3, n3, HID, _C.EPI_STORE, out_dtype=_C.F32)
    return h1, h2, p3


# (pixels, K1 = padded 9*Cin, N3 = padded 9*Cout): level 1 affine (54 -> 64, 108 -> 112), level 1 additive (54 -> 64),
# a ragged last tile, more tiles than SMs, level 2 affine (108 -> 128, 216 -> 224: GEMM3 deferred), K1 = 256
SHAPES = [(2048, 64, 112), (1024, 64, 64), (1000, 64, 112), (148 * 128 * 2 + 128, 64, 112), (768, 128, 224),
          (300, 128, 128), (512, 256, 128), (256, 192, 16)]


@pytest.mark.parametrize("m,k1,n3", SHAPES)
@pytest.mark.parametrize("save", [False, True])
def test_cnet_forward_matches_three_gemms(m, k1, n3, save):
    if not _C.has_tcgen05():
        pytest.skip("needs sm_100")
    assert K.cnet_fused_supported(False, k1, HID, n3)
    a1, w1, w2, w3, b1, l1, b2, l2 = _mk(m, k1, n3, 7 + m)
    h1r, h2r, p3r = _three_gemms(a1, w1, w2, w3, b1, l1, b2, l2, n3, k1)
    p3, h1, h2 = K.cnet_forward(a1, w1, w2, w3, HID, n3, b1, l1, 3.0, b2, l2, 3.0, save=save)
    torch.cuda.synchronize()
    if save:
        assert torch.equal(h1, h1r), "h1 differs: max %g" % (h1.float() - h1r.float()).abs().max().item()
        assert torch.equal(h2, h2r), "h2 differs: max %g" % (h2.float() - h2r.float()).abs().max().item()
    else:
        assert h1 is None and h2 is None
    assert torch.equal(p3, p3r), "p3 differs: max %g" % (p3 - p3r).abs().max().item()
    # fp32 restatement (bf16 operands, fp32 accumulate, bf16 rounding of the hidden activations)
    f = lambda t: t.float()
    h1t = torch.relu((f(a1) @ f(w1).t() + b1) * torch.exp(3.0 * l1)).bfloat16()
    h2t = torch.relu((f(h1t) @ f(w2).t() + b2) * torch.exp(3.0 * l2)).bfloat16()
    p3t = f(h2t) @ f(w3).t()
    err = (p3 - p3t).abs().max().item()
    assert err <= 2e-2 * max(1.0, p3t.abs().max().item()), err       # one-ulp bf16 flips of h1/h2 propagate


def _mk_bwd(m, k3, k1p, seed):
    g = torch.Generator().manual_seed(seed)
    dev = "cuda"
    d3 = (torch.randn(m, k3, generator=g) * 0.5).to(dev).bfloat16()
    w3t = (torch.randn(HID, k3, generator=g) * 0.05).to(dev).bfloat16()
    w2t = (torch.randn(HID, HID, generator=g) * 0.05).to(dev).bfloat16()
    w1t = (torch.randn(k1p, HID, generator=g) * 0.05).to(dev).bfloat16()
    h2 = torch.relu(torch.randn(m, HID, generator=g)).to(dev).bfloat16()
    h1 = torch.relu(torch.randn(m, HID, generator=g)).to(dev).bfloat16()
    l2 = (torch.randn(HID, generator=g) * 0.1).to(dev)
    l1 = (torch.randn(HID, generator=g) * 0.1).to(dev)
    return d3, w3t, w2t, w1t, h2, h1, l2, l1


# (pixels, K3 = padded 9*Cout, K1p): level 1 affine (108 -> 128, 64), level 1 additive (54 -> 64), ragged,
# many tiles, level 2 (216 -> 256, 128)
BWD_SHAPES = [(2048, 128, 64), (1024, 64, 64), (1000, 128, 64), (148 * 128 + 256, 128, 64), (768, 256, 128)]


@pytest.mark.parametrize("m,k3,k1p", BWD_SHAPES)
def test_cnet_backward_matches_three_gemms(m, k3, k1p):
    if not _C.has_tcgen05():
        pytest.skip("needs sm_100")
    assert K.cnet_fused_supported(True, k3, HID, k1p)
    d3, w3t, w2t, w1t, h2, h1, l2, l1 = _mk_bwd(m, k3, k1p, 11 + m)
    db2r = torch.zeros(HID, device="cuda")
    db1r = torch.zeros(HID, device="cuda")
    d2r = K.gemm(d3, w3t, HID, k3, _C.EPI_RELU_BWD, None, l2, 3.0, y=h2, dlogs=None, dbias=db2r, out_dtype=_C.BF16)
    d1r = K.gemm(d2r, w2t, HID, HID, _C.EPI_RELU_BWD, None, l1, 3.0, y=h1, dlogs=None, dbias=db1r, out_dtype=_C.BF16)
    da1r = K.gemm(d1r, w1t, k1p, HID, _C.EPI_STORE, out_dtype=_C.BF16)
    db2 = torch.zeros(HID, device="cuda")
    db1 = torch.zeros(HID, device="cuda")
    d2, d1, da1 = K.cnet_backward(d3, w3t, w2t, w1t, HID, k1p, l2, 3.0, l1, 3.0, h2, h1, dbias2=db2, dbias1=db1)
    torch.cuda.synchronize()
    assert torch.equal(d2, d2r), "d2 differs: max %g" % (d2.float() - d2r.float()).abs().max().item()
    assert torch.equal(d1, d1r), "d1 differs: max %g" % (d1.float() - d1r.float()).abs().max().item()
    assert torch.equal(da1, da1r), "da1 differs: max %g" % (da1.float() - da1r.float()).abs().max().item()
    # bias gradients = column sums of the STORED (bf16) gradients here; the GEMM epilogue sums before rounding
    for got, stored, ref in ((db2, d2, db2r), (db1, d1, db1r)):
        exact = stored.float().sum(0)
        assert torch.allclose(got, exact, rtol=1e-4, atol=1e-3 * max(1.0, exact.abs().max().item()))
        assert torch.allclose(got, ref, rtol=2e-2, atol=2e-2 * max(1.0, ref.abs().max().item()))


# (N, H, W, C): level 1 / 2 shapes of the 64x64 and 32x32 models, a ragged pixel count, with and without the ones column
IMPL_SHAPES = [(2, 32, 32, 12, -1), (3, 16, 16, 24, -1), (1, 10, 6, 12, -1), (2, 32, 32, 12, 54), (5, 16, 16, 12, 54)]


@pytest.mark.parametrize("n,h,w,c,ones", IMPL_SHAPES)
@pytest.mark.parametrize("save", [False, True])
def test_cnet_forward_implicit_matches_explicit_im2col(n, h, w, c, ones, save):
    """conv1 as an implicit GEMM (in-kernel 3x3 gather, network/module.py:252) == glowk_im2col_rows + the explicit path."""
    if not _C.has_tcgen05():
        pytest.skip("needs sm_100")
    cin = c // 2
    k1p = K.round_up(9 * cin, 64)
    n3 = K.round_up(9 * c, 16)
    if not K.cnet_fused_supported(False, k1p, HID, n3):
        pytest.skip("shape not served by the fused kernel")
    g = torch.Generator().manual_seed(n * 1000 + h)
    z = torch.randn(n * h * w, c, generator=g).cuda()
    _, w1, w2, w3, b1, l1, b2, l2 = _mk(8, k1p, n3, 3)
    w1[:, 9 * cin:] = 0
    a1r = K.im2col_rows(z, n, h, w, 0, cin, 3, _C.BF16, k1p, ones_col=ones)
    p3r, h1r, h2r = K.cnet_forward(a1r, w1, w2, w3, HID, n3, b1, l1, 3.0, b2, l2, 3.0, save=True)
    p3, a1, h1, h2 = K.cnet_forward_implicit(z, n, h, w, 0, cin, k1p, w1, w2, w3, HID, n3, b1, l1, 3.0, b2, l2, 3.0,
                                             save=save, ones_col=ones)
    torch.cuda.synchronize()
    assert torch.equal(p3, p3r), (p3 - p3r).abs().max().item()
    if save:
        assert torch.equal(a1, a1r) and torch.equal(h1, h1r) and torch.equal(h2, h2r)
    else:
        assert a1 is None and h1 is None and h2 is None


@pytest.mark.parametrize("n,h,w,c", [(2, 32, 32, 12), (3, 16, 16, 24), (1, 10, 6, 12)])
def test_cnet_backward_implicit_matches_explicit_im2col(n, h, w, c):
    """dgrad3's operand gathered in-kernel (flipped 3x3 im2col of du) == glowk_im2col_rows(flip=1) + the explicit chain."""
    if not _C.has_tcgen05():
        pytest.skip("needs sm_100")
    cout, cin = c, c // 2
    k3p, k1p = K.round_up(9 * cout, 64), K.round_up(9 * cin, 64)
    if not K.cnet_fused_supported(True, k3p, HID, k1p):
        pytest.skip("shape not served by the fused kernel")
    m = n * h * w
    g = torch.Generator().manual_seed(n * 77 + h)
    du = (torch.randn(m, cout, generator=g) * 0.5).cuda()
    _, w3t, w2t, w1t, h2, h1, l2, l1 = _mk_bwd(m, k3p, k1p, 5)
    w3t[:, 9 * cout:] = 0
    d3r = K.im2col_rows(du, n, h, w, 0, cout, 3, _C.BF16, k3p, flip=True)
    db2r = torch.zeros(HID, device="cuda")
    d2r, d1r, da1r = K.cnet_backward(d3r, w3t, w2t, w1t, HID, k1p, l2, 3.0, l1, 3.0, h2, h1, dbias2=db2r)
    db2 = torch.zeros(HID, device="cuda")
    d3, d2, d1, da1 = K.cnet_backward_implicit(du, n, h, w, cout, k3p, w3t, w2t, w1t, HID, k1p, l2, 3.0, l1, 3.0, h2, h1,
                                               dbias2=db2)
    torch.cuda.synchronize()
    assert torch.equal(d3, d3r) and torch.equal(d2, d2r) and torch.equal(d1, d1r) and torch.equal(da1, da1r)
    assert torch.allclose(db2, db2r, rtol=1e-4, atol=1e-4)


def _encode_relu_bits(hact, m):
    """Host restatement of the bit-mask layout of glowk_cnet_*_implicit_masked (include/glowk.h): one 64-bit word per
    (tile of 128 rows, group of 64 columns, row); column pair jp of the group -> 32-bit word jp // 16, bit
    15 - jp % 16 (even column) / 31 - jp % 16 (odd column)."""
    tiles = (m + 127) // 128
    pos = torch.zeros(tiles * 128, 512, dtype=torch.bool)
    pos[:m] = hact[:m, :512].float().cpu() > 0
    pos = pos.view(tiles, 128, 8, 2, 16, 2)                    # tile, row, group, word, jp % 16, parity
    sh_even = torch.tensor([15 - j for j in range(16)], dtype=torch.int64)
    words = (pos[..., 0].long() << sh_even).sum(-1) + (pos[..., 1].long() << (sh_even + 16)).sum(-1)   # tile,row,group,word
    full = words[..., 0] + (words[..., 1] << 32)               # u0 = low half of the 64-bit word
    return full.permute(0, 2, 1).contiguous().view(-1)         # [tile][group][row]


@pytest.mark.parametrize("n,h,w,c", [(2, 32, 32, 12), (3, 16, 16, 24), (1, 10, 6, 12)])
def test_cnet_relu_bit_masks(n, h, w, c):
    """Training forward writes the ReLU masks of h1 / h2 as bits (documented layout); the backward chain fed with them is
    bit-identical to the chain fed with the bf16 activations."""
    if not _C.has_tcgen05():
        pytest.skip("needs sm_100")
    cin, cout = c // 2, c
    k1p, n3 = K.round_up(9 * cin, 64), K.round_up(9 * c, 16)
    k3p = K.round_up(9 * cout, 64)
    if not (K.cnet_fused_supported(False, k1p, HID, n3) and K.cnet_fused_supported(True, k3p, HID, k1p)):
        pytest.skip("shape not served by the fused kernels")
    m = n * h * w
    g = torch.Generator().manual_seed(n * 31 + h)
    z = torch.randn(m, c, generator=g).cuda()
    _, w1, w2, w3, b1, l1, b2, l2 = _mk(8, k1p, n3, 3)
    w1[:, 9 * cin:] = 0
    masks = K.cnet_relu_masks(m, z.device)
    for t in masks:
        t.fill_(-1)
    p3, a1, h1, h2 = K.cnet_forward_implicit(z, n, h, w, 0, cin, k1p, w1, w2, w3, HID, n3, b1, l1, 3.0, b2, l2, 3.0,
                                             save=True, masks=masks)
    p3r, _, h1r, h2r = K.cnet_forward_implicit(z, n, h, w, 0, cin, k1p, w1, w2, w3, HID, n3, b1, l1, 3.0, b2, l2, 3.0,
                                               save=True)
    torch.cuda.synchronize()
    assert torch.equal(p3, p3r) and torch.equal(h1, h1r) and torch.equal(h2, h2r)
    tiles = (m + 127) // 128
    rows_ok = (torch.arange(tiles * 128) < m).view(tiles, 1, 128).expand(tiles, 8, 128).reshape(-1)
    for got, act in ((masks[0], h1), (masks[1], h2)):
        ref = _encode_relu_bits(act, m)
        assert torch.equal(got.cpu()[rows_ok], ref[rows_ok])
    # backward: bits vs bf16 masks
    du = (torch.randn(m, cout, generator=g) * 0.5).cuda()
    _, w3t, w2t, w1t, _, _, lb2, lb1 = _mk_bwd(m, k3p, k1p, 5)
    w3t[:, 9 * cout:] = 0
    db2r, db2 = torch.zeros(HID, device="cuda"), torch.zeros(HID, device="cuda")
    ref = K.cnet_backward_implicit(du, n, h, w, cout, k3p, w3t, w2t, w1t, HID, k1p, lb2, 3.0, lb1, 3.0, h2, h1, dbias2=db2r)
    got = K.cnet_backward_implicit(du, n, h, w, cout, k3p, w3t, w2t, w1t, HID, k1p, lb2, 3.0, lb1, 3.0, h2, h1, dbias2=db2,
                                   masks=masks)
    torch.cuda.synchronize()
    for a, b, name in zip(got, ref, ("d3col", "d2", "d1", "da1")):
        assert torch.equal(a, b), name
    assert torch.allclose(db2, db2r, rtol=1e-5, atol=1e-5)
