"""Tensor-level wrappers of the glowk C ABI (one function per entry point).

Everything here takes/returns torch CUDA tensors, allocates outputs with torch's
caching allocator and launches on torch's current stream.  No arithmetic is done
in Python/torch: the math lives in pytorch_glow_b200/csrc/*.cu.
"""
import torch

from . import _C
from ._C import F32, BF16, call, ptr, check_cuda

TORCH_DTYPE = {F32: torch.float32, BF16: torch.bfloat16}


def round_up(v, m):
    return (v + m - 1) // m * m


def _f32c(t):
    assert t.dtype == torch.float32, "expected float32, got %s" % t.dtype
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------ ActNorm
def actnorm(x, bias, logs, logscale_factor=3.0, reverse=False, out=None):
    """y = (x+bias)*exp(f*logs) / inverse.  network/module.py:34-84."""
    check_cuda(x, bias, logs)
    x = _f32c(x)
    n, c, h, w = x.shape
    y = torch.empty_like(x) if out is None else out
    call("glowk_actnorm", ptr(x), ptr(y), ptr(bias), ptr(logs), float(logscale_factor), n, c, h * w,
         int(bool(reverse)))
    return y


def actnorm_init_nchw(x, scale=1.0, logscale_factor=3.0, batch_variance=False, reverse=False):
    """(bias, logs) [C] from a [N,C,H,W] fp32 batch.  network/module.py:86-120 (+ 44-45, 62-63, 112-113: the
    batch_variance option and a first call in the reverse direction)."""
    check_cuda(x)
    x = _f32c(x)
    n, c, h, w = x.shape
    bias = torch.empty(c, device=x.device, dtype=torch.float32)
    logs = torch.empty_like(bias)
    if batch_variance or reverse:
        call("glowk_actnorm_init_ex", ptr(x), n, c, h * w, c * h * w, h * w, 1, float(scale), float(logscale_factor),
             int(bool(batch_variance)), int(bool(reverse)), ptr(bias), ptr(logs))
        return bias, logs
    call("glowk_actnorm_init", ptr(x), F32, n, c, h * w, c * h * w, h * w, 1, float(scale),
         float(logscale_factor), ptr(bias), ptr(logs))
    return bias, logs


def actnorm_init_rows(rows, n_cols, scale=1.0, logscale_factor=3.0, batch_variance=False, reverse=False):
    """Same statistics over a pixel-major fp32 matrix [P][ld] (columns 0..n_cols-1)."""
    check_cuda(rows)
    assert rows.dtype == torch.float32 and rows.dim() == 2 and rows.is_contiguous()
    p, ld = rows.shape
    bias = torch.empty(n_cols, device=rows.device, dtype=torch.float32)
    logs = torch.empty_like(bias)
    if batch_variance or reverse:
        call("glowk_actnorm_init_ex", ptr(rows), 1, n_cols, p, 0, 1, ld, float(scale), float(logscale_factor),
             int(bool(batch_variance)), int(bool(reverse)), ptr(bias), ptr(logs))
        return bias, logs
    call("glowk_actnorm_init", ptr(rows), F32, 1, n_cols, p, 0, 1, ld, float(scale), float(logscale_factor),
         ptr(bias), ptr(logs))
    return bias, logs


# ------------------------------------------------------------------ invertible 1x1 conv
def invconv_prepare(weight, need_inverse):
    """(log|det W| as a 1-element tensor, W^-1 or None).  network/module.py:357,365."""
    check_cuda(weight)
    w = _f32c(weight)
    c = w.shape[0]
    logabsdet = torch.empty(1, device=w.device, dtype=torch.float32)
    winv = torch.empty_like(w) if need_inverse else None
    call("glowk_invconv_prepare", ptr(w), c, ptr(logabsdet), ptr(winv))
    return logabsdet, winv


def invconv_prepare_batched(weights, need_inverse):
    """weights: [B][C][C] -> (log|det| [B], inverses [B][C][C] or None) in one launch."""
    check_cuda(weights)
    w = _f32c(weights)
    b, c, _ = w.shape
    logabsdet = torch.empty(b, device=w.device, dtype=torch.float32)
    winv = torch.empty_like(w) if need_inverse else None
    call("glowk_invconv_prepare_batched", ptr(w), b, c, ptr(logabsdet), ptr(winv))
    return logabsdet, winv


def invconv_lu_assemble(p, l, u, sign_s, log_s, need_inverse):
    """W = P L (U + diag(sign_s exp(log_s))), optional W^-1, sum(log_s)."""
    check_cuda(p, l, u, sign_s, log_s)
    c = p.shape[0]
    w = torch.empty(c, c, device=p.device, dtype=torch.float32)
    winv = torch.empty_like(w) if need_inverse else None
    logabsdet = torch.empty(1, device=p.device, dtype=torch.float32)
    call("glowk_invconv_lu_assemble", ptr(_f32c(p)), ptr(_f32c(l)), ptr(_f32c(u)), ptr(_f32c(sign_s)),
         ptr(_f32c(log_s)), c, ptr(w), ptr(winv), ptr(logabsdet))
    return w, winv, logabsdet


def actnorm_mix(x, weight=None, indices=None, bias=None, logs=None, logscale_factor=3.0, reverse=False):
    """Fused ActNorm + 1x1 channel mix (or permutation gather).  model.py:94-103 / 142-152."""
    check_cuda(x, weight, indices, bias, logs)
    x = _f32c(x)
    n, c, h, w = x.shape
    z = torch.empty_like(x)
    call("glowk_actnorm_mix", ptr(x), ptr(z), ptr(weight), ptr(indices), ptr(bias), ptr(logs),
         float(logscale_factor), n, c, h * w, int(bool(reverse)))
    return z


# ------------------------------------------------------------------ squeeze
def squeeze2d(x, factor=2, reverse=False):
    """network/module.py:551-591 (bit-exact)."""
    check_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 4
    n, c, h, w = x.shape
    if factor == 1:
        return x
    # accept a channel-sliced view (Split2d's z1) without a copy: only the batch stride may differ
    if not (x.stride(3) == 1 and x.stride(2) == w and x.stride(1) == h * w):
        x = x.contiguous()
    sn = x.stride(0) if n > 1 else c * h * w
    if not reverse:
        if h % factor or w % factor:
            raise ValueError("Squeeze2d: H, W = %d, %d not divisible by factor %d" % (h, w, factor))
        y = torch.empty(n, c * factor * factor, h // factor, w // factor, device=x.device, dtype=x.dtype)
    else:
        if c < factor * factor or c % (factor * factor):
            raise ValueError("Squeeze2d: C = %d not divisible by factor^2" % c)
        y = torch.empty(n, c // (factor * factor), h * factor, w * factor, device=x.device, dtype=x.dtype)
    call("glowk_squeeze2d", x.data_ptr(), ptr(y), n, c, h, w, sn, int(factor), int(bool(reverse)))
    return y


# ------------------------------------------------------------------ rows <-> NCHW, weight packing
def im2col(src, c0, cin, ksize, dtype, ld, flip=False):
    """rows[p][tap*cin+ci] from channels c0..c0+cin of an NCHW fp32 tensor (SAME zero padding)."""
    check_cuda(src)
    assert src.dtype == torch.float32 and src.dim() == 4
    n, ctot, h, w = src.shape
    if not (src.stride(3) == 1 and src.stride(2) == w and src.stride(1) == h * w):
        src = src.contiguous()
    sn = src.stride(0) if n > 1 else ctot * h * w
    rows = torch.empty(n * h * w, ld, device=src.device, dtype=TORCH_DTYPE[dtype])
    call("glowk_im2col", src.data_ptr(), n, sn, c0, cin, h, w, int(ksize), int(bool(flip)), ptr(rows), dtype, ld)
    return rows


def im2col_rows(src_rows, n, h, w, c0, cin, ksize, dtype, ld, flip=False, ones_col=-1):
    """ones_col >= 0: also write 1.0 into that zero-padding column (glowk_im2col_rows_ones)."""
    check_cuda(src_rows)
    assert src_rows.dtype == torch.float32 and src_rows.dim() == 2 and src_rows.is_contiguous()
    rows = torch.empty(n * h * w, ld, device=src_rows.device, dtype=TORCH_DTYPE[dtype])
    if ones_col >= 0:
        call("glowk_im2col_rows_ones", ptr(src_rows), src_rows.shape[1], n, c0, cin, h, w, int(ksize),
             int(bool(flip)), ptr(rows), dtype, ld, int(ones_col))
    else:
        call("glowk_im2col_rows", ptr(src_rows), src_rows.shape[1], n, c0, cin, h, w, int(ksize), int(bool(flip)),
             ptr(rows), dtype, ld)
    return rows


def rows_to_nchw(rows, n, c, h, w):
    check_cuda(rows)
    dt = BF16 if rows.dtype == torch.bfloat16 else F32
    out = torch.empty(n, c, h, w, device=rows.device, dtype=torch.float32)
    call("glowk_rows_to_nchw", ptr(rows), dt, rows.shape[1], ptr(out), n, c, h * w)
    return out


def tapsum_to_nchw(p_rows, dst, c0, c, flip=False, accumulate=False):
    check_cuda(p_rows, dst)
    n, ctot, h, w = dst.shape
    assert dst.is_contiguous() and p_rows.dtype == torch.float32
    call("glowk_tapsum_to_nchw", ptr(p_rows), p_rows.shape[1], ptr(dst), n, ctot, c0, c, h, w, int(bool(flip)),
         int(bool(accumulate)))
    return dst


def pack_conv_weight(weight, layout, dtype, rows, ld):
    """fp32 [O][I][k][k] -> GEMM B operand (see include/glowk.h for the four layouts)."""
    check_cuda(weight)
    w = _f32c(weight)
    o, i, k, _ = w.shape
    dst = torch.empty(rows, ld, device=w.device, dtype=TORCH_DTYPE[dtype])
    call("glowk_pack_conv_weight", ptr(w), o, i, k, int(layout), ptr(dst), dtype, rows, ld)
    return dst


# ------------------------------------------------------------------ GEMMs
def gemm(a, b, n, k, epilogue=_C.EPI_STORE, bias=None, logs=None, logscale_factor=3.0, y=None,
         dlogs=None, dbias=None, out_dtype=F32, ldo=None, out=None, cluster=None):
    """out[M][n] = epilogue(a[M][:k] . b[:n][:k]^T).  a, b: 2-D row-major, same dtype."""
    check_cuda(a, b)
    assert a.dim() == 2 and b.dim() == 2 and a.dtype == b.dtype
    dt = BF16 if a.dtype == torch.bfloat16 else F32
    m = a.shape[0]
    ldo = n if ldo is None else ldo
    if out is None:
        out = torch.empty(m, ldo, device=a.device, dtype=TORCH_DTYPE[out_dtype])
    cm, cn = cluster or (0, 0)
    call("glowk_gemm_ex", ptr(a), a.shape[1], ptr(b), b.shape[1], dt, m, n, k, int(epilogue), ptr(bias), ptr(logs),
         float(logscale_factor), ptr(y), 0 if y is None else y.shape[1], ptr(dlogs), ptr(dbias), ptr(out),
         out_dtype, ldo, int(cm), int(cn))
    return out


def gemm_wgrad(a, b, mo, no, dw):
    """dw[mo][no] += a[P][:mo]^T . b[P][:no]  (fp32 accumulate into dw)."""
    check_cuda(a, b, dw)
    assert a.dtype == b.dtype and dw.dtype == torch.float32 and dw.is_contiguous()
    dt = BF16 if a.dtype == torch.bfloat16 else F32
    call("glowk_gemm_wgrad", ptr(a), a.shape[1], ptr(b), b.shape[1], dt, a.shape[0], mo, no, ptr(dw),
         dw.shape[-1] if dw.dim() == 2 else no)
    return dw


def invconv_lu_grads(dw, p, l, u, sign_s, log_s, dl, du, dlog_s):
    """Accumulate the gradients of the LU parameters from dW (glowk_invconv_lu_grads)."""
    check_cuda(dw, p, l, u, sign_s, log_s, dl, du, dlog_s)
    call("glowk_invconv_lu_grads", ptr(dw), ptr(p), ptr(l), ptr(u), ptr(sign_s), ptr(log_s), l.shape[0], ptr(dl), ptr(du),
         ptr(dlog_s))


def cnet_fused_supported(backward, k1, hidden, n3):
    """True iff glowk_cnet_forward / glowk_cnet_backward serve this shape on the current device."""
    return bool(_C.lib().glowk_cnet_fused_supported(int(bool(backward)), int(k1), int(hidden), int(n3)))


def cnet_forward(a1, w1, w2, w3, hidden, n3, bias1, logs1, f1, bias2, logs2, f2, ldp3=None, save=False, ldh=None,
                 masks=None):
    """The coupling net's three convs in one tcgen05 kernel (glowk_cnet_forward).  a1: [M][k1p] bf16 im2col rows.
    Returns (p3 [M][ldp3] fp32, h1, h2) with h1/h2 [M][ldh] bf16 when `save`, else None.  masks: see
    cnet_forward_implicit."""
    check_cuda(a1, w1, w2, w3)
    assert a1.dtype == torch.bfloat16 and w1.dtype == torch.bfloat16
    m, k1 = a1.shape
    ldp3 = n3 if ldp3 is None else ldp3
    ldh = hidden if ldh is None else ldh
    p3 = torch.empty(m, ldp3, device=a1.device, dtype=torch.float32)
    h1 = torch.empty(m, ldh, device=a1.device, dtype=torch.bfloat16) if save else None
    h2 = torch.empty(m, ldh, device=a1.device, dtype=torch.bfloat16) if save else None
    args = (ptr(a1), k1, ptr(w1), w1.shape[1], ptr(w2), w2.shape[1], ptr(w3), w3.shape[1], m, k1,
            hidden, n3, ptr(bias1), ptr(logs1), float(f1), ptr(bias2), ptr(logs2), float(f2), ptr(p3), ldp3, ptr(h1),
            ptr(h2), ldh)
    if masks is not None:
        assert save, "the bit masks are a by-product of the training forward"
        call("glowk_cnet_forward_masked", *args, ptr(masks[0]), ptr(masks[1]))
    else:
        call("glowk_cnet_forward", *args)
    return p3, h1, h2


def cnet_relu_masks(m, device):
    """Two buffers for the ReLU bit masks of h1 / h2 over m pixels (glowk_cnet_*_implicit_masked)."""
    nbytes = int(_C.lib().glowk_cnet_relu_mask_bytes(m))
    return (torch.empty(nbytes // 8, dtype=torch.int64, device=device),
            torch.empty(nbytes // 8, dtype=torch.int64, device=device))


def cnet_forward_implicit(z, n, h, w, c0, cin, k1p, w1, w2, w3, hidden, n3, bias1, logs1, f1, bias2, logs2, f2, ldp3=None,
                          save=False, ldh=None, ones_col=-1, masks=None):
    """glowk_cnet_forward_implicit: conv1's im2col operand is gathered in-kernel from the rows z [n*h*w][ld] fp32.
    Returns (p3, a1, h1, h2); a1 / h1 / h2 are None unless `save` (training).  masks = (m1, m2) from cnet_relu_masks:
    also write the ReLU masks of h1 / h2 as bits for cnet_backward_implicit."""
    check_cuda(z, w1, w2, w3)
    assert z.dtype == torch.float32 and z.dim() == 2 and w1.dtype == torch.bfloat16
    m = n * h * w
    assert z.shape[0] == m
    ldp3 = n3 if ldp3 is None else ldp3
    ldh = hidden if ldh is None else ldh
    dev = z.device
    p3 = torch.empty(m, ldp3, device=dev, dtype=torch.float32)
    a1 = torch.empty(m, k1p, device=dev, dtype=torch.bfloat16) if save else None
    h1 = torch.empty(m, ldh, device=dev, dtype=torch.bfloat16) if save else None
    h2 = torch.empty(m, ldh, device=dev, dtype=torch.bfloat16) if save else None
    args = (ptr(z), z.shape[1], c0, cin, n, h, w, int(ones_col), ptr(a1), k1p, ptr(w1),
            w1.shape[1], ptr(w2), w2.shape[1], ptr(w3), w3.shape[1], k1p, hidden, n3, ptr(bias1), ptr(logs1), float(f1),
            ptr(bias2), ptr(logs2), float(f2), ptr(p3), ldp3, ptr(h1), ptr(h2), ldh)
    if masks is not None:
        assert save, "the bit masks are a by-product of the training forward"
        call("glowk_cnet_forward_implicit_masked", *args, ptr(masks[0]), ptr(masks[1]))
    else:
        call("glowk_cnet_forward_implicit", *args)
    return p3, a1, h1, h2


def cnet_backward(d3col, w3t, w2t, w1t, hidden, k1p, logs2, f2, logs1, f1, h2, h1, dbias2=None, dbias1=None):
    """The dgrad chain of the coupling net in one tcgen05 kernel (glowk_cnet_backward).
    Returns (d2, d1 [M][ldh] bf16, da1 [M][k1p] bf16)."""
    check_cuda(d3col, w3t, w2t, w1t, h2, h1)
    m, k3 = d3col.shape
    ldh = h1.shape[1]
    assert h2.shape[1] == ldh and d3col.dtype == torch.bfloat16
    d2 = torch.empty(m, ldh, device=d3col.device, dtype=torch.bfloat16)
    d1 = torch.empty(m, ldh, device=d3col.device, dtype=torch.bfloat16)
    da1 = torch.empty(m, k1p, device=d3col.device, dtype=torch.bfloat16)
    call("glowk_cnet_backward", ptr(d3col), k3, ptr(w3t), w3t.shape[1], ptr(w2t), w2t.shape[1], ptr(w1t), w1t.shape[1],
         m, k3, hidden, k1p, ptr(logs2), float(f2), ptr(logs1), float(f1), ptr(h2), ptr(h1), ptr(d2), ptr(d1), ldh,
         ptr(da1), k1p, ptr(dbias2), ptr(dbias1))
    return d2, d1, da1


def cnet_backward_implicit(du, n, h, w, cout, k3p, w3t, w2t, w1t, hidden, k1p, logs2, f2, logs1, f1, h2, h1, dbias2=None,
                           dbias1=None, masks=None):
    """glowk_cnet_backward_implicit: the dgrad chain with its first operand (flipped im2col of du) gathered in-kernel.
    Returns (d3col [M][k3p] bf16 -- the conv3 wgrad operand --, d2, d1 [M][ldh] bf16, da1 [M][k1p] bf16)."""
    check_cuda(du, w3t, w2t, w1t, h2, h1)
    assert du.dtype == torch.float32 and du.dim() == 2
    m = n * h * w
    assert du.shape[0] == m
    ldh = h1.shape[1]
    dev = du.device
    d3col = torch.empty(m, k3p, device=dev, dtype=torch.bfloat16)
    d2 = torch.empty(m, ldh, device=dev, dtype=torch.bfloat16)
    d1 = torch.empty(m, ldh, device=dev, dtype=torch.bfloat16)
    da1 = torch.empty(m, k1p, device=dev, dtype=torch.bfloat16)
    args = (ptr(du), du.shape[1], cout, n, h, w, ptr(d3col), k3p, ptr(w3t), w3t.shape[1],
            ptr(w2t), w2t.shape[1], ptr(w1t), w1t.shape[1], k3p, hidden, k1p, ptr(logs2), float(f2), ptr(logs1), float(f1),
            ptr(h2), ptr(h1), ptr(d2), ptr(d1), ldh, ptr(da1), k1p, ptr(dbias2), ptr(dbias1))
    if masks is not None:          # (m1, m2) as written by cnet_forward_implicit: the chain applies m2 first, then m1
        call("glowk_cnet_backward_implicit_masked", *args, ptr(masks[1]), ptr(masks[0]))
    else:
        call("glowk_cnet_backward_implicit", *args)
    return d3col, d2, d1, da1


# ------------------------------------------------------------------ coupling / logdet / prior
def coupling(p_rows, bias3, logs3, z, affine, reverse, logscale_factor=3.0, save_h=False):
    """In-place coupling on z[:, C/2:] from the tap-GEMM output p_rows.  model.py:105-115 / 131-140.

    Returns (partials [N][nblk] or None, h rows or None)."""
    check_cuda(p_rows, bias3, logs3, z)
    assert z.is_contiguous() and z.dtype == torch.float32 and p_rows.dtype == torch.float32
    n, c, h, w = z.shape
    cout = c if affine else c // 2
    nblk = _C.coupling_nblk(h * w)
    partials = torch.empty(n, nblk, device=z.device, dtype=torch.float32) if affine else None
    hs = torch.empty(n * h * w, cout, device=z.device, dtype=torch.float32) if save_h else None
    call("glowk_coupling", ptr(p_rows), p_rows.shape[1], ptr(bias3), ptr(logs3), float(logscale_factor), ptr(z),
         ptr(partials), ptr(hs), n, c, h, w, int(bool(affine)), int(bool(reverse)))
    return partials, hs


def logdet_finish(logdet_in, n, hw, logs=None, logabsdet=None, partials=None, logscale_factor=3.0, sign=1.0,
                  device=None):
    """logdet_out[n] = logdet_in[n] + sign*HW*(sum f*logs + log|det W|) + sum_b partials[n][b]."""
    dev = device if device is not None else (logdet_in.device if logdet_in is not None else logs.device)
    out = torch.empty(n, device=dev, dtype=torch.float32)
    c = 0 if logs is None else logs.numel()
    nblk = 0 if partials is None else partials.shape[1]
    call("glowk_logdet_finish", ptr(logdet_in), ptr(out), ptr(logs), c, float(logscale_factor), ptr(logabsdet),
         ptr(partials), nblk, hw, float(sign), n)
    return out


def gaussian_logp(h_rows, x, c0, cz, logdet_in=None):
    """logdet_in + sum_{c,p} log N(x[:, c0:c0+cz]; mean, exp(logs)) with (mean, logs) = cross-split of h_rows
    (None => standard normal).  network/module.py:437-467."""
    check_cuda(h_rows, x, logdet_in)
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c, h, w = x.shape
    out = torch.empty(n, device=x.device, dtype=torch.float32)
    call("glowk_gaussian_logp", ptr(h_rows), 0 if h_rows is None else h_rows.shape[1], ptr(x), n, c, h * w, c0, cz,
         ptr(logdet_in), ptr(out))
    return out


_tickets = {}


def _ticket(device):
    """One zero-initialised counter per device for glowk_nll_head's last-CTA reduction (the kernel leaves it at zero).
    Launches that use it must be stream-ordered per device: one loss head at a time per GPU, which is what a training
    / evaluation process does; two models driven concurrently on different streams of ONE device need their own."""
    t = _tickets.get(device)
    if t is None:
        t = _tickets[device] = torch.zeros(4, device=device, dtype=torch.int32)
    return t


def nll_head(z, ld, c0, denom, want_loss=True):
    """Glow's loss head (model.py:425-450, 496-498; N(0, I) top prior): z [N, ...] top latent, ld [N] the flow's
    logdet started from zero -> (nll [N] in bits/dim, mean loss as a 0-dim tensor or None)."""
    check_cuda(z, ld)
    assert z.dtype == torch.float32 and z.is_contiguous()
    n = z.shape[0]
    nll = torch.empty(n, device=z.device, dtype=torch.float32)
    loss = torch.empty((), device=z.device, dtype=torch.float32) if want_loss else None
    call("glowk_nll_head", ptr(z), ptr(ld), float(c0), float(denom), n, z[0].numel() if n else 0, ptr(nll), ptr(loss),
         ptr(_ticket(z.device)) if want_loss else 0)
    return nll, loss


def nll_head_bwd(z, denom, g_loss=None, g_nll=None, dz_in=None):
    """Adjoint of nll_head -> (dz, dld)."""
    check_cuda(z, g_loss, g_nll, dz_in)
    n = z.shape[0]
    dz = torch.empty_like(z)
    dld = torch.empty(n, device=z.device, dtype=torch.float32)
    call("glowk_nll_head_bwd", ptr(z), ptr(g_loss), ptr(g_nll), ptr(dz_in), float(denom), n, z[0].numel() if n else 0,
         ptr(dz), ptr(dld))
    return dz, dld


def split2d_sample(h_rows, z1, eps):
    """cat(z1, mean + exp(logs)*eps).  network/module.py:482-483, 532-536."""
    check_cuda(h_rows, z1, eps)
    z1 = _f32c(z1)
    eps = _f32c(eps)
    n, ch, h, w = z1.shape
    assert eps.numel() == z1.numel(), "eps must have the shape of z1 %s, got %s" % (tuple(z1.shape), tuple(eps.shape))
    out = torch.empty(n, 2 * ch, h, w, device=z1.device, dtype=torch.float32)
    call("glowk_split2d_sample", ptr(h_rows), h_rows.shape[1], ptr(z1), ptr(eps), ptr(out), n, ch, h * w)
    return out


# ------------------------------------------------------------------ backward (training) wrappers
def coupling_bwd(y, hrows, dy, dld, logs3, affine, dlogs3, dbias3, logscale_factor=3.0):
    """Adjoint of `coupling` (+ Conv2dZeros scale).  Returns (dz [N,C,H,W], du rows [P][Cout])."""
    check_cuda(y, hrows, dy, dld, logs3, dlogs3, dbias3)
    n, c, h, w = y.shape
    cout = c if affine else c // 2
    dy = _f32c(dy)
    dz = torch.empty_like(y)
    du = torch.empty(n * h * w, cout, device=y.device, dtype=torch.float32)
    call("glowk_coupling_bwd", ptr(y), ptr(hrows), ptr(dy), ptr(dld), ptr(logs3), float(logscale_factor), ptr(dz),
         ptr(du), ptr(dlogs3), ptr(dbias3), n, c, h, w, int(bool(affine)))
    return dz, du


def split2d_bwd(x, hrows, dz1, dld, logs_p, dlogs_p, dbias_p, logscale_factor=3.0):
    """Adjoint of Split2d forward.  Returns (dx [N,C,H,W], du rows [P][C])."""
    check_cuda(x, hrows, dz1, dld, logs_p)
    n, c, h, w = x.shape
    dx = torch.empty_like(x)
    du = torch.empty(n * h * w, c, device=x.device, dtype=torch.float32)
    sn = 0
    if dz1 is not None:
        if not (dz1.stride(3) == 1 and dz1.stride(2) == w and dz1.stride(1) == h * w):
            dz1 = dz1.contiguous()
        sn = dz1.stride(0) if n > 1 else (c // 2) * h * w
    call("glowk_split2d_bwd", ptr(x), ptr(hrows), hrows.shape[1], None if dz1 is None else dz1.data_ptr(), sn,
         ptr(dld), ptr(logs_p), float(logscale_factor), ptr(dx), ptr(du), c, ptr(dlogs_p), ptr(dbias_p), n, c, h * w)
    return dx, du


def actnorm_mix_bwd(x, dz, weight=None, indices=None, bias=None, logs=None, dw=None, dlogs=None, dbias=None,
                    logscale_factor=3.0):
    """Adjoint of the forward `actnorm_mix`; accumulates dw / dlogs / dbias, returns dx."""
    check_cuda(x, dz, weight, indices, bias, logs, dw, dlogs, dbias)
    n, c, h, w = x.shape
    dz = _f32c(dz)
    dx = torch.empty_like(x)
    call("glowk_actnorm_mix_bwd", ptr(x), ptr(dz), ptr(weight), ptr(indices), ptr(bias), ptr(logs),
         float(logscale_factor), ptr(dx), ptr(dw), ptr(dlogs), ptr(dbias), n, c, h * w)
    return dx


def logdet_param_grad(dld, hw, dlogs=None, winv=None, dw=None, logscale_factor=3.0):
    check_cuda(dld, dlogs, winv, dw)
    c = dlogs.numel() if dlogs is not None else winv.shape[0]
    call("glowk_logdet_param_grad", ptr(dld), dld.numel(), hw, float(logscale_factor), ptr(dlogs), c, ptr(winv), ptr(dw))


def unpack_weight_grad(src, o, i, ksize, layout, grad, accumulate=True):
    check_cuda(src, grad)
    assert src.dtype == torch.float32 and grad.is_contiguous()
    call("glowk_unpack_weight_grad", ptr(src), src.shape[1], o, i, int(ksize), int(layout), ptr(grad),
         int(bool(accumulate)))


def optim_workspace(device):
    return torch.zeros(int(_C.lib().glowk_optim_workspace_floats()), device=device, dtype=torch.float32)


def optim_clip_norm(grads, clip_value, max_norm, workspace):
    """In-place clip_grad_value_ then norm/coef into workspace[0:2] (trainer.py:142-147)."""
    call("glowk_optim_clip_norm", ptr(grads), grads.numel(), float(clip_value or 0.0), float(max_norm or 0.0),
         ptr(workspace))


def optim_adam(params, grads, exp_avg, exp_avg_sq, workspace, step, lr, beta1, beta2, eps, sched=None):
    call("glowk_optim_adam", ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), params.numel(), ptr(workspace),
         ptr(sched), float(lr), float(beta1), float(beta2), float(eps), int(step))


def optim_adamax(params, grads, exp_avg, exp_inf, workspace, step, lr, beta1, beta2, eps, sched=None):
    call("glowk_optim_adamax", ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_inf), params.numel(), ptr(workspace),
         ptr(sched), float(lr), float(beta1), float(beta2), float(eps), int(step))


def optim_schedule(step_dev, sched_dev, base_lr, warmup_steps, min_lr, beta1, beta2):
    """sched_dev <- [lr, 1-beta1^t, sqrt(1-beta2^t)] for the iteration counted by the device int64 step_dev; ++step_dev."""
    assert step_dev.dtype == torch.int64 and sched_dev.dtype == torch.float32
    call("glowk_optim_schedule", ptr(step_dev), ptr(sched_dev), float(base_lr), int(warmup_steps),
         -1.0 if min_lr is None else float(min_lr), float(beta1), float(beta2))


# ------------------------------------------------------------------ pixel-major ("rows") flow state
NCHW, ROWS = 0, 1


def rows_max_channels():
    return int(_C.lib().glowk_rows_max_channels())


def rows_max_channels_wide():
    return int(_C.lib().glowk_rows_max_channels_wide())


def rows_actnorm_bwd(da, x, bias, logs, dlogs, dbias, logscale_factor=3.0, out=None):
    """dx = da*s (in place by default), dbias += sum da*s, dlogs += f*sum da*(x+bias)*s on rows [P][C]."""
    check_cuda(da, x, bias, logs, dlogs, dbias)
    out = da if out is None else out
    call("glowk_rows_actnorm_bwd", ptr(da), ptr(x), ptr(bias), ptr(logs), float(logscale_factor), ptr(out), ptr(dlogs),
         ptr(dbias), da.shape[0], da.shape[1])
    return out


def rows_squeeze(src, src_layout, src_ld, dst, dst_layout, dst_ld, n, c, h, w, factor, reverse, add=None):
    """Squeeze2d / unsqueeze between layouts (module.py:551-591); see glowk_rows_squeeze.
    add: tensor laid out like src, added on the fly (the dequantisation noise, model.py:421-423; forward only)."""
    check_cuda(src, dst, add)
    if add is not None:
        assert not reverse and add.shape == src.shape and add.dtype == torch.float32 and add.is_contiguous()
        call("glowk_rows_squeeze_add", src.data_ptr(), add.data_ptr(), int(src_layout), int(src_ld), dst.data_ptr(),
             int(dst_layout), int(dst_ld), n, c, h, w, int(factor))
        return dst
    call("glowk_rows_squeeze", src.data_ptr(), int(src_layout), int(src_ld), dst.data_ptr(), int(dst_layout),
         int(dst_ld), n, c, h, w, int(factor), int(bool(reverse)))
    return dst


def rows_actnorm_mix(x, weight=None, indices=None, bias=None, logs=None, logscale_factor=3.0, reverse=False):
    check_cuda(x, weight, indices, bias, logs)
    p, c = x.shape
    z = torch.empty_like(x)
    call("glowk_rows_actnorm_mix", ptr(x), ptr(z), ptr(weight), ptr(indices), ptr(bias), ptr(logs),
         float(logscale_factor), p, c, int(bool(reverse)))
    return z


def rows_coupling_rev_mix(p_rows, bias3, logs3, x, n, h, w, affine, logscale_factor3, weight=None, indices=None,
                          bias=None, logs=None, logscale_factor=3.0):
    """Reverse FlowStep behind the coupling net in one launch: inverse coupling (from the tap-GEMM output p_rows) +
    W^-1 mix / inverse permutation + ActNorm^-1 on the rows x (left untouched); returns the new rows."""
    check_cuda(p_rows, bias3, logs3, x, weight, indices, bias, logs)
    z = torch.empty_like(x)
    call("glowk_rows_coupling_rev_mix", ptr(p_rows), p_rows.shape[1], ptr(bias3), ptr(logs3), float(logscale_factor3),
         ptr(x), ptr(z), ptr(weight), ptr(indices), ptr(bias), ptr(logs), float(logscale_factor), n, x.shape[1], h, w,
         int(bool(affine)))
    return z


def rows_coupling_nblk(hw, c):
    return int(_C.lib().glowk_rows_coupling_nblk(hw, c))


def rows_coupling(p_rows, bias3, logs3, z, n, h, w, affine, reverse, logscale_factor=3.0, save_h=False,
                  ld_in=None, want_ld=False, an_logs=None, an_f=3.0, logabsdet=None, sign=1.0, partials=None,
                  tickets=None):
    """Tap gather-sum + coupling in place on z[:, C/2:] (+ this step's logdet).  Returns (ld_out, h rows)."""
    check_cuda(p_rows, bias3, logs3, z)
    c = z.shape[1]
    cout = c if affine else c // 2
    hs = torch.empty(n * h * w, cout, device=z.device, dtype=torch.float32) if save_h else None
    ld_out = torch.empty(n, device=z.device, dtype=torch.float32) if want_ld else None
    call("glowk_rows_coupling", ptr(p_rows), p_rows.shape[1], ptr(bias3), ptr(logs3), float(logscale_factor), ptr(z),
         ptr(hs), n, c, h, w, int(bool(affine)), int(bool(reverse)), ptr(ld_in), ptr(ld_out), ptr(an_logs),
         float(an_f), ptr(logabsdet), float(sign), ptr(partials), ptr(tickets))
    return ld_out, hs


def rows_coupling_bwd(y, hrows, dy, dld, logs3, n, hw, affine, dlogs3, dbias3, logscale_factor=3.0):
    check_cuda(y, hrows, dy, dld, logs3, dlogs3, dbias3)
    c = y.shape[1]
    cout = c if affine else c // 2
    dz = torch.empty_like(y)
    du = torch.empty(y.shape[0], cout, device=y.device, dtype=torch.float32)
    call("glowk_rows_coupling_bwd", ptr(y), ptr(hrows), ptr(dy), ptr(dld), ptr(logs3), float(logscale_factor),
         ptr(dz), ptr(du), ptr(dlogs3), ptr(dbias3), n, c, hw, int(bool(affine)))
    return dz, du


def rows_actnorm_mix_bwd(x, dz, n, h, w, da1=None, cin=0, weight=None, indices=None, bias=None, logs=None, dw=None,
                         dlogs=None, dbias=None, logscale_factor=3.0, dld=None, winv=None):
    check_cuda(x, dz, da1, weight, indices, bias, logs, dw, dlogs, dbias, dld, winv)
    c = x.shape[1]
    dx = torch.empty_like(x)
    adt = BF16 if (da1 is not None and da1.dtype == torch.bfloat16) else F32
    call("glowk_rows_actnorm_mix_bwd_ex", ptr(x), ptr(dz), ptr(da1), adt, 0 if da1 is None else da1.shape[1], int(cin),
         ptr(weight), ptr(indices), ptr(bias), ptr(logs), float(logscale_factor), ptr(dx), ptr(dw), ptr(dlogs),
         ptr(dbias), n, c, h, w, ptr(dld), ptr(winv))
    return dx


def rows_gaussian_logp(h_rows, x, n, hw, c0, cz, logdet_in=None):
    check_cuda(h_rows, x, logdet_in)
    out = torch.empty(n, device=x.device, dtype=torch.float32)
    call("glowk_rows_gaussian_logp", ptr(h_rows), 0 if h_rows is None else h_rows.shape[1], ptr(x), x.shape[1], n, hw,
         c0, cz, ptr(logdet_in), ptr(out))
    return out


def rows_split2d_sample(h_rows, z1, ldz1, eps, n, ch, hw):
    check_cuda(h_rows, z1, eps)
    assert eps.numel() == n * ch * hw, "eps must be [N, C/2, H, W] = %s, got %s" % ((n, ch, hw), tuple(eps.shape))
    out = torch.empty(n * hw, 2 * ch, device=z1.device, dtype=torch.float32)
    call("glowk_rows_split2d_sample", ptr(h_rows), h_rows.shape[1], z1.data_ptr(), int(ldz1), ptr(_f32c(eps)), ptr(out),
         n, ch, hw)
    return out


def rows_split2d_bwd(x, hrows, dld, logs_p, dx, dlogs_p, dbias_p, n, hw, logscale_factor=3.0):
    check_cuda(x, hrows, dld, logs_p, dx, dlogs_p, dbias_p)
    c = x.shape[1]
    du = torch.empty(x.shape[0], c, device=x.device, dtype=torch.float32)
    call("glowk_rows_split2d_bwd", ptr(x), ptr(hrows), hrows.shape[1], ptr(dld), ptr(logs_p), float(logscale_factor),
         ptr(dx), ptr(du), c, ptr(dlogs_p), ptr(dbias_p), n, c, hw)
    return du


def rows_tapsum(p_rows, dst, c0, c, n, h, w, flip=False, accumulate=False):
    check_cuda(p_rows, dst)
    call("glowk_rows_tapsum", ptr(p_rows), p_rows.shape[1], ptr(dst), dst.shape[1], c0, c, n, h, w, int(bool(flip)),
         int(bool(accumulate)))
    return dst


def pack_conv_weights_batched(jobs_dev, njobs, total_blocks, dtype):
    call("glowk_pack_conv_weights_batched", ptr(jobs_dev), njobs, total_blocks, int(dtype))


def unpack_weight_grads_batched(jobs_dev, njobs, total_blocks):
    call("glowk_unpack_weight_grads_batched", ptr(jobs_dev), njobs, total_blocks)


def conv_actnorm_finish_batched(jobs_dev, njobs, max_n):
    """dbias += db; dlogs += f*(<W, dW> + bias*db) per output channel of every job (see include/glowk.h)."""
    call("glowk_conv_actnorm_finish_batched", ptr(jobs_dev), njobs, max_n)
