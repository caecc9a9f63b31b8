"""Autograd bindings: hand-written backward kernels behind torch.autograd.Function.

Design: the forward of a flow layer runs the same kernels as inference and keeps the handful of
tensors its adjoint needs (`ctx` dicts below); the backward launches the adjoint kernels and
ACCUMULATES parameter gradients straight into ``param.grad`` (allocated on demand), so a whole
``FlowModel.encode`` is one autograd node and nothing is re-materialised by torch.  The reference
relies on torch autograd over ~170 ATen ops per step (network/trainer.py:138-140).
"""
import torch

from . import _C, config
from . import functional as K
from .functional import round_up


def _gbuf(p):
    """param.grad as a flat fp32 accumulator (created zero-filled if absent)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad.view(-1)


# ------------------------------------------------------------------ FlowStep
def flowstep_forward_save(step, x, ld_vec):
    """FlowStep.normal_flow (network/model.py:82-117) keeping what the adjoint needs.
    x: [N,C,H,W] contiguous; ld_vec: [N] or None.  Returns (y, ld_out or None, ctx)."""
    n, c, h, w = x.shape
    an = step.actnorm
    if an.needs_init:
        an.initialize_from_nchw(x)
    if step.permutation == 'invconv':
        wmat, winv, logabsdet = step.invconv.prepared(need_inverse=True)
        idx = None
    else:
        wmat, winv, logabsdet = None, None, None
        idx = step.perm_module.device_indices(x.device, False)
    b, l = an.bias.detach().reshape(-1), an.logs.detach().reshape(-1)
    z = K.actnorm_mix(x, wmat, idx, b, l, an.logscale_factor, reverse=False)
    save = {}
    p3 = step.f.tap_rows(z, step.conv_dtype, save=save)
    c3 = step.f[4]
    affine = step.coupling == 'affine'
    partials, hrows = K.coupling(p3, c3.bias.detach(), c3.logs.detach().reshape(-1), z, affine, False,
                                 c3.logscale_factor, save_h=True)
    ld_out = None
    if ld_vec is not None:
        ld_out = K.logdet_finish(ld_vec, n, h * w, logs=l, logabsdet=logabsdet, partials=partials,
                                 logscale_factor=an.logscale_factor, sign=1.0)
    ctx = dict(x=x, y=z, hrows=hrows, a1=save["a1"], h1=save["h1"], h2=save["h2"], wmat=wmat, winv=winv, idx=idx)
    return z, ld_out, ctx


def flowstep_backward(step, ctx, dy, dld):
    """Adjoint of flowstep_forward_save.  dy: grad wrt y; dld: grad wrt ld_out ([N]) or None.
    Accumulates every parameter gradient of the step; returns dx."""
    x, y = ctx["x"], ctx["y"]
    n, c, h, w = x.shape
    net = step.f
    c1, c2, c3 = net[0], net[2], net[4]
    an, an1, an2 = step.actnorm, c1.actnorm, c2.actnorm
    hid = net.hidden_channels
    kh = round_up(hid, 64)
    affine = step.coupling == 'affine'
    cout = net.out_channels
    h1, h2, a1 = ctx["h1"], ctx["h2"], ctx["a1"]
    dt = _C.BF16 if h1.dtype == torch.bfloat16 else _C.F32
    dev = x.device

    # (1) coupling + Conv2dZeros scale
    dz, du = K.coupling_bwd(y, ctx["hrows"], dy, dld, c3.logs.detach().reshape(-1), affine, _gbuf(c3.logs),
                            _gbuf(c3.bias), c3.logscale_factor)
    # (2) conv3 (tap form): dP3 = flipped im2col of du
    k3p = round_up(9 * cout, 64)
    d3col = K.im2col_rows(du, n, h, w, 0, cout, 3, dt, k3p, flip=True)
    tmp3 = torch.zeros(k3p, kh, device=dev, dtype=torch.float32)
    K.gemm_wgrad(d3col, h2, k3p, hid, tmp3)
    K.unpack_weight_grad(tmp3, cout, hid, 3, 1, _gbuf(c3.weight))
    d2 = K.gemm(d3col, net.packed("w3t", dt), hid, k3p, _C.EPI_RELU_BWD, None, an2.logs.detach().reshape(-1),
                an2.logscale_factor, y=h2, dlogs=_gbuf(an2.logs), dbias=_gbuf(an2.bias), out_dtype=dt, ldo=kh)
    # (3) conv2 (1x1)
    K.gemm_wgrad(d2, h1, hid, hid, _gbuf(c2.weight).view(hid, hid))
    d1 = K.gemm(d2, net.packed("w2t", dt), hid, hid, _C.EPI_RELU_BWD, None, an1.logs.detach().reshape(-1),
                an1.logscale_factor, y=h1, dlogs=_gbuf(an1.logs), dbias=_gbuf(an1.bias), out_dtype=dt, ldo=kh)
    # (4) conv1 (im2col form)
    k1p = net.k1p
    tmp1 = torch.zeros(kh, k1p, device=dev, dtype=torch.float32)
    K.gemm_wgrad(d1, a1, hid, k1p, tmp1)
    K.unpack_weight_grad(tmp1, hid, net.in_channels, 3, 0, _gbuf(c1.weight))
    da1 = K.gemm(d1, net.packed("w1t", dt), k1p, hid, _C.EPI_STORE, out_dtype=_C.F32)
    K.tapsum_to_nchw(da1, dz, 0, net.in_channels, flip=True, accumulate=True)
    # (5) ActNorm + mix
    gw = _gbuf(step.invconv.weight) if (step.permutation == 'invconv' and not step.invconv.lu_decomposition) else None
    if step.permutation == 'invconv' and gw is None:
        gw = torch.zeros(c * c, device=dev, dtype=torch.float32)
    dx = K.actnorm_mix_bwd(x, dz, ctx["wmat"], ctx["idx"], an.bias.detach().reshape(-1), an.logs.detach().reshape(-1),
                           gw, _gbuf(an.logs), _gbuf(an.bias), an.logscale_factor)
    if dld is not None:
        K.logdet_param_grad(dld, h * w, _gbuf(an.logs), ctx["winv"], gw, an.logscale_factor)
    if step.permutation == 'invconv' and step.invconv.lu_decomposition:
        step.invconv.accumulate_lu_grads(gw.view(c, c))
    return dx


# ------------------------------------------------------------------ Split2d
def split2d_forward_save(sp, x, ld_vec):
    """Split2d forward (network/module.py:526-530).  Returns (z1 view, ld_out, ctx)."""
    n, c, h, w = x.shape
    ch = c // 2
    conv = sp.conv2d_zeros
    from . import config
    dt = config.resolve_conv_dtype(64, sp.conv_dtype)
    kp = round_up(9 * ch, 64)
    a = K.im2col(x, 0, ch, 3, dt, kp)
    wp = conv._packs.get(("w0", dt), conv.weight,
                         lambda: K.pack_conv_weight(conv.weight.detach(), 0, dt, round_up(c, 16), kp))
    hrows = K.gemm(a, wp, c, kp, _C.EPI_ZEROS, conv.bias.detach(), conv.logs.detach().reshape(-1),
                   conv.logscale_factor, out_dtype=_C.F32)
    ld_out = K.gaussian_logp(hrows, x, ch, ch, ld_vec)
    return x[:, :ch], ld_out, dict(x=x, a=a, hrows=hrows, dt=dt, kp=kp)


def split2d_backward(sp, ctx, dz1, dld):
    x, hrows, a, dt, kp = ctx["x"], ctx["hrows"], ctx["a"], ctx["dt"], ctx["kp"]
    n, c, h, w = x.shape
    ch = c // 2
    conv = sp.conv2d_zeros
    dev = x.device
    if dld is None:
        dld = torch.zeros(n, device=dev, dtype=torch.float32)
    dx, du = K.split2d_bwd(x, hrows, dz1, dld, conv.logs.detach().reshape(-1), _gbuf(conv.logs), _gbuf(conv.bias),
                           conv.logscale_factor)
    cp = round_up(c, 64)
    duc = K.im2col_rows(du, n, h, w, 0, c, 1, dt, cp)                       # convert + zero-pad to the GEMM tiling
    tmp = torch.zeros(cp, kp, device=dev, dtype=torch.float32)
    K.gemm_wgrad(duc, a, cp, kp, tmp)
    K.unpack_weight_grad(tmp, c, ch, 3, 0, _gbuf(conv.weight))
    wt = conv._packs.get(("w0t", dt), conv.weight,
                         lambda: K.pack_conv_weight(conv.weight.detach(), 2, dt, kp, cp))
    da = K.gemm(duc, wt, kp, cp, _C.EPI_STORE, out_dtype=_C.F32)
    K.tapsum_to_nchw(da, dx, 0, ch, flip=True, accumulate=True)
    return dx


# ------------------------------------------------------------------ whole-model node
class FlowEncodeFunction(torch.autograd.Function):
    """FlowModel.encode (network/model.py:263-276) as a single autograd node."""

    @staticmethod
    def forward(ctx, flow, z, logdet, *params):
        from .model import FlowStep
        from .module import Split2d, Squeeze2d
        from . import rows_path
        z = z.detach().contiguous()
        ld = None if logdet is None else logdet.detach().contiguous()
        ctx.flow = flow
        ctx.rows = config.use_rows_path and rows_path.supported(flow, z)
        tape = []
        if ctx.rows:
            z, ld = rows_path.encode(flow, z, ld, tape)
            ctx.tape = tape
            ctx.has_ld = ld is not None
            return z if ld is None else (z, ld)
        # hybrid: the leading levels on the pixel-major kernels, the wide tail (C > 96) on the per-layer NCHW kernels
        ctx.head = None
        tail = flow.layers
        hd = rows_path.head(flow, z) if config.use_rows_path else None
        if hd is not None:
            view, k = hd
            head_tape = []
            z, ld = rows_path.encode(view, z, ld, head_tape)
            ctx.head = (view, head_tape)
            tail = list(flow.layers)[k:]
        for layer in tail:
            if isinstance(layer, Squeeze2d):
                z = K.squeeze2d(z, layer.factor, reverse=False)
                tape.append((layer, None))
            elif isinstance(layer, FlowStep):
                z, ld, c = flowstep_forward_save(layer, z, ld)
                tape.append((layer, c))
            elif isinstance(layer, Split2d):
                z, ld, c = split2d_forward_save(layer, z.contiguous(), ld)
                tape.append((layer, c))
            else:
                raise TypeError("unexpected layer %r" % type(layer))
        ctx.tape = tape
        ctx.has_ld = ld is not None
        if ld is None:
            return z.contiguous() if not z.is_contiguous() else z
        return z, ld

    @staticmethod
    def backward(ctx, dz, dld=None):
        from .model import FlowStep
        from .module import Split2d
        from . import rows_path
        if dz is None:
            raise RuntimeError("FlowModel.encode: the latent z must take part in the loss")
        dz = dz.contiguous()
        if dld is not None:
            dld = dld.contiguous()
        if ctx.rows:
            dx = rows_path.backward(ctx.flow, ctx.tape, dz, dld)
            ctx.tape = None
            return (None, dx, dld) + (None,) * (len(ctx.needs_input_grad) - 3)
        for layer, c in reversed(ctx.tape):
            if c is None:
                dz = K.squeeze2d(dz, layer.factor, reverse=True)
            elif isinstance(layer, FlowStep):
                dz = flowstep_backward(layer, c, dz, dld)
            elif isinstance(layer, Split2d):
                dz = split2d_backward(layer, c, dz, dld)
        ctx.tape = None
        if ctx.head is not None:
            view, head_tape = ctx.head
            dz = rows_path.backward(view, head_tape, dz.contiguous(), dld)
            ctx.head = None
        return (None, dz, dld) + (None,) * (len(ctx.needs_input_grad) - 3)


def flow_encode_autograd(flow, z, logdet):
    """Entry used by FlowModel.encode when gradients are required.  logdet: None | number | tensor."""
    n = z.shape[0]
    scalar_like = False
    if logdet is None:
        ld = None
    elif torch.is_tensor(logdet) and logdet.dim() == 1:
        ld = logdet.to(torch.float32)
    else:
        val = float(logdet)
        ld = torch.full((n,), val, device=z.device, dtype=torch.float32)
        scalar_like = True
    params = [p for p in flow.parameters() if p.requires_grad]
    out = FlowEncodeFunction.apply(flow, z, ld, *params)
    if ld is None:
        return out, None
    return out[0], out[1]


class TopPriorFunction(torch.autograd.Function):
    """objective + log N(z; 0, I) (network/model.py:435-438 with an all-zero h_top)."""

    @staticmethod
    def forward(ctx, z, objective):
        ctx.save_for_backward(z)
        return K.gaussian_logp(None, z.contiguous(), 0, z.shape[1], objective.contiguous())

    @staticmethod
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        return -z * g.view(-1, 1, 1, 1), g


class GlowNLLFunction(torch.autograd.Function):
    """Glow.normal_flow + Glow.generative_loss with the plain N(0, I) top prior as ONE autograd node
    (network/model.py:409-452, 496-498): dequantisation add inside the entry squeeze, the flow on the pixel-major
    kernels with its logdet started at zero, and the loss head (objective start value, top prior, bits/dim scaling,
    batch mean) in one kernel.  Returns (z, nll [N], loss); parameter gradients are accumulated into .grad."""

    @staticmethod
    def forward(ctx, flow, n_bins, x, noise, *params):
        from . import rows_path
        import math
        x = x.detach().contiguous()
        d_x = x[0].numel()
        tape = []
        z, ld = rows_path.encode(flow, x, None, tape, add=None if noise is None else noise.detach().contiguous(),
                                 want_ld=True)
        ctx.denom = math.log(2.) * d_x
        nll, loss = K.nll_head(z, ld, -math.log(n_bins) * d_x, ctx.denom)
        ctx.flow, ctx.tape = flow, tape
        # z is an OUTPUT of this node: keeping it as a plain attribute would tie node -> z -> node into a reference
        # cycle, and a node that outlives its iteration keeps the parameters' AccumulateGrad nodes (and the stream
        # they were created on) alive -- fatal for the next CUDA-graph capture.  save_for_backward breaks the cycle.
        ctx.save_for_backward(z)
        ctx.set_materialize_grads(False)
        return z, nll, loss

    @staticmethod
    def backward(ctx, dz, dnll, dloss):
        from . import rows_path
        if dz is None and dnll is None and dloss is None:
            return (None,) * len(ctx.needs_input_grad)
        f = lambda t: None if t is None else t.contiguous().float()
        (z,) = ctx.saved_tensors
        dz_top, dld = K.nll_head_bwd(z, ctx.denom, g_loss=f(dloss), g_nll=f(dnll), dz_in=f(dz))
        if dloss is None and dnll is None:
            dld.zero_()                                     # only z took part in the loss
            dz_top = f(dz)
        dx = rows_path.backward(ctx.flow, ctx.tape, dz_top, dld)
        ctx.tape = None
        # the dequantised input is x + noise: both receive the same gradient
        return (None, None, dx if ctx.needs_input_grad[2] else None, dx if ctx.needs_input_grad[3] else None) + \
            (None,) * (len(ctx.needs_input_grad) - 4)


# ------------------------------------------------------------------ stand-alone layer nodes
class _StepFunction(torch.autograd.Function):
    """A FlowStep called as a layer: on the pixel-major kernels (fused coupling net, packed weight gradients) behind a
    layout change when the step's shape allows it, else on the per-layer NCHW kernels."""

    @staticmethod
    def forward(ctx, step, x, logdet, *params):
        from . import rows_path
        x = x.detach().contiguous()
        ld = None if logdet is None else logdet.detach().contiguous()
        ctx.rows = step._rows_route(x)
        if ctx.rows:
            y, ld, c = rows_path.step_forward_nchw(step, x, ld, save=True)
        else:
            y, ld, c = flowstep_forward_save(step, x, ld)
        ctx.step, ctx.c = step, c
        return (y, ld) if ld is not None else y

    @staticmethod
    def backward(ctx, dy, dld=None):
        from . import rows_path
        dld_c = None if dld is None else dld.contiguous()
        if ctx.rows:
            dx = rows_path.step_backward_nchw(ctx.step, ctx.c, dy.contiguous(), dld_c)
        else:
            dx = flowstep_backward(ctx.step, ctx.c, dy.contiguous(), dld_c)
        ctx.c = None
        return (None, dx, dld) + (None,) * (len(ctx.needs_input_grad) - 3)


def _ld_vec(logdet, n, device):
    if logdet is None:
        return None
    if torch.is_tensor(logdet) and logdet.dim() == 1:
        return logdet.to(torch.float32)
    return torch.full((n,), float(logdet), device=device, dtype=torch.float32)


def flowstep_autograd(step, x, logdet):
    ld = _ld_vec(logdet, x.shape[0], x.device)
    params = [p for p in step.parameters() if p.requires_grad]
    out = _StepFunction.apply(step, x, ld, *params)
    return (out, None) if ld is None else (out[0], out[1])


class _SplitFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sp, x, logdet, *params):
        z1, ld, c = split2d_forward_save(sp, x.detach().contiguous(), logdet.detach().contiguous())
        ctx.sp, ctx.c = sp, c
        return z1.contiguous(), ld

    @staticmethod
    def backward(ctx, dz1, dld):
        dx = split2d_backward(ctx.sp, ctx.c, dz1, dld)
        return (None, dx, dld) + (None,) * (len(ctx.needs_input_grad) - 3)


def split2d_autograd(sp, x, logdet):
    ld = _ld_vec(0. if logdet is None else logdet, x.shape[0], x.device)
    params = [p for p in sp.parameters() if p.requires_grad]
    return _SplitFunction.apply(sp, x, ld, *params)


class SqueezeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, factor, reverse):
        ctx.factor, ctx.reverse = factor, reverse
        return K.squeeze2d(x.detach(), factor, reverse)

    @staticmethod
    def backward(ctx, g):
        return K.squeeze2d(g.contiguous(), ctx.factor, not ctx.reverse), None, None


class _MixFunction(torch.autograd.Function):
    """ActNorm / Invertible1x1Conv / Permutation2d as stand-alone differentiable layers (forward direction)."""

    @staticmethod
    def forward(ctx, x, weight, indices, bias, logs, f, owner):
        x = x.detach().contiguous()
        ctx.args = (x, None if weight is None else weight.detach(), indices,
                    None if bias is None else bias.detach().reshape(-1),
                    None if logs is None else logs.detach().reshape(-1), f, owner)
        return K.actnorm_mix(x, ctx.args[1], indices, ctx.args[3], ctx.args[4], f, reverse=False)

    @staticmethod
    def backward(ctx, dz):
        x, w, idx, b, l, f, owner = ctx.args
        gw = _gbuf(owner["weight"]) if w is not None else None
        gl = _gbuf(owner["logs"]) if l is not None else None
        gb = _gbuf(owner["bias"]) if b is not None else None
        dx = K.actnorm_mix_bwd(x, dz.contiguous(), w, idx, b, l, gw, gl, gb, f)
        return dx, None, None, None, None, None, None


def actnorm_autograd(an, x, logdet, reverse):
    if reverse:
        raise NotImplementedError("gradients through the reverse direction are not needed by the reference's callers")
    idx = torch.arange(an.num_channels, device=x.device)
    y = _MixFunction.apply(x, None, idx, an.bias, an.logs, an.logscale_factor, {"bias": an.bias, "logs": an.logs})
    if logdet is None:
        return y, None
    d = torch.sum(an.logs * an.logscale_factor) * (x.shape[2] * x.shape[3])     # sample-independent scalar term
    return y, logdet + d


def invconv_autograd(ic, x, logdet, reverse):
    if reverse or ic.lu_decomposition:
        raise NotImplementedError("stand-alone autograd covers the dense forward direction; use FlowStep for LU")
    z = _MixFunction.apply(x, ic.weight, None, None, None, 3.0, {"weight": ic.weight})
    if logdet is None:
        return z, None
    return z, logdet + torch.log(torch.abs(torch.det(ic.weight))) * (x.shape[2] * x.shape[3])


class PermuteFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pm, reverse):
        ctx.pm, ctx.reverse = pm, reverse
        return K.actnorm_mix(x.detach().contiguous(), indices=pm.device_indices(x.device, reverse), reverse=False)

    @staticmethod
    def backward(ctx, g):
        return K.actnorm_mix(g.contiguous(), indices=ctx.pm.device_indices(g.device, not ctx.reverse), reverse=False), None, None
