"""Host-side mirror of pytorch-glow's flow layers (network/module.py).

Same class names, constructor signatures, call conventions, attributes and
``state_dict()`` keys as the reference, so reference snapshots load and the
reference's trainer / inferer / tests can drive these classes unchanged.  The
arithmetic runs in libglowk.so (hand-written sm_100a CUDA); there is no CPU path.

Citations are file:line in corenel/pytorch-glow.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import _C, config, ops
from . import functional as K
from .functional import round_up


# ------------------------------------------------------------------ logdet conventions
def _logdet_in(logdet, n, device):
    """Normalise the reference's logdet argument (None | number | 0-dim | [N] tensor).

    Returns (vector or None, scalar_like).  scalar_like inputs become a 1-element vector and the
    caller turns the result back into a 0-dim tensor (reference broadcasting semantics)."""
    if logdet is None:
        return None, False
    if not torch.is_tensor(logdet):
        return torch.full((1,), float(logdet), device=device, dtype=torch.float32), True
    if logdet.dim() == 0:
        return logdet.detach().to(device=device, dtype=torch.float32).reshape(1), True
    assert logdet.dim() == 1 and logdet.shape[0] == n, "logdet must be a [N] tensor"
    return logdet.to(device=device, dtype=torch.float32).contiguous(), False


def _logdet_out(vec, scalar_like):
    return vec.reshape(()) if scalar_like else vec


_WEIGHT_GENERATION = [0]


def bump_weight_generation():
    """Invalidate every packed-weight / LU cache.  Called by optimizers that update parameters through
    raw pointers (train.FusedTrainStep), which torch's per-tensor version counter cannot see."""
    _WEIGHT_GENERATION[0] += 1


class _PackCache:
    """Packed (GEMM-layout) copies of a parameter, rebuilt when the parameter changes."""

    def __init__(self):
        self._d = {}

    def get(self, key, param, build, extra=()):
        """`extra`: further validity components (versions of sibling parameters); a stale entry is OVERWRITTEN under
        the same key, never accumulated."""
        tag = (param.data_ptr(), param._version, param.device, _WEIGHT_GENERATION[0]) + tuple(extra)
        hit = self._d.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        val = build()
        self._d[key] = (tag, val)
        return val

    def fresh(self, key, param):
        hit = self._d.get(key)
        return hit is not None and hit[0] == (param.data_ptr(), param._version, param.device, _WEIGHT_GENERATION[0])

    def put(self, key, param, val):
        self._d[key] = ((param.data_ptr(), param._version, param.device, _WEIGHT_GENERATION[0]), val)

    def clear(self):
        self._d.clear()


def _wants_grad(mod, x):
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in mod.parameters()))


def _no_standalone_autograd(name):
    raise NotImplementedError(
        "pytorch_glow_b200.%s has no stand-alone autograd: its forward runs the CUDA kernels on detached weights.  "
        "Gradients of the coupling network are produced by the fused FlowStep / FlowModel nodes "
        "(FlowStep.forward, FlowModel.encode, Glow.forward); call it under torch.no_grad() or train through those." % name)


# ------------------------------------------------------------------ ActNorm
class ActNorm(nn.Module):
    """Activation normalisation (network/module.py:9-149)."""

    def __init__(self, num_channels, scale=1., logscale_factor=3., batch_variance=False):
        super().__init__()
        self.num_channels = num_channels
        self.scale = scale
        self.logscale_factor = logscale_factor
        self.batch_variance = batch_variance
        self.bias_inited = False
        self.logs_inited = False
        self.register_parameter('bias', nn.Parameter(torch.zeros(1, self.num_channels, 1, 1)))
        self.register_parameter('logs', nn.Parameter(torch.zeros(1, self.num_channels, 1, 1)))

    @property
    def needs_init(self):
        """True iff the next training-mode call performs the data-dependent init (module.py:93-94)."""
        return self.training and not (self.bias_inited and self.logs_inited)

    def initialize_from_nchw(self, x, reverse=False):
        """module.py:86-120 on a [N,C,H,W] batch.  reverse: the first training-mode call came in the reverse direction
        (module.py:143-146: logs from the raw input, then the bias from the scaled input)."""
        with torch.no_grad():
            b, l = K.actnorm_init_nchw(x, self.scale, self.logscale_factor, self.batch_variance, reverse)
            self._store_init(b, l)

    def initialize_from_rows(self, rows):
        """Same, on a pixel-major fp32 matrix [P][>=C] (conv output before the ActNorm)."""
        with torch.no_grad():
            b, l = K.actnorm_init_rows(rows, self.num_channels, self.scale, self.logscale_factor, self.batch_variance)
            self._store_init(b, l)

    def _store_init(self, b, l):
        self.bias.data.copy_(b.view_as(self.bias))
        self.logs.data.copy_(l.view_as(self.logs))
        self.bias_inited = True
        self.logs_inited = True

    def logdet_term(self, logdet, n, hw, reverse, device):
        vec, scalar_like = _logdet_in(logdet, n, device)
        if vec is None:
            return None
        out = K.logdet_finish(vec, vec.shape[0], hw, logs=self.logs.detach().reshape(-1),
                              logscale_factor=self.logscale_factor, sign=-1.0 if reverse else 1.0)
        return _logdet_out(out, scalar_like)

    def forward(self, x, logdet=None, reverse=False):
        assert len(x.shape) == 4
        assert x.shape[1] == self.num_channels, \
            'Input shape should be NxCxHxW, however channels are {} instead of {}'.format(x.shape[1], self.num_channels)
        assert x.device == self.bias.device and x.device == self.logs.device, \
            'Expect input device {} instead of {}'.format(self.bias.device, x.device)
        if self.needs_init:
            self.initialize_from_nchw(x, reverse=reverse)
        if torch.is_grad_enabled() and (x.requires_grad or self.bias.requires_grad):
            from .autograd import actnorm_autograd
            return actnorm_autograd(self, x, logdet, reverse)
        y = K.actnorm(x, self.bias.detach().reshape(-1), self.logs.detach().reshape(-1), self.logscale_factor, reverse)
        return y, self.logdet_term(logdet, x.shape[0], x.shape[2] * x.shape[3], reverse, x.device)


class LinearZeros(nn.Linear):
    """network/module.py:152-185.  Only used with y_condition (off in every BASELINE config): plain torch."""

    def __init__(self, in_features, out_features, bias=True, logscale_factor=3.):
        super().__init__(in_features, out_features, bias)
        self.logscale_factor = logscale_factor
        self.weight.data.zero_()
        self.bias.data.zero_()
        self.register_parameter('logs', nn.Parameter(torch.zeros(out_features)))

    def forward(self, x):
        output = super().forward(x)
        return output * torch.exp(self.logs * self.logscale_factor)


# ------------------------------------------------------------------ convolutions of the coupling net
def _conv_rows(x, weight, ksize, dtype, epilogue, bias, logs, logscale_factor, cache, key):
    """NCHW fp32 -> pixel-major GEMM (im2col, taps folded into K) -> rows [P][Cout] fp32."""
    n, cin, h, w = x.shape
    cout = weight.shape[0]
    kp = round_up(ksize * ksize * cin, 64)
    rows = K.im2col(x, 0, cin, ksize, dtype, kp)
    wp = cache.get((key, dtype), weight,
                   lambda: K.pack_conv_weight(weight.detach(), 0, dtype, round_up(cout, 16), kp))
    return K.gemm(rows, wp, cout, kp, epilogue, bias, logs, logscale_factor, out_dtype=_C.F32)


class Conv2d(nn.Conv2d):
    """Conv (SAME, no bias) + ActNorm (network/module.py:188-260)."""

    @staticmethod
    def get_padding(padding_type, kernel_size, stride):
        assert padding_type in ['SAME', 'VALID'], "Unsupported padding type: {}".format(padding_type)
        if isinstance(kernel_size, int):
            kernel_size = [kernel_size, kernel_size]
        if padding_type == 'SAME':
            assert stride == 1, "'SAME' padding only supports stride=1"
            return tuple((k - 1) // 2 for k in kernel_size)
        return tuple(0 for _ in kernel_size)

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1, padding_type='SAME',
                 do_weightnorm=False, do_actnorm=True, dilation=1, groups=1):
        padding = self.get_padding(padding_type, kernel_size, stride)
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                         bias=(not do_actnorm))
        self.do_weight_norm = do_weightnorm
        self.do_actnorm = do_actnorm
        self.padding_type = padding_type
        self.weight.data.normal_(mean=0.0, std=0.05)
        if self.do_actnorm:
            self.actnorm = ActNorm(out_channels)
        else:
            self.bias.data.zero_()
        self._packs = _PackCache()

    def _check_supported(self):
        k = self.kernel_size
        if not (k[0] == k[1] and k[0] in (1, 3) and self.padding_type == 'SAME' and self.groups == 1
                and tuple(self.dilation) == (1, 1) and tuple(self.stride) == (1, 1)):
            raise NotImplementedError("glowk convs: square 1x1/3x3, stride 1, SAME padding, groups=1, dilation=1")
        return k[0]

    def forward(self, x, conv_dtype=None):
        ks = self._check_supported()
        _C.check_cuda(x)
        if _wants_grad(self, x):
            _no_standalone_autograd("Conv2d")
        n, _, h, w = x.shape
        dt = config.resolve_conv_dtype(64, conv_dtype)   # K and N are padded to the tensor-core tiling
        f = 3.0
        if self.do_actnorm:
            an = self.actnorm
            f = an.logscale_factor
            if an.needs_init:
                pre = _conv_rows(x, self.weight, ks, dt, _C.EPI_STORE, None, None, f, self._packs, "w")
                an.initialize_from_rows(pre)
            bias, logs = an.bias.detach().reshape(-1), an.logs.detach().reshape(-1)
        else:
            bias = self.bias.detach()
            logs = torch.zeros_like(bias)
        rows = _conv_rows(x, self.weight, ks, dt, _C.EPI_ACTNORM, bias, logs, f, self._packs, "w")
        return K.rows_to_nchw(rows, n, self.out_channels, h, w)


class Conv2dZeros(nn.Conv2d):
    """Zero-initialised conv (+bias) * exp(logs*factor) (network/module.py:263-297)."""

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1, padding_type='SAME',
                 logscale_factor=3, dilation=1, groups=1, bias=True):
        padding = Conv2d.get_padding(padding_type, kernel_size, stride)
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.logscale_factor = logscale_factor
        self.padding_type = padding_type
        self.bias.data.zero_()
        self.weight.data.zero_()
        self.register_parameter("logs", nn.Parameter(torch.zeros(out_channels, 1, 1)))
        self._packs = _PackCache()

    def forward_rows(self, x, c0, cin, conv_dtype=None):
        """Rows [P][Cout] fp32 of this conv applied to channels c0..c0+cin of NCHW x (no NCHW output)."""
        assert tuple(self.kernel_size) == (3, 3) and self.padding_type == 'SAME'
        dt = config.resolve_conv_dtype(64, conv_dtype)
        cout = self.out_channels
        kp = round_up(9 * cin, 64)
        rows = K.im2col(x, c0, cin, 3, dt, kp)
        wp = self._packs.get(("w0", dt), self.weight,
                             lambda: K.pack_conv_weight(self.weight.detach(), 0, dt, round_up(cout, 16), kp))
        return K.gemm(rows, wp, cout, kp, _C.EPI_ZEROS, self.bias.detach(), self.logs.detach().reshape(-1),
                      self.logscale_factor, out_dtype=_C.F32)

    def forward(self, x, conv_dtype=None):
        _C.check_cuda(x)
        if _wants_grad(self, x):
            # `Glow.prior` with ablation.learn_top (network/model.py:371-373) trains this conv on the all-zero h_top:
            # a [B, 2C, H_top, W_top] side input off the flow's hot path.  It stays differentiable through ATen (like
            # LinearZeros, which the reference also keeps in torch) instead of silently detaching its weights.
            out = torch.nn.functional.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
            return out * torch.exp(self.logs * self.logscale_factor)
        n, cin, h, w = x.shape
        if tuple(self.kernel_size) == (3, 3):
            rows = self.forward_rows(x, 0, cin, conv_dtype)
        else:
            dt = config.resolve_conv_dtype(64, conv_dtype)
            rows = _conv_rows(x, self.weight, self.kernel_size[0], dt, _C.EPI_ZEROS, self.bias.detach(),
                              self.logs.detach().reshape(-1), self.logscale_factor, self._packs, "w")
        return K.rows_to_nchw(rows, n, self.out_channels, h, w)


class CouplingNet(nn.Sequential):
    """The coupling network `f()` (network/module.py:300-319): Conv2d 3x3 + ReLU + Conv2d 1x1 + ReLU +
    Conv2dZeros 3x3, with the reference's Sequential indices (0, 2, 4) so state_dict keys match.

    The three convs run as pixel-major GEMMs: conv1 = im2col (taps in K), conv2 = plain,
    conv3 = nine pointwise GEMMs folded into N followed by a 3x3 tap gather-sum."""

    def __init__(self, in_channels, hidden_channels, out_channels):
        super().__init__(
            Conv2d(in_channels, hidden_channels),
            nn.ReLU(inplace=True),
            Conv2d(hidden_channels, hidden_channels, kernel_size=1),
            nn.ReLU(inplace=True),
            Conv2dZeros(hidden_channels, out_channels))
        self.in_channels = in_channels
        self.hidden_channels = hidden_channels
        self.out_channels = out_channels
        self.k1p = round_up(9 * in_channels, 64)
        self.n3 = 9 * out_channels
        self.n3p = round_up(self.n3, 16)
        self._packs = _PackCache()

    def dtype(self, override=None):
        return config.resolve_conv_dtype(self.hidden_channels, override)

    def fused(self, backward):
        """True iff the fused coupling-net kernels (glowk_cnet_forward / glowk_cnet_backward) serve this net's
        shapes on the current device; GLOWK_CNET_FUSED=0 forces the three-GEMM path (A/B runs)."""
        key = "_fused_bwd" if backward else "_fused_fwd"
        ok = self.__dict__.get(key)
        if ok is None:
            if os.environ.get("GLOWK_CNET_FUSED", "1") == "0":
                ok = False
            elif backward:
                ok = K.cnet_fused_supported(True, round_up(self.n3, 64), self.hidden_channels, self.k1p)
            else:
                ok = K.cnet_fused_supported(False, self.k1p, self.hidden_channels, self.n3p)
            self.__dict__[key] = ok
        return ok

    def packed(self, which, dt):
        """GEMM-layout weight copies, cached per parameter version (see glowk_pack_conv_weight)."""
        c1, c2, c3 = self[0], self[2], self[4]
        hid, hp = self.hidden_channels, round_up(self.hidden_channels, 16)
        if which == "w1":      # [hid][k1p], k = tap*Cin + ci
            return self._packs.get(("w1", dt), c1.weight, lambda: K.pack_conv_weight(c1.weight.detach(), 0, dt, hp, self.k1p))
        if which == "w2":      # [hid][hid]
            return self._packs.get(("w2", dt), c2.weight, lambda: K.pack_conv_weight(c2.weight.detach(), 0, dt, hp, round_up(hid, 64)))
        if which == "w3":      # [9*Cout (pad 16)][hid], row = tap*Cout + co
            return self._packs.get(("w3", dt), c3.weight, lambda: K.pack_conv_weight(c3.weight.detach(), 1, dt, self.n3p, round_up(hid, 64)))
        if which == "w1t":     # [k1p][hid]   (dgrad of conv1: B operand [N=k1p][K=hid])
            return self._packs.get(("w1t", dt), c1.weight, lambda: K.pack_conv_weight(c1.weight.detach(), 2, dt, self.k1p, round_up(hid, 64)))
        if which == "w2t":     # [hid(in)][hid(out)]
            return self._packs.get(("w2t", dt), c2.weight, lambda: K.pack_conv_weight(c2.weight.detach(), 2, dt, hp, round_up(hid, 64)))
        if which == "w3t":     # [hid][9*Cout pad 64]   (dgrad of conv3: B operand [N=hid][K=9*Cout])
            return self._packs.get(("w3t", dt), c3.weight, lambda: K.pack_conv_weight(c3.weight.detach(), 3, dt, hp, round_up(self.n3, 64)))
        raise KeyError(which)

    def pack_specs(self, dt, backward):
        """[(key, parameter, layout, rows, ld)] of the GEMM-layout weight copies this net uses (see `packed`)."""
        c1, c2, c3 = self[0], self[2], self[4]
        hid, hp, kh = self.hidden_channels, round_up(self.hidden_channels, 16), round_up(self.hidden_channels, 64)
        specs = [("w1", c1.weight, 0, hp, self.k1p), ("w2", c2.weight, 0, hp, kh), ("w3", c3.weight, 1, self.n3p, kh)]
        if backward:
            specs += [("w1t", c1.weight, 2, self.k1p, kh), ("w2t", c2.weight, 2, hp, kh),
                      ("w3t", c3.weight, 3, hp, round_up(self.n3, 64))]
        return specs

    def tap_rows(self, z, conv_dtype=None, save=None):
        """P3 rows [P][n3p] fp32 (conv3 before the tap gather-sum) from channels 0..Cin-1 of NCHW z."""
        dt = self.dtype(conv_dtype)
        a1 = K.im2col(z, 0, self.in_channels, 3, dt, self.k1p)
        return self.tap_rows_from_a1(a1, dt, save)

    def tap_rows_from_rows(self, z, n, h, w, dt, save=None, ones_col=-1):
        """P3 rows [P][n3p] fp32 from channels 0..Cin-1 of the pixel-major flow state z [P][C] fp32.  On the fused
        bf16 path conv1 is an implicit GEMM: the kernel gathers its im2col operand itself (no glowk_im2col_rows, no
        a1 in HBM when sampling).  `save`, if a dict, receives a1 / h1 / h2 for the backward pass."""
        an1, an2 = self[0].actnorm, self[2].actnorm
        # (A/B switch GLOWK_CNET_IMPLICIT_TRAIN=0: the training forward takes its conv1 operand from glowk_im2col_rows)
        implicit_ok = save is None or os.environ.get("GLOWK_CNET_IMPLICIT_TRAIN", "1") != "0"
        if (dt == _C.BF16 and not (an1.needs_init or an2.needs_init) and self.fused(False) and implicit_ok
                and self.in_channels % 2 == 0 and os.environ.get("GLOWK_CNET_IMPLICIT", "1") != "0"):
            # training: the ReLU masks of h1 / h2 also leave as bits -- the backward chain reads those instead of the
            # bf16 activations (cnet_backward_implicit; GLOWK_CNET_BITMASK=0: bf16 masks)
            masks = None
            if save is not None and self.fused(True) and os.environ.get("GLOWK_CNET_BITMASK", "1") != "0":
                masks = K.cnet_relu_masks(n * h * w, z.device)
            p3, a1, h1, h2 = K.cnet_forward_implicit(
                z, n, h, w, 0, self.in_channels, self.k1p, self.packed("w1", dt), self.packed("w2", dt),
                self.packed("w3", dt), self.hidden_channels, self.n3p, an1.bias.detach().reshape(-1),
                an1.logs.detach().reshape(-1), an1.logscale_factor, an2.bias.detach().reshape(-1),
                an2.logs.detach().reshape(-1), an2.logscale_factor, ldp3=self.n3p, save=save is not None,
                ldh=round_up(self.hidden_channels, 64), ones_col=ones_col, masks=masks)
            if save is not None:
                save.update(a1=a1, h1=h1, h2=h2, masks=masks)
            return p3
        a1 = K.im2col_rows(z, n, h, w, 0, self.in_channels, 3, dt, self.k1p, ones_col=ones_col)
        return self.tap_rows_from_a1(a1, dt, save, want_masks=True)

    def tap_rows_from_a1(self, a1, dt, save=None, want_masks=False):
        """The three GEMMs of the coupling net on a1 = im2col(z1) ([P][k1p]); returns P3 rows [P][n3p] fp32.

        Performs the data-dependent ActNorm init of the two hidden ActNorms on the first training
        call (module.py:86-120, 238-239).  `save`, if a dict, receives a1/h1/h2 for the backward pass."""
        c1, c2, c3 = self[0], self[2], self[4]
        hid = self.hidden_channels
        kh = round_up(hid, 64)
        w1 = self.packed("w1", dt)
        an1, an2 = c1.actnorm, c2.actnorm
        if dt == _C.BF16 and not (an1.needs_init or an2.needs_init) and self.fused(False):
            # one tcgen05 kernel for the three convs: h1 stays in tensor memory, h2 in shared memory
            # (csrc/cnet_fused_sm100.cu); bit-identical to the three GEMMs below
            masks = None
            if (want_masks and save is not None and self.fused(True) and self.out_channels % 2 == 0
                    and os.environ.get("GLOWK_CNET_IMPLICIT", "1") != "0" and os.environ.get("GLOWK_CNET_BITMASK", "1") != "0"):
                masks = K.cnet_relu_masks(a1.shape[0], a1.device)          # read by cnet_backward_implicit (rows path)
            p3, h1, h2 = K.cnet_forward(a1, w1, self.packed("w2", dt), self.packed("w3", dt), hid, self.n3p,
                                        an1.bias.detach().reshape(-1), an1.logs.detach().reshape(-1),
                                        an1.logscale_factor, an2.bias.detach().reshape(-1),
                                        an2.logs.detach().reshape(-1), an2.logscale_factor, ldp3=self.n3p,
                                        save=save is not None, ldh=kh, masks=masks)
            if save is not None:
                save.update(a1=a1, h1=h1, h2=h2, masks=masks)
            return p3
        if an1.needs_init:
            an1.initialize_from_rows(K.gemm(a1, w1, hid, self.k1p, _C.EPI_STORE, out_dtype=_C.F32))
        h1 = K.gemm(a1, w1, hid, self.k1p, _C.EPI_ACTNORM_RELU, an1.bias.detach().reshape(-1),
                    an1.logs.detach().reshape(-1), an1.logscale_factor, out_dtype=dt, ldo=kh)
        w2 = self.packed("w2", dt)
        if an2.needs_init:
            an2.initialize_from_rows(K.gemm(h1, w2, hid, hid, _C.EPI_STORE, out_dtype=_C.F32))
        h2 = K.gemm(h1, w2, hid, hid, _C.EPI_ACTNORM_RELU, an2.bias.detach().reshape(-1),
                    an2.logs.detach().reshape(-1), an2.logscale_factor, out_dtype=dt, ldo=kh)
        # K = hid (not the padded pitch): pad columns of h1/h2 are never read (TMA clips at K)
        p3 = K.gemm(h2, self.packed("w3", dt), self.n3, hid, _C.EPI_STORE, out_dtype=_C.F32, ldo=self.n3p)
        if save is not None:
            save.update(a1=a1, h1=h1, h2=h2)
        return p3

    def forward(self, x, conv_dtype=None):
        """Stand-alone NCHW -> NCHW evaluation (drop-in for the reference's nn.Sequential)."""
        _C.check_cuda(x)
        if _wants_grad(self, x):
            _no_standalone_autograd("f() / CouplingNet")
        x = x.contiguous()
        n, _, h, w = x.shape
        p3 = self.tap_rows(x, conv_dtype)
        out = torch.empty(n, self.out_channels, h, w, device=x.device, dtype=torch.float32)
        K.tapsum_to_nchw(p3, out, 0, self.out_channels)
        c3 = self[4]
        # (u + bias) * exp(logs*f): reuse the ActNorm kernel's forward form
        return K.actnorm(out, c3.bias.detach(), c3.logs.detach().reshape(-1), c3.logscale_factor, False, out=out)


def f(in_channels, hidden_channels, out_channels):
    """network/module.py:300-319."""
    return CouplingNet(in_channels, hidden_channels, out_channels)


# ------------------------------------------------------------------ invertible 1x1 conv, permutation
class Invertible1x1Conv(nn.Module):
    """network/module.py:322-369.  lu_decomposition=True is the LU parameterisation the reference
    leaves unimplemented (module.py:336-337): W = P L (U + diag(sign_s exp(log_s)))."""

    def __init__(self, num_channels, lu_decomposition=False):
        super().__init__()
        self.num_channels = num_channels
        self.lu_decomposition = lu_decomposition
        w_shape = [num_channels, num_channels]
        w_init = np.linalg.qr(np.random.randn(*w_shape))[0].astype('float32')   # module.py:341
        if not lu_decomposition:
            self.register_parameter('weight', nn.Parameter(torch.Tensor(w_init)))
        else:
            self._register_lu(torch.from_numpy(w_init))
        self._cache = _PackCache()

    def _register_lu(self, w):
        plu, piv = torch.linalg.lu_factor(w.double())
        pm, lm, um = torch.lu_unpack(plu, piv)
        s = torch.diagonal(um)
        self.register_buffer('p', pm.float())
        self.register_buffer('sign_s', torch.sign(s).float())
        self.register_parameter('l', nn.Parameter(torch.tril(lm, -1).float()))
        self.register_parameter('u', nn.Parameter(torch.triu(um, 1).float()))
        self.register_parameter('log_s', nn.Parameter(torch.log(torch.abs(s)).float()))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # import a dense reference snapshot (`invconv.weight`) into the LU parameterisation
        key = prefix + 'weight'
        if self.lu_decomposition and key in state_dict:
            w = state_dict.pop(key).detach().cpu().float()
            tmp = Invertible1x1Conv.__new__(Invertible1x1Conv)
            nn.Module.__init__(tmp)
            tmp._register_lu(w)
            for k in ('p', 'sign_s', 'l', 'u', 'log_s'):
                state_dict[prefix + k] = getattr(tmp, k).detach()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def prepared(self, need_inverse):
        """(W, W^-1 or None, log|det W| [1]) on the device, cached per parameter version."""
        if not self.lu_decomposition:
            w = self.weight
            ld, winv = self._cache.get(("dense", need_inverse), w, lambda: K.invconv_prepare(w.detach(), need_inverse))
            return w.detach(), winv, ld
        # fixed key: a torch optimizer bumps the versions of l / u / log_s every step, and a key that contained them
        # would leave one dead (W, W^-1, logdet) triple per step behind; the versions are validity tags
        extra = (self.l._version, self.u._version, self.l.data_ptr(), self.u.data_ptr())
        w, winv, ld = self._cache.get(("lu", need_inverse), self.log_s, lambda: K.invconv_lu_assemble(
            self.p, self.l.detach(), self.u.detach(), self.sign_s, self.log_s.detach(), need_inverse), extra=extra)
        return w, winv, ld

    def dense_weight(self):
        return self.prepared(False)[0]

    def accumulate_lu_grads(self, dw):
        """Chain dL/dW (CxC, includes the logdet term) into the LU parameters: W = P Lf Uf (one kernel)."""
        for prm in (self.l, self.u, self.log_s):
            if prm.grad is None:
                prm.grad = torch.zeros_like(prm)
        K.invconv_lu_grads(dw.contiguous(), self.p, self.l.detach(), self.u.detach(), self.sign_s, self.log_s.detach(),
                           self.l.grad, self.u.grad, self.log_s.grad)

    def forward(self, x, logdet=None, reverse=False):
        _C.check_cuda(x)
        n, c, h, w = x.shape
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import invconv_autograd
            return invconv_autograd(self, x, logdet, reverse)
        wmat, winv, ld = self.prepared(need_inverse=reverse)
        z = K.actnorm_mix(x, weight=winv if reverse else wmat, reverse=False)
        vec, scalar_like = _logdet_in(logdet, n, x.device)
        if vec is None:
            return z, None
        out = K.logdet_finish(vec, vec.shape[0], h * w, logabsdet=ld, sign=-1.0 if reverse else 1.0)
        return z, _logdet_out(out, scalar_like)


class Permutation2d(nn.Module):
    """network/module.py:372-397 (numpy-RNG shuffle at construction; indices are not in state_dict)."""

    def __init__(self, num_channels, shuffle=False):
        super().__init__()
        self.num_channels = num_channels
        self.indices = np.arange(self.num_channels - 1, -1, -1, dtype=np.int64)
        if shuffle:
            np.random.shuffle(self.indices)
        self.indices_inverse = np.zeros(self.num_channels, dtype=np.int64)
        for i in range(self.num_channels):
            self.indices_inverse[self.indices[i]] = i
        self._dev = {}

    def set_indices(self, indices):
        """Adopt an existing permutation (e.g. from a reference model, whose indices are not saved)."""
        self.indices = np.asarray(indices, dtype=np.int64).copy()
        self.indices_inverse = np.zeros(self.num_channels, dtype=np.int64)
        for i in range(self.num_channels):
            self.indices_inverse[self.indices[i]] = i
        self._dev = {}

    def device_indices(self, device, reverse):
        key = (str(device), bool(reverse))
        t = self._dev.get(key)
        if t is None:
            t = torch.from_numpy(self.indices_inverse if reverse else self.indices).to(device)
            self._dev[key] = t
        return t

    def forward(self, x, reverse=False):
        assert len(x.shape) == 4
        _C.check_cuda(x)
        if torch.is_grad_enabled() and x.requires_grad:
            from .autograd import PermuteFunction
            return PermuteFunction.apply(x, self, reverse)
        return K.actnorm_mix(x, indices=self.device_indices(x.device, reverse), reverse=False)


# ------------------------------------------------------------------ GaussianDiag, Split2d, Squeeze2d
class GaussianDiag:
    """network/module.py:400-483."""

    log_2pi = float(np.log(2 * np.pi))

    @staticmethod
    def eps(shape_tensor, eps_std=None):
        # torch's own generator on purpose: keeps the noise stream identical to the reference
        eps_std = eps_std or 1.
        if shape_tensor.is_cuda and torch.cuda.is_current_stream_capturing():
            # torch.normal(Tensor, Tensor) validates std on the host (a device sync): not capturable.  Same
            # generator stream and the same arithmetic (N(0,1) draw times std).
            return torch.randn_like(shape_tensor) * eps_std
        return torch.normal(mean=torch.zeros_like(shape_tensor), std=torch.ones_like(shape_tensor) * eps_std)

    @staticmethod
    def flatten_sum(tensor):
        assert len(tensor.shape) == 4
        return ops.reduce_sum(tensor, dim=[1, 2, 3])

    @staticmethod
    def logps(mean, logs, x):
        return -0.5 * (GaussianDiag.log_2pi + 2. * logs + ((x - mean) ** 2) / torch.exp(2. * logs))

    @staticmethod
    def logp(mean, logs, x):
        # generic (differentiable) form; Split2d and Glow call the fused kernel glowk_gaussian_logp directly
        return GaussianDiag.flatten_sum(GaussianDiag.logps(mean, logs, x))

    @staticmethod
    def sample(mean, logs, eps_std=None):
        eps = GaussianDiag.eps(mean, eps_std)
        return mean + torch.exp(logs) * eps


class Split2d(nn.Module):
    """network/module.py:486-536."""

    def __init__(self, num_channels):
        super().__init__()
        self.num_channels = num_channels
        self.conv2d_zeros = Conv2dZeros(num_channels // 2, num_channels)
        self.conv_dtype = None      # override of config.conv_dtype ("fp32" | "bf16" | None)

    def pack_specs(self, dt, backward):
        conv, c = self.conv2d_zeros, self.num_channels
        kp = round_up(9 * (c // 2), 64)
        specs = [(conv, "w0", conv.weight, 0, round_up(c, 16), kp)]
        if backward:
            specs.append((conv, "w0t", conv.weight, 2, kp, round_up(c, 64)))
        return specs

    def prior_rows(self, x, conv_dtype=None):
        """h rows [P][C] fp32: Conv2dZeros(z1) with (mean, logs) interleaved ('cross' split)."""
        return self.conv2d_zeros.forward_rows(x, 0, self.num_channels // 2, conv_dtype or self.conv_dtype)

    def prior(self, z):
        h = self.conv2d_zeros(z)
        return ops.split_channel(h, 'cross')

    def forward(self, x, logdet=0., reverse=False, eps_std=None, eps=None):
        _C.check_cuda(x)
        ch = self.num_channels // 2
        if not reverse:
            assert x.shape[1] == self.num_channels
            if torch.is_grad_enabled() and (x.requires_grad or self.conv2d_zeros.weight.requires_grad):
                from .autograd import split2d_autograd
                return split2d_autograd(self, x, logdet)
            x = x.contiguous()
            vec, _ = _logdet_in(logdet, x.shape[0], x.device)
            if vec is not None and vec.shape[0] == 1 and x.shape[0] != 1:
                vec = vec.expand(x.shape[0]).contiguous()
            h = self.prior_rows(x)
            out = K.gaussian_logp(h, x, ch, ch, vec)
            return x[:, :ch], out
        assert x.shape[1] == ch
        x = x.contiguous()
        h = self.prior_rows(x)
        if eps is None:
            eps = GaussianDiag.eps(x, eps_std)
        return K.split2d_sample(h, x, eps), logdet


class Squeeze2d(nn.Module):
    """network/module.py:539-612."""

    def __init__(self, factor=2):
        super().__init__()
        self.factor = factor

    @staticmethod
    def unsqueeze(x, factor=2):
        assert factor >= 1
        if factor == 1:
            return x
        _, nc, nh, nw = x.shape
        assert nc >= factor ** 2 and nc % factor ** 2 == 0
        if torch.is_grad_enabled() and x.requires_grad:
            from .autograd import SqueezeFunction
            return SqueezeFunction.apply(x, factor, True)
        return K.squeeze2d(x, factor, reverse=True)

    @staticmethod
    def squeeze(x, factor=2):
        assert factor >= 1
        if factor == 1:
            return x
        _, nc, nh, nw = x.shape
        assert nh % factor == 0 and nw % factor == 0
        if torch.is_grad_enabled() and x.requires_grad:
            from .autograd import SqueezeFunction
            return SqueezeFunction.apply(x, factor, False)
        return K.squeeze2d(x, factor, reverse=False)

    def forward(self, x, logdet=None, reverse=False):
        if not reverse:
            output = self.squeeze(x, self.factor)
        else:
            output = self.unsqueeze(x, self.factor)
        return output, logdet
