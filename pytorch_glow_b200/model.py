"""Host-side mirror of pytorch-glow's FlowStep / FlowModel / Glow (network/model.py).

`FlowModel` is what BASELINE.json's north_star calls "FlowNet":
``forward(x) -> (z, logdet)`` / ``reverse(z) -> x``.  Class names, constructor
arguments, call conventions, ``output_shapes`` and ``state_dict()`` keys follow the
reference; every flow layer executes as fused sm_100a kernels from libglowk.so.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import _C, config, module, ops, rows_path
from . import functional as K
from .module import _logdet_in, _logdet_out


class FlowStep(nn.Module):
    """One step of flow: ActNorm -> {1x1 conv | permutation} -> coupling (network/model.py:10-173)."""

    flow_permutation_list = ['invconv', 'reverse', 'shuffle']
    flow_coupling_list = ['additive', 'affine']

    def __init__(self, in_channels, hidden_channels, permutation='invconv', coupling='additive',
                 actnorm_scale=1., lu_decomposition=False):
        super().__init__()
        assert permutation in self.flow_permutation_list, 'Unsupported flow permutation: {}'.format(permutation)
        assert coupling in self.flow_coupling_list, 'Unsupported flow coupling: {}'.format(coupling)
        self.permutation = permutation
        self.coupling = coupling
        self.in_channels = in_channels
        self.conv_dtype = None      # per-step override of config.conv_dtype ("fp32" | "bf16" | None)

        self.actnorm = module.ActNorm(num_channels=in_channels, scale=actnorm_scale)
        if permutation == 'invconv':
            self.invconv = module.Invertible1x1Conv(num_channels=in_channels, lu_decomposition=lu_decomposition)
        elif permutation == 'reverse':
            self.reverse = module.Permutation2d(num_channels=in_channels, shuffle=False)
        else:
            self.shuffle = module.Permutation2d(num_channels=in_channels, shuffle=True)
        if coupling == 'additive':
            self.f = module.f(in_channels // 2, hidden_channels, in_channels // 2)
        else:
            self.f = module.f(in_channels // 2, hidden_channels, in_channels)

    # -- helpers ---------------------------------------------------------------------------------
    @property
    def perm_module(self):
        return None if self.permutation == 'invconv' else getattr(self, self.permutation)

    def _mix_args(self, device, reverse):
        """(weight, indices, log|det W| or None) for glowk_actnorm_mix in the given direction."""
        if self.permutation == 'invconv':
            w, winv, ld = self.invconv.prepared(need_inverse=reverse)
            return (winv if reverse else w), None, ld
        return None, self.perm_module.device_indices(device, reverse), None

    def _rows_route(self, x):
        """A stand-alone call (the reference's FlowModel driving this package's FlowStep) runs on the pixel-major
        kernels -- fused coupling net included -- behind a layout change in and out.  GLOWK_LAYER_ROWS=0: the
        per-layer NCHW kernels (csrc/flow_kernels.cu), also the path of channel counts the rows kernels do not take."""
        return os.environ.get("GLOWK_LAYER_ROWS", "1") != "0" and rows_path.step_supported(self, x)

    def _rows_logdet_in(self, logdet, n, device):
        vec, scalar_like = _logdet_in(logdet, n, device)
        if vec is not None and vec.shape[0] == 1 and n != 1:
            vec = vec.expand(n).contiguous()
        return vec, scalar_like

    def _rows_logdet_out(self, ld, scalar_like):
        if ld is None:
            return None
        keep_scalar = scalar_like and self.coupling != 'affine'      # (reference broadcasting semantics)
        return _logdet_out(ld[:1] if keep_scalar else ld, keep_scalar)

    # -- forward (model.py:82-117) ------------------------------------------------------------------
    def normal_flow(self, x, logdet=None):
        _C.check_cuda(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import flowstep_autograd
            return flowstep_autograd(self, x, logdet)
        x = x.contiguous()
        n, c, h, w = x.shape
        if self._rows_route(x):
            vec, scalar_like = self._rows_logdet_in(logdet, n, x.device)
            z, ld, _ = rows_path.step_forward_nchw(self, x, vec)
            return z, self._rows_logdet_out(ld, scalar_like)
        an = self.actnorm
        if an.needs_init:
            an.initialize_from_nchw(x)
        wmat, idx, logabsdet = self._mix_args(x.device, False)
        # (1) ActNorm + 1x1 mix / permutation, one pass over x
        z = K.actnorm_mix(x, wmat, idx, an.bias.detach().reshape(-1), an.logs.detach().reshape(-1),
                          an.logscale_factor, reverse=False)
        # (2) coupling network on z1 as three GEMMs (conv3 left in tap form)
        p3 = self.f.tap_rows(z, self.conv_dtype)
        # (3) tap gather-sum + Conv2dZeros scale + coupling, in place on z2, per-CTA logdet partials
        c3 = self.f[4]
        affine = self.coupling == 'affine'
        partials, _ = K.coupling(p3, c3.bias.detach(), c3.logs.detach().reshape(-1), z, affine, False,
                                 c3.logscale_factor)
        vec, scalar_like = _logdet_in(logdet, n, x.device)
        if vec is None:
            return z, None
        if affine and vec.shape[0] == 1 and n != 1:
            vec, scalar_like = vec.expand(n).contiguous(), False
        out = K.logdet_finish(vec, vec.shape[0], h * w, logs=an.logs.detach().reshape(-1), logabsdet=logabsdet,
                              partials=partials, logscale_factor=an.logscale_factor, sign=1.0)
        return z, _logdet_out(out, scalar_like and not affine)

    # -- reverse (model.py:119-154); like the reference it clobbers z2 of its input --------
    def reverse_flow(self, x, logdet=None):
        _C.check_cuda(x)
        x = x.contiguous()
        n, c, h, w = x.shape
        an = self.actnorm
        if self._rows_route(x) and not (an.needs_init and logdet is not None):
            vec, scalar_like = self._rows_logdet_in(logdet, n, x.device)
            out, ld = rows_path.step_reverse_nchw(self, x, vec)
            return out, self._rows_logdet_out(ld, scalar_like)
        c3 = self.f[4]
        affine = self.coupling == 'affine'
        p3 = self.f.tap_rows(x, self.conv_dtype)
        partials, _ = K.coupling(p3, c3.bias.detach(), c3.logs.detach().reshape(-1), x, affine, True,
                                 c3.logscale_factor)
        wmat, idx, logabsdet = self._mix_args(x.device, True)
        if an.needs_init:
            # first training-mode call in the reverse direction (network/module.py:143-146): un-mix, then let the
            # ActNorm initialise itself from that tensor (logs from the raw second moment, then the bias)
            y = K.actnorm_mix(x, wmat, idx, None, None, an.logscale_factor, reverse=True)
            an.initialize_from_nchw(y, reverse=True)
            out_x = K.actnorm(y, an.bias.detach().reshape(-1), an.logs.detach().reshape(-1), an.logscale_factor, True)
        else:
            out_x = K.actnorm_mix(x, wmat, idx, an.bias.detach().reshape(-1), an.logs.detach().reshape(-1),
                                  an.logscale_factor, reverse=True)
        vec, scalar_like = _logdet_in(logdet, n, x.device)
        if vec is None:
            return out_x, None
        if affine and vec.shape[0] == 1 and n != 1:
            vec, scalar_like = vec.expand(n).contiguous(), False
        out = K.logdet_finish(vec, vec.shape[0], h * w, logs=an.logs.detach().reshape(-1), logabsdet=logabsdet,
                              partials=partials, logscale_factor=an.logscale_factor, sign=-1.0)
        return out_x, _logdet_out(out, scalar_like and not affine)

    def forward(self, x, logdet=None, reverse=False):
        assert x.shape[1] % 2 == 0
        if not reverse:
            return self.normal_flow(x, logdet)
        return self.reverse_flow(x, logdet)


class FlowModel(nn.Module):
    """Multi-scale flow (network/model.py:176-314): [Squeeze2d, FlowStep x K, Split2d] x (L-1),
    Squeeze2d, FlowStep x K."""

    def __init__(self, in_shape, hidden_channels, K, L, permutation='invconv', coupling='additive',
                 actnorm_scale=1., lu_decomposition=False):
        super().__init__()
        self.K = K
        self.L = L
        assert len(in_shape) == 3
        assert in_shape[2] == 1 or in_shape[2] == 3
        nh, nw, nc = in_shape
        self.layers = nn.ModuleList()
        self.output_shapes = []
        for i in range(L):
            self.layers.append(module.Squeeze2d(factor=2))
            nc, nh, nw = nc * 4, nh // 2, nw // 2
            self.output_shapes.append([-1, nc, nh, nw])
            for _ in range(K):
                self.layers.append(FlowStep(in_channels=nc, hidden_channels=hidden_channels,
                                            permutation=permutation, coupling=coupling,
                                            actnorm_scale=actnorm_scale, lu_decomposition=lu_decomposition))
                self.output_shapes.append([-1, nc, nh, nw])
            if i < L - 1:
                self.layers.append(module.Split2d(num_channels=nc))
                nc = nc // 2
                self.output_shapes.append([-1, nc, nh, nw])

    def set_conv_dtype(self, mode):
        """'fp32' | 'bf16' | None for every FlowStep (see config.conv_dtype)."""
        for layer in self.layers:
            if isinstance(layer, (FlowStep, module.Split2d)):
                layer.conv_dtype = mode

    def adopt_permutations(self, other):
        """Copy the (unsaved, numpy-RNG) channel permutations from another FlowModel-like object."""
        for mine, theirs in zip(self.layers, other.layers):
            if isinstance(mine, FlowStep) and mine.permutation != 'invconv':
                mine.perm_module.set_indices(getattr(theirs, mine.permutation).indices)

    def prepare_invconvs(self, need_inverse):
        """Factorise every stale dense invconv weight now, one launch per channel count
        (glowk_invconv_prepare_batched), instead of one LU per FlowStep call (module.py:357,365)."""
        groups = {}
        for layer in self.layers:
            if isinstance(layer, FlowStep) and layer.permutation == 'invconv' and not layer.invconv.lu_decomposition:
                ic = layer.invconv
                if ic.weight.is_cuda and not ic._cache.fresh(("dense", need_inverse), ic.weight):
                    groups.setdefault((ic.num_channels, ic.weight.device), []).append(ic)
        for (c, _), mods in groups.items():
            if len(mods) == 1:
                continue                      # the per-module path handles it
            w = torch.stack([m.weight.detach() for m in mods])
            ld, winv = K.invconv_prepare_batched(w, need_inverse)
            for i, m in enumerate(mods):
                m._cache.put(("dense", need_inverse), m.weight, (ld[i:i + 1], None if winv is None else winv[i]))

    def encode(self, z, logdet=0.):
        if torch.is_grad_enabled() and getattr(self, "_is_replica", False):
            # torch.nn.DataParallel replicas hold their weights as plain attributes (replicate() empties
            # `_parameters`): the fused backward, which accumulates straight into param.grad, cannot reach the master
            # module through them.  Inference under DataParallel works; training is one process per GPU.
            raise RuntimeError(
                "pytorch_glow_b200.FlowModel cannot be TRAINED under torch.nn.DataParallel (network/trainer.py:117-120): "
                "its backward accumulates into param.grad of the module it runs on, and DataParallel replicas have no "
                "parameters.  Use one process per GPU with pytorch_glow_b200.train.FusedTrainStep (flat gradient "
                "arena + NCCL all-reduce), see INTEGRATION.md section 2; torch.no_grad() inference under DataParallel is fine.")
        use_autograd = torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in self.parameters()))
        if z.is_cuda:
            self.prepare_invconvs(need_inverse=use_autograd)      # the adjoint needs W^-T (logdet term)
        if use_autograd:
            from .autograd import flow_encode_autograd
            _C.check_cuda(z)
            return flow_encode_autograd(self, z, logdet)      # the whole encode is one autograd node
        vector_ld = logdet is None or (torch.is_tensor(logdet) and logdet.dim() == 1 and logdet.shape[0] == z.shape[0])
        if config.use_rows_path and vector_ld and rows_path.supported(self, z):
            ld = None if logdet is None else logdet.to(torch.float32).contiguous()
            return rows_path.encode(self, z, ld)              # pixel-major flow state (rows_path.py)
        tail = self.layers
        hd = rows_path.head(self, z) if (config.use_rows_path and vector_ld) else None
        if hd is not None:                                    # wide tail levels (C > 96): per-layer NCHW kernels
            ld = None if logdet is None else logdet.to(torch.float32).contiguous()
            z, logdet = rows_path.encode(hd[0], z, ld)
            tail = list(self.layers)[hd[1]:]
        for layer in tail:
            z, logdet = layer(z, logdet, reverse=False)
        return z, logdet

    def decode(self, z, eps_std=None, eps_list=None):
        """model.py:278-294.  `eps_list` optionally supplies the Split2d noise (deepest split first)."""
        k = 0
        if z.is_cuda:
            self.prepare_invconvs(need_inverse=True)
        if config.use_rows_path and rows_path.supported(self, z):
            with torch.no_grad():
                return rows_path.decode(self, z, eps_std, eps_list)
        tail = list(self.layers)
        hd = None
        if config.use_rows_path and z.is_cuda:
            kh = rows_path._prefix_len(self)
            if 0 < kh < len(tail) and isinstance(tail[kh - 1], module.Split2d):
                hd, tail = kh, tail[kh:]
        for layer in reversed(tail):
            if isinstance(layer, module.Split2d):
                e = None if eps_list is None else eps_list[k]
                k += 1
                z, logdet = layer(z, logdet=0., reverse=True, eps_std=eps_std, eps=e)
            else:
                z, logdet = layer(z, logdet=0., reverse=True)
        if hd is not None:                                    # the leading levels on the pixel-major kernels
            view = rows_path.head(self, z)
            if view is not None:
                with torch.no_grad():
                    return rows_path.decode(view[0], z, eps_std, None if eps_list is None else eps_list[k:])
            for layer in reversed(list(self.layers)[:hd]):    # (tensor not eligible: finish on the NCHW kernels)
                if isinstance(layer, module.Split2d):
                    e = None if eps_list is None else eps_list[k]
                    k += 1
                    z, logdet = layer(z, logdet=0., reverse=True, eps_std=eps_std, eps=e)
                else:
                    z, logdet = layer(z, logdet=0., reverse=True)
        return z

    def forward(self, z, logdet=0., eps_std=None, reverse=False):
        if not reverse:
            return self.encode(z, logdet)
        return self.decode(z, eps_std)


FlowNet = FlowModel   # the name BASELINE.json's north_star uses


class Glow(nn.Module):
    """network/model.py:317-550: dequantisation, objective, top prior, bits/dim; thin caller of FlowModel."""

    bce_criterion = nn.BCEWithLogitsLoss()
    ce_criterion = nn.CrossEntropyLoss()

    def __init__(self, hps):
        super().__init__()
        self.hps = hps
        self.flow = FlowModel(in_shape=hps.model.image_shape, hidden_channels=hps.model.hidden_channels,
                              K=hps.model.K, L=hps.model.L, permutation=hps.ablation.flow_permutation,
                              coupling=hps.ablation.flow_coupling, actnorm_scale=hps.model.actnorm_scale,
                              lu_decomposition=hps.ablation.lu_decomposition)
        if hps.ablation.learn_top:
            nc = self.flow.output_shapes[-1][1]
            self.learn_top = module.Conv2dZeros(in_channels=2 * nc, out_channels=2 * nc)
        if hps.ablation.y_condition:
            nc = self.flow.output_shapes[-1][1]
            self.y_emb = module.LinearZeros(hps.dataset.num_classes, nc * 2)
            self.classifier = module.LinearZeros(nc, hps.dataset.num_classes)
        num_device = max(1, len(_graph_devices(hps)))
        assert hps.optim.num_batch_train % num_device == 0
        self.register_parameter('h_top', nn.Parameter(torch.zeros(
            [hps.optim.num_batch_train // num_device, self.flow.output_shapes[-1][1] * 2,
             self.flow.output_shapes[-1][2], self.flow.output_shapes[-1][3]])))

    @property
    def batch_h_top(self):
        return self.h_top.shape[0]

    @property
    def _plain_top_prior(self):
        return not (self.hps.ablation.learn_top or self.hps.ablation.y_condition)

    def prior(self, y_onehot=None):
        nc = self.h_top.shape[1]
        h = self.h_top.detach().clone()
        if self.hps.ablation.learn_top:
            h = self.learn_top(h)
        if self.hps.ablation.y_condition:
            assert y_onehot is not None
            h = h + self.y_emb(y_onehot).view(-1, nc, 1, 1)
        return ops.split_channel(h, 'simple')

    def _fused_head(self, x):
        """The loss head runs as one node (autograd.GlowNLLFunction) when the top prior is the plain N(0, I) and the
        whole flow is on the pixel-major kernels (GLOWK_FUSED_HEAD=0: the layer-by-layer composition below)."""
        from . import rows_path
        return (self._plain_top_prior and config.use_rows_path and os.environ.get("GLOWK_FUSED_HEAD", "1") != "0"
                and not getattr(self.flow, "_is_replica", False) and rows_path.supported(self.flow, x))

    def normal_flow(self, x, y_onehot, noise=None):
        """model.py:409-452.  `noise` (U(0, 1/n_bins), same shape as x) may be supplied for parity runs."""
        n_bins = 2 ** self.hps.model.n_bits_x
        if noise is None:
            noise = torch.nn.init.uniform_(torch.empty(*x.shape, device=x.device), 0, 1. / n_bins)
        if self._fused_head(x):
            train = torch.is_grad_enabled() and any(p.requires_grad for p in self.flow.parameters())
            self.flow.prepare_invconvs(need_inverse=train)   # one batched LU kernel (the adjoint needs W^-T)
            if train:
                from .autograd import GlowNLLFunction
                params = [p for p in self.flow.parameters() if p.requires_grad]
                z, nll, loss = GlowNLLFunction.apply(self.flow, n_bins, x, noise, *params)
            else:
                from . import rows_path
                d_x = x[0].numel()
                with torch.no_grad():
                    z, ld = rows_path.encode(self.flow, x.contiguous(), None, add=noise.contiguous(), want_ld=True)
                    nll, loss = K.nll_head(z, ld, -float(np.log(n_bins)) * d_x, float(np.log(2.)) * d_x)
            nll._glowk_mean = loss                           # picked up by generative_loss(nll)
            return z, nll, None
        z = x + noise
        logdet_factor = x.shape[1] * ops.count_pixels(x)
        objective = torch.full((x.shape[0],), float(-np.log(n_bins)) * logdet_factor, device=x.device,
                               dtype=torch.float32)
        z, objective = self.flow(z, logdet=objective, reverse=False)
        if self._plain_top_prior and torch.is_grad_enabled() and z.requires_grad:
            from .autograd import TopPriorFunction
            objective = TopPriorFunction.apply(z, objective)
        elif self._plain_top_prior:
            objective = K.gaussian_logp(None, z.contiguous(), 0, z.shape[1], objective)   # N(0,1) top prior
        else:
            mean, logs = self.prior(y_onehot)
            objective = objective + module.GaussianDiag.logp(mean, logs, z)
        if self.hps.ablation.y_condition and self.hps.model.weight_y > 0:
            y_logits = self.classifier(ops.reduce_mean(z, dim=[2, 3]))
        else:
            y_logits = None
        nll = (-objective) / float(np.log(2.) * logdet_factor)
        return z, nll, y_logits

    def reverse_flow(self, z, y_onehot, eps_std=None):
        with torch.no_grad():
            mean, logs = self.prior(y_onehot)
            if z is None:
                z = module.GaussianDiag.sample(mean, logs, eps_std)
            return self.flow(z, eps_std=eps_std, reverse=True)

    def forward(self, x=None, y_onehot=None, z=None, eps_std=None, reverse=False):
        if not reverse:
            return self.normal_flow(x, y_onehot)
        return self.reverse_flow(z, y_onehot, eps_std)

    @staticmethod
    def generative_loss(nll):
        """model.py:496-498.  For the nll tensor normal_flow returned from its fused loss head the batch mean was
        computed by the same kernel."""
        fused = getattr(nll, "_glowk_mean", None)
        return fused if fused is not None else torch.mean(nll)

    @staticmethod
    def single_class_loss(y_logits, y):
        if y_logits is None:
            return 0
        return Glow.ce_criterion(y_logits, y.long())

    @staticmethod
    def multi_class_loss(y_logits, y_onehot):
        if y_logits is None:
            return 0
        return Glow.bce_criterion(y_logits, y_onehot.float())

    def set_actnorm_inited(self, inited=True):
        for name, m in self.named_modules():
            if m.__class__.__name__.find("ActNorm") >= 0:
                m.bias_inited = inited
                m.logs_inited = inited


def _graph_devices(hps):
    """Number of model replicas the profile asks for (reference: misc/util.py:34-75 get_devices).
    One process per GPU here, so a multi-device list still means "global batch / len(list)" per rank."""
    devs = list(getattr(getattr(hps, 'device', None), 'graph', None) or ['cuda:0'])
    if any('cpu' in str(d) for d in devs):
        return ['cpu']
    return devs
