"""FlowModel.encode / decode / backward on a pixel-major ("rows") flow state.

Between the NCHW tensors of the reference API (network/model.py:263-294) the flow state lives as
x[p][c] (p = (n*H + y)*W + x, fp32) -- the layout of every coupling-network GEMM operand -- so each
flow kernel (csrc/flow_rows_kernels.cu) reads and writes whole contiguous pixels.  Squeeze2d doubles
as the layout change at the model's entry and exit, Split2d's halves are addressed in place through
row pitches, and the conv weight packing / weight-gradient unpacking of ALL coupling networks is one
launch each (`PackPlan`, `GradPlan`).

The per-layer modules (module.py / model.py) keep the reference's NCHW call conventions and kernels;
this file is what FlowModel uses when the whole model fits the rows kernels (`supported`).
"""
import os

import numpy as np
import torch

from . import _C, config, module
from . import functional as K
from .functional import NCHW, ROWS, round_up

_JOB = np.dtype([("w", "<u8"), ("packed", "<u8"), ("O", "<i4"), ("I", "<i4"), ("ks", "<i4"), ("layout", "<i4"),
                 ("rows", "<i4"), ("ld", "<i4"), ("block0", "<i8")])
assert _JOB.itemsize == 48
# struct FinishJob of flow_rows_kernels.cu (glowk_conv_actnorm_finish_batched)
_FJOB = np.dtype([("w", "<u8"), ("dw", "<u8"), ("bias", "<u8"), ("db", "<u8"), ("dbias", "<u8"), ("dlogs", "<u8"),
                  ("N", "<i4"), ("K", "<i4"), ("ldw", "<i4"), ("lddw", "<i4"), ("f", "<f4"), ("db_stride", "<i4")])
assert _FJOB.itemsize == 72


def _gbuf(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad.view(-1)


def _steps_and_splits(flow):
    from .model import FlowStep
    steps = [l for l in flow.layers if isinstance(l, FlowStep)]
    splits = [l for l in flow.layers if isinstance(l, module.Split2d)]
    return steps, splits


def _step_ok(step):
    """Channel counts the rows kernels take: <= 96 for every permutation type (ActNorm + mix kernels with the
    weight in shared memory); up to 384 for the 1x1 conv, whose ActNorm + mix then run as fp32 GEMMs (_mix_wide_*)."""
    c = step.in_channels
    if c % 4:
        return False
    return c <= K.rows_max_channels() or (c <= K.rows_max_channels_wide() and step.permutation == 'invconv')


def _split_ok(sp):
    return sp.num_channels % 4 == 0 and sp.num_channels <= K.rows_max_channels_wide()


def supported(flow, z):
    """True iff FlowModel `flow` can run on the rows kernels for input z (else: per-layer NCHW path)."""
    from .model import FlowStep
    if not (torch.is_tensor(z) and z.is_cuda and z.dtype == torch.float32 and z.dim() == 4):
        return False
    if not (0 < z.shape[0] <= 65535) or len(flow.layers) == 0:
        return False
    layers = list(flow.layers)
    if not isinstance(layers[0], module.Squeeze2d) or isinstance(layers[-1], module.Split2d):
        return False
    for i, layer in enumerate(layers):
        if isinstance(layer, FlowStep):
            if not _step_ok(layer):
                return False
        elif isinstance(layer, module.Split2d):
            if not _split_ok(layer) or not isinstance(layers[i + 1], module.Squeeze2d):
                return False
        elif not isinstance(layer, module.Squeeze2d):
            return False
    return True


class _FlowView:
    """The leading levels of a FlowModel as a flow of their own (same layer objects; own plan / workspace caches)."""

    def __init__(self, layers):
        self.layers = layers


def _prefix_len(flow):
    """Number of leading layers that form whole levels (Squeeze2d, FlowStep x K [, Split2d]) the rows kernels run:
    channel counts a multiple of 4 and <= glowk_rows_max_channels()."""
    from .model import FlowStep
    layers = list(flow.layers)
    i = good = 0
    while i < len(layers) and isinstance(layers[i], module.Squeeze2d):
        j = i + 1
        while j < len(layers) and isinstance(layers[j], FlowStep) and _step_ok(layers[j]):
            j += 1
        if j == i + 1 or (j < len(layers) and isinstance(layers[j], FlowStep)):
            break                                   # no step, or a step the kernels cannot run
        if j < len(layers) and isinstance(layers[j], module.Split2d):
            if not _split_ok(layers[j]):
                break
            j += 1
        elif j < len(layers):
            break
        good = i = j
    return good


def head(flow, z):
    """(view, k) if only the first k layers (whole levels, ending in a Split2d) can run on the rows kernels --
    the CelebA-HQ shape family, whose levels 5 and 6 have 192 / 384 channels -- else None.  FlowModel.encode /
    decode and the autograd node then run the head on the pixel-major path and the tail on the per-layer NCHW
    kernels; the hand-over tensor is the NCHW z1 of the head's last Split2d."""
    if not (torch.is_tensor(z) and z.is_cuda and z.dtype == torch.float32 and z.dim() == 4 and 0 < z.shape[0] <= 65535):
        return None
    k = _prefix_len(flow)
    if k <= 0 or k >= len(flow.layers) or not isinstance(flow.layers[k - 1], module.Split2d):
        return None
    # nn.DataParallel replicas share this dict (shallow __dict__ copies) but own their layers: key on the layer object
    cache = flow.__dict__.setdefault("_rows_head", {})
    key = (k, id(flow.layers[0]))
    view = cache.get(key)
    if view is None or view.layers[0] is not flow.layers[0]:
        if len(cache) >= 16:
            cache.clear()
        view = cache[key] = _FlowView(list(flow.layers)[:k])
    return view, k


# ------------------------------------------------------------------ per-model device workspaces
class _Workspace:
    """Deterministic-reduction scratch of glowk_rows_coupling (tickets stay zero between launches)."""

    def __init__(self, device, n, nblk):
        self.n, self.nblk = n, nblk
        self.tickets = torch.zeros(n, dtype=torch.int32, device=device)
        self.partials = torch.empty(n * nblk, dtype=torch.float32, device=device)


def _workspace(flow, device, n, nblk):
    cache = flow.__dict__.setdefault("_rows_ws", {})
    key = (str(device), n)
    ws = cache.get(key)
    if ws is None or ws.nblk < nblk:
        ws = _Workspace(device, n, max(nblk, 32))
        cache[key] = ws
    return ws


class PackPlan:
    """GEMM-layout copies of every conv weight of a FlowModel in one arena, refreshed by ONE kernel."""

    def __init__(self, flow, backward, device):
        steps, splits = _steps_and_splits(flow)
        entries = []          # (cache, key, param, layout, rows, ld, dt)
        for st in steps:
            dt = st.f.dtype(st.conv_dtype)
            for key, prm, layout, rows, ld in st.f.pack_specs(dt, backward):
                entries.append((st.f._packs, key, prm, layout, rows, ld, dt))
        for sp in splits:
            dt = config.resolve_conv_dtype(64, sp.conv_dtype)
            for conv, key, prm, layout, rows, ld in sp.pack_specs(dt, backward):
                entries.append((conv._packs, key, prm, layout, rows, ld, dt))
        self.entries = entries
        self.sig = tuple(e[2].data_ptr() for e in entries)
        self.groups = {}
        for dt in sorted(set(e[6] for e in entries)):
            mine = [e for e in entries if e[6] == dt]
            numel = sum(e[4] * e[5] for e in mine)
            arena = torch.zeros(numel, device=device, dtype=K.TORCH_DTYPE[dt])     # padding stays zero for good
            jobs = np.zeros(len(mine), dtype=_JOB)
            views, off, blk = [], 0, 0
            esz = arena.element_size()
            for i, (cache, key, prm, layout, rows, ld, _) in enumerate(mine):
                v = arena[off:off + rows * ld].view(rows, ld)
                views.append(v)
                jobs[i] = (prm.data_ptr(), arena.data_ptr() + off * esz, prm.shape[0], prm.shape[1], prm.shape[2],
                           layout, rows, ld, blk)
                off += rows * ld
                blk += ((prm.shape[0] + 31) // 32) * ((prm.shape[1] + 31) // 32)     # one CTA per 32x32 channel tile
            jobs_dev = torch.from_numpy(jobs.view(np.uint8).copy()).to(device)
            self.groups[dt] = (mine, views, arena, jobs_dev, len(mine), blk)

    def valid(self, flow):
        return self.sig == tuple(e[2].data_ptr() for e in self.entries)

    def fresh(self):
        return all(cache.fresh((key, dt), prm) for cache, key, prm, _, _, _, dt in self.entries)

    def run(self):
        for dt, (mine, views, arena, jobs_dev, njobs, blocks) in self.groups.items():
            K.pack_conv_weights_batched(jobs_dev, njobs, blocks, dt)
            for (cache, key, prm, _, _, _, _), v in zip(mine, views):
                cache.put((key, dt), prm, v)


def prepare_packs(flow, backward, device):
    plans = flow.__dict__.setdefault("_pack_plans", {})
    key = (bool(backward), str(device))
    plan = plans.get(key)
    if plan is None or not plan.valid(flow):
        plan = PackPlan(flow, backward, device)
        plans[key] = plan
    if not plan.fresh():
        plan.run()


class GradPlan:
    """fp32 scratch for the packed-layout weight gradients of the coupling-net convs / the Split2d prior conv of
    every layer (zeroed once per backward) and ONE kernel scattering them into the [O][I][k][k] .grad tensors.
    The scratch also holds this pass's ActNorm bias gradients of conv1 / conv2 (`db`): on the bf16 path the dgrad
    epilogue no longer reduces sum g*y, and `finish` recovers dlogs from W, this pass's dW and db in one batched
    kernel (glowk_conv_actnorm_finish_batched)."""

    def __init__(self, flow, device):
        steps, splits = _steps_and_splits(flow)
        # level of every layer = number of Squeeze2d layers up to it, minus one (levels finish their backward pass
        # top-down: `finish_level` lets the gradients of a finished level leave for the all-reduce while the lower
        # levels are still being differentiated, train.FusedTrainStep)
        self.level_of, lvl = {}, -1
        for layer in flow.layers:
            if isinstance(layer, module.Squeeze2d):
                lvl += 1
            self.level_of[id(layer)] = max(lvl, 0)
        self.n_levels = max(lvl, 0) + 1
        items = []            # (owner, tag, param, layout, rows, ld)
        for st in steps:
            net = st.f
            kh = round_up(net.hidden_channels, 64)
            k3p = round_up(9 * net.out_channels, 64)
            items.append((st, "w3", net[4].weight, 1, k3p, kh))
            items.append((st, "w1", net[0].weight, 0, kh, net.k1p))
            items.append((st, "w2", net[2].weight, 0, kh, kh))
        for sp in splits:
            c = sp.num_channels
            items.append((sp, "w0", sp.conv2d_zeros.weight, 0, round_up(c, 64), round_up(9 * (c // 2), 64)))
        self.items = items
        self.sig = tuple(_gbuf(it[2]).data_ptr() for it in items)
        numel = sum(it[4] * it[5] for it in items)
        ndb = sum(2 * round_up(st.f.hidden_channels, 64) for st in steps)
        self.arena = torch.zeros(numel + ndb, device=device, dtype=torch.float32)
        self.views = {}
        off = numel
        for st in steps:
            kh = round_up(st.f.hidden_channels, 64)
            self.views[(id(st), "db1")] = self.arena[off:off + kh]
            self.views[(id(st), "db2")] = self.arena[off + kh:off + 2 * kh]
            off += 2 * kh
        self._fin = [[] for _ in range(self.n_levels)]
        self._fin_sig = [None] * self.n_levels
        self._fin_dev = [None] * self.n_levels
        self._done = [False] * self.n_levels
        # weight-gradient GEMMs run on a side stream (they are off the backward pass's critical path: nothing reads
        # dW before finish_level): their tails and launch gaps fill with the next step's kernels
        self.side = torch.cuda.Stream(device=device) if (device.type == "cuda" and os.environ.get("GLOWK_WGRAD_STREAM", "1") != "0") else None
        self._keep, self._forked = [], False
        self.side_max_pixels = int(os.environ.get("GLOWK_WGRAD_STREAM_MAXP", "200000"))
        per_level = [[] for _ in range(self.n_levels)]
        off = 0
        for owner, tag, prm, layout, rows, ld in items:
            self.views[(id(owner), tag)] = self.arena[off:off + rows * ld].view(rows, ld)
            per_level[self.level_of[id(owner)]].append((_gbuf(prm).data_ptr(), self.arena.data_ptr() + off * 4, prm.shape[0],
                                                        prm.shape[1], prm.shape[2], layout, rows, ld, prm.numel()))
            off += rows * ld
        self.level_jobs = []                 # per level: (device job array, njobs, blocks) of the unpack kernel
        for lv in range(self.n_levels):
            jobs = np.zeros(len(per_level[lv]), dtype=_JOB)
            blk = 0
            for i, (g, a, o, ii, ks, layout, rows, ld, numel_p) in enumerate(per_level[lv]):
                jobs[i] = (g, a, o, ii, ks, layout, rows, ld, blk)
                blk += (numel_p + 255) // 256
            dev_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(device) if len(jobs) else None
            self.level_jobs.append((dev_jobs, len(jobs), blk))

    def valid(self):
        return self.sig == tuple(_gbuf(it[2]).data_ptr() for it in self.items)

    def begin(self):
        self.arena.zero_()
        self._fin = [[] for _ in range(self.n_levels)]
        self._done = [False] * self.n_levels
        self._keep, self._forked = [], False

    def wgrads(self, jobs, keep):
        """Launch the weight-gradient GEMMs `jobs` = [(a, b, mo, no, dw)] of one layer.  With a side stream: ordered
        after everything launched so far on the current stream, concurrently with what follows; `keep` holds the
        operands' storage until the streams are joined (the caching allocator must not recycle them earlier)."""
        # Measured on B200 (gpurun_out/r2 A/B): +5.8 % at 64 img/GPU (the small kernels of a step leave gaps and
        # tails that the weight gradients fill), -1.1 % at 512 img/GPU (two persistent 148-CTA kernels at a time only
        # compete for the SMs): used for layers with at most `side_max_pixels` pixels.
        if self.side is None or jobs[0][0].shape[0] > self.side_max_pixels:
            for a, b, mo, no, dw in jobs:
                K.gemm_wgrad(a, b, mo, no, dw)
            return
        main = torch.cuda.current_stream()
        if len(self._keep) >= 4:                # bound the memory held back: let the side stream catch up
            self.join()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            for a, b, mo, no, dw in jobs:
                K.gemm_wgrad(a, b, mo, no, dw)
        self._keep.append(keep)
        self._forked = True

    def join(self):
        if self.side is not None and self._forked:
            torch.cuda.current_stream().wait_stream(self.side)
            self._keep, self._forked = [], False

    def view(self, owner, tag):
        return self.views[(id(owner), tag)]

    def defer_dlogs(self, owner, w_packed, dw, actnorm, db, n, k, ones_col=-1):
        """Register one conv + ActNorm layer of `owner` for the batched dbias / dlogs finish (bf16 path).
        ones_col >= 0: this pass's bias gradient is column `ones_col` of dw (the im2col operand carried a ones column)."""
        if ones_col >= 0:
            db_ptr, stride = dw.data_ptr() + 4 * ones_col, dw.shape[1]
        else:
            db_ptr, stride = db.data_ptr(), 1
        self._fin[self.level_of[id(owner)]].append(
            (w_packed.data_ptr(), dw.data_ptr(), actnorm.bias.data_ptr(), db_ptr, _gbuf(actnorm.bias).data_ptr(),
             _gbuf(actnorm.logs).data_ptr(), n, k, w_packed.shape[1], dw.shape[1], float(actnorm.logscale_factor), stride))

    def finish_level(self, lv):
        """Every layer of level `lv` has been differentiated: apply its deferred ActNorm gradients and scatter its
        packed weight gradients into the .grad tensors (two launches).  After this the level's gradients are final."""
        if self._done[lv]:
            return
        self._done[lv] = True
        self.join()                              # every weight gradient of the level is complete
        fin = self._fin[lv]
        if fin:
            sig = tuple(fin)
            if sig != self._fin_sig[lv]:       # pointers are stable across steps: built once, before graph capture
                jobs = np.array(fin, dtype=_FJOB)
                self._fin_dev[lv] = torch.from_numpy(jobs.view(np.uint8).copy()).to(self.arena.device)
                self._fin_sig[lv] = sig
            K.conv_actnorm_finish_batched(self._fin_dev[lv], len(fin), max(j[6] for j in fin))
        jobs_dev, njobs, blocks = self.level_jobs[lv]
        if njobs:
            K.unpack_weight_grads_batched(jobs_dev, njobs, blocks)

    def finish(self):
        for lv in range(self.n_levels - 1, -1, -1):
            self.finish_level(lv)


def grad_plan(flow, device):
    plans = flow.__dict__.setdefault("_grad_plans", {})
    plan = plans.get(str(device))
    if plan is None or not plan.valid():
        plan = GradPlan(flow, device)
        plans[str(device)] = plan
    return plan


# ------------------------------------------------------------------ FlowStep
def _mix_params(step, device, reverse, need_inverse):
    if step.permutation == 'invconv':
        wmat, winv, logabsdet = step.invconv.prepared(need_inverse=need_inverse or reverse)
        return (winv if reverse else wmat), None, logabsdet, wmat, winv
    return None, step.perm_module.device_indices(device, reverse), None, None, None


def _is_wide(c):
    return c > K.rows_max_channels()


def _mix_wide_forward(x, wm, bias, logs, f, reverse):
    """ActNorm + 1x1 conv on rows for C > 96 (model.py:94-99 / 148-152): the C x C weight does not fit the shared
    memory of rows_mix_kernel, and at these levels (8x8 / 4x4 pixels) the mix is a small fp32 GEMM anyway."""
    p, c = x.shape
    if not reverse:
        a = K.actnorm(x.view(p, c, 1, 1), bias, logs, f, reverse=False).view(p, c)
        return K.gemm(a, wm, c, c, _C.EPI_STORE, out_dtype=_C.F32)
    t = K.gemm(x, wm, c, c, _C.EPI_STORE, out_dtype=_C.F32)                  # wm = W^-1
    return K.actnorm(t.view(p, c, 1, 1), bias, logs, f, reverse=True).view(p, c)


def _mix_wide_backward(x, dz, n, h, w, da1, cin, wmat, winv, an, gw, dld):
    """Adjoint of _mix_wide_forward (+ conv1 dgrad tap gather, + logdet parameter gradients)."""
    p, c = x.shape
    K.rows_tapsum(da1, dz, 0, cin, n, h, w, flip=True, accumulate=True)       # dz[:, :cin] += conv1 dgrad
    bias, logs = an.bias.detach().reshape(-1), an.logs.detach().reshape(-1)
    a = K.actnorm(x.view(p, c, 1, 1), bias, logs, an.logscale_factor, reverse=False).view(p, c)
    K.gemm_wgrad(dz, a, c, c, gw.view(c, c))                                  # dW += dz^T a
    da = K.gemm(dz, wmat.t().contiguous(), c, c, _C.EPI_STORE, out_dtype=_C.F32)   # da = dz W
    dx = K.rows_actnorm_bwd(da, x, bias, logs, _gbuf(an.logs), _gbuf(an.bias), an.logscale_factor)
    if dld is not None:
        K.logdet_param_grad(dld, h * w, _gbuf(an.logs), winv, gw, an.logscale_factor)
    return dx


def _ones_col(net, dt):
    """First zero-padding column of the conv1 im2col operand (or -1): written as 1.0 on the bf16 training path so
    that conv1's weight-gradient GEMM also yields the bias gradient of its ActNorm (glowk_im2col_rows_ones)."""
    k = 9 * net.in_channels
    return k if (dt == _C.BF16 and k < net.k1p and net.in_channels % 2 == 0) else -1


def _step_forward(step, x, n, c, h, w, ld, ws, save, want_ld=False):
    """FlowStep.normal_flow (network/model.py:82-117) on rows x [P][C].  want_ld with ld None: the logdet starts at 0."""
    an = step.actnorm
    if an.needs_init:
        an.initialize_from_rows(x)
    wm, idx, logabsdet, wmat, winv = _mix_params(step, x.device, False, save)
    b, l = an.bias.detach().reshape(-1), an.logs.detach().reshape(-1)
    if _is_wide(c):
        z = _mix_wide_forward(x, wm, b, l, an.logscale_factor, False)
    else:
        z = K.rows_actnorm_mix(x, wm, idx, b, l, an.logscale_factor, reverse=False)
    net = step.f
    dt = net.dtype(step.conv_dtype)
    # activation recompute (config.recompute_activations): nothing of the coupling net is kept, see _recompute
    keep = save and not (config.recompute_activations and not an.needs_init
                         and not (net[0].actnorm.needs_init or net[2].actnorm.needs_init))
    sv = {} if keep else None
    p3 = net.tap_rows_from_rows(z, n, h, w, dt, sv, ones_col=_ones_col(net, dt) if keep else -1)
    c3 = net[4]
    affine = step.coupling == 'affine'
    ld_out, hrows = K.rows_coupling(p3, c3.bias.detach(), c3.logs.detach().reshape(-1), z, n, h, w, affine, False,
                                    c3.logscale_factor, save_h=keep, ld_in=ld, want_ld=want_ld or ld is not None, an_logs=l,
                                    an_f=an.logscale_factor, logabsdet=logabsdet, sign=1.0, partials=ws.partials,
                                    tickets=ws.tickets)
    ctx = None
    if keep:
        ctx = dict(x=x, y=z, hrows=hrows, a1=sv["a1"], h1=sv["h1"], h2=sv["h2"], masks=sv.get("masks"), wmat=wmat,
                   winv=winv, idx=idx)
    elif save:
        ctx = dict(x=x, y=z, recompute=dt, wmat=wmat, winv=winv, idx=idx)
    return z, ld_out, ctx


def _recompute(step, ctx, n, h, w):
    """Rebuild what _step_forward did not keep (config.recompute_activations) from the step's OUTPUT rows: the coupling
    leaves z1 -- the coupling net's input -- untouched, so the fused forward on y[:, :C/2] reproduces a1 / h1 / h2 / the
    ReLU masks bit for bit, and the tap sum of its P3 rows + the Conv2dZeros scale reproduce the coupling's (shift, scale)
    pre-activations with the coupling kernel's own arithmetic."""
    net = step.f
    dt = ctx.pop("recompute")
    sv = {}
    p3 = net.tap_rows_from_rows(ctx["y"], n, h, w, dt, sv, ones_col=_ones_col(net, dt))
    c3 = net[4]
    cout = net.out_channels
    u = torch.empty(n * h * w, cout, device=p3.device, dtype=torch.float32)
    K.rows_tapsum(p3, u, 0, cout, n, h, w)
    hrows = K.actnorm(u.view(n * h * w, cout, 1, 1), c3.bias.detach().reshape(-1), c3.logs.detach().reshape(-1),
                      c3.logscale_factor, False, out=u.view(n * h * w, cout, 1, 1)).view(n * h * w, cout)
    ctx.update(hrows=hrows, a1=sv["a1"], h1=sv["h1"], h2=sv["h2"], masks=sv.get("masks"))


def _step_reverse(step, x, n, c, h, w, ws, ld=None, want_ld=False):
    """FlowStep.reverse_flow (network/model.py:119-154) on rows; x is clobbered.  Returns the new rows, or
    (rows, ld_out) when the reverse direction's logdet is asked for (ld given or want_ld)."""
    an = step.actnorm
    net = step.f
    dt = net.dtype(step.conv_dtype)
    p3 = net.tap_rows_from_rows(x, n, h, w, dt)
    c3 = net[4]
    wm, idx, logabsdet, _, _ = _mix_params(step, x.device, True, False)
    want_ld = want_ld or ld is not None
    affine = step.coupling == 'affine'
    b3, l3 = c3.bias.detach(), c3.logs.detach().reshape(-1)

    def coupling_inverse():
        # in place on x; with the logdet: -(HW * (sum f*logs + log|det W|) + sum log scale)   (model.py:131-152)
        ld_out, _ = K.rows_coupling(p3, b3, l3, x, n, h, w, affine, True, c3.logscale_factor, ld_in=ld, want_ld=want_ld,
                                    an_logs=an.logs.detach().reshape(-1) if want_ld else None, an_f=an.logscale_factor,
                                    logabsdet=logabsdet if want_ld else None, sign=-1.0,
                                    partials=ws.partials if want_ld else None, tickets=ws.tickets if want_ld else None)
        return ld_out

    if an.needs_init:
        # first training-mode call arrives in the reverse direction (network/module.py:143-146 with 44-45, 62-63):
        # the ActNorm initialises from the un-mixed tensor -- logs from its raw second moment, then the bias
        assert not want_ld, "the logdet of an uninitialised ActNorm is not defined before its init"
        coupling_inverse()
        if _is_wide(c):
            y = K.gemm(x, wm, c, c, _C.EPI_STORE, out_dtype=_C.F32)
        else:
            y = K.rows_actnorm_mix(x, wm, idx, None, None, an.logscale_factor, reverse=True)
        b, l = K.actnorm_init_rows(y, c, an.scale, an.logscale_factor, an.batch_variance, reverse=True)
        an._store_init(b, l)
        return K.actnorm(y.view(n * h * w, c, 1, 1), b, l, an.logscale_factor, reverse=True).view(n * h * w, c)
    ab, al = an.bias.detach().reshape(-1), an.logs.detach().reshape(-1)
    if not want_ld and not _is_wide(c) and os.environ.get("GLOWK_REV_FUSED", "1") != "0":
        # inverse coupling + W^-1 mix + ActNorm^-1 in ONE launch (the coupled rows never leave shared memory)
        return K.rows_coupling_rev_mix(p3, b3, l3, x, n, h, w, affine, c3.logscale_factor, wm, idx, ab, al,
                                       an.logscale_factor)
    ld_out = coupling_inverse()
    if _is_wide(c):
        out = _mix_wide_forward(x, wm, ab, al, an.logscale_factor, True)
    else:
        out = K.rows_actnorm_mix(x, wm, idx, ab, al, an.logscale_factor, reverse=True)
    return (out, ld_out) if want_ld else out


def _step_backward(step, ctx, dy, dld, n, c, h, w, plan):
    """Adjoint of _step_forward; accumulates every parameter gradient of the step, returns dx rows."""
    net = step.f
    c1, c2, c3 = net[0], net[2], net[4]
    an, an1, an2 = step.actnorm, c1.actnorm, c2.actnorm
    hid = net.hidden_channels
    kh = round_up(hid, 64)
    affine = step.coupling == 'affine'
    cout = net.out_channels
    if "recompute" in ctx:
        _recompute(step, ctx, n, h, w)
    h1, h2, a1 = ctx["h1"], ctx["h2"], ctx["a1"]
    dt = _C.BF16 if h1.dtype == torch.bfloat16 else _C.F32
    dev = dy.device
    # (1) coupling + Conv2dZeros scale
    dz, du = K.rows_coupling_bwd(ctx["y"], ctx["hrows"], dy, dld, c3.logs.detach().reshape(-1), n, h * w, affine,
                                 _gbuf(c3.logs), _gbuf(c3.bias), c3.logscale_factor)
    # (2) conv3 (tap form): dP3 = flipped im2col of du
    k3p = round_up(9 * cout, 64)
    k1p = net.k1p
    wide = _is_wide(c)
    fused = dt == _C.BF16 and not wide and net.fused(True)
    implicit = fused and cout % 2 == 0 and os.environ.get("GLOWK_CNET_IMPLICIT", "1") != "0"
    if not implicit:
        d3col = K.im2col_rows(du, n, h, w, 0, cout, 3, dt, k3p, flip=True)
    # bf16 (tcgen05) path: the ReLU-backward epilogue only reduces this pass's bias gradient into scratch; dlogs of
    # the two hidden ActNorms come from W, dW and db in GradPlan.finish (see glowk_conv_actnorm_finish_batched)
    defer = dt == _C.BF16
    dl2, db2 = (None, plan.view(step, "db2")) if defer else (_gbuf(an2.logs), _gbuf(an2.bias))
    dl1, db1 = (None, plan.view(step, "db1")) if defer else (_gbuf(an1.logs), _gbuf(an1.bias))
    ones = _ones_col(net, dt) if defer else -1        # a1 carries a ones column: dbias of an1 = that column of dW1
    if ones >= 0:
        db1 = None
    if implicit:
        # dgrad3 -> ReLU' / ActNorm scale -> dgrad2 -> ReLU' / ActNorm scale -> dgrad1 in ONE kernel, dgrad3's operand
        # (flipped im2col of du) gathered in-kernel; d3col / d2 / d1 are stored once for the wgrad GEMMs
        d3col, d2, d1, da1 = K.cnet_backward_implicit(
            du, n, h, w, cout, k3p, net.packed("w3t", dt), net.packed("w2t", dt), net.packed("w1t", dt), hid, k1p,
            an2.logs.detach().reshape(-1), an2.logscale_factor, an1.logs.detach().reshape(-1), an1.logscale_factor,
            h2, h1, dbias2=db2, dbias1=db1, masks=ctx.get("masks"))
    elif fused:
        d2, d1, da1 = K.cnet_backward(d3col, net.packed("w3t", dt), net.packed("w2t", dt), net.packed("w1t", dt), hid,
                                      k1p, an2.logs.detach().reshape(-1), an2.logscale_factor,
                                      an1.logs.detach().reshape(-1), an1.logscale_factor, h2, h1, dbias2=db2, dbias1=db1)
    else:
        d2 = K.gemm(d3col, net.packed("w3t", dt), hid, k3p, _C.EPI_RELU_BWD, None, an2.logs.detach().reshape(-1),
                    an2.logscale_factor, y=h2, dlogs=dl2, dbias=db2, out_dtype=dt, ldo=kh)
    # (3) conv2 (1x1)
    dw2 = plan.view(step, "w2")
    if not fused:
        d1 = K.gemm(d2, net.packed("w2t", dt), hid, hid, _C.EPI_RELU_BWD, None, an1.logs.detach().reshape(-1),
                    an1.logscale_factor, y=h1, dlogs=dl1, dbias=db1, out_dtype=dt, ldo=kh)
    # (4) the three weight gradients (conv3 tap form, conv2, conv1 im2col form): side stream, see GradPlan.wgrads
    plan.wgrads([(d3col, h2, k3p, hid, plan.view(step, "w3")), (d2, h1, hid, hid, dw2),
                 (d1, a1, hid, k1p, plan.view(step, "w1"))], (d3col, d2, d1, h1, h2, a1))
    if defer:
        plan.defer_dlogs(step, net.packed("w2", dt), dw2, an2, db2, hid, hid)
        plan.defer_dlogs(step, net.packed("w1", dt), plan.view(step, "w1"), an1, db1, hid, k1p, ones_col=ones)
    # bf16 path: the nine-tap partial gradients are stored in bf16 (they are products of bf16 operands already) and
    # summed in fp32 by the mix adjoint
    if not fused:
        da1 = K.gemm(d1, net.packed("w1t", dt), k1p, hid, _C.EPI_STORE, out_dtype=_C.F32 if wide else dt)
    # (5) ActNorm + mix
    dense = step.permutation == 'invconv' and not step.invconv.lu_decomposition
    gw = _gbuf(step.invconv.weight) if dense else None
    if step.permutation == 'invconv' and gw is None:
        gw = torch.zeros(c * c, device=dev, dtype=torch.float32)
    if wide:
        dx = _mix_wide_backward(ctx["x"], dz, n, h, w, da1, net.in_channels, ctx["wmat"], ctx["winv"], an, gw, dld)
    else:
        dx = K.rows_actnorm_mix_bwd(ctx["x"], dz, n, h, w, da1=da1, cin=net.in_channels, weight=ctx["wmat"],
                                    indices=ctx["idx"], bias=an.bias.detach().reshape(-1),
                                    logs=an.logs.detach().reshape(-1), dw=gw, dlogs=_gbuf(an.logs),
                                    dbias=_gbuf(an.bias), logscale_factor=an.logscale_factor, dld=dld,
                                    winv=ctx["winv"])
    if step.permutation == 'invconv' and step.invconv.lu_decomposition:
        step.invconv.accumulate_lu_grads(gw.view(c, c))
    return dx


# ------------------------------------------------------------------ Split2d
def _split_conv(sp, x, n, c, h, w):
    ch = c // 2
    conv = sp.conv2d_zeros
    dt = config.resolve_conv_dtype(64, sp.conv_dtype)
    kp = round_up(9 * ch, 64)
    a = K.im2col_rows(x, n, h, w, 0, ch, 3, dt, kp)
    wp = conv._packs.get(("w0", dt), conv.weight,
                         lambda: K.pack_conv_weight(conv.weight.detach(), 0, dt, round_up(c, 16), kp))
    hrows = K.gemm(a, wp, c, kp, _C.EPI_ZEROS, conv.bias.detach(), conv.logs.detach().reshape(-1),
                   conv.logscale_factor, out_dtype=_C.F32)
    return a, hrows, dt, kp


def _split_forward(sp, x, n, c, h, w, ld, save):
    """Split2d forward (network/module.py:526-530) on rows x [P][C]; z1 stays in place as x[:, :C/2]."""
    a, hrows, dt, kp = _split_conv(sp, x, n, c, h, w)
    ld_out = K.rows_gaussian_logp(hrows, x, n, h * w, c // 2, c // 2, ld)
    return ld_out, (dict(x=x, a=a, hrows=hrows, dt=dt, kp=kp) if save else None)


def _split_reverse(sp, z1, n, ch, h, w, eps):
    """Split2d reverse (module.py:532-536): rows z1 [P][C/2] -> rows [P][C]."""
    _, hrows, _, _ = _split_conv(sp, z1, n, 2 * ch, h, w)
    return K.rows_split2d_sample(hrows, z1, ch, eps, n, ch, h * w)


def _split_backward(sp, ctx, dx, dld, n, c, h, w, plan):
    """dx: rows [P][C] whose channels 0..C/2-1 already hold the gradient of the returned z1."""
    x, hrows, a, dt, kp = ctx["x"], ctx["hrows"], ctx["a"], ctx["dt"], ctx["kp"]
    ch = c // 2
    conv = sp.conv2d_zeros
    if dld is None:
        dld = torch.zeros(n, device=dx.device, dtype=torch.float32)
    du = K.rows_split2d_bwd(x, hrows, dld, conv.logs.detach().reshape(-1), dx, _gbuf(conv.logs), _gbuf(conv.bias), n,
                            h * w, conv.logscale_factor)
    cp = round_up(c, 64)
    duc = K.im2col_rows(du, n, h, w, 0, c, 1, dt, cp)                       # convert + zero-pad to the GEMM tiling
    K.gemm_wgrad(duc, a, cp, kp, plan.view(sp, "w0"))
    wt = conv._packs.get(("w0t", dt), conv.weight,
                         lambda: K.pack_conv_weight(conv.weight.detach(), 2, dt, kp, cp))
    da = K.gemm(duc, wt, kp, cp, _C.EPI_STORE, out_dtype=_C.F32)
    K.rows_tapsum(da, dx, 0, ch, n, h, w, flip=True, accumulate=True)
    return dx


# ------------------------------------------------------------------ one FlowStep called as a layer (NCHW in / out)
def _layer_view(layer):
    """A one-layer flow around `layer`: owner of its pack / gradient plans and workspaces."""
    v = layer.__dict__.get("_rows_view")
    if v is None or v.layers[0] is not layer:          # (DataParallel replicas share __dict__ copies)
        v = layer.__dict__["_rows_view"] = _FlowView([layer])
    return v


def step_supported(step, x):
    """True iff a stand-alone FlowStep call on the NCHW tensor x can run on the pixel-major kernels (a layout change
    on the way in and out: the reference's FlowStep/FlowModel classes calling this package's layers, INTEGRATION.md
    section 1, second patch line)."""
    return (config.use_rows_path and torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
            and 0 < x.shape[0] <= 65535 and x.shape[1] == step.in_channels and _step_ok(step)
            and x.numel() < (1 << 31))


def _to_rows(x):
    n, c, h, w = x.shape
    rows = torch.empty(n * h * w, c, device=x.device, dtype=torch.float32)
    K.rows_squeeze(x.contiguous(), NCHW, c * h * w, rows, ROWS, c, n, c, h, w, 1, False)
    return rows


def _to_nchw(rows, n, c, h, w):
    out = torch.empty(n, c, h, w, device=rows.device, dtype=torch.float32)
    K.rows_squeeze(rows, ROWS, c, out, NCHW, c * h * w, n, c, h, w, 1, False)
    return out


def step_forward_nchw(step, x, ld, save=False, want_ld=False):
    """FlowStep.normal_flow on an NCHW tensor through the rows kernels -> (z NCHW, ld_out, ctx | None)."""
    n, c, h, w = x.shape
    view = _layer_view(step)
    prepare_packs(view, save, x.device)
    ws = _workspace(view, x.device, n, K.rows_coupling_nblk(h * w, c))
    z, ld_out, ctx = _step_forward(step, _to_rows(x), n, c, h, w, ld, ws, save, want_ld)
    return _to_nchw(z, n, c, h, w), ld_out, ctx


def step_reverse_nchw(step, z, ld=None):
    """FlowStep.reverse_flow on an NCHW tensor through the rows kernels -> (x NCHW, ld_out | None)."""
    n, c, h, w = z.shape
    view = _layer_view(step)
    prepare_packs(view, False, z.device)
    ws = _workspace(view, z.device, n, K.rows_coupling_nblk(h * w, c))
    out = _step_reverse(step, _to_rows(z), n, c, h, w, ws, ld)
    ld_out = None
    if ld is not None:
        out, ld_out = out
    return _to_nchw(out, n, c, h, w), ld_out


def step_backward_nchw(step, ctx, dy, dld):
    """Adjoint of step_forward_nchw: accumulates the step's parameter gradients into .grad, returns dx NCHW."""
    n, c, h, w = dy.shape
    plan = grad_plan(_layer_view(step), dy.device)
    plan.begin()
    dx = _step_backward(step, ctx, _to_rows(dy), dld, n, c, h, w, plan)
    ctx.clear()
    plan.finish()
    return _to_nchw(dx, n, c, h, w)


# ------------------------------------------------------------------ whole model
def _max_nblk(flow, h, w):
    from .model import FlowStep
    best, c = 1, None
    for layer in flow.layers:
        if isinstance(layer, module.Squeeze2d):
            h, w = h // layer.factor, w // layer.factor
        elif isinstance(layer, FlowStep):
            best = max(best, K.rows_coupling_nblk(h * w, layer.in_channels))
    return best


def encode(flow, z, ld, tape=None, add=None, want_ld=False):
    """FlowModel.encode (network/model.py:263-276).  z: NCHW fp32; ld: [N] fp32 or None.
    With `tape` (a list) every layer records what its adjoint needs.  Returns (z_out NCHW, ld_out).
    add: NCHW tensor added to z inside the entry squeeze (Glow's dequantisation noise); want_ld with ld None: the
    logdet is accumulated from zero (Glow's loss head adds the objective's constant start value)."""
    from .model import FlowStep
    z = z.contiguous()
    n, c, h, w = z.shape
    dev = z.device
    save = tape is not None
    prepare_packs(flow, save, dev)
    ws = _workspace(flow, dev, n, _max_nblk(flow, h, w))
    cur, layout, pitch = z, NCHW, c * h * w
    for layer in flow.layers:
        if isinstance(layer, module.Squeeze2d):
            f = layer.factor
            if h % f or w % f:
                raise ValueError("Squeeze2d: H, W = %d, %d not divisible by factor %d" % (h, w, f))
            dst = torch.empty(n * (h // f) * (w // f), c * f * f, device=dev, dtype=torch.float32)
            K.rows_squeeze(cur, layout, pitch, dst, ROWS, c * f * f, n, c, h, w, f, False, add=add)
            add = None
            if save:
                tape.append(("squeeze", layer, (c, h, w, layout, pitch)))
            c, h, w = c * f * f, h // f, w // f
            cur, layout, pitch = dst, ROWS, c
        elif isinstance(layer, FlowStep):
            assert layout == ROWS and pitch == c
            cur, ld, ctx = _step_forward(layer, cur, n, c, h, w, ld, ws, save, want_ld)
            if save:
                tape.append(("step", layer, ctx))
        else:
            assert layout == ROWS and pitch == c
            ld, ctx = _split_forward(layer, cur, n, c, h, w, ld, save)
            if save:
                tape.append(("split", layer, ctx))
            c = c // 2                      # z1 = first half of the same rows (pitch unchanged)
    out = torch.empty(n, c, h, w, device=dev, dtype=torch.float32)
    K.rows_squeeze(cur, layout, pitch, out, NCHW, c * h * w, n, c, h, w, 1, False)
    return out, ld


def backward(flow, tape, dz, dld):
    """Adjoint of `encode` over its tape.  dz: NCHW grad of z_out; dld: [N] grad of ld_out or None.
    Accumulates parameter gradients into .grad; returns the NCHW gradient of the input."""
    dz = dz.contiguous()
    n, c, h, w = dz.shape
    dev = dz.device
    plan = grad_plan(flow, dev)
    plan.begin()
    # a flow that ends in a Split2d (the head of a hybrid model) returned z1 only: dz fills the first half of the
    # [P][2c] gradient rows, the Split2d adjoint writes the second half
    pitch = 2 * c if (tape and tape[-1][0] == "split") else c
    cur = torch.empty(n * h * w, pitch, device=dev, dtype=torch.float32)
    K.rows_squeeze(dz, NCHW, c * h * w, cur, ROWS, pitch, n, c, h, w, 1, False)
    hook = flow.__dict__.get("_level_done_hook")      # train.FusedTrainStep: per-level gradient all-reduce
    for kind, layer, ctx in reversed(tape):
        if kind == "step":
            cur = _step_backward(layer, ctx, cur, dld, n, c, h, w, plan)
            ctx.clear()                     # saved (or recomputed) activations of this step go back to the allocator now
        elif kind == "split":
            c = c * 2                       # `cur` is the [P][C] buffer the squeeze adjoint below left half-filled
            cur = _split_backward(layer, ctx, cur, dld, n, c, h, w, plan)
            ctx.clear()
        else:
            lv = plan.level_of.get(id(layer))
            if lv is not None:              # the level that starts at this Squeeze2d is fully differentiated
                plan.finish_level(lv)
                if hook is not None:
                    hook(lv, plan.n_levels)
            c0, h0, w0, layout, pitch = ctx
            if layout == NCHW:
                dst = torch.empty(n, c0, h0, w0, device=dev, dtype=torch.float32)
            else:
                dst = torch.empty(n * h0 * w0, pitch, device=dev, dtype=torch.float32)
            K.rows_squeeze(cur, ROWS, c, dst, layout, pitch, n, c0, h0, w0, layer.factor, True)
            cur, c, h, w = dst, c0, h0, w0
    plan.finish()
    return cur


def decode(flow, z, eps_std=None, eps_list=None):
    """FlowModel.decode (network/model.py:278-294).  z: NCHW latent; returns the NCHW image batch."""
    from .model import FlowStep
    z = z.contiguous()
    n, c, h, w = z.shape
    dev = z.device
    prepare_packs(flow, False, dev)
    ws = _workspace(flow, dev, n, 1)
    cur = torch.empty(n * h * w, c, device=dev, dtype=torch.float32)
    K.rows_squeeze(z, NCHW, c * h * w, cur, ROWS, c, n, c, h, w, 1, False)
    k = 0
    layers = list(flow.layers)
    for li in range(len(layers) - 1, -1, -1):
        layer = layers[li]
        if isinstance(layer, FlowStep):
            cur = _step_reverse(layer, cur, n, c, h, w, ws)
        elif isinstance(layer, module.Split2d):
            if eps_list is not None:
                e = eps_list[k]
            else:
                e = module.GaussianDiag.eps(torch.empty(n, c, h, w, device=dev, dtype=torch.float32), eps_std)
            k += 1
            cur = _split_reverse(layer, cur, n, c, h, w, e)
            c = c * 2
        else:
            f = layer.factor
            if c < f * f or c % (f * f):
                raise ValueError("Squeeze2d: C = %d not divisible by factor^2" % c)
            c0, h0, w0 = c // (f * f), h * f, w * f
            if li == 0:
                dst = torch.empty(n, c0, h0, w0, device=dev, dtype=torch.float32)
                K.rows_squeeze(cur, ROWS, c, dst, NCHW, c0 * h0 * w0, n, c0, h0, w0, f, True)
            else:
                dst = torch.empty(n * h0 * w0, c0, device=dev, dtype=torch.float32)
                K.rows_squeeze(cur, ROWS, c, dst, ROWS, c0, n, c0, h0, w0, f, True)
            cur, c, h, w = dst, c0, h0, w0
    return cur
