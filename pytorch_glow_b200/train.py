"""Data-parallel training step (the caller side of the hot path, network/trainer.py:84-150).

One process per GPU.  All trainable parameters live in ONE flat fp32 arena with a matching flat
gradient arena, so that
  * the backward kernels accumulate straight into the arena (no per-tensor autograd adds),
  * gradient averaging across ranks is a single NCCL all-reduce over NVLink (the reference's
    DataParallel re-broadcasts 176 MB of parameters and reduces 176 MB of gradients to device 0
    every step, trainer.py:117-120),
  * clip_grad_value_(5) -> clip_grad_norm_(100) -> Adam (trainer.py:142-150) is three kernels over
    the arena (glowk_optim_*), and
  * the whole step can be captured in CUDA graphs (launch-bound otherwise: ~2000 kernels/step).
"""
import math
import os

import gc

import torch
import torch.distributed as dist

from . import functional as K
from . import module as _module


def noam_lr(base_lr, global_step, warmup_steps=4000, min_lr=None):
    """misc/lr_scheduler.py:18-37 (the schedule profile/celeba.json selects)."""
    step_num = global_step + 1.0
    lr = base_lr * warmup_steps ** 0.5 * min(step_num * warmup_steps ** -1.5, step_num ** -0.5)
    if global_step >= warmup_steps and min_lr is not None:
        lr = max(lr, min_lr)
    return lr


class _DivAfter:
    """Work handle of a SUM all-reduce that still has to be divided by the world size (gloo has no AVG)."""

    def __init__(self, work, t, n):
        self.work, self.t, self.n = work, t, n

    def wait(self):
        self.work.wait()
        self.t.div_(self.n)


def allreduce_mean_(t, group=None, async_op=False):
    """In-place mean over the ranks of `group` (NCCL: one AVG all-reduce over NVLink; gloo has no AVG: SUM then
    divide).  Gradients are averaged BEFORE clipping so that every rank clips identical tensors.
    async_op: returns a handle whose wait() orders the current stream after the collective."""
    if dist.get_backend(group) == "nccl":
        w = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        return w if async_op else t
    w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if async_op:
        return _DivAfter(w, t, dist.get_world_size(group))
    t.div_(dist.get_world_size(group))
    return t


def arena_level_ranges(glow, arena):
    """[(lo, hi)] slices of the flat gradient arena per flow level (level = number of Squeeze2d layers up to the
    layer, minus one), in level order.  named_parameters() walks flow.layers in order, so every level is one
    contiguous slice; parameters outside `flow.layers` come after them."""
    from .module import Squeeze2d
    level_of_layer, lvl = {}, -1
    for i, layer in enumerate(glow.flow.layers):
        if isinstance(layer, Squeeze2d):
            lvl += 1
        level_of_layer[i] = max(lvl, 0)
    n_levels = max(lvl, 0) + 1
    ranges = [[None, None] for _ in range(n_levels)]
    end_flow = 0
    for name, p, off in zip(arena.names, arena.params, arena.offsets):
        parts = name.split(".")
        if len(parts) > 2 and parts[0] == "flow" and parts[1] == "layers":
            lv = level_of_layer[int(parts[2])]
            size = (p.numel() + 3) // 4 * 4
            r = ranges[lv]
            r[0] = off if r[0] is None else min(r[0], off)
            r[1] = off + size if r[1] is None else max(r[1], off + size)
            end_flow = max(end_flow, off + size)
    out = [(r[0] or 0, r[1] or 0) if r[0] is not None else (0, 0) for r in ranges]
    for (lo, hi), (lo2, hi2) in zip(out, out[1:]):
        assert hi == lo2, "flow levels are not contiguous in the arena"
    assert not out or out[0][0] == 0
    return out


def shard_batch(x, rank, world_size):
    """This rank's contiguous slice of a global batch (DataParallel's scatter on dim 0, trainer.py:118-123)."""
    n = x.shape[0]
    if n % world_size:
        raise ValueError("global batch %d is not divisible by %d ranks (model.py:350-353)" % (n, world_size))
    per = n // world_size
    return x[rank * per:(rank + 1) * per]


class FlatArena:
    """Re-homes a module's trainable parameters (and their .grad) into flat fp32 buffers."""

    def __init__(self, model, skip=("h_top",)):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n.split(".")[-1] not in skip]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # 16-byte aligned slots
        self.offsets, self.numel = offs, total
        dev = self.params[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, offs):
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view(p.shape)
            p.grad = self.grad[o:o + n].view(p.shape)
        _module.bump_weight_generation()

    def rebind_grads(self):
        """(Re)attach .grad views -- e.g. after someone called zero_grad(set_to_none=True)."""
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)


class FusedTrainStep:
    """forward + backward + (all-reduce) + clip + Adam for a pytorch_glow_b200.Glow.

    Mirrors one iteration of Trainer.train (network/trainer.py:84-150) for the generative loss.
    `use_graphs=True` captures the iteration in two CUDA graphs around the (eager) NCCL all-reduce.
    """

    def __init__(self, glow, lr=1e-3, betas=(0.9, 0.9999), eps=1e-8, max_grad_clip=5.0, max_grad_norm=100.0,
                 warmup_steps=4000, min_lr=1e-4, use_graphs=False, process_group=None, world_size=1, overlap=None,
                 optimizer="adam"):
        self.glow = glow
        if optimizer not in ("adam", "adamax"):            # network/builder.py:10-13
            raise ValueError("optimizer must be 'adam' or 'adamax'")
        self.optimizer = optimizer
        self.base_lr, self.betas, self.eps = lr, betas, eps
        self.max_grad_clip, self.max_grad_norm = max_grad_clip, max_grad_norm
        self.warmup_steps, self.min_lr = warmup_steps, min_lr
        self.world_size, self.pg = world_size, process_group
        for p in glow.parameters():
            if p is glow.h_top:
                p.requires_grad_(False)           # never receives a gradient
        self.arena = FlatArena(glow)
        dev = self.arena.flat.device
        self.exp_avg = torch.zeros_like(self.arena.flat)
        self.exp_avg_sq = torch.zeros_like(self.arena.flat)
        self.ws = K.optim_workspace(dev)
        self._one = torch.ones((), device=dev, dtype=torch.float32)      # seed of loss.backward
        # [lr, 1-beta1^t, sqrt(1-beta2^t)] are computed ON THE DEVICE from a device step counter inside the
        # (captured) optimizer step: no host memory is read when the GPU gets there, however far the CPU ran ahead
        self.sched_dev = torch.zeros(4, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self.global_step = 0
        self._inited_ok = False
        # Gradient all-reduce overlapped with the backward pass (the reference's DataParallel reduces
        # after the whole backward, trainer.py:117-123,140): the flow differentiates its levels top-down, and as soon
        # as a level is done its slice of the gradient arena leaves on NCCL's stream while the lower (larger-image,
        # slower) levels are still running.  With CUDA graphs the collectives are captured into the iteration graph
        # (one graph, no host round trip between backward, all-reduce and optimizer).  GLOWK_DDP_OVERLAP=0: one
        # blocking all-reduce between two graphs (round-1 behaviour).
        if overlap is None:
            overlap = os.environ.get("GLOWK_DDP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap) and world_size > 1
        self._pending = []
        self.level_ranges = arena_level_ranges(glow, self.arena) if self.overlap else None
        self.use_graphs = use_graphs
        self._g_fb = self._g_opt = None
        self._static_x = None
        self._static_loss = None

    # -- step 0 of the reference trainer: data-dependent ActNorm init on the first shard (trainer.py:112-115)
    def init_actnorm(self, x):
        self.glow.train()
        with torch.no_grad():
            self.glow(x=x)
        _module.bump_weight_generation()
        if self.world_size > 1:
            dist.broadcast(self.arena.flat, src=0, group=self.pg)    # rank 0's initialisation wins

    @property
    def lr(self):
        """Learning rate of the NEXT iteration (host mirror of the device schedule, trainer.py:89-92)."""
        return noam_lr(self.base_lr, self.global_step, self.warmup_steps, self.min_lr)

    def _check_inited(self):
        """The data-dependent ActNorm init must have run (init_actnorm / a loaded snapshot) before an iteration is
        captured or stepped: a graph warm-up would otherwise initialise from the warm-up batch and then restore the
        pre-init arena while the `inited` flags stay set."""
        for m in self.glow.modules():
            if m.__class__.__name__.find("ActNorm") >= 0 and not (m.bias_inited and m.logs_inited):
                raise RuntimeError("FusedTrainStep.step: ActNorm layers are not initialised -- call init_actnorm(x) on "
                                   "the first batch (trainer.py:112-115) or load a snapshot first")

    # -- torch.optim.Adam wire format (builder.py:91-93 `optimizer.load_state_dict(state['optimizer'])`)
    def state_dict(self):
        """Adam state in torch.optim.Adam's layout, parameter indices in `glow.parameters()` order."""
        index = {id(p): i for i, p in enumerate(self.glow.parameters())}
        state = {}
        second = "exp_avg_sq" if self.optimizer == "adam" else "exp_inf"      # torch.optim.Adamax' key
        step_t = float(self.global_step)
        for p, o in zip(self.arena.params, self.arena.offsets):
            n = p.numel()
            state[index[id(p)]] = {"step": torch.tensor(step_t),
                                   "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                                   second: self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(index)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts what `state_dict()` / a torch.optim.Adam over `glow.parameters()` produced; restores the moments
        and the iteration count (Noam schedule + bias corrections resume where they stopped)."""
        index = {id(p): i for i, p in enumerate(self.glow.parameters())}
        st = sd.get("state", {})
        steps = []
        for p, o in zip(self.arena.params, self.arena.offsets):
            e = st.get(index[id(p)])
            if e is None:
                e = st.get(str(index[id(p)]))
            n = p.numel()
            if e is None:                       # a parameter that never received a gradient has no entry
                self.exp_avg[o:o + n].zero_(); self.exp_avg_sq[o:o + n].zero_()
                continue
            self.exp_avg[o:o + n].copy_(e["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(e["exp_avg_sq" if "exp_avg_sq" in e else "exp_inf"].reshape(-1))
            steps.append(int(float(e["step"])))
        if steps:
            self.set_global_step(max(steps))
        groups = sd.get("param_groups") or []
        if groups:
            self.betas = tuple(groups[0].get("betas", self.betas))
            self.eps = groups[0].get("eps", self.eps)

    def set_global_step(self, step):
        self.global_step = int(step)
        self.step_dev.fill_(int(step))

    def _level_done(self, level, n_levels):
        """rows_path.backward finished a level: start the all-reduce of that level's gradient slice."""
        lo, hi = self.level_ranges[level]
        if hi > lo:
            self._pending.append(allreduce_mean_(self.arena.grad[lo:hi], self.pg, async_op=True))

    def _forward_backward(self, x):
        self.arena.grad.zero_()
        self.arena.rebind_grads()
        if self.overlap:
            self._pending = []
            self.glow.flow.__dict__["_level_done_hook"] = self._level_done
        try:
            z, nll, _ = self.glow(x=x)
            loss = self.glow.generative_loss(nll)
            loss.backward(self._one)                 # explicit seed: no ones_like fill kernel in the step
        finally:
            self.glow.flow.__dict__.pop("_level_done_hook", None)
        if self.overlap:
            covered = sum(hi - lo for lo, hi in self.level_ranges)
            if len(self._pending) != sum(1 for lo, hi in self.level_ranges if hi > lo):
                # the flow did not run on the level-wise path (hybrid / NCHW fall-back): reduce everything now
                for w in self._pending:
                    w.wait()
                self._pending = [allreduce_mean_(self.arena.grad, self.pg, async_op=True)]
            elif covered < self.arena.numel:     # parameters outside the flow (learn_top, y_emb, classifier)
                self._pending.append(allreduce_mean_(self.arena.grad[covered:], self.pg, async_op=True))
            for w in self._pending:              # the current stream waits for NCCL's stream
                w.wait()
            self._pending = []
        return loss.detach()

    def _optimizer(self):
        K.optim_schedule(self.step_dev, self.sched_dev, self.base_lr, self.warmup_steps, self.min_lr, self.betas[0],
                         self.betas[1])
        K.optim_clip_norm(self.arena.grad, self.max_grad_clip, self.max_grad_norm, self.ws)
        step_fn = K.optim_adam if self.optimizer == "adam" else K.optim_adamax     # exp_avg_sq doubles as Adamax' exp_inf
        step_fn(self.arena.flat, self.arena.grad, self.exp_avg, self.exp_avg_sq, self.ws, self.global_step + 1,
                0.0, self.betas[0], self.betas[1], self.eps, sched=self.sched_dev)

    def _allreduce(self):
        if self.world_size > 1 and not self.overlap:
            allreduce_mean_(self.arena.grad, self.pg)

    def step(self, x):
        """One training iteration on the device batch x [B,3,H,W] in [0,1).  Returns the loss (bits/dim)
        as a 0-dim device tensor."""
        self.glow.train()
        if not self._inited_ok:
            self._check_inited()
            self._inited_ok = True
        if not self.use_graphs:
            loss = self._forward_backward(x)
            self._allreduce()
            self._optimizer()
            _module.bump_weight_generation()
        else:
            if self._g_fb is None:
                self._capture(x)
            self._static_x.copy_(x, non_blocking=True)
            self._g_fb.replay()                   # overlap: forward + backward + per-level all-reduces + optimizer
            if self._g_opt is not None:
                self._allreduce()
                self._g_opt.replay()
            _module.bump_weight_generation()      # eager callers (sampling, eval) must re-pack the new weights
            loss = self._static_loss
        self.global_step += 1
        return loss

    @property
    def grad_norm(self):
        return self.ws[0]

    def _capture(self, x):
        self._static_x = torch.empty_like(x)
        self._static_x.copy_(x)
        # warm-up on a side stream (allocator + lazy init), then capture; weights caches are invalidated so
        # that every pack / LU kernel is part of the graph and re-runs on each replay
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        saved = (self.arena.flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone())
        with torch.cuda.stream(s):
            for _ in range(2):
                _module.bump_weight_generation()
                self._forward_backward(self._static_x)
                self._optimizer()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gc.collect()      # no autograd node of the warm-up (created on the side stream) may survive into the capture
        self.arena.flat.copy_(saved[0]); self.exp_avg.copy_(saved[1]); self.exp_avg_sq.copy_(saved[2])
        self.step_dev.copy_(saved[3])
        _module.bump_weight_generation()
        from . import _C
        c0 = _C.launch_count
        self._g_fb = torch.cuda.CUDAGraph()
        if self.overlap:
            # NCCL collectives are captured with the kernels: one graph for the whole iteration
            with torch.cuda.graph(self._g_fb):
                self._static_loss = self._forward_backward(self._static_x)
                self._optimizer()
            self._g_opt = None
        else:
            with torch.cuda.graph(self._g_fb):
                self._static_loss = self._forward_backward(self._static_x)
            self._g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g_opt):
                self._optimizer()
        self.captured_calls = _C.launch_count - c0     # glowk C-ABI launches replayed per step
        # the captures above only recorded; nothing has executed yet


class GraphedSampler:
    """`Glow(z=None, eps_std=..., reverse=True)` (network/inferer.py:53-60, model.py:454-471) captured in ONE CUDA
    graph: a 64x64 K=32 L=3 reverse pass is ~600 kernels of 10-130 us, i.e. launch-bound when driven from Python.
    Every replay draws fresh prior / Split2d noise (torch's CUDA generator is graph-safe).  The weights are read
    through caches filled at capture time: call `recapture()` after they change."""

    def __init__(self, glow, eps_std=0.7, y_onehot=None):
        self.glow, self.eps_std, self.y_onehot = glow, eps_std, y_onehot
        self._graph = None
        self.out = None

    def recapture(self):
        self.glow.eval()
        _module.bump_weight_generation()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):                       # allocator warm-up, weight packing, LU of the 1x1 convs
                self.glow(z=None, y_onehot=self.y_onehot, eps_std=self.eps_std, reverse=True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph), torch.no_grad():
            self.out = self.glow(z=None, y_onehot=self.y_onehot, eps_std=self.eps_std, reverse=True)
        return self

    def __call__(self):
        """Images [B,3,H,W] of one reverse pass (a static buffer: clone it to keep it across calls)."""
        if self._graph is None:
            self.recapture()
        self._graph.replay()
        return self.out
