"""Harness that drives the reference implementation for bench.py's baseline arms.

`load()` imports the staged reference (baseline/_ref, see stage_reference.py) and returns its modules; when it is not
staged the callers fall back to the oracle port (oracle/glow_oracle.py), which is pinned to the reference by
tests/golden.  Training under torch >= 2 needs the out-of-place restatement of FlowStep.normal_flow (the
reference's `z2 += f(z1)` on a view trips autograd's version check; the forward is bit-identical) -- applied here as
a monkey-patch, the reference files stay unmodified.

The iteration mirrors network/trainer.py:84-150: lr from noam_decay, zero_grad, loss = mean(nll), backward,
clip_grad_value_(5), clip_grad_norm_(100), Adam step.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "network"))


def load():
    """(network.model, network.module, misc.ops, misc.lr_scheduler) of the staged reference."""
    if not available():
        raise RuntimeError("the reference is not staged under baseline/_ref (run baseline/stage_reference.py in the "
                           "build container)")
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)      # the reference's regex literals under Python 3.12
    for p in (os.path.join(REF, "_shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from network import model as rmodel
    from network import module as rmodule
    from misc import ops as rops
    from misc import lr_scheduler as rsched
    _patch_normal_flow(rmodel, rops)
    return rmodel, rmodule, rops, rsched


def _patch_normal_flow(rmodel, rops):
    if getattr(rmodel.FlowStep, "_glowk_patched", False):
        return

    def normal_flow(self, x, logdet=None):
        z, logdet = self.actnorm(x, logdet=logdet, reverse=False)
        if self.permutation == 'invconv':
            z, logdet = self.invconv(z, logdet, reverse=False)
        elif self.permutation == 'reverse':
            z = self.reverse(z, reverse=False)
        else:
            z = self.shuffle(z, reverse=False)
        z1, z2 = rops.split_channel(z, 'simple')
        if self.coupling == 'additive':
            z2 = z2 + self.f(z1)
        else:
            shift, scale = rops.split_channel(self.f(z1), 'cross')
            scale = torch.sigmoid(scale + 2.)
            z2 = (z2 + shift) * scale
            logdet = rops.reduce_sum(torch.log(scale), dim=[1, 2, 3]) + logdet
        return rops.cat_channel(z1, z2), logdet

    rmodel.FlowStep.normal_flow = normal_flow
    rmodel.FlowStep._glowk_patched = True


class ReferenceTrainer:
    """One replica of the reference's training loop body on `device` ('cpu' or 'cuda:0'), synthetic images."""

    def __init__(self, hps, device, tf32=False, seed=2384):
        rmodel, _, _, rsched = load()
        self.device = torch.device(device)
        if self.device.type == "cuda":
            torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
            torch.backends.cudnn.allow_tf32 = bool(tf32)
        hps.device.graph = [device]
        np.random.seed(seed)
        torch.manual_seed(seed)
        self.graph = rmodel.Glow(hps).to(self.device)
        self.rmodel = rmodel
        args = hps.optim.optimizer_args
        self.opt = torch.optim.Adam(self.graph.parameters(), lr=args.lr, betas=tuple(args.betas), eps=args.eps)
        self.sched = lambda step: rsched.noam_decay(args.lr, step, **dict(hps.optim.lr_scheduler_args))
        self.clip_value, self.clip_norm = hps.ablation.max_grad_clip, hps.ablation.max_grad_norm
        self.step_no = 0

    def init_actnorm(self, x):
        self.graph.train()
        with torch.no_grad():
            self.graph(x=x)

    def step(self, x):
        """trainer.py:89-150 for the generative loss; returns the loss tensor."""
        self.graph.train()
        lr = self.sched(self.step_no)
        for g in self.opt.param_groups:
            g['lr'] = lr
        self.opt.zero_grad()
        z, nll, _ = self.graph(x=x)
        loss = self.rmodel.Glow.generative_loss(nll)
        loss.backward()
        torch.nn.utils.clip_grad_value_(self.graph.parameters(), self.clip_value)
        torch.nn.utils.clip_grad_norm_(self.graph.parameters(), self.clip_norm)
        self.opt.step()
        self.step_no += 1
        return loss.detach()

    def sample(self, eps_std=0.7):
        self.graph.eval()
        with torch.no_grad():
            return self.graph(z=None, eps_std=eps_std, reverse=True)


def time_train(hps, device, batch, shape, steps, warmup, tf32=False, pinned_e2e=False):
    """img/s of ReferenceTrainer.step on synthetic [0,1) images; device time when on CUDA (events), wall time on CPU.
    pinned_e2e: each step copies its batch from pinned host memory and reads the loss back."""
    tr = ReferenceTrainer(hps, device, tf32=tf32)
    dev = tr.device
    g = torch.Generator().manual_seed(1234)
    xs = [torch.rand(batch, shape[2], shape[0], shape[1], generator=g) for _ in range(2)]
    if dev.type == "cuda":
        xs = [x.pin_memory() for x in xs]
    xd = [x.to(dev) for x in xs]
    tr.init_actnorm(xd[0])
    last = None
    for i in range(warmup):
        last = tr.step(xd[i % 2])
    if dev.type == "cuda":
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            if pinned_e2e:
                last = float(tr.step(xs[i % 2].to(dev, non_blocking=True)))
            else:
                last = tr.step(xd[i % 2])
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
    else:
        t0 = time.perf_counter()
        for i in range(steps):
            last = tr.step(xd[i % 2])
        sec = time.perf_counter() - t0
    return batch * steps / sec, sec / steps, float(last)


def time_sample(hps, device, batch, steps, warmup, tf32=False):
    tr = ReferenceTrainer(hps, device, tf32=tf32)
    for m in tr.graph.modules():
        if m.__class__.__name__.find("ActNorm") >= 0:
            m.bias_inited = m.logs_inited = True
    dev = tr.device
    for _ in range(warmup):
        tr.sample()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        x = tr.sample()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    assert tuple(x.shape)[0] == batch
    return batch * steps / sec, sec / steps
