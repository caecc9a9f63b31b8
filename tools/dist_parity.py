"""On-hardware data-parallel parity (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_parity.py [--out profiles/r2_dist_parity_2gpu.jsonl]

Checks, for the reference's data-parallel semantics (network/trainer.py:112-123,138-150: scatter the batch on dim 0,
average the gradients, clip, Adam):
  1. after 5 iterations every rank holds BIT-IDENTICAL parameter / Adam arenas (bf16 convs, CUDA graphs, per-level
     all-reduce overlapped with the backward pass -- the benchmarked configuration);
  2. the same with one blocking all-reduce (GLOWK_DDP_OVERLAP=0 path) gives the same loss trajectory;
  3. N ranks x B images == 1 rank x N*B images: loss and averaged gradients of an fp32 iteration agree to 1e-5
     (same ActNorm initialisation, same dequantisation noise per image);
  4. step time with and without the overlap at a small per-GPU batch.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_glow_b200 as G  # noqa: E402
from pytorch_glow_b200.hps import make_hps  # noqa: E402
from pytorch_glow_b200.train import FusedTrainStep, shard_batch  # noqa: E402

SHAPE = (64, 64, 3)


def build(K, L, batch, device, conv_dtype):
    np.random.seed(2384); torch.manual_seed(2384)
    glow = G.Glow(make_hps(SHAPE, K=K, L=L, hidden_channels=512, batch=batch)).to(device)
    glow.flow.set_conv_dtype(conv_dtype)
    return glow


class NoiseFeed:
    """Feeds a prescribed dequantisation noise tensor to Glow.forward (network/model.py:421-423 draws it with
    torch.nn.init.uniform_), so that sharded and un-sharded runs dequantise every image identically."""

    def __init__(self, device, shape):
        self.buf = torch.empty(shape, device=device)
        self.orig = torch.nn.init.uniform_

    def __enter__(self):
        buf = self.buf

        def fake(t, a=0., b=1.):
            t.copy_(buf)
            return t
        torch.nn.init.uniform_ = fake
        return self

    def __exit__(self, *a):
        torch.nn.init.uniform_ = self.orig


def run_steps(ts, xs, noises, device, steps):
    losses = []
    with NoiseFeed(device, xs[0].shape) as nf:
        for t in range(steps):
            nf.buf.copy_(noises[t])
            losses.append(float(ts.step(xs[t])))
    return losses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    rec = {"world": world}
    K, L, B, steps = 4, 3, 8, a.steps
    g = torch.Generator().manual_seed(7)
    xg = [torch.rand(B * world, 3, 64, 64, generator=g) for _ in range(steps)]
    ng = [torch.rand(B * world, 3, 64, 64, generator=g) / 256 for _ in range(steps)]
    xs = [shard_batch(x, rank, world).to(device) for x in xg]
    ns = [shard_batch(n, rank, world).to(device) for n in ng]

    # ---- 1 + 2: bf16, graphs, overlapped vs blocking all-reduce
    results = {}
    for overlap in (True, False):
        glow = build(K, L, B, device, "bf16")
        ts = FusedTrainStep(glow, use_graphs=True, world_size=world, overlap=overlap)
        ts.init_actnorm(xs[0])
        losses = run_steps(ts, xs, ns, device, steps)
        torch.cuda.synchronize()
        flat = [torch.empty_like(ts.arena.flat) for _ in range(world)]
        dist.all_gather(flat, ts.arena.flat)
        m = [torch.empty_like(ts.exp_avg) for _ in range(world)]
        dist.all_gather(m, ts.exp_avg)
        same = all(torch.equal(flat[0], f) for f in flat) and all(torch.equal(m[0], f) for f in m)
        results[overlap] = (losses, same, ts.arena.flat.clone())
        del ts, glow
    rec["bf16_graphs_overlap_arenas_bit_identical_across_ranks"] = results[True][1]
    rec["bf16_graphs_blocking_arenas_bit_identical_across_ranks"] = results[False][1]
    rec["loss_overlap"] = results[True][0]
    rec["loss_blocking"] = results[False][0]
    rec["overlap_vs_blocking_max_param_rel_diff"] = float((results[True][2] - results[False][2]).abs().max() /
                                                          results[False][2].abs().max())

    # ---- 3: N x B == 1 x N*B (fp32 convs, eager), same ActNorm init, same noise per image.  Compared: the loss and
    # the AVERAGED GRADIENTS of one iteration (1e-5), then the parameters after `steps` iterations (informational:
    # Adam divides by sqrt(v), so coordinates with near-zero gradients take full lr-sized steps of either sign)
    glow = build(K, L, B, device, "fp32")
    ts = FusedTrainStep(glow, use_graphs=False, world_size=world)
    ts.init_actnorm(xs[0])                               # rank 0's shard initialises, broadcast (trainer.py:112-115)
    init_sd = {k: v.detach().clone() for k, v in glow.state_dict().items() if k != "h_top"}
    with NoiseFeed(device, xs[0].shape) as nf:
        nf.buf.copy_(ns[0])
        l0 = ts._forward_backward(xs[0]).clone()
        ts._allreduce()
    dist.all_reduce(l0, op=dist.ReduceOp.AVG)
    dp_grad = ts.arena.grad.clone()
    losses_dp = run_steps(ts, xs, ns, device, steps)
    dp_flat = ts.arena.flat.clone()
    names, offs, params = ts.arena.names, ts.arena.offsets, ts.arena.params
    del ts
    if rank == 0:
        big = build(K, L, B * world, device, "fp32")
        big.load_state_dict(init_sd, strict=False)
        big.set_actnorm_inited()
        ts1 = FusedTrainStep(big, use_graphs=False, world_size=1)
        with NoiseFeed(device, xg[0].shape) as nf:
            nf.buf.copy_(ng[0].to(device))
            l1 = ts1._forward_backward(xg[0].to(device))
        gworst = 0.0
        for o, p in zip(offs, params):
            nump = p.numel()
            a_, b_ = dp_grad[o:o + nump], ts1.arena.grad[o:o + nump]
            gworst = max(gworst, float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-12)))
        rec["fp32_N_x_B_vs_1_x_NB_loss"] = [float(l0), float(l1)]
        rec["fp32_N_x_B_vs_1_x_NB_worst_grad_rel_err"] = gworst
        losses_1 = run_steps(ts1, [x.to(device) for x in xg], [n.to(device) for n in ng], device, steps)
        assert ts1.arena.names == names
        worst = 0.0
        for o, p in zip(offs, params):
            nump = p.numel()
            a_, b_ = dp_flat[o:o + nump], ts1.arena.flat[o:o + nump]
            worst = max(worst, float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-12)))
        rec["fp32_N_x_B_vs_1_x_NB_worst_param_rel_err_after_steps"] = worst
        del ts1, big
    dist.barrier()

    # ---- 4: step time, small per-GPU batch, full-depth model
    for b in (64,):
        for overlap in (False, True):
            glow = build(32, 3, b, device, "bf16")
            ts = FusedTrainStep(glow, use_graphs=True, world_size=world, overlap=overlap)
            x = torch.rand(b, 3, 64, 64, device=device)
            ts.init_actnorm(x)
            for _ in range(4):
                ts.step(x)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 20
            for _ in range(n):
                ts.step(x)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / n], device=device)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            rec["ms_per_step_B%d_%s" % (b, "overlap" if overlap else "blocking")] = float(ms)
            del ts, glow
            torch.cuda.empty_cache()
    if rank == 0:
        rec["when"] = time.strftime("%Y-%m-%d %H:%M:%S")
        print(json.dumps(rec))
        if a.out:
            with open(a.out, "a") as f:
                f.write(json.dumps(rec) + "\n")
        ok = rec["bf16_graphs_overlap_arenas_bit_identical_across_ranks"] and rec["bf16_graphs_blocking_arenas_bit_identical_across_ranks"]
        ok = ok and rec["fp32_N_x_B_vs_1_x_NB_worst_grad_rel_err"] < 1e-5
        ok = ok and abs(rec["fp32_N_x_B_vs_1_x_NB_loss"][0] - rec["fp32_N_x_B_vs_1_x_NB_loss"][1]) < 1e-5
        print("DIST PARITY", "OK" if ok else "FAILED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
