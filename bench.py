#!/usr/bin/env python
"""Headline benchmark: training (and sampling) images/s of the 64x64 CelebA-shaped Glow
(K=32, L=3, hidden 512, affine coupling, invertible 1x1 conv -- profile/celeba.json:47-70 of the
reference) on N B200s of one node, synthetic images, one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU arithmetic (oracle port) on the host cores

Prints ONE JSON line (rank 0).  A "step" is one full training iteration of network/trainer.py:84-150:
dequantise -> encode -> bits/dim -> backward -> gradient all-reduce -> clip(5)/clip-norm(100) -> Adam.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (image_shape HWC, K, L, hidden, coupling, default per-GPU batch, fwd GFLOP/img)
    "celeba64": ((64, 64, 3), 32, 3, 512, "affine", 512, 32.06),
    "cifar32": ((32, 32, 3), 32, 3, 512, "affine", 256, 8.02),
    "cifar32_additive": ((32, 32, 3), 32, 3, 512, "additive", 256, 7.22),
    "celebahq256": ((256, 256, 3), 32, 6, 512, "affine", 32, 537.6),
    "tiny": ((32, 32, 3), 4, 3, 64, "affine", 16, None),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="celeba64", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = workload default)")
    ap.add_argument("--sample-batch", type=int, default=256, help="per-GPU sampling batch (BASELINE config 4)")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sample", action="store_true")
    ap.add_argument("--conv-dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch-eager reference on the same GPU")
    ap.add_argument("--ref-gpu-batch", type=int, default=64, help="batch of the torch-eager reference on the GPU "
                    "(its fp32 activations are ~0.35 GB/img: 512 does not fit)")
    ap.add_argument("--sweep", default=None, help="comma-separated per-GPU batches for the batch sweep "
                    "(default at N=1: 25,64,128,256; 'none' to skip)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: --batch is the GLOBAL batch, split over the ranks")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off ...)")
    ap.add_argument("--profile-step", action="store_true",
                    help="for ncu --profile-from-start off: run ONE eager train step (and one sampling pass) "
                         "between cudaProfilerStart/Stop after warm-up, print nothing, exit")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not (t0 - 0.1 <= t <= t1 + 0.3):
                continue
            f = [c.strip() for c in line.split(",")]
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except (ValueError, IndexError):
                continue
            for nme, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU baseline (oracle)
def cpu_train_baseline(workload, batch, steps, warmup, state_dict=None, device="cpu"):
    """The reference's arithmetic (CPU oracle port, torch fp32 on the host cores): full train iterations
    on a bounded batch.  Returns (img/s, cores, seconds per step)."""
    from oracle import glow_oracle as O
    shape, K, L, hidden, coupling, _, _ = WORKLOADS[workload]
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    if state_dict is None:
        import pytorch_glow_b200 as G
        from pytorch_glow_b200.hps import make_hps
        np.random.seed(2384); torch.manual_seed(2384)
        state_dict = G.Glow(make_hps(shape, K=K, L=L, hidden_channels=hidden, coupling=coupling, batch=batch)).state_dict()
    p = {k: v.detach().float().to(device).clone().requires_grad_(k != "h_top") for k, v in state_dict.items()}
    names = [k for k in p if k != "h_top"]
    ms = {k: torch.zeros_like(p[k]) for k in names}
    vs = {k: torch.zeros_like(p[k]) for k in names}
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(batch, shape[2], shape[0], shape[1], generator=g).to(device)
    times = []
    for t in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        noise = (torch.rand(x.shape, generator=g) / 256).to(device)
        for k in names:
            p[k].grad = None
        _, nll = O.glow_nll(x, noise, p, shape, K, L, "invconv", coupling)
        loss = O.generative_loss(nll)
        loss.backward()
        O.clip_grads_([p[k].grad for k in names], 5.0, 100.0)
        with torch.no_grad():
            for k in names:
                O.adam_step_(p[k], p[k].grad, ms[k], vs[k], t + 1, O.noam_lr(1e-3, t, 4000, 1e-4))
        if device != "cpu":
            torch.cuda.synchronize()
        if t >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, cores, sec


def _ref_hps(workload, batch, device):
    from pytorch_glow_b200.hps import make_hps
    shape, K, L, hidden, coupling, _, _ = WORKLOADS[workload]
    return make_hps(shape, K=K, L=L, hidden_channels=hidden, coupling=coupling, batch=batch, devices=(device,)), shape


def reference_cpu(workload, batch, steps, warmup, state_dict=None):
    """Train iterations of the reference on the host cores: the REAL reference when it is staged under
    baseline/_ref (kind "reference"), else the oracle port (kind "port").  -> (img/s, cores, s/step, kind)."""
    from baseline import ref_harness as R
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    if R.available():
        hps, shape = _ref_hps(workload, batch, "cpu")
        ips, sec, _ = R.time_train(hps, "cpu", batch, shape, steps, warmup)
        return ips, cores, sec, "reference"
    ips, cores, sec = cpu_train_baseline(workload, batch, steps, warmup, state_dict)
    return ips, cores, sec, "port"


def reference_gpu(workload, batch, steps, warmup, device="cuda:0"):
    """The reference's modules in torch eager (cuDNN / cuBLAS) on the same GPU -- the GPU bar SURVEY section 2 names --
    with TF32 off (the reference's default numerics) and on.  Falls back to the oracle port on the device."""
    from baseline import ref_harness as R
    out = {"per_gpu_batch": batch, "steps": steps, "warmup": warmup}
    if R.available():
        out["kind"] = "reference (baseline/_ref, torch eager, F3 patch)"
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            hps, shape = _ref_hps(workload, batch, device)
            ips, sec, loss = R.time_train(hps, device, batch, shape, steps, warmup, tf32=tf32)
            e2e, _, _ = R.time_train(hps, device, batch, shape, steps, 1, tf32=tf32, pinned_e2e=True) if tag == "tf32" else (None, None, None)
            out[tag] = {"train_img_s": ips, "ms_per_step": sec * 1e3, "loss_bits_per_dim": loss}
            if e2e:
                out[tag]["train_e2e_img_s"] = e2e
            torch.cuda.empty_cache()
        sb = min(256, 4 * batch)
        hps, _ = _ref_hps(workload, sb, device)
        sips, ssec = R.time_sample(hps, device, sb, max(2, steps), 1, tf32=True)
        out["tf32"]["sample_img_s"] = sips
        out["sample_batch"] = sb
        torch.cuda.empty_cache()
    else:
        out["kind"] = "port (oracle/glow_oracle.py on the device)"
        for tag, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            ips, _, sec = cpu_train_baseline(workload, batch, steps, warmup, None, device=device)
            out[tag] = {"train_img_s": ips, "ms_per_step": sec * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.impl == "reference-gpu":
        steps, warmup = max(2, min(args.steps, 5)), max(1, min(args.warmup, 2))
        torch.cuda.set_device(0)
        r = reference_gpu(args.workload, args.ref_gpu_batch, steps, warmup)
        best = r["tf32"]["train_img_s"]
        line = {
            "impl": "reference-gpu", "metric": "train_images_per_sec", "value": best, "unit": "img/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": r["tf32"]["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "per_gpu_batch": args.ref_gpu_batch,
                       "device": "cuda:0, torch eager (cuDNN)"},
            "gpu_eager_baseline": r,
            "e2e": {"value": r["tf32"].get("train_e2e_img_s", best), "unit": "img/s",
                    "h2d_bytes_per_step": args.ref_gpu_batch * 3 * 64 * 64 * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))   # bounded: ~4 s per CPU step
    ips, cores, sec, kind = reference_cpu(args.workload, args.cpu_batch, steps, warmup)
    sample = "%d train iterations (+%d warm-up) of %s on a batch of %d" % (
        steps, warmup, "the reference (baseline/_ref, F3 patch)" if kind == "reference" else "the oracle port", args.cpu_batch)
    line = {
        "impl": "reference", "metric": "train_images_per_sec", "value": ips, "unit": "img/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "per_gpu_batch": args.cpu_batch, "device": "host CPU"},
        "cpu_baseline": {"value": ips, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(w):
    shape, K, L, hidden, coupling, _, _ = WORKLOADS[w]
    return "%s: Glow %dx%dx%d K=%d L=%d hidden=%d %s coupling, invconv, train iteration" % (
        w, shape[2], shape[0], shape[1], K, L, hidden, coupling)


# ------------------------------------------------------------------------------------------ GPU arm
def timed(fn, steps, dist_on, device):
    """K calls of fn bracketed by barrier + synchronize on both sides; device time by CUDA events; max over ranks."""
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    w1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if dist_on:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / 1e3, w0, w1


def kernel_rooflines(device, B, shape, hidden, peaks):
    """Stand-alone timing of the level-1 kernels of one flow step against their roofline.

    bound "hbm": achieved = algorithmic bytes / time vs the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs);
    bound "tensor": achieved = 2*M*N*K / time vs the measured bf16 burst rate.  Buffers rotate over >= 3 copies
    (> 126 MB L2 in total) so that no launch finds its inputs in L2."""
    from pytorch_glow_b200 import _C
    from pytorch_glow_b200 import functional as KF
    hbm = peaks.get("hbm_gbs", 6500.0)
    tfl = peaks.get("bf16_tflops", 1590.0)
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    H, W = shape[0] // 2, shape[1] // 2
    C = shape[2] * 4
    M = B * H * W
    R = 3
    bf = torch.bfloat16
    rnd = lambda *s: torch.randn(*s, device=device)
    hs = [rnd(M, hidden).to(bf) for _ in range(R)]                  # h1 / h2 / d2 stand-ins (3 x 2*M*512 bytes)
    outs = [torch.empty(M, hidden, device=device, dtype=bf) for _ in range(R)]
    w2 = (rnd(hidden, hidden) * 0.05).to(bf)
    bias = torch.zeros(hidden, device=device); logs = torch.zeros(hidden, device=device)
    dlogs = torch.zeros(hidden, device=device); dbias = torch.zeros(hidden, device=device)
    dw = torch.zeros(hidden, hidden, device=device)
    n3p = (9 * C + 15) // 16 * 16
    p3 = [rnd(M, n3p) * 0.1 for _ in range(R)]
    zs = [rnd(M, C) for _ in range(R)]
    b3 = torch.zeros(C, device=device); l3 = torch.zeros(C, device=device)
    nblk = KF.rows_coupling_nblk(H * W, C)
    tickets = torch.zeros(B, dtype=torch.int32, device=device)
    parts = torch.empty(B * nblk, device=device)
    ld_in = torch.zeros(B, device=device)
    wmix = torch.eye(C, device=device) + 0.01 * rnd(C, C)
    mb = torch.zeros(C, device=device)

    def t_us(fn, reps=9):
        for i in range(R):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i % R)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    gemm_flops = 2.0 * M * hidden * hidden
    # fused coupling-net kernels (csrc/cnet_fused_sm100.cu) at level 1: conv1 K = 9*C/2 -> k1p, conv3 N = 9*C -> n3p
    cin = C // 2
    k1p = (9 * cin + 63) // 64 * 64
    k3p = (9 * C + 63) // 64 * 64
    w1 = (rnd(hidden, k1p) * 0.05).to(bf); w1[:, 9 * cin:] = 0
    w3 = (rnd(n3p, hidden) * 0.05).to(bf)
    w3t = (rnd(hidden, k3p) * 0.05).to(bf); w2t = (rnd(hidden, hidden) * 0.05).to(bf); w1t = (rnd(k1p, hidden) * 0.05).to(bf)
    dus = [rnd(M, C) * 0.5 for _ in range(2)]                        # gradient of Conv2dZeros' output (affine: C_out = C)
    fwd_flops = 2.0 * M * (9 * cin * hidden + hidden * hidden + hidden * 9 * C)
    bwd_flops = 2.0 * M * (k3p * hidden + hidden * hidden + hidden * k1p)
    fused = KF.cnet_fused_supported(False, k1p, hidden, n3p) and KF.cnet_fused_supported(True, k3p, hidden, k1p)
    masks = KF.cnet_relu_masks(M, device) if fused else None        # ReLU masks of h1 / h2 as bits (128 B per pixel)
    # (id, description, launch, algorithmic bytes, algorithmic flops or None)
    cases = []
    if fused:
        cases += [
            ("cnet_bwd", "cnet_chain_kernel<BWD>: dgrad3 -> ReLU'/ActNorm -> dgrad2 -> ReLU'/ActNorm -> dgrad1 fused (M=%d, "
             "K3=%d, hidden=%d, K1p=%d); reads du (in-kernel flipped im2col) + the two ReLU masks as bits, writes d3col, d2, d1, dA1" % (M, k3p, hidden, k1p),
             lambda i: KF.cnet_backward_implicit(dus[i % 2], B, H, W, C, k3p, w3t, w2t, w1t, hidden, k1p, logs, 3.0, logs, 3.0,
                                                 hs[i], hs[(i + 1) % R], dbias2=dbias, masks=masks),
             4.0 * M * C + 128.0 * M + 2.0 * M * (k3p + 2 * hidden + k1p), bwd_flops),
            ("cnet_fwd_train", "cnet_chain_kernel<FWD>, training: implicit conv1 -> conv2 -> conv3 fused, a1 / h1 / h2 stored "
             "once for the backward pass (+ their ReLU masks as bits) (M=%d)" % M,
             lambda i: KF.cnet_forward_implicit(zs[i], B, H, W, 0, cin, k1p, w1, w2, w3, hidden, n3p, bias, logs, 3.0, bias, logs,
                                                3.0, save=True, ones_col=9 * cin, masks=masks),
             M * (4.0 * cin + 4.0 * n3p + 2.0 * k1p + 4.0 * hidden + 128.0), fwd_flops),
            ("cnet_fwd_sample", "cnet_chain_kernel<FWD>, sampling: implicit conv1 -> conv2 -> conv3 fused, hidden activations "
             "never leave the SM (M=%d)" % M,
             lambda i: KF.cnet_forward_implicit(zs[i], B, H, W, 0, cin, k1p, w1, w2, w3, hidden, n3p, bias, logs, 3.0, bias, logs, 3.0),
             M * (4.0 * cin + 4.0 * n3p), fwd_flops),
        ]
    cases += [
        ("dgrad2", "gemm_tc_kernel<RELU_BWD,bf16>: dgrad of conv2 + ReLU/ActNorm backward epilogue (M=%d N=K=%d); levels the "
         "fused kernel does not serve (conv3 N > 256)" % (M, hidden),
         lambda i: KF.gemm(hs[i], w2, hidden, hidden, _C.EPI_RELU_BWD, None, logs, 3.0, y=hs[(i + 1) % R], dlogs=None,
                           dbias=None, out_dtype=_C.BF16, out=outs[i]),
         3.0 * M * hidden * 2, gemm_flops),
        ("wgrad2", "wgrad_tc_kernel: dW2 += d2^T h1 (P=%d, 512x512)" % M,
         lambda i: KF.gemm_wgrad(hs[i], hs[(i + 1) % R], hidden, hidden, dw),
         2.0 * M * hidden * 2, gemm_flops),
        ("conv2", "gemm_tc_kernel<ACTNORM_RELU,bf16>: conv2 1x1 + ActNorm + ReLU (M=%d N=K=%d)" % (M, hidden),
         lambda i: KF.gemm(hs[i], w2, hidden, hidden, _C.EPI_ACTNORM_RELU, bias, logs, 3.0, out_dtype=_C.BF16, out=outs[i]),
         2.0 * M * hidden * 2, gemm_flops),
        ("coupling", "rows_coupling_kernel: tap gather-sum + affine coupling + logdet (M=%d C=%d)" % (M, C),
         lambda i: KF.rows_coupling(p3[i], b3, l3, zs[i], B, H, W, True, False, 3.0, save_h=True, ld_in=ld_in,
                                    want_ld=True, an_logs=mb, logabsdet=ld_in[:1], partials=parts, tickets=tickets),
         4.0 * M * (9 * C + C // 2 * 2 + C), None),
        ("actnorm_mix", "rows_mix_kernel: ActNorm + invertible 1x1 conv (M=%d C=%d)" % (M, C),
         lambda i: KF.rows_actnorm_mix(zs[i], wmix, None, mb, mb, 3.0, False),
         8.0 * M * C, None),
    ]
    out = []
    for kid, name, fn, nbytes, flops in cases:
        us = t_us(fn)
        hbm_ach = nbytes / us / 1e3                      # GB/s
        ten_ach = flops / us / 1e6 if flops else None    # TFLOP/s
        hbm_frac = hbm_ach / hbm
        ten_frac = ten_ach / tfl if flops else None
        # the roofline that bounds the launch = the one whose floor (bytes / peak or flops / peak) is the longer time
        bound = "tensor" if (flops and flops / tfl / 1e6 > nbytes / hbm / 1e3) else "hbm"
        ach, peak, unit, frac = (ten_ach, tfl, "TFLOP/s", ten_frac) if bound == "tensor" else (hbm_ach, hbm, "GB/s", hbm_frac)
        out.append({"id": kid, "kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                    "frac": frac, "traffic": None, "us_per_launch": us, "algorithmic_bytes": nbytes,
                    "algorithmic_flops": flops, "hbm_frac": hbm_frac, "hbm_gbs": hbm_ach, "tensor_frac": ten_frac,
                    "tensor_tflops": ten_ach,
                    "peak_source": src + (" hbm_gbs (copy bandwidth)" if bound == "hbm" else " bf16_tflops (burst; kernel timed alone)")})
    return out


def run_b200(args):
    import torch.distributed as dist
    import pytorch_glow_b200 as G
    from pytorch_glow_b200 import _C, config
    from pytorch_glow_b200 import functional as KF
    from pytorch_glow_b200.hps import make_hps
    from pytorch_glow_b200.train import FusedTrainStep

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        dist.init_process_group("nccl", device_id=device)
    _C.lib()                                            # fail loudly if the extension is missing
    config.conv_dtype = args.conv_dtype

    shape, K, L, hidden, coupling, default_b, gflop = WORKLOADS[args.workload]
    B = args.batch or default_b
    if args.strong:
        if B % world:
            raise SystemExit("--strong: global batch %d is not divisible by %d ranks" % (B, world))
        B //= world
    seed = 2384                                         # profile/celeba.json:66
    np.random.seed(seed); torch.manual_seed(seed)       # identical initial replicas on every rank
    glow = G.Glow(make_hps(shape, K=K, L=L, hidden_channels=hidden, coupling=coupling, batch=B)).to(device)
    n_params = sum(p.numel() for n, p in glow.named_parameters() if n != "h_top")
    ts = FusedTrainStep(glow, use_graphs=not args.no_graphs, world_size=world)

    # synthetic images in [0,1) like ToTensor() output (train.py:41-45); distinct per rank and per step
    gen = torch.Generator().manual_seed(1234 + rank)
    nb = 4
    x_host = [torch.rand(B, shape[2], shape[0], shape[1], generator=gen).pin_memory() for _ in range(nb)]
    x_dev = [x.to(device) for x in x_host]
    x_stage = torch.empty_like(x_dev[0])

    ts.init_actnorm(x_dev[0])                           # trainer.py:112-115
    launches0 = _C.launch_count
    for i in range(max(args.warmup, 3)):
        ts.step(x_dev[i % nb])
    torch.cuda.synchronize()
    calls_per_step = None
    if not args.no_graphs:
        calls_per_step = getattr(ts, "captured_calls", None)

    if args.profile_step:
        rt = torch.cuda.cudart()
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        ts.step(x_dev[0])
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        return

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    losses = []

    def resident(i):
        losses.append(ts.step(x_dev[i % nb]))

    def end_to_end(i):
        x_stage.copy_(x_host[i % nb], non_blocking=True)        # H2D of this step's inputs (pinned)
        losses.append(float(ts.step(x_stage)))                   # D2H read of the step's loss

    for attempt in range(2):
        sampler.start()
        c0 = _C.launch_count
        if args.profiler_range:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        sec, w0, w1 = timed(resident, args.steps, dist_on, device)
        if args.profiler_range:
            torch.cuda.cudart().cudaProfilerStop()
        c1 = _C.launch_count
        clocks = sampler.stop(w0, w1)
        if not (set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}):
            break
        sampler = ClockSampler(sampler.idx)
    sec_e2e, _, _ = timed(end_to_end, args.steps, dist_on, device)
    last_loss = float(losses[-1])

    imgs = B * world * args.steps
    value, e2e = imgs / sec, imgs / sec_e2e
    if args.no_graphs:
        gpu_launches = c1 - c0
    else:
        gpu_launches = (ts.captured_calls if hasattr(ts, "captured_calls") else 0) * args.steps

    # ---- sampling (BASELINE config 4): reverse pass, eps_std 0.7, no collective
    sample = None
    if not args.no_sample:
        SB = args.sample_batch
        glow_s = G.Glow(make_hps(shape, K=K, L=L, hidden_channels=hidden, coupling=coupling, batch=SB))
        sd = {k: v for k, v in glow.state_dict().items() if k != "h_top"}
        glow_s.load_state_dict(sd, strict=False)
        glow_s.set_actnorm_inited()
        glow_s = glow_s.to(device).eval()

        from pytorch_glow_b200.train import GraphedSampler
        sampler_g = GraphedSampler(glow_s, eps_std=0.7)

        def sample_eager(i):
            with torch.no_grad():
                return glow_s(z=None, eps_std=0.7, reverse=True)

        def sample_fn(i):
            return sampler_g()
        for _ in range(2):
            sample_eager(0)
        nrep = max(2, args.steps // 2)
        esec, _, _ = timed(sample_eager, nrep, dist_on, device)
        for _ in range(2):
            sample_fn(0)
        ssec, _, _ = timed(sample_fn, nrep, dist_on, device)
        # end to end: every pass ends with the device->host copy of its images into pinned memory (what
        # infer.py / Inferer.sample hand to make_grid, network/inferer.py:53-60)
        host_img = torch.empty(SB, shape[2], shape[0], shape[1]).pin_memory()

        def sample_e2e(i):
            host_img.copy_(sampler_g(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        s2sec, _, _ = timed(sample_e2e, nrep, dist_on, device)
        sample = {"value": SB * world * nrep / ssec, "unit": "img/s", "per_gpu_batch": SB,
                  "eps_std": 0.7, "collective": "none", "cuda_graph": True,
                  "eager_value": SB * world * nrep / esec,
                  "e2e": {"value": SB * world * nrep / s2sec, "unit": "img/s", "h2d_bytes_per_step": 0,
                          "d2h_bytes_per_step": host_img.numel() * 4 * world}}
        del sampler_g
        del glow_s

    # ---- rooflines of the kernels that dominate the step (level 1: C=12, 32x32, M = B*1024 pixels), each timed
    # alone with CUDA events on the launch stream over rotating buffers larger than L2.  "roofline" is the kernel
    # with the largest share of the step in profiles/ (the fused backward chain of the coupling network);
    # "roofline_all" lists the others, each with BOTH its HBM and its tensor fraction.  Algorithmic bytes / flops
    # per launch: DESIGN.md section 4.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    roof, roof_all = None, None
    if rank == 0 and args.conv_dtype == "bf16":
        roof_all = kernel_rooflines(device, B, shape, hidden, peaks)
        roof = dict(roof_all[0])
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tr.get(roof["id"])
            if ent and ent.get("per_gpu_batch") == B:
                roof["traffic"] = ent["dram_bytes"]
                roof["traffic_source"] = ent.get("source")
        except (OSError, ValueError):
            pass

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload through the oracle port
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd_cpu = {k: v.detach().cpu() for k, v in glow.state_dict().items()}
        ips, cores, sec_cpu, kind = reference_cpu(args.workload, args.cpu_batch, 2, 1, sd_cpu)
        what = "the reference (baseline/_ref, F3 patch)" if kind == "reference" else "oracle/glow_oracle.py"
        cpu = {"value": ips, "unit": "img/s", "cores": cores, "kind": kind,
               "sample": "2 train iterations (+1 warm-up) of %s on a batch of %d (%.1f s/iteration)" % (what, args.cpu_batch, sec_cpu)}

    # ---- batch sweep (same model, other per-GPU batches; 25 = profile/celeba.json:27 num_batch_train 50 over its
    # two devices) and the torch-eager reference on this GPU.  The headline trainer is released first.
    sweep = None
    pts = args.sweep if args.sweep is not None else ("25,64,128,256" if (world == 1 and args.workload == "celeba64") else "none")
    if pts != "none":
        del ts
        glow = None
        torch.cuda.empty_cache()
        sweep = [{"per_gpu_batch": B, "value": value, "ms_per_step": sec / args.steps * 1e3}]
        for b in [int(v) for v in pts.split(",") if int(v) != B]:
            np.random.seed(seed); torch.manual_seed(seed)
            g2 = G.Glow(make_hps(shape, K=K, L=L, hidden_channels=hidden, coupling=coupling, batch=b)).to(device)
            t2 = FusedTrainStep(g2, use_graphs=not args.no_graphs, world_size=world)
            xb = [torch.rand(b, shape[2], shape[0], shape[1], generator=gen).to(device) for _ in range(2)]
            t2.init_actnorm(xb[0])
            for i in range(3):
                t2.step(xb[i % 2])
            n2 = max(5, args.steps)
            sec2, _, _ = timed(lambda i: t2.step(xb[i % 2]), n2, dist_on, device)
            sweep.append({"per_gpu_batch": b, "value": b * world * n2 / sec2, "ms_per_step": sec2 / n2 * 1e3})
            del t2, g2, xb
            torch.cuda.empty_cache()
        sweep.sort(key=lambda r: r["per_gpu_batch"])
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline and args.workload in ("celeba64", "cifar32", "tiny"):
        try:
            ts = None
            torch.cuda.empty_cache()
            gpu_eager = reference_gpu(args.workload, args.ref_gpu_batch, 3, 2, device="cuda:%d" % local_rank)
        except Exception as e:                        # a baseline must never take the headline down
            gpu_eager = {"error": repr(e)[:300]}

    if rank == 0:
        act_gb = None
        line = {
            "metric": "train_images_per_sec", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": args.conv_dtype, "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "global_batch": B * world, "per_gpu_batch": B,
                       "parallelism": "dp%d" % world, "cuda_graphs": not args.no_graphs,
                       "l2": "per-step working set (saved activations ~%.1f GB/GPU) >> 126 MB L2; 4 rotating input batches" % (B * 0.1),
                       "grad_allreduce_mb": n_params * 4 / 1e6 if world > 1 else 0,
                       "flow_state_dtype": "fp32", "conv_operands": args.conv_dtype},
            "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": x_host[0].numel() * 4 * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": sec_e2e / args.steps * 1e3},
            "gpu_launches": gpu_launches, "clocks": clocks, "loss_bits_per_dim": last_loss,
            "roofline": roof, "roofline_all": roof_all, "cpu_baseline": cpu, "sample": sample,
            "batch_sweep": sweep, "gpu_eager_baseline": gpu_eager,
        }
        if gpu_eager and "tf32" in gpu_eager:
            line["vs_gpu_eager_tf32"] = value / gpu_eager["tf32"]["train_img_s"]
            if sample and gpu_eager["tf32"].get("sample_img_s"):
                line["sample_vs_gpu_eager_tf32"] = sample["value"] / gpu_eager["tf32"]["sample_img_s"]
        if gflop:
            tf = value * gflop * 3 / 1e3          # fwd + dgrad + wgrad
            line["tensor_tflops_train"] = tf
            line["tensor_frac_of_sustained"] = tf / world / peaks.get("bf16_tflops_sustained", 1400.0)
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl in ("reference", "reference-gpu"):
        run_reference(a)
    else:
        run_b200(a)
