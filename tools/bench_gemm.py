#!/usr/bin/env python
"""Micro-benchmark + correctness check of the tcgen05 GEMMs at the coupling-network shapes.

    python tools/bench_gemm.py fwd 2x2       # one (kind, cluster) per process: a trap must not poison the rest
Kinds: fwd (conv2 512x512 ACTNORM_RELU), bwd (RELU_BWD), c1 (K=64), c3 (N=112 fp32 out), wgrad.
Prints one line per case: us/launch, TFLOP/s, GB/s of algorithmic operand+output bytes, max rel err.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pytorch_glow_b200 import _C  # noqa: E402
from pytorch_glow_b200 import functional as K  # noqa: E402

DEV = "cuda:0"


def timeit(fn, reps=10, nbuf=3):
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def run(kind, cluster, M=65536):
    torch.manual_seed(0)
    nb = 3
    hid = 512
    if kind in ("fwd", "bwd", "bwd3", "c1", "c3"):
        n, k = {"fwd": (512, 512), "bwd": (512, 512), "bwd3": (512, 128), "c1": (512, 64), "c3": (108, 512)}[kind]
        a = [(torch.randn(M, k, device=DEV) * 0.5).bfloat16() for _ in range(nb)]
        w = (torch.randn((n + 15) // 16 * 16, k, device=DEV) * 0.05).bfloat16()
        bias = torch.randn(n, device=DEV) * 0.1
        logs = torch.randn(n, device=DEV) * 0.05
        if kind in ("bwd", "bwd3"):
            y = [torch.randn(M, n, device=DEV).clamp_min(0).bfloat16() for _ in range(nb)]
            outs = [torch.empty(M, n, device=DEV, dtype=torch.bfloat16) for _ in range(nb)]
            dl, db = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
            nogy = bool(os.environ.get("NOGY"))       # dlogs recovered from dW (glowk_conv_actnorm_finish_batched)
            fn = lambda i: K.gemm(a[i], w, n, k, _C.EPI_RELU_BWD, None, logs, 3.0, y=y[i], dlogs=None if nogy else dl, dbias=None if os.environ.get("NOGB") else db,
                                  out_dtype=_C.BF16, out=outs[i], cluster=cluster)
            fn(0)
            acc = a[0].float() @ w[:n].float().t()
            g = torch.where(y[0].float() > 0, acc, torch.zeros_like(acc))
            ref = g * torch.exp(3 * logs)
            err = float((outs[0].float() - ref).abs().max() / ref.abs().max())
            e2 = float((dl - 3 * (g * y[0].float()).sum(0)).abs().max() / (3 * (g * y[0].float()).sum(0)).abs().max())
            e3 = float((db - torch.exp(3 * logs) * g.sum(0)).abs().max() / (torch.exp(3 * logs) * g.sum(0)).abs().max())
            err = max(err, 0.0 if nogy else e2, 0.0 if os.environ.get("NOGB") else e3)
            byts = M * k * 2 + 2 * M * n * 2
        else:
            odt = _C.F32 if kind == "c3" else _C.BF16
            ldo = (n + 15) // 16 * 16
            outs = [torch.empty(M, ldo, device=DEV, dtype=K.TORCH_DTYPE[odt]) for _ in range(nb)]
            epi = _C.EPI_STORE if kind == "c3" else _C.EPI_ACTNORM_RELU
            fn = lambda i: K.gemm(a[i], w, n, k, epi, bias, logs, 3.0, out_dtype=odt, ldo=ldo, out=outs[i], cluster=cluster)
            fn(0)
            acc = a[0].float() @ w[:n].float().t()
            ref = acc if kind == "c3" else ((acc + bias) * torch.exp(3 * logs)).clamp_min(0)
            err = float((outs[0][:, :n].float() - ref).abs().max() / ref.abs().max())
            byts = M * k * 2 + M * n * (4 if kind == "c3" else 2)
        flops = 2.0 * M * n * k
    else:  # wgrad: dW[512][512] += d[M][512]^T h[M][512]
        a = [(torch.randn(M, hid, device=DEV) * 0.3).bfloat16() for _ in range(nb)]
        b = [(torch.randn(M, hid, device=DEV) * 0.3).bfloat16() for _ in range(nb)]
        dw = torch.zeros(hid, hid, device=DEV)
        fn = lambda i: K.gemm_wgrad(a[i], b[i], hid, hid, dw)
        fn(0)
        ref = a[0].float().t() @ b[0].float()
        err = float((dw - ref).abs().max() / ref.abs().max())
        flops = 2.0 * M * hid * hid
        byts = 2 * M * hid * 2
    torch.cuda.synchronize()
    t = timeit(fn)
    print("%-5s cluster=%s M=%d: %8.1f us  %7.1f TFLOP/s  %7.1f GB/s(alg)  max rel err %.2e" % (
        kind, cluster, M, t * 1e6, flops / t / 1e12, byts / t / 1e9, err), flush=True)
    if int(os.environ.get("GLOWK_GEMM_DEBUG", "0")) & 64:
        import ctypes
        buf = (ctypes.c_ulonglong * 16)()
        _C.lib().glowk_debug_gemm_trace(ctypes.cast(buf, ctypes.c_void_p))
        v = list(buf)
        tiles = max(v[9], 1)
        print("      CTA0 cycles: producer wait-empty %d / total %d | MMA wait-operands %d, wait-accumulator %d / total %d | "
              "epilogue warp0 wait-acc %d, wait-y %d, wait-staging %d / total %d | tiles %d (%.0f cycles per tile)" % (
                  v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[4] / tiles), flush=True)


if __name__ == "__main__":
    kind = sys.argv[1]
    cl = tuple(int(v) for v in sys.argv[2].split("x")) if len(sys.argv) > 2 else None
    for M in ([65536, 16384, 4096] if len(sys.argv) <= 3 else [int(sys.argv[3])]):
        run(kind, cl, M)
