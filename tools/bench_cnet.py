"""Fused coupling-net kernels vs the three-GEMM path they replace, timed alone with CUDA events (inputs > L2).

    python tools/bench_cnet.py [--m 524288] [--k1 64] [--n3 112] [--k3 128] [--iters 10]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_glow_b200 import _C  # noqa: E402
from pytorch_glow_b200 import functional as K  # noqa: E402

HID = 512


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def trace(label):
    import ctypes
    buf = (ctypes.c_ulonglong * 16)()
    _C.lib().glowk_debug_cnet_trace(buf)
    v = list(buf)
    tiles = max(1, v[9])
    names = ["a_full", "acc_empty1", "ring1", "h1_full", "acc_empty2", "ring2", "h2_full", "ring3", "MMA_total", "tiles",
             "prod_wait", "prod_total", "epi_acc1", "epi_acc2", "epi_c3", "epi_total"]
    print("  trace %s (cycles per tile, CTA 0, %d tiles): " % (label, tiles) +
          "  ".join("%s=%d" % (n, x // tiles) for n, x in zip(names, v) if n != "tiles"))


def timeline(label):
    import ctypes
    buf = (ctypes.c_ulonglong * 64)()
    _C.lib().glowk_debug_cnet_timeline(buf)
    v = list(buf)
    t0 = v[0]
    rel = lambda i: (v[i] - t0) if v[i] else -1
    print("  timeline %s (cycles from tile start)" % label)
    print("    MMA : a_full %d | GEMM1 free/issued %s | h1_full %d" % (rel(1), [(rel(2 + 2 * c), rel(3 + 2 * c)) for c in range(4)], rel(10)))
    print("          GEMM2 free/issued/g3 %s | c3 commit %d" % ([(rel(11 + 3 * c), rel(12 + 3 * c), rel(13 + 3 * c)) for c in range(4)], rel(24)))
    print("    EPI1: seen/released/published %s" % [(rel(32 + 3 * c), rel(33 + 3 * c), rel(34 + 3 * c)) for c in range(4)])
    print("    EPI2: seen/released/published %s" % [(rel(44 + 3 * c), rel(45 + 3 * c), rel(46 + 3 * c)) for c in range(4)])
    print("    EPI3: c3_full seen %d drained %d" % (rel(56), rel(57)))
    print("    EPI1 chunk 1: math done %d, tmem st done %d | EPI2 chunk 1: math done %d, buffer free %d, stored %d, fenced %d"
          % (rel(58), rel(59), rel(60), rel(61), rel(62), rel(63)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=524288)
    ap.add_argument("--k1", type=int, default=64)
    ap.add_argument("--n3", type=int, default=112)
    ap.add_argument("--k3", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--skip-bwd", action="store_true")
    ap.add_argument("--fused-only", action="store_true")
    a = ap.parse_args()
    m, k1, n3, k3 = a.m, a.k1, a.n3, a.k3
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    a1 = (rn(m, k1) * 0.5).bfloat16()
    w1, w2, w3 = (rn(HID, k1) * 0.05).bfloat16(), (rn(HID, HID) * 0.05).bfloat16(), (rn(n3, HID) * 0.05).bfloat16()
    b1, l1, b2, l2 = rn(HID) * 0.1, rn(HID) * 0.1, rn(HID) * 0.1, rn(HID) * 0.1
    flops_f = 2.0 * m * (k1 * HID + HID * HID + HID * n3)

    def three():
        h1 = K.gemm(a1, w1, HID, k1, _C.EPI_ACTNORM_RELU, b1, l1, 3.0, out_dtype=_C.BF16)
        h2 = K.gemm(h1, w2, HID, HID, _C.EPI_ACTNORM_RELU, b2, l2, 3.0, out_dtype=_C.BF16)
        return K.gemm(h2, w3, n3, HID, _C.EPI_STORE, out_dtype=_C.F32)

    t3 = 1.0 if a.fused_only else timeit(three, a.iters)
    print("fwd three GEMMs      : %8.1f us  %7.1f TFLOP/s" % (t3, flops_f / t3 * 1e-6))
    if K.cnet_fused_supported(False, k1, HID, n3):
        for save in (False, True):
            tf = timeit(lambda: K.cnet_forward(a1, w1, w2, w3, HID, n3, b1, l1, 3.0, b2, l2, 3.0, save=save), a.iters)
            print("fwd fused (save=%d)   : %8.1f us  %7.1f TFLOP/s  (x%.2f)" % (save, tf, flops_f / tf * 1e-6, t3 / tf))
            if os.environ.get("GLOWK_CNET_DEBUG"):
                trace("fwd save=%d" % save)
                if int(os.environ["GLOWK_CNET_DEBUG"]) & 16:
                    timeline("fwd save=%d" % save)
    else:
        print("fwd fused: shape not supported")
    if a.skip_bwd:
        return
    k1p = k1
    d3 = (rn(m, k3) * 0.5).bfloat16()
    w3t, w2t, w1t = (rn(HID, k3) * 0.05).bfloat16(), (rn(HID, HID) * 0.05).bfloat16(), (rn(k1p, HID) * 0.05).bfloat16()
    h2 = torch.relu(rn(m, HID)).bfloat16()
    h1 = torch.relu(rn(m, HID)).bfloat16()
    db2, db1 = torch.zeros(HID, device=dev), torch.zeros(HID, device=dev)
    flops_b = 2.0 * m * (k3 * HID + HID * HID + HID * k1p)

    def three_b():
        d2 = K.gemm(d3, w3t, HID, k3, _C.EPI_RELU_BWD, None, l2, 3.0, y=h2, dlogs=None, dbias=db2, out_dtype=_C.BF16)
        d1 = K.gemm(d2, w2t, HID, HID, _C.EPI_RELU_BWD, None, l1, 3.0, y=h1, dlogs=None, dbias=None, out_dtype=_C.BF16)
        return K.gemm(d1, w1t, k1p, HID, _C.EPI_STORE, out_dtype=_C.BF16)

    t3 = timeit(three_b, a.iters)
    print("bwd three GEMMs      : %8.1f us  %7.1f TFLOP/s" % (t3, flops_b / t3 * 1e-6))
    if K.cnet_fused_supported(True, k3, HID, k1p):
        tf = timeit(lambda: K.cnet_backward(d3, w3t, w2t, w1t, HID, k1p, l2, 3.0, l1, 3.0, h2, h1, dbias2=db2), a.iters)
        print("bwd fused (dbias2)   : %8.1f us  %7.1f TFLOP/s  (x%.2f)" % (tf, flops_b / tf * 1e-6, t3 / tf))
        tf = timeit(lambda: K.cnet_backward(d3, w3t, w2t, w1t, HID, k1p, l2, 3.0, l1, 3.0, h2, h1), a.iters)
        print("bwd fused (no dbias) : %8.1f us  %7.1f TFLOP/s  (x%.2f)" % (tf, flops_b / tf * 1e-6, t3 / tf))
        if os.environ.get("GLOWK_CNET_DEBUG"):
            trace("bwd")
            if int(os.environ["GLOWK_CNET_DEBUG"]) & 16:
                timeline("bwd")
        # the path the training step runs: operands gathered in-kernel, ReLU masks as bits or as the bf16 activations
        n_img, hh = m // 1024, 32
        if m % 1024 == 0 and k1 == 64 and k3 == 128:
            c = 12
            zz = rn(m, c)
            w1i = (rn(HID, 64) * 0.05).bfloat16(); w1i[:, 54:] = 0
            w2i, w3i = (rn(HID, HID) * 0.05).bfloat16(), (rn(112, HID) * 0.05).bfloat16()
            bz = torch.zeros(HID, device=dev)
            masks = K.cnet_relu_masks(m, dev)
            for label, mk in (("bf16 masks", None), ("bit masks ", masks)):
                tfw = timeit(lambda: K.cnet_forward_implicit(zz, n_img, hh, hh, 0, c // 2, 64, w1i, w2i, w3i, HID, 112, bz, l1, 3.0,
                                                             bz, l2, 3.0, save=True, ones_col=54, masks=mk), a.iters)
                du = rn(m, c) * 0.5
                w3ti = w3t.clone(); w3ti[:, 108:] = 0
                tbw = timeit(lambda: K.cnet_backward_implicit(du, n_img, hh, hh, c, 128, w3ti, w2t, w1t, HID, 64, l2, 3.0, l1, 3.0,
                                                              h2, h1, dbias2=db2, masks=mk), a.iters)
                print("implicit, %s : fwd(train) %8.1f us   bwd %8.1f us" % (label, tfw, tbw))
            if os.environ.get("GLOWK_CNET_DEBUG"):
                trace("bwd bit masks")
                if int(os.environ["GLOWK_CNET_DEBUG"]) & 16:
                    timeline("bwd bit masks")
    else:
        print("bwd fused: shape not supported")


if __name__ == "__main__":
    main()
