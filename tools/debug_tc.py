"""Diagnostic for the tcgen05 GEMMs: runs structured inputs and prints where results deviate.
Usage (GPU box): python tools/debug_tc.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_glow_b200 import _C  # noqa: E402
from pytorch_glow_b200 import functional as K  # noqa: E402

dev = "cuda:0"
print("has_tcgen05", _C.has_tcgen05(), torch.cuda.get_device_name(0))


def report(name, out, ref, blk=(32, 16)):
    err = (out.double() - ref.double()).abs()
    print("%s: shape %s max_err %.4g ref_max %.4g nan %d" % (name, tuple(out.shape), float(err.max()),
                                                             float(ref.abs().max()), int(torch.isnan(out).sum())))
    if float(err.max()) > 1e-2 * float(ref.abs().max()) + 1e-3:
        m, n = err.shape
        bm, bn = blk
        for i in range(0, min(m, 256), bm):
            print("  rows %4d: " % i + " ".join("%8.2g" % float(err[i:i + bm, j:j + bn].max()) for j in range(0, min(n, 256), bn)))
        bad = (err > 1e-2 * float(ref.abs().max()) + 1e-3).nonzero()
        print("  first bad:", bad[:8].tolist())
        i, j = bad[0].tolist()
        print("  out[%d,%d:%d+8]" % (i, j, j), out[i, j:j + 8].tolist())
        print("  ref[%d,%d:%d+8]" % (i, j, j), ref[i, j:j + 8].tolist())


def run_gemm(m, n, k, seed=0, pattern="rand"):
    g = torch.Generator().manual_seed(seed)
    if pattern == "rand":
        a = torch.randn(m, k, generator=g)
        b = torch.randn(n, k, generator=g)
    else:  # structured: a[i,kk] = i+1 if kk==0 ; b[j,kk] = j+1 if kk == 0  -> out[i,j] = (i+1)(j+1)
        a = torch.zeros(m, k)
        b = torch.zeros(n, k)
        a[:, 0] = (torch.arange(m) % 64 + 1).float()
        b[:, 0] = (torch.arange(n) % 64 + 1).float()
    a, b = a.bfloat16(), b.bfloat16()
    ref = a.double() @ b.double().t()
    try:
        out = K.gemm(a.to(dev), b.to(dev), n, k, _C.EPI_STORE, out_dtype=_C.F32, ldo=(n + 3) // 4 * 4)
        torch.cuda.synchronize()
        report("gemm %s m=%d n=%d k=%d" % (pattern, m, n, k), out[:, :n].cpu(), ref)
    except Exception as e:  # noqa: BLE001
        print("gemm m=%d n=%d k=%d FAILED: %s" % (m, n, k, e))
        raise


def run_wgrad(p, mo, no, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(p, mo, generator=g).bfloat16()
    b = torch.randn(p, no, generator=g).bfloat16()
    ref = a.double().t() @ b.double()
    dw = torch.zeros(mo, no, device=dev)
    K.gemm_wgrad(a.to(dev), b.to(dev), mo, no, dw)
    torch.cuda.synchronize()
    report("wgrad p=%d mo=%d no=%d" % (p, mo, no), dw.cpu(), ref)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "gemm"):
        run_gemm(128, 16, 64, pattern="struct")
        run_gemm(128, 64, 64)
        run_gemm(128, 64, 128)
        run_gemm(256, 256, 512)
        run_gemm(1000, 112, 512)
        run_gemm(4096, 512, 512)
    if which in ("all", "wgrad"):
        run_wgrad(64, 128, 64)
        run_wgrad(128, 128, 64)
        run_wgrad(1000, 512, 512)
