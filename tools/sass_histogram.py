#!/usr/bin/env python
"""Opcode histogram of the Blackwell-specific SASS in libglowk.so (whole library + per tensor-core kernel).

    python tools/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_glow_b200", "libglowk.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|STTM|LDTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UTCBAR|UTCCP|HMMA|HGMMA|FFMA2|HMNMX2|ELECT|LDGSTS)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    total = collections.Counter()
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = PAT.search(line)
        if m and cur:
            total[m.group(1)] += 1
            per[cur][m.group(1)] += 1
    print("# SASS opcode histogram of pytorch_glow_b200/libglowk.so (cuobjdump -sass; built by __graft_entry__.build() with")
    print("# nvcc -gencode arch=compute_100a,code=sm_100a).  tcgen05.mma -> UTCHMMA, tcgen05.ld / tcgen05.st -> LDTM / STTM,")
    print("# TMA -> UTMALDG / UTMASTG / UTMAREDG / UBLKCP, tcgen05.commit -> UTCBAR, fma.rn.f32x2 -> FFMA2, cp.async -> LDGSTS.")
    print("# No HMMA (mma.sync / wmma) and no HGMMA (wgmma) anywhere.")
    for k, v in total.most_common():
        print("%8d  %s" % (v, k))
    print("\n# kernels that issue tcgen05 / TMA instructions")
    demangle = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
    for (k, c), name in zip(per.items(), demangle):
        if c.get("UTCHMMA") or c.get("UTMALDG") or c.get("UBLKCP"):
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("glowk::", "")
            print("%-72s %s" % (name[:72], "  ".join("%s=%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    main()
