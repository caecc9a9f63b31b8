#!/usr/bin/env python
"""Per-source-line summary of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
instructions executed, stall samples, shared/global excess.  usage: ncu_lines.py REP KERNEL_REGEX [TOP]"""
import csv
import subprocess
import sys


def main(rep, kern, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if len(r) > 8]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    isx, smp = ci["Instructions Executed"], ci["# Samples"]
    lines = []
    for r in rows[1:]:
        if r[0] == "Line No" or r[0] == "":
            continue
        try:
            lines.append((int(r[isx] or 0), int(r[smp] or 0), r[0], r[1].strip()[:110]))
        except ValueError:
            continue
    ti, ts = sum(l[0] for l in lines), sum(l[1] for l in lines)
    print("total warp-instructions %d, samples %d" % (ti, ts))
    print("--- by instructions executed")
    for l in sorted(lines, key=lambda l: -l[0])[:top]:
        print("%5.1f%% inst %5.1f%% smp  L%-4s %s" % (100.0 * l[0] / max(ti, 1), 100.0 * l[1] / max(ts, 1), l[2], l[3]))
    print("--- by stall samples")
    for l in sorted(lines, key=lambda l: -l[1])[:top]:
        print("%5.1f%% inst %5.1f%% smp  L%-4s %s" % (100.0 * l[0] / max(ti, 1), 100.0 * l[1] / max(ts, 1), l[2], l[3]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
