#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the text kept under profiles/: per kernel the duration, DRAM
bytes, throughput percentages, tensor-pipe utilisation, occupancy and the top stall reasons.
usage: ncu_summary.py REP [REP ...] > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("smsp__cycles_active.avg", None),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", None),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", None),
    ("sm__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", None),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read bytes"),
    ("smsp__inst_executed.avg.per_cycle_active", "IPC per SMSP"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main(reps):
    for rep in reps:
        hdr, units, rows = raw(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows:
            print("=" * 100)
            print("report: %s" % rep.split("/")[-1])
            print("kernel: %s" % r[ix["Kernel Name"]][:160])
            for key, label in WANT:
                for h in hdr:
                    if h == key or h.startswith(key + " ") or h.endswith("." + key):
                        print("  %-58s %14s %s" % (label or key, r[ix[h]], units[ix[h]]))
                        break
            tc = [h for h in hdr if "pipe_tensor" in h or "pipe_tc" in h or "utc" in h.lower()]
            for h in tc[:12]:
                if r[ix[h]] not in ("", "n/a"):
                    print("  %-58s %14s %s" % (h[-58:], r[ix[h]], units[ix[h]]))
            stalls = [(float(r[ix[h]] or 0), h) for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and r[ix[h]] not in ("", "n/a")]
            stalls = [(v, h) for v, h in stalls if "ratio" in h]
            for v, h in sorted(stalls, reverse=True)[:6]:
                print("  stall %-52s %14.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")[:52], v))


if __name__ == "__main__":
    main(sys.argv[1:])
