"""Where the captured training step's wall time goes between kernels: one graph replay under torch.profiler (CUPTI
kernel records), then busy time, idle gaps, and the kernels after which the longest / most frequent gaps occur.

    python tools/step_gaps.py [--batch 512]
"""
import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_glow_b200 as G  # noqa: E402
from pytorch_glow_b200.hps import make_hps  # noqa: E402
from pytorch_glow_b200.train import FusedTrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    np.random.seed(0); torch.manual_seed(0)
    glow = G.Glow(make_hps((64, 64, 3), K=32, L=3, hidden_channels=512, batch=a.batch)).to(dev)
    glow.flow.set_conv_dtype("bf16")
    ts = FusedTrainStep(glow, use_graphs=True)
    x = torch.rand(a.batch, 3, 64, 64, device=dev)
    ts.init_actnorm(x)
    for _ in range(4):
        ts.step(x)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts.step(x)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    if not ev:
        print("no CUDA kernel records"); return
    t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
    busy = 0.0
    cur_end = t0
    gaps = collections.defaultdict(lambda: [0, 0.0])
    for i, e in enumerate(ev):
        s, en = e.time_range.start, e.time_range.end
        if s > cur_end:
            prev = ev[i - 1].name.split("(")[0][-60:] if i else "-"
            g = gaps[prev + "  ->  " + e.name.split("(")[0][-50:]]
            g[0] += 1; g[1] += s - cur_end
        busy += max(0.0, en - max(s, cur_end))
        cur_end = max(cur_end, en)
    wall = t1 - t0
    print("kernels %d  wall %.2f ms  busy (union) %.2f ms  idle %.2f ms (%.1f %%)" % (len(ev), wall / 1e3, busy / 1e3, (wall - busy) / 1e3, 100 * (wall - busy) / wall))
    print("sum of kernel durations %.2f ms" % (sum(e.time_range.end - e.time_range.start for e in ev) / 1e3))
    for k, (n, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:14]:
        print("  %8.1f us in %4d gaps (%.1f us each)  %s" % (t, n, t / n, k))


if __name__ == "__main__":
    main()
