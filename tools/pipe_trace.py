#!/usr/bin/env python
"""hunt sporadic stalls of the stream-pool pipeline: many repetitions with per-call wall times; prints the slow repetitions"""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import pipe_probe  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402


def main():
    batch, _, _ = bench.build_workload(0, 1.0, os.cpu_count() or 8)
    gp = G.default_params(library_type=H.FR_FIRST)
    views = pipe_probe.views_of(batch, 4, False)
    ctxs = [G.Context(0) for _ in range(4)]
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    for rep in range(reps):
        work = list(range(len(views))) * 3
        nxt = [0]
        lock = threading.Lock()
        log = []

        def run(ci):
            ctx = ctxs[ci]
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(work):
                    return
                v, keep = views[work[i]]
                t = [time.perf_counter()]
                bt = ctx.upload(v, keepalive=keep); t.append(time.perf_counter())
                bt.evidence(gp); t.append(time.perf_counter())
                bt.fragments(); t.append(time.perf_counter())
                bt.graph(gp); t.append(time.perf_counter())
                bt.cluster(gp); t.append(time.perf_counter())
                bt.bridge(gp); t.append(time.perf_counter())
                bt.update(); t.append(time.perf_counter())
                bt.counts(); t.append(time.perf_counter())
                bt.free(); t.append(time.perf_counter())
                log.append((ci, i, np.diff(t) * 1e3))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=run, args=(k,)) for k in range(4)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for c in ctxs:
            c.sync()
        dt = (time.perf_counter() - t0) * 1e3 / 3
        slow = dt > 25
        print("rep %d: %.2f ms/step%s" % (rep, dt, "  <-- SLOW" if slow else ""), flush=True)
        if slow:
            for ci, i, d in sorted(log, key=lambda x: -x[2].sum())[:6]:
                print("   ctx %d item %d: upload %.1f evid %.1f frag %.1f graph %.1f clus %.1f brid %.1f upd %.1f counts %.1f free %.1f" % ((ci, i) + tuple(d)), flush=True)


if __name__ == "__main__":
    main()
