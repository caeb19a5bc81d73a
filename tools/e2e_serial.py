#!/usr/bin/env python
"""serial phase costs of the end-to-end path: ONE thread, one context, the configs[1] batch cut into N sub-batches; per
sub-batch upload -> sync -> bridge_all -> sync -> results -> free with wall-clock times of every phase (nothing overlaps).
Tells fixed per-sub-batch overhead from size-proportional work.  usage: e2e_serial.py [chunks ...]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402


def compact_views(batch, chunks):
    out = []
    for ch in batch.split(chunks):
        ch.a["bundle_strand"] = np.ascontiguousarray(ch.a["strand"][np.minimum(ch.a["bundle_hit_off"][:-1], max(ch.n_hits - 1, 0))])
        pin = {}
        for f, a in ch.compact().items():
            v = a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int16) if a.dtype == np.uint16 else a)
            pin[f] = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
        out.append((H.compact_struct(pin, ch.n_cigar, ptr=lambda t: t.data_ptr()), pin))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("chunks", type=int, nargs="*", default=[1, 4])
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    ns = argparse.Namespace(config=a.config, scale=None)
    cfg = bench.config_of(ns)
    batch, _ = bench.build_workload(cfg, 0, os.cpu_count() or 8)
    parts = bench.device_batches(batch)
    gp = bench.gpu_params(cfg, G)
    for chunks in a.chunks:
        views = []
        for part in parts:
            views.extend(compact_views(part, max(1, chunks // len(parts))))
        ctx = G.Context(0)
        acc = np.zeros(5)
        syncs = launches = 0
        for rep in range(a.reps + 2):
            if rep == 2:
                acc[:] = 0
                syncs, launches = ctx.syncs, ctx.launches
            for v, keep in views:
                t0 = time.perf_counter()
                bt = ctx.upload(v, keepalive=keep)
                ctx.sync()
                t1 = time.perf_counter()
                bt.bridge_all(gp)
                ctx.sync()
                t2 = time.perf_counter()
                bt.counts()
                t3 = time.perf_counter()
                res = bt.results(G.RESULT_ALL)
                t4 = time.perf_counter()
                bt.free()
                t5 = time.perf_counter()
                acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
        acc *= 1e3 / a.reps
        print("[serial] %2d sub-batches: per STEP upload %.2f  stages %.2f  counts %.2f  results %.2f  free %.2f  = %.2f ms;  %d launches, %d drains per step; D2H %.0f MB per sub-batch"
              % (len(views), acc[0], acc[1], acc[2], acc[3], acc[4], acc.sum(), (ctx.launches - launches) // a.reps, (ctx.syncs - syncs) // a.reps, res.bytes / 1e6), flush=True)
        ctx.close()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
