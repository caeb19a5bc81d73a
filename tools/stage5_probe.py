#!/usr/bin/env python
"""where the stage-5 leg (bundle_group::resolve over all region groups) spends its time: splice fetch, host re-packing, device pair
filter, host union-find (AGPU_STAGE5_TRACE prints the last two from inside agpu_group_resolve_batch).
usage: stage5_probe.py [--config 3]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("AGPU_STAGE5_TRACE", "1")
import bench  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    cfg = bench.config_of(argparse.Namespace(config=a.config, scale=None))
    batch, _ = bench.build_workload(cfg, 0, os.cpu_count() or 8)
    parts = bench.device_batches(batch)
    gp = bench.gpu_params(cfg, G)
    ctx = G.Context(0)
    bts = []
    for part in parts:
        dev = {f: torch.from_numpy(bench._tview(part.a[f])).to("cuda") for f in bench.FIELDS}
        b = H.BatchIn()
        b.n_bundles, b.n_hits, b.n_cigar = part.n_bundles, part.n_hits, part.n_cigar
        for f in bench.FIELDS:
            setattr(b, f, dev[f].data_ptr())
        x = ctx.adopt(b, keepalive=dev)
        x.evidence(gp)
        bts.append(x)
    gp5 = G.default_params(library_type=cfg["library_type"], **cfg["group"])
    groups = bench.region_groups(batch)
    order = np.concatenate(groups)
    group_off = np.zeros(len(groups) + 1, np.int32)
    np.cumsum([len(g) for g in groups], out=group_off[1:])
    for it in range(a.reps):
        if it == a.reps - 1:
            ctx.profile(True)
            ctx.profile_reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        offs, vals, base = [np.zeros(1, np.int64)], [], 0
        for x in bts:
            o, v = x.fetch_splices()
            offs.append(o[1:] + base)
            vals.append(v)
            base += int(o[-1]) if len(o) else 0
        t1 = time.perf_counter()
        off = np.concatenate(offs)
        val = np.concatenate(vals)
        loff, lval = G.reorder_lists(off, val, order)
        t2 = time.perf_counter()
        cl_of, ncl = G.group_resolve_arrays(ctx, group_off, loff, lval, gp5)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print("[stage5-probe] config %d: %d groups, %d bundles, %d splice values: fetch %.2f ms, host re-pack %.2f ms, agpu_group_resolve_batch %.2f ms, clusters %d"
              % (a.config, len(groups), batch.n_bundles, len(val), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), int(ncl.sum())), flush=True)


    for name, (ms, cnt) in sorted(ctx.profile_read().items(), key=lambda kv: -kv[1][0])[:14]:
        print("[stage5-probe] kernel %-24s %8.3f ms  %4d launches" % (name, ms, cnt))


if __name__ == "__main__":
    main()
