#!/usr/bin/env python
"""per-stage wall times of one batch (each ABI stage ends with a stream synchronisation), with and without per-kernel profiling"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    batch, _, _ = bench.build_workload(0, scale, os.cpu_count() or 8)
    gp = G.default_params(library_type=H.FR_FIRST)
    ctx = G.Context(0)
    bt = ctx.upload(batch.view(), keepalive=batch)
    for prof in (False, True, False):
        ctx.profile(prof)
        for rep in range(3):
            bt.reset()
            ctx.sync()
            t = [time.perf_counter()]
            for st in ("evidence", "fragments", "graph", "cluster", "bridge", "update"):
                if st in ("fragments", "update"):
                    getattr(bt, st)()
                else:
                    getattr(bt, st)(gp)
                ctx.sync()
                t.append(time.perf_counter())
            d = np.diff(t) * 1e3
        print("profiling=%s  " % prof + "  ".join("%s %.2f" % (n, x) for n, x in zip(("evid", "frag", "graph", "clus", "brid", "upd"), d)) + "  total %.2f ms" % d.sum(), flush=True)
        if prof:
            pr = ctx.profile_read()
            ks = sum(v[0] for v in pr.values()) / 3
            print("   kernel time per step %.2f ms over %d launches" % (ks, sum(v[1] for v in pr.values()) // 3))
            ctx.profile_reset()
    bt.free()
    ctx.close()


if __name__ == "__main__":
    main()
