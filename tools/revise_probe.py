#!/usr/bin/env python
"""wall and per-kernel times of the boundary revision (agpu_batch_revise) and the phase set on the bench workload"""
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
    ctx = G.Context(0)
    bt = ctx.upload(batch.view(), keepalive=batch)
    for ratio in (2.0, 1.1):
        gp = G.default_params(library_type=H.FR_FIRST, min_boundary_log_ratio=ratio)
        bt.reset()
        bt.bridge_all(gp)
        bt.graph(gp)
        for prof in (False, True):
            ctx.profile(prof)
            ts = []
            for rep in range(4):
                ctx.sync()
                t0 = time.perf_counter()
                bt.revise(gp, fetch=False)
                ctx.sync()
                ts.append((time.perf_counter() - t0) * 1e3)
            print("ratio %.1f profiling=%s revise wall ms: %s" % (ratio, prof, " ".join("%.2f" % x for x in ts)), flush=True)
            if prof:
                pr = ctx.profile_read()
                for k, v in sorted(pr.items(), key=lambda kv: -kv[1][0]):
                    print("    %-28s %.3f ms/call x %d" % (k, v[0] / max(v[1], 1), v[1] // 4))
                ctx.profile_reset()
        rv = bt.revise(gp)
        print("   edges added %d, vertices marked %d, reserved %.2f GB" % (sum(len(r["rev_edge_d"]) for r in rv),
              sum(int((r["rev_vert"] > 0).sum()) for r in rv), ctx.reserved / 2 ** 30), flush=True)
    bt.free()
    ctx.close()


if __name__ == "__main__":
    main()
