#!/usr/bin/env python
"""the bench's end-to-end leg alone (Pipeline.run over compact sub-batches, bridge_all), a few passes; prints ms per step"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import pipe_stages  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402
from aletsch_b200.pipeline import Pipeline  # noqa: E402


def main():
    batch, _, _ = bench.build_workload(0, 1.0, os.cpu_count() or 8)
    gp = G.default_params(library_type=H.FR_FIRST)
    views = pipe_stages.compact_views(batch, 4)
    pipe = Pipeline(0, n_streams=4)
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s0 = pipe.syncs
        res = pipe.run(views * 3, gp)
        pipe.sync()
        dt = (time.perf_counter() - t0) / 3
        print("pass %d: %.2f ms/step, %d hits, %d bridged, %.1f drains/step" % (rep, dt * 1e3, sum(r["hits"] for r in res) // 3,
              sum(r["bridged"] for r in res) // 3, (pipe.syncs - s0) / 3), flush=True)
    pipe.close()


if __name__ == "__main__":
    main()
