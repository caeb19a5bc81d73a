#!/usr/bin/env python
"""does bulk host->device traffic slow the (device-resident) pipeline down?  resident sub-batches through the stream pool,
alone and next to a thread that keeps a pinned->device copy in flight on its own stream"""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import pipe_probe  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402
from aletsch_b200.pipeline import Pipeline  # noqa: E402


def main():
    batch, _, _ = bench.build_workload(0, 1.0, os.cpu_count() or 8)
    gp = G.default_params(library_type=H.FR_FIRST)
    views = pipe_probe.views_of(batch, 4, True)
    pipe = Pipeline(0, n_streams=4)
    pipe.run(views, gp, resident=True)
    src = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    for traffic, duty in ((False, 0), (True, 1.0), (True, 0.5), (False, 0)):
        stop = [False]
        moved = [0]

        def pump():
            with torch.cuda.stream(side):
                while not stop[0]:
                    t0 = time.perf_counter()
                    dst.copy_(src, non_blocking=True)
                    side.synchronize()
                    moved[0] += src.numel()
                    if duty < 1.0:
                        time.sleep((time.perf_counter() - t0) * (1 - duty) / duty)
        th = threading.Thread(target=pump) if traffic else None
        if th:
            th.start()
        steps = 6
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.run(views * steps, gp, resident=True)
        pipe.sync()
        dt = time.perf_counter() - t0
        stop[0] = True
        if th:
            th.join()
        print("background H2D %s (duty %.1f): %.2f ms/step, %.1f GB/s moved meanwhile" % (traffic, duty, dt / steps * 1e3, moved[0] / dt / 1e9), flush=True)
    pipe.close()


if __name__ == "__main__":
    main()
