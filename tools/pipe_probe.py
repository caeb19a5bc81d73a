#!/usr/bin/env python
"""where does the pipelined end-to-end time go: the same sub-batches through the stream pool with host (pinned) inputs and with
device-resident inputs (no H2D), for a few pool shapes"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402
from aletsch_b200.pipeline import Pipeline  # noqa: E402

FIELDS = ["bundle_hit_off", "bundle_tid", "bundle_sample", "pos", "mpos", "isize", "xs", "qid", "cigar_off", "cigar", "bundle_strand"]


def views_of(batch, chunks, device):
    out = []
    for ch in batch.split(chunks):
        ch.a["bundle_strand"] = np.ascontiguousarray(ch.a["strand"][np.minimum(ch.a["bundle_hit_off"][:-1], max(ch.n_hits - 1, 0))])
        keep = {}
        b = H.BatchIn()
        b.n_bundles, b.n_hits, b.n_cigar = ch.n_bundles, ch.n_hits, ch.n_cigar
        for f in FIELDS:
            a = ch.a[f]
            v = a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else a)
            t = torch.from_numpy(v)
            keep[f] = t.to("cuda") if device else t.pin_memory()
            setattr(b, f, keep[f].data_ptr())
        out.append((b, keep))
    return out


def main():
    batch, _, _ = bench.build_workload(0, 1.0, os.cpu_count() or 8)
    gp = G.default_params(library_type=H.FR_FIRST)
    steps = 4
    for chunks, streams in ((1, 1), (4, 1), (4, 2), (4, 4), (8, 4)):
        for device in (True, False):
            views = views_of(batch, chunks, device)
            pipe = Pipeline(0, n_streams=streams)
            pipe.run(views, gp, resident=device)
            pipe.sync()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe.run(views * steps, gp, resident=device)
            pipe.sync()
            dt = (time.perf_counter() - t0) / steps
            print("chunks %d streams %d %-8s %.2f ms/step" % (chunks, streams, "resident" if device else "host", dt * 1e3), flush=True)
            pipe.close()
            del views
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
