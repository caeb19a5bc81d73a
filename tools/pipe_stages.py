#!/usr/bin/env python
"""mean wall time of every phase of the stream-pool pipeline in steady state (compact or lean upload), per thread"""
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
    batch, _, _ = bench.build_workload(0, 1.0, os.cpu_count() or 8)
    gp = G.default_params(library_type=H.FR_FIRST)
    nthr = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    for mode in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("compact", "lean")):
        views = compact_views(batch, chunks) if mode == "compact" else pipe_probe.views_of(batch, chunks, False)
        ctxs = [G.Context(0) for _ in range(nthr)]
        for rep in range(3):
            steps = 6
            work = list(range(len(views))) * steps
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
                    log.append(np.diff(t) * 1e3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ths = [threading.Thread(target=run, args=(k,)) for k in range(nthr)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            for c in ctxs:
                c.sync()
            dt = (time.perf_counter() - t0) * 1e3 / steps
            m = np.mean(np.array(log), axis=0)
            print("%s rep %d: %.2f ms/step | per sub-batch: upload %.2f evid %.2f frag %.2f graph %.2f clus %.2f brid %.2f upd %.2f counts %.2f free %.2f = %.2f"
                  % ((mode, rep, dt) + tuple(m) + (m.sum(),)), flush=True)
        for c in ctxs:
            c.close()
        del views
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
