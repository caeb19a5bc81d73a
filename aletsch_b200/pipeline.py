"""Host-side pipelining of batches through the C ABI: the call a user of the path makes for a region's worth of
samples.  The reference enters bundle::bridge once per bundle from `-t N` pool threads (meta/incubator.cc:615-635);
here a few host threads each own an agpu_ctx (= one CUDA stream) and take whole batches (e.g. one sample's bundles
of a region) from a queue: upload from pinned host memory -> agpu_batch_bridge_all -> results back.  While one
stream computes, another stream's host->device copy is in flight on the copy engine, so the PCIe transfer of
batch i+1 hides behind the kernels of batch i.  The ABI is re-entrant across contexts."""
import os
import threading
import time

from . import gpu as G


class Pipeline:
    def __init__(self, device=0, n_streams=3, lib_path=None, blocking_sync=None, prefetch=True):
        """n_streams host threads.  prefetch: every thread owns a second context and queues the upload of its next batch there
        (agpu_upload_async) before it runs the stages of the current one, so its copy hides behind its own kernels too."""
        self.n_threads = max(1, n_streams)
        self.prefetch = prefetch
        # AGPU_PIPE_TRACE=1: wall times of every phase of every sub-batch (upload, stages, results, free) in self.trace
        self.trace = [] if os.environ.get("AGPU_PIPE_TRACE") else None
        self.ctxs = [G.Context(device, lib_path=lib_path) for _ in range(self.n_threads * (2 if prefetch else 1))]
        if prefetch:
            for c in self.ctxs:
                c.upload_async(True)
        if blocking_sync is None:
            # spinning waits are the fastest as long as every waiting thread has a core of its own; with one stream pool per
            # rank and several ranks per box they do not
            ranks = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
            cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            # (a rank bound to its GPU's NUMA node -- affinity.bind_to_device -- sees only that node's CPUs, shared with the
            # other ranks of the node; os.cpu_count() would still report the whole box)
            blocking_sync = ranks * (self.n_threads + 1) > (os.cpu_count() or 1) // 2 or (self.n_threads + 1) > cpus // 2
        for c in self.ctxs:
            c.blocking_sync(blocking_sync)

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    @property
    def launches(self):
        return sum(c.launches for c in self.ctxs)

    @property
    def syncs(self):
        return sum(c.syncs for c in self.ctxs)

    def sync(self):
        for c in self.ctxs:
            c.sync()

    def run(self, views, params, consume=None, resident=False, results=G.RESULT_ALL):
        """views: list of (agpu_batch_in with HOST pointers, keepalive).  Each batch goes through upload + bridge_all and its
        results come back to the host: by default agpu_batch_results(results) -- everything bundle::bridge leaves behind
        (mmap, hcst, frgs, fcst, splice graph, pereads clusters, bridge paths) packed on the device and copied into the context's
        pinned buffers -- handed to consume(i, batch, agpu_results) while the views are valid (default: keep the counters and the
        device -> host byte count).  results=0 skips the fetch (counters only).  Returns the per-batch results in input order.
        resident=True: the views hold DEVICE pointers (agpu_batch_adopt, no copy)."""
        out = [None] * len(views)
        nxt = [0]
        lock = threading.Lock()
        errs = []

        def claim():
            with lock:
                i = nxt[0]
                nxt[0] += 1
            return i if i < len(views) and not errs else None

        tr = self.trace

        def begin(ctx, i):
            v, keep = views[i]
            t0 = time.perf_counter()
            bt = ctx.adopt(v, keepalive=keep) if resident else ctx.upload(v, keepalive=keep)
            if tr is not None:
                tr.append((i, "upload", threading.get_ident(), t0, time.perf_counter()))
            return bt

        def work(mine):
            pending = cur = None
            try:
                i = claim()
                if i is None:
                    return
                turn = 0
                cur = (i, begin(mine[0], i))
                while cur is not None:
                    if len(mine) > 1:
                        # queue the next batch's upload on the other context, then run this one's stages
                        j = claim()
                        pending = (j, begin(mine[1 - turn], j)) if j is not None else None
                    i, bt = cur
                    try:
                        t0 = time.perf_counter()
                        bt.bridge_all(params)
                        t1 = time.perf_counter()
                        cnt = None if consume else bt.counts()
                        res = bt.results(results) if results else None
                        if tr is not None:
                            tr.append((i, "stages", threading.get_ident(), t0, t1))
                            tr.append((i, "results", threading.get_ident(), t1, time.perf_counter()))
                        if consume:
                            out[i] = consume(i, bt, res)
                        else:
                            out[i] = cnt
                            out[i]["d2h_bytes"] = int(res.bytes) if res is not None else 0
                    finally:
                        t0 = time.perf_counter()
                        bt.free()
                        if tr is not None:
                            tr.append((i, "free", threading.get_ident(), t0, time.perf_counter()))
                        cur = None
                    if len(mine) > 1:
                        cur, pending, turn = pending, None, 1 - turn
                    else:
                        j = claim()
                        cur = (j, begin(mine[0], j)) if j is not None else None
            except Exception as e:      # noqa: BLE001 -- re-raised on the caller's thread
                errs.append(e)
            finally:
                # nothing stays resident on an error path: the prefetched batch and a current one whose stages never ran
                for x in (pending, cur):
                    if x is not None:
                        try:
                            x[1].free()
                        except Exception:       # noqa: BLE001
                            pass

        per = 2 if self.prefetch and not resident else 1
        groups = [self.ctxs[k * per:(k + 1) * per] for k in range(self.n_threads)] if per == 2 else [[c] for c in self.ctxs[:self.n_threads]]
        ths = [threading.Thread(target=work, args=(g,)) for g in groups[:max(1, min(len(groups), len(views)))]]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]
        # every context keeps as much scratch as the hungriest one has needed so far: whichever sub-batch a thread takes
        # next, it will not have to grow its arena in the middle of the pipeline (idempotent once the sizes agree)
        # (best effort: a failed reservation must not lose the finished results)
        need = max(c.reserved for c in self.ctxs)
        for c in self.ctxs:
            if c.reserved < need:
                try:
                    c.reserve(need)
                except G.AgpuError:
                    break
        # the same for the pinned result buffers (growing one in the middle of a run synchronises the device)
        if results:
            try:
                for c in self.ctxs[1:]:
                    self.ctxs[0].pinned_match(c)
                for c in self.ctxs[1:]:
                    c.pinned_match(self.ctxs[0])
            except G.AgpuError:
                pass
        return out
