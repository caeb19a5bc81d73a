#!/usr/bin/env python
"""bench.py -- throughput of the per-bundle read-evidence hot path (bundle::bridge, meta/bundle.cc:55-88)
on synthetic sorted-BAM records of BASELINE.json's shape.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's own C++ (oracle/_ref) on the host cores

Workload at N = 1: configs[1] of BASELINE.json -- 10 synthetic paired_end samples x 5M pairs on one 100 Mb
chromosome, bridging on.  For N > 1 every rank gets its own chromosome of that shape (weak scaling; bundles
are independent, there is no data-path collective).  A step = one pass of the hot path over the batch:
evidence (CIGAR -> coverage / chains) -> mate pairing -> splice graphs -> paired-read clusters -> bridging DP +
vote -> update_bridges.  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from aletsch_b200 import hostlib as H   # noqa: E402

SAMPLES = 10
PAIRS = 5_000_000
CHROM_LEN = 100_000_000
SEED = 20260101 + 1          # synth-v1, config index 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_workload(rank, scale, threads):
    """records of SAMPLES samples on this rank's chromosome, packed into one batch of bundles"""
    t0 = time.time()
    cfg = H.default_config(H.SYNTH_PAIRED, chrom_len=CHROM_LEN, seed=SEED + 1000 * rank)
    syn = H.Synth(cfg)
    pairs = max(1000, int(PAIRS * scale))
    recs = [None] * SAMPLES
    n_records = [0] * SAMPLES

    def work(k):
        recs[k] = syn.sample(k, pairs, threads=2)
        n_records[k] = recs[k]["n"]

    sem = threading.Semaphore(max(1, threads // 2))

    def guarded(k):
        with sem:
            work(k)

    ths = [threading.Thread(target=guarded, args=(k,)) for k in range(SAMPLES)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    t1 = time.time()
    batch = H.pack(recs, H.default_packer_params(H.FR_FIRST))
    t2 = time.time()
    log("[bench] rank %d: %d records generated in %.1fs, packed %d bundles / %d admitted hits in %.1fs" %
        (rank, sum(n_records), t1 - t0, batch.n_bundles, batch.n_hits, t2 - t1))
    return batch, sum(n_records), pairs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = "/tmp/agpu_clocks_%d_%d.csv" % (os.getpid(), index)

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        return out


def algorithmic_bytes(kernel, cnt, n_cigar, n_mblocks, n_splice_pairs):
    """bytes a kernel has to read + write once per STEP (all its launches of a step together),
    from the real counts of the batch"""
    Hh, S, F, C = cnt["hits"], cnt["segments"], cnt["fragments"], cnt["clusters"]
    J, V, E, NBD, M = cnt["junctions"], cnt["vertices"], cnt["edges"], cnt["borders"], cnt["cluster_members"]
    kernel = kernel.replace("(side)", "")
    if kernel == "k_hit_cigar":
        # in: pos, rpos, cigar_off (12 B/hit) + CIGAR ops; out: nspl, hash, bundle id (16 B/hit) + splice coordinates
        # + one read-modify-write of a 4-byte bitmap word per block end
        return Hh * (12 + 16) + 4 * n_cigar + 8 * n_splice_pairs + 16 * n_mblocks
    if kernel == "k_cov_add":
        # in: pos, cigar_off, bundle id (12 B/hit) + CIGAR ops; per block end a bitmap word, a rank word and a 4-byte RMW
        return Hh * 12 + 4 * n_cigar + 2 * n_mblocks * (4 + 4 + 8)
    if kernel == "k_covc_emit":
        return NBD * 8 + 12 * S
    if kernel == "k_qid_insert":
        return Hh * (8 + 4 + 8 + 4) + Hh * 16       # qid, bundle id in; slot index, next out; one 16-byte slot touch
    if kernel == "k_pair":
        return Hh * (8 + 4 + 16 + 4) + F * (2 * 12 + 8)   # slot index, bundle id, slot, next per hit; pos/mpos/isize of both mates, mate out
    if kernel == "k_hcst_insert":
        return Hh * (4 + 8 + 4 + 1 + 8) + 8 * n_splice_pairs
    if kernel == "k_frag_align":
        return F * (12 + 2 * (4 + 4 + 4) + 16 + 8 + 4)
    if kernel == "k_graph_build":
        # in: segments (12 B) and chain coordinates / counts; out: junctions (9 ints), partial exons (6 ints + 3 doubles),
        # vertices (6 ints + 3 doubles), edges (3 ints + 1 double), both adjacency tables
        return 12 * S + 4 * cnt["splice_ints"] + 12 * cnt["chains"] + 36 * J + 48 * (V - 2 * cnt.get("bundles", 0)) + 48 * V + 20 * E + 16 * E + 8 * V
    if kernel == "k_vote":
        # both passes: per cluster vp1, vp2, bundle, two chain ids, two bounds, pier (32 B) in, 8 result words (36 B) out, the pick
        # and two offsets (20 B) back in; the coordinates of the chosen chain / whole read and written once
        return C * (32 + 36 + 20) + 8 * cnt["bridge_whole_ints"] + 8 * cnt["bridge_chain_ints"]
    if kernel in ("k_group_partition", "k_group_partition_warp"):
        return M * (8 + 16 + 8)                     # h1 / h2, four keys, member + cluster flag out
    if kernel == "k_update":
        return 2 * (C * 24 + M * (4 + 8 + 8 + 2)) + M * 4 + cnt["bridged"] * 20
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from aletsch_b200 import gpu as G

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ncpu = os.cpu_count() or 8
    batch, n_records, pairs = build_workload(rank, args.scale, max(2, ncpu // max(1, world)))
    gp = G.default_params(library_type=H.FR_FIRST)

    stream = torch.cuda.current_stream()
    ctx = G.Context(local, stream=stream.cuda_stream)

    # pinned host copies (e2e path) and device-resident copies (kernel path)
    fields = ["bundle_hit_off", "bundle_tid", "bundle_sample", "pos", "rpos", "mpos", "isize", "flag", "strand", "xs", "qid", "cigar_off", "cigar"]
    dev = {}
    for f in fields:
        a = batch.a[f]
        v = a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else (a.view(np.int16) if a.dtype == np.uint16 else a))
        dev[f] = torch.from_numpy(v).to("cuda")
    torch.cuda.synchronize()

    def view_of(tensors):
        b = H.BatchIn()
        b.n_bundles, b.n_hits, b.n_cigar = batch.n_bundles, batch.n_hits, batch.n_cigar
        for f in fields:
            setattr(b, f, tensors[f].data_ptr())
        return b

    n_hits = batch.n_hits
    cig = batch.a["cigar"]
    ops = cig & 0xF
    n_mblocks = int(np.count_nonzero(ops == 0))
    n_cigar = int(batch.n_cigar)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: inputs already in HBM, a step = reset + bridge_all -------------------
    bt = ctx.adopt(view_of(dev), keepalive=dev)
    counts = None
    for _ in range(args.warmup):
        bt.reset()
        bt.bridge_all(gp)
    ctx.sync()
    counts = bt.counts()
    launches0 = ctx.launches
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        bt.reset()
        bt.bridge_all(gp)
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    # the same K steps again with CUDA events around every kernel launch (per-kernel durations for the roofline; the
    # event records cost ~2% so they stay out of the region `value` is taken from)
    ctx.profile(True)
    ctx.profile_reset()
    ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ep0.record(stream)
    for _ in range(args.steps):
        bt.reset()
        bt.bridge_all(gp)
    ep1.record(stream)
    barrier()
    ms_prof = ep0.elapsed_time(ep1)
    clk = clocks.stop()
    prof = ctx.profile_read()
    ctx.profile(False)
    counts = bt.counts()
    per_bundle = bt.bundle_counts()          # [NB, 4]: segments, fragments, clusters, bridged pairs
    stage5 = stage5_gpu(ctx, bt, batch, args) if not args.no_stage5 else None
    group_leg = group_bridge_gpu(ctx, bt, gp, stage5_gpu.clusters, args) if stage5 is not None else None
    phase_leg = phase_set_gpu(ctx, bt, gp, args) if stage5 is not None else None
    bt.free()

    # ---- end to end: pinned host buffers -> upload -> bridge_all -> counters back to the host ----------
    # the public call: aletsch_b200.pipeline.Pipeline.run over the batch cut into contiguous sub-batches, a few
    # host threads with one CUDA stream each, so the H2D copy of one sub-batch overlaps the kernels of another
    from aletsch_b200.pipeline import Pipeline
    torch.cuda.synchronize()
    del dev
    torch.cuda.empty_cache()
    chunks = batch.split(args.chunks)
    views = []
    h2d_bytes = 0
    for ch in chunks:
        pin = {}
        if args.upload == "compact":
            # the compact link format (agpu_batch_packed): 16-bit position deltas / mate offsets / insert sizes / CIGAR units
            # with escape lists, decoded on the device into the arrays of the plain upload (include/aletsch_gpu.h)
            ch.a["bundle_strand"] = np.ascontiguousarray(ch.a["strand"][np.minimum(ch.a["bundle_hit_off"][:-1], max(ch.n_hits - 1, 0))])
            arrays = ch.compact()
            for f, a in arrays.items():
                v = a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int16) if a.dtype == np.uint16 else a)
                pin[f] = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
                h2d_bytes += pin[f].numel() * pin[f].element_size()
            views.append((H.compact_struct(pin, ch.n_cigar, ptr=lambda t: t.data_ptr()), pin))
            continue
        # lean upload: flag[] is never read on the device, rpos[] is re-derived from the CIGAR there, and the strand that all
        # hits of a bundle share goes up once per bundle (agpu_batch_in: rpos / flag / strand may be NULL)
        ch.a["bundle_strand"] = np.ascontiguousarray(ch.a["strand"][np.minimum(ch.a["bundle_hit_off"][:-1], max(ch.n_hits - 1, 0))])
        lean = [f for f in fields if f not in ("rpos", "flag", "strand")] + ["bundle_strand"]
        for f in lean:
            a = ch.a[f]
            v = a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else (a.view(np.int16) if a.dtype == np.uint16 else a))
            pin[f] = torch.from_numpy(v).pin_memory()
            h2d_bytes += pin[f].numel() * pin[f].element_size()
        b = H.BatchIn()
        b.n_bundles, b.n_hits, b.n_cigar = ch.n_bundles, ch.n_hits, ch.n_cigar
        for f in lean:
            setattr(b, f, pin[f].data_ptr())
        views.append((b, pin))
    pipe = Pipeline(local, n_streams=args.streams, prefetch=not args.no_prefetch)
    for _ in range(min(args.warmup, 2)):
        pipe.run(views, gp)
    pipe.sync()
    barrier()
    launches_e2e0 = pipe.launches
    syncs_e2e0 = pipe.syncs
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    d2h_bytes = 0
    # all K steps' sub-batches go through the stream pool back to back (every step uploads its inputs again and reads its
    # counters back; there is no barrier between steps, the K steps are bracketed as a whole)
    res = pipe.run(views * args.steps, gp)
    d2h_bytes = 6 * 4 * batch.n_bundles
    pipe.sync()
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    launches_e2e = pipe.launches - launches_e2e0
    syncs_e2e = pipe.syncs - syncs_e2e0
    assert sum(r["hits"] for r in res) == n_hits * args.steps and sum(r["bridged"] for r in res) == counts["bridged"] * args.steps, "pipelined result differs"
    pipe.close()

    # max over ranks, totals over ranks
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_hits), float(counts["bridged"]), float(h2d_bytes), float(d2h_bytes)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    hits_all, bridged_all = float(tot[0]), float(tot[1])

    if rank == 0:
        steps = args.steps
        value = hits_all * steps / (ms_dev / 1e3)
        e2e_value = hits_all * steps / (ms_e2e / 1e3)
        # dominant kernel by accumulated device time
        total_k = sum(v[0] for v in prof.values())
        top = sorted(prof.items(), key=lambda kv: -kv[1][0])
        for name, (ms, cnt) in top[:14]:
            log("[bench] kernel %-22s %9.3f ms/step  (%d launches/step, %.1f%% of kernel time)" %
                (name, ms / steps, cnt // steps, 100 * ms / max(total_k, 1e-9)))
        # bridged pairs / time of stage 4 + update = the bridging kernels' accumulated time (rank 0's batch)
        stage4 = ("k_bridge_vertices", "k_piers", "k_cluster_pier", "k_group_cand", "k_bridge_job_counts", "k_bridge_job_fill", "k_bridge_dp",
                  "k_pier_bridges", "k_vote", "k_vote_type1", "k_vote_type2", "k_update", "k_fcst_insert", "k_merge_entries", "k_scatter_handle")
        stage4_ms = sum(ms for name, (ms, cnt) in prof.items() if name.replace("(side)", "") in stage4) / steps
        dom = None
        n_splice_pairs = int(np.count_nonzero(ops == 3))
        counts["bundles"] = batch.n_bundles
        # the dominant kernel: side-stream launches of the same kernel count with it
        merged = {}
        for name, (ms, cnt) in prof.items():
            m = merged.setdefault(name.replace("(side)", ""), [0.0, 0])
            m[0] += ms
            m[1] += cnt
        for name, (ms, cnt) in sorted(merged.items(), key=lambda kv: -kv[1][0]):
            ab = algorithmic_bytes(name, counts, n_cigar, n_mblocks, n_splice_pairs)
            if ab is not None:
                dom = (name, ms, cnt, ab)
                break
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this build (profiles/), if any
        traffic = {}
        try:
            import glob
            tf = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
            if tf:
                traffic = json.load(open(tf[-1]))
        except (OSError, ValueError):
            traffic = {}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        roof = None
        if dom:
            name, ms, cnt, ab = dom          # ab: bytes per step over all launches of the kernel in a step
            per_launch_ms = ms / cnt
            launches_per_step = cnt / steps
            achieved = ab / launches_per_step / (per_launch_ms / 1e3) / 1e9
            tk = traffic.get("kernels", {}).get(name)
            traffic_per_launch = (tk["dram_read_bytes"] + tk["dram_write_bytes"]) / tk["launches"] if tk else None
            roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "algorithmic_bytes_per_launch": ab / launches_per_step, "ms_per_launch": per_launch_ms,
                    "launches_per_step": launches_per_step,
                    "share_of_kernel_time": ms / max(total_k, 1e-9), "ms_per_step_profiled": ms_prof / steps, "traffic": traffic_per_launch,
                    "traffic_source": os.path.basename(tf[-1]) if (tk and tf) else None}
        out = {"metric": "hits_per_sec", "value": value, "unit": "hits/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
               "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
               "data": "synthetic", "impl": "ours",
               "config": {"workload": "configs[1]: %d synthetic paired_end samples x %d pairs, 1 chromosome of %d bp per GPU, bridging on"
                          % (SAMPLES, pairs, CHROM_LEN), "generator": "synth-v1 seed %d" % SEED,
                          "records_per_gpu": n_records, "admitted_hits_per_gpu": n_hits, "bundles_per_gpu": batch.n_bundles,
                          "l2": "inputs (%.1f GB) and scratch far larger than the 126 MB L2, no flush needed" % (h2d_bytes / 1e9), "scale": args.scale},
               "bridged_pairs_per_sec": bridged_all * steps / (ms_dev / 1e3), "bridged_pairs_per_step": bridged_all,
               "bridged_pairs_per_sec_stage4": bridged_all / max(stage4_ms / 1e3, 1e-12), "stage4_ms_per_step": stage4_ms,
               "counts": counts, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": "hits/s", "h2d_bytes_per_step": float(tot[2]), "d2h_bytes_per_step": float(tot[3]),
                       "ms_per_step": ms_e2e / steps, "upload": args.upload, "prefetch": not args.no_prefetch, "stream_drains_per_step": syncs_e2e / steps, "sub_batches": len(views), "streams": args.streams,
                       "gpu_launches_per_step": int(launches_e2e // steps)},
               "roofline": roof, "clocks": clk}
        if stage5 is not None:
            out["stage5"] = stage5
        if group_leg is not None:
            out["group_bridge"] = group_leg
        if phase_leg is not None:
            out["phase_set"] = phase_leg
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(batch, min(10, ncpu), args.cpu_seconds, gpu_bridged=per_bundle[:, 3])
            if stage5 is not None:
                out["stage5"]["cpu_baseline"] = stage5_cpu(batch, min(10, ncpu), max(2.0, args.cpu_seconds / 4))
            if group_leg is not None:
                out["group_bridge"]["cpu_baseline"] = group_bridge_cpu(batch, stage5_gpu.clusters, min(10, ncpu), max(2.0, args.cpu_seconds / 4))
        emit(out)
    if world > 1:
        dist.destroy_process_group()


REGION = 1_000_000      # region_partition_length (util/parameters.cc:42)


def region_groups(batch):
    """bundle groups as the reference forms them: all samples' bundles of one (chromosome, 1 Mb region, strand)
    (meta/incubator.cc:312-314, :461-471), members in (sample, bundle) order"""
    a = batch.a
    first = np.minimum(a["bundle_hit_off"][:-1], max(batch.n_hits - 1, 0))
    key = (a["bundle_tid"].astype(np.int64) << 40) | ((a["pos"][first].astype(np.int64) // REGION) << 8) | a["strand"][first].astype(np.int64)
    order = np.lexsort((np.arange(batch.n_bundles), a["bundle_sample"], key))
    cuts = np.nonzero(np.diff(key[order]))[0] + 1
    return [g for g in np.split(order, cuts) if len(g)]


def stage5_gpu(ctx, bt, batch, args):
    """bundle_group::resolve over every region group of the batch: splice signatures compacted on the device, all groups'
    pair counts in one launch (agpu_group_resolve_batch), size-capped union-find on the host; configs[1]: -c 20 -s 0.2"""
    stage5_gpu.clusters = []
    import torch
    from aletsch_b200 import gpu as G
    gp5 = G.default_params(library_type=H.FR_FIRST, max_group_size=20, min_grouping_similarity=0.2)
    groups = region_groups(batch)
    order = np.concatenate(groups) if groups else np.zeros(0, np.int64)
    group_off = np.zeros(len(groups) + 1, np.int32)
    np.cumsum([len(g) for g in groups], out=group_off[1:])
    clusters = 0
    t_all = []
    for it in range(1 + max(1, args.steps)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        off, val = bt.fetch_splices()
        loff, lval = G.reorder_lists(off, val, order)
        cl_of, ncl = G.group_resolve_arrays(ctx, group_off, loff, lval, gp5)
        torch.cuda.synchronize()
        if it > 0:
            t_all.append(time.perf_counter() - t0)
        clusters = int(ncl.sum())
    # the clusters with at least two bundles (input of assembler::bridge), members in gset order
    multi = []
    for gi, g in enumerate(groups):
        a, b = int(group_off[gi]), int(group_off[gi + 1])
        byc = {}
        for l in range(a, b):
            byc.setdefault(int(cl_of[l]), []).append(int(order[l]))
        multi.extend(v for _, v in sorted(byc.items()) if len(v) >= 2)
    stage5_gpu.clusters = multi
    dt = float(np.mean(t_all))
    pairs = int(sum(len(g) * (len(g) - 1) // 2 for g in groups))
    return {"bundle_groups": len(groups), "bundles": int(batch.n_bundles), "pairs": pairs, "clusters": int(clusters), "ms": dt * 1e3,
            "bundles_per_sec": batch.n_bundles / dt, "pairs_per_sec": pairs / dt, "params": "-c 20 -s 0.2",
            "timing": "host wall clock around splice fetch + agpu_group_resolve_batch (device pair counts + host union-find), synchronised"}


def group_bridge_gpu(ctx, bt, gp, clusters, args):
    """assembler::bridge (meta/assembler.cc:977-1018) over the clusters stage 5 found: combined bundles, their splice graphs,
    every member re-clustered / re-bridged / updated against them.  Timed after a fresh per-bundle pass, host wall clock
    (the call ends with a stream synchronisation)."""
    import torch
    if not clusters:
        return None
    t_all = []
    extra = 0
    for it in range(1 + max(1, min(args.steps, 3))):
        bt.reset()
        bt.bridge_all(gp)
        before = int(bt.bundle_counts()[:, 3].sum())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bt.group_bridge(clusters, gp)
        torch.cuda.synchronize()
        if it > 0:
            t_all.append(time.perf_counter() - t0)
        extra = int(bt.bundle_counts()[:, 3].sum()) - before
    dt = float(np.mean(t_all))
    members = int(sum(len(c) for c in clusters))
    return {"clusters": len(clusters), "member_bundles": members, "ms": dt * 1e3, "member_bundles_per_sec": members / dt,
            "bridged_pairs_added": extra}


def phase_set_gpu(ctx, bt, gp, args):
    """bundle_base::build_phase_set (rnacore/bundle_base.cc:338-418) for every bundle: graphs rebuilt from the bridged evidence,
    then the phasing paths; device part only (agpu_batch_phase_set ends with a stream synchronisation)"""
    import torch
    t_all = []
    for it in range(1 + max(1, min(args.steps, 3))):
        bt.reset()
        bt.bridge_all(gp)
        bt.graph(gp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bt._run("phase_set")
        torch.cuda.synchronize()
        if it > 0:
            t_all.append(time.perf_counter() - t0)
    ph = bt.phase_set()
    dt = float(np.mean(t_all))
    n_ph = int(sum(len(p["phase_cnt"]) for p in ph))
    n_el = int(sum(int(p["phase_cnt"].sum()) for p in ph))
    # the boundary revision of the same transform(bd, gr, true) call (identify_boundaries + remove_false_boundaries)
    t_rv = []
    for it in range(1 + max(1, min(args.steps, 3))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bt.revise(gp, fetch=False)
        torch.cuda.synchronize()
        if it > 0:
            t_rv.append(time.perf_counter() - t0)
    rv = bt.revise(gp)
    return {"ms": dt * 1e3, "distinct_phases": n_ph, "phase_elements": n_el, "phase_elements_per_sec": n_el / dt,
            "revise_ms": float(np.mean(t_rv)) * 1e3, "revise_edges_added": int(sum(len(r["rev_edge_d"]) for r in rv)),
            "revise_vertices_marked": int(sum(int((r["rev_vert"] > 0).sum()) for r in rv))}


def group_bridge_cpu(batch, clusters, threads, budget_s):
    """the reference's assembler::bridge on a bounded sample of the same clusters (per-bundle bridging done untimed first)"""
    import orclib
    kind = "reference"
    try:
        orclib.Checker("ref")
    except (OSError, FileNotFoundError):
        kind = "port"
    op = orclib.default_params(library_type=H.FR_FIRST)
    sizes = np.diff(batch.a["bundle_hit_off"])
    total = int(sum(int(sizes[k]) for c in clusters for k in c))
    target_hits = int(60_000 * threads * budget_s)
    step = max(1, int(np.ceil(total / max(target_hits, 1))))
    sample = clusters[::step]
    lock = threading.Lock()
    nxt = [0]
    spent = [0.0]
    added = [0]

    def worker():
        chk = orclib.Checker("ref" if kind == "reference" else "orc")
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(sample):
                return
            hs = [chk.new_bundle(batch.bundle(int(k)), op) for k in sample[i]]
            for h in hs:
                chk.run_quiet(h, "fragments")
                chk.run_quiet(h, "bridge")
            t0 = time.perf_counter()
            tot, _ = chk.group_bridge(hs)
            dt = time.perf_counter() - t0
            for h in hs:
                chk.free_bundle(h)
            with lock:
                spent[0] += dt
                added[0] += max(int(tot), 0)
    ths = [threading.Thread(target=worker) for _ in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    members = int(sum(len(c) for c in sample))
    wall = spent[0] / threads
    return {"kind": kind, "cores": threads, "sample": "every %d-th cluster: %d clusters, %d member bundles" % (step, len(sample), members),
            "member_bundles_per_sec": members / max(wall, 1e-9), "bridged_pairs_added": added[0],
            "note": "includes copying the checker's result arrays out (chk.group_bridge dumps the combined bundle and every member)"}


def stage5_cpu(batch, threads, budget_s):
    """the reference's bundle_group::resolve on a bounded sample of the same region groups (bundle objects built untimed)"""
    import orclib
    kind = "reference"
    try:
        orclib.Checker("ref")
    except (OSError, FileNotFoundError):
        kind = "port"
    op = orclib.default_params(library_type=H.FR_FIRST, max_group_size=20, min_grouping_similarity=0.2)
    groups = region_groups(batch)
    sizes = np.diff(batch.a["bundle_hit_off"])
    target_hits = int(60_000 * threads * budget_s)          # building the bundle objects dominates; bound that
    step = max(1, int(np.ceil(sizes.sum() / max(target_hits, 1))))
    sample = groups[::step]
    lock = threading.Lock()
    nxt = [0]
    spent = [0.0]

    def worker():
        chk = orclib.Checker("ref" if kind == "reference" else "orc")
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(sample):
                return
            hs = [chk.new_bundle(batch.bundle(int(k)), op) for k in sample[i]]
            t0 = time.perf_counter()
            chk.group_resolve(hs, op)
            dt = time.perf_counter() - t0
            for h in hs:
                chk.free_bundle(h)
            with lock:
                spent[0] += dt
    # bundle_group::resolve prints per-group statistics lines (meta/bundle_group.cc:360-393): keep them off our stdout
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        ths = [threading.Thread(target=worker) for _ in range(threads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    nbun = int(sum(len(g) for g in sample))
    pairs = int(sum(len(g) * (len(g) - 1) // 2 for g in sample))
    wall = spent[0] / threads            # thread-seconds inside bundle_group::resolve spread over the pool
    return {"kind": kind, "cores": threads, "sample": "every %d-th region group: %d groups, %d bundles" % (step, len(sample), nbun),
            "bundles_per_sec": nbun / max(wall, 1e-9), "pairs_per_sec": pairs / max(wall, 1e-9)}


def cpu_baseline(batch, threads, budget_s, all_cores=False, gpu_bridged=None):
    """the reference's own C++ (oracle/_ref) over a bounded sample of the same bundles, one bundle per task
    on a pool of `threads` workers (the granularity of aletsch -t N, meta/incubator.cc:615-635)"""
    import orclib
    kind = "reference"
    try:
        chk0 = orclib.Checker("ref")
    except (OSError, FileNotFoundError):
        chk0 = orclib.Checker("orc")
        kind = "port"
    del chk0
    op = orclib.default_params(library_type=H.FR_FIRST)
    sizes = np.diff(batch.a["bundle_hit_off"])
    # bounded sample: every k-th bundle until the estimated work fits the budget (~60k hits/s per core)
    target_hits = int(60_000 * threads * budget_s)
    step = max(1, int(np.ceil(sizes.sum() / max(target_hits, 1))))
    sample = list(range(0, batch.n_bundles, step))
    bundles = [batch.bundle(k) for k in sample]
    hits = int(sum(len(b["pos"]) for b in bundles))
    lock = threading.Lock()
    nxt = [0]
    bridged = [0]
    per = [0] * len(bundles)

    def worker():
        chk = orclib.Checker("ref" if kind == "reference" else "orc")
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(bundles):
                return
            h = chk.new_bundle(bundles[i], op)
            chk.run_quiet(h, "fragments")
            c = chk.run_quiet(h, "bridge")
            chk.free_bundle(h)
            per[i] = max(c, 0)
            with lock:
                bridged[0] += max(c, 0)

    t0 = time.time()
    ths = [threading.Thread(target=worker) for _ in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    out = {"value": hits / dt, "unit": "hits/s", "cores": threads, "kind": kind,
           "sample": "every %d-th bundle of the same batch: %d bundles, %d hits, %.1f s wall" % (step, len(bundles), hits, dt),
           "bridged_pairs_per_sec": bridged[0] / dt}
    if gpu_bridged is not None:
        # full-size cross-check: the bridged-pair count of every sampled bundle, CUDA path vs this CPU run
        bad = [int(k) for k, c in zip(sample, per) if int(gpu_bridged[k]) != int(c)]
        out["bridged_count_check"] = {"bundles": len(sample), "mismatches": len(bad), "first": bad[:5]}
    return out


def run_reference(args):
    """reference arm: the reference's own CPU implementation of the path on all host cores"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 8
    batch, n_records, pairs = build_workload(0, args.scale, ncpu)
    per_step = max(2.0, args.cpu_seconds / max(1, args.steps + args.warmup))
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(batch, ncpu, per_step)
        if i >= args.warmup:
            vals.append(r)
        last = r
    value = float(np.mean([v["value"] for v in vals]))
    bps = float(np.mean([v["bridged_pairs_per_sec"] for v in vals]))
    hits_sample = float(last["sample"].split("bundles, ")[1].split(" hits")[0])
    out = {"metric": "hits_per_sec", "value": value, "unit": "hits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * hits_sample / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
           "data": "synthetic", "impl": "reference",
           "config": {"workload": "configs[1]: %d synthetic paired_end samples x %d pairs, 1 chromosome of %d bp per GPU, bridging on"
                      % (SAMPLES, pairs, CHROM_LEN), "generator": "synth-v1 seed %d" % SEED, "scale": args.scale},
           "bridged_pairs_per_sec": bps,
           "cpu_baseline": {"value": value, "unit": "hits/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
           "e2e": {"value": value, "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_REAL_STDOUT = None


def quiet_stdout():
    """native libraries under us write to fd 1 (NCCL's version banner, the reference's statistics lines): send fd 1 to
    stderr for the whole run and keep the real stdout for the ONE JSON line"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the 5M pairs per sample (development only; 1.0 = the named config)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU work budget of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage5", action="store_true", help="skip the bundle_group::resolve (stage 5) leg")
    ap.add_argument("--chunks", type=int, default=4, help="sub-batches of the end-to-end pipeline")
    ap.add_argument("--no-prefetch", action="store_true", help="end-to-end leg without the double-buffered asynchronous upload")
    ap.add_argument("--upload", choices=["compact", "lean"], default="compact",
                    help="host->device format of the end-to-end leg: agpu_batch_packed (default) or the lean agpu_batch_in")
    ap.add_argument("--streams", type=int, default=4, help="host threads / CUDA streams of the end-to-end pipeline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
