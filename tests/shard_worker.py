"""One rank of the sample-sharded run (tests/test_shard_nccl.py, tests/test_shard_gloo.py): launched by torchrun (NCCL, one GPU per
rank) or spawned over gloo on the kernel-logic build.  Every rank ingests ITS OWN samples (sample % world == rank), runs
bundle::bridge on its bundles, and the splice signatures are all-gathered for bundle_group::resolve (shard.resolve_region_groups).
Rank 0 also runs the whole batch on its own device and checks that (a) every rank's per-bundle results equal the single-device
ones and (b) the clusters every rank computed equal the single-device clusters."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(backend, lib_path=None, samples=6, templates=15000, out=None):
    import torch
    import torch.distributed as dist
    import parity
    from aletsch_b200 import gpu as G, hostlib as H, shard
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank)) if backend == "nccl" else 0
    if backend == "nccl":
        torch.cuda.set_device(local)
    batch, lt = parity.make_batch(H.SYNTH_PAIRED, templates, samples=samples, chrom_len=3_000_000, seed=20260407)
    gp, _ = parity.params_pair(lt, max_group_size=4, min_grouping_similarity=0.2)
    keys_all = (batch.a["bundle_sample"].astype(np.int64) << 32) | np.arange(batch.n_bundles, dtype=np.int64)
    mine = np.nonzero(batch.a["bundle_sample"] % world == rank)[0]
    sub = batch.select([int(k) for k in mine])
    ctx = G.Context(local, lib_path=lib_path)
    bt = ctx.upload(sub.view(), keepalive=sub)
    bt.bridge_all(gp)
    counts = bt.bundle_counts()
    off, val = bt.fetch_splices()
    bt.free()
    keys_s, reg_s, cl_of, owner_s, nbytes = shard.resolve_region_groups(ctx, off, val, keys_all[mine], shard.bundle_region_keys(sub), gp)
    # everything rank 0 needs for the comparison travels through the same process group
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "mine": mine.tolist(), "counts": counts.tolist(),
                                      "clusters": list(zip(keys_s.tolist(), reg_s.tolist(), cl_of.tolist(), owner_s.tolist()))})
    result = {"rank": rank, "world": world, "backend": backend, "ok": True, "bundles": int(batch.n_bundles), "bytes_received": nbytes}
    if rank == 0:
        full = ctx.upload(batch.view(), keepalive=batch)
        full.bridge_all(gp)
        want_counts = full.bundle_counts()
        foff, fval = full.fetch_splices()
        full.free()
        wk, wr, wc, _, _ = (lambda r: r)(shard_single(ctx, foff, fval, keys_all, shard.bundle_region_keys(batch), gp))
        want = list(zip(wk.tolist(), wr.tolist(), wc.tolist()))
        seen = set()
        for g in gathered:
            got = [(k, r, c) for k, r, c, _ in g["clusters"]]
            assert got == want, "rank %d resolved different clusters" % g["rank"]
            for k, c in zip(g["mine"], g["counts"]):
                assert c == want_counts[k].tolist(), "bundle %d differs when its sample is ingested on rank %d" % (k, g["rank"])
                seen.add(k)
            for k, r, c, o in g["clusters"]:
                assert o == (k >> 32) % world, "owner of bundle key %d" % k
        assert seen == set(range(batch.n_bundles))
        sizes = {}
        for k, r, c in want:
            sizes[(r, c)] = sizes.get((r, c), 0) + 1
        result["clusters"] = len(sizes)
        result["largest_cluster"] = max(sizes.values()) if sizes else 0
        assert result["largest_cluster"] >= 2, "the test batch must produce multi-bundle clusters"
    ctx.close()
    dist.barrier()
    if out and rank == 0:
        json.dump(result, open(out, "w"))
    return result


def shard_single(ctx, off, val, keys, region_key, params):
    """the same resolve without a process group (single device)"""
    from aletsch_b200 import gpu as G
    order = np.lexsort((keys, region_key))
    loff, lval = G.reorder_lists(np.asarray(off, np.int64), np.asarray(val, np.int32), order)
    reg_s = np.asarray(region_key)[order]
    cuts = np.concatenate([[0], np.nonzero(np.diff(reg_s))[0] + 1, [len(reg_s)]]) if len(reg_s) else np.zeros(1, np.int64)
    cl_of, _ = G.group_resolve_arrays(ctx, np.asarray(cuts, np.int32), loff, lval, params)
    return np.asarray(keys)[order], reg_s, cl_of[:len(order)], None, 0


if __name__ == "__main__":
    import torch.distributed as dist
    import torch
    backend = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    out = sys.argv[2] if len(sys.argv) > 2 else None
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    else:
        dist.init_process_group("gloo")
    r = run(backend, out=out)
    if dist.get_rank() == 0:
        print("shard_worker ok: %s" % json.dumps(r))
    dist.destroy_process_group()
