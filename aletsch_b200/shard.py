"""Multi-GPU plumbing of the path: one process per GPU, torch.distributed.

Stages 1-4 are per bundle and need no collective: work units (regions, i.e. all bundles of one
(chromosome, 1 Mb region) key, meta/incubator.cc:357-381) are dealt to ranks by longest-processing-
time-first on their hit counts.  The only exchange of the path is for stage 5 when samples are
ingested on different ranks: the per-bundle splice lists (bundle::splices, the "junction
signatures") of a region group are all-gathered so that every rank runs the identical
bundle_group::resolve (meta/bundle_group.cc:26-56) and learns which bundles it has to receive.
Backend: NCCL over NVLink on GPUs; the CPU tests run the same code over gloo.
"""
import numpy as np
import torch
import torch.distributed as dist


def assign_units(weights, world):
    """LPT: units (descending weight, ties by index) go to the currently lightest rank (ties by rank).
    Returns rank_of[unit]; deterministic, identical on every rank."""
    w = np.asarray(weights, np.int64)
    order = sorted(range(len(w)), key=lambda i: (-int(w[i]), i))
    load = [0] * world
    rank_of = np.zeros(len(w), np.int32)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        rank_of[i] = r
        load[r] += int(w[i])
    return rank_of


def my_units(weights, world, rank):
    return np.nonzero(assign_units(weights, world) == rank)[0]


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_signatures(lists, keys=None):
    """all-gather of variable-length sorted int32 lists.

    lists: this rank's splice lists (one per local bundle); keys: optional int64 per list (e.g.
    sample << 32 | bundle index) carried along.  Returns (all_lists, owner_rank, all_keys) in rank order, the
    same on every rank.  Two collectives: the per-rank (n_lists, n_values) header, then one padded payload
    [lengths | keys | values] -- signatures are KBs per bundle, latency not bandwidth matters.
    """
    world = dist.get_world_size()
    dev = _device()
    n = len(lists)
    lens = np.array([len(x) for x in lists], np.int64)
    vals = np.concatenate([np.asarray(x, np.int64) for x in lists]) if n and lens.sum() else np.zeros(0, np.int64)
    k = np.asarray(keys, np.int64) if keys is not None else np.zeros(n, np.int64)
    head = torch.tensor([n, int(lens.sum())], dtype=torch.int64, device=dev)
    heads = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(heads, head)
    heads = [h.cpu().numpy() for h in heads]
    cap = max(int(2 * h[0] + h[1]) for h in heads)
    mine = torch.zeros(max(cap, 1), dtype=torch.int64, device=dev)
    payload = np.concatenate([lens, k, vals])
    if len(payload):
        mine[:len(payload)] = torch.from_numpy(payload).to(dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    all_lists, owner, all_keys = [], [], []
    for r in range(world):
        nr, nv = int(heads[r][0]), int(heads[r][1])
        p = parts[r].cpu().numpy()
        ln, kk, vv = p[:nr], p[nr:2 * nr], p[2 * nr:2 * nr + nv]
        off = np.concatenate([[0], np.cumsum(ln)])
        for i in range(nr):
            all_lists.append(vv[off[i]:off[i + 1]].astype(np.int32))
            owner.append(r)
            all_keys.append(int(kk[i]))
    return all_lists, np.array(owner, np.int32), np.array(all_keys, np.int64)


def resolve_groups(ctx, lists, params, keys=None):
    """bundle_group::resolve over the bundles of ALL ranks: gather the signatures, sort them into the canonical
    (key) order so that every rank sees the same gset order (gset order must be fixed), and run
    agpu_group_resolve on this rank's GPU.  Returns (groups as lists of (owner_rank, key)), identical on every rank."""
    from . import gpu as G
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        all_lists, owner, all_keys = gather_signatures(lists, keys)
    else:
        all_lists = [np.asarray(x, np.int32) for x in lists]
        owner = np.zeros(len(lists), np.int32)
        all_keys = np.asarray(keys, np.int64) if keys is not None else np.arange(len(lists), dtype=np.int64)
    order = sorted(range(len(all_lists)), key=lambda i: (int(all_keys[i]), int(owner[i]), i))
    groups = G.group_resolve(ctx, [all_lists[i] for i in order], params)
    return [[(int(owner[order[i]]), int(all_keys[order[i]])) for i in g] for g in groups]


REGION = 1_000_000      # region_partition_length (util/parameters.cc:42)


def gather_packed(off, val, keys, extra=None):
    """all-gather of packed variable-length int32 lists (vectorised: no per-list Python work).

    off[n+1] / val: this rank's lists; keys[n] (int64) and extra[n] (int64, optional) travel with them.  Two collectives: the
    per-rank (n_lists, n_values) header, then one padded int64 payload [lengths | keys | extra | values].  Returns
    (off_all, val_all, keys_all, extra_all, owner_rank) concatenated in rank order, identical on every rank, plus the bytes this
    rank received."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    off = np.asarray(off, np.int64)
    val = np.asarray(val, np.int32)
    n = len(off) - 1
    keys = np.asarray(keys, np.int64)
    extra = np.zeros(n, np.int64) if extra is None else np.asarray(extra, np.int64)
    lens = np.diff(off)
    if world == 1:
        return off, val, keys, extra, np.zeros(n, np.int32), 0
    dev = _device()
    head = torch.tensor([n, int(lens.sum())], dtype=torch.int64, device=dev)
    heads = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(heads, head)
    heads = [h.cpu().numpy() for h in heads]
    cap = max(int(3 * h[0] + h[1]) for h in heads)
    payload = np.concatenate([lens, keys, extra, val.astype(np.int64)])
    mine = torch.zeros(max(cap, 1), dtype=torch.int64, device=dev)
    if len(payload):
        mine[:len(payload)] = torch.from_numpy(payload).to(dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    lens_all, keys_all, extra_all, vals_all, owner = [], [], [], [], []
    for r in range(world):
        nr, nv = int(heads[r][0]), int(heads[r][1])
        p = parts[r].cpu().numpy()
        lens_all.append(p[:nr]); keys_all.append(p[nr:2 * nr]); extra_all.append(p[2 * nr:3 * nr])
        vals_all.append(p[3 * nr:3 * nr + nv].astype(np.int32))
        owner.append(np.full(nr, r, np.int32))
    lens_all = np.concatenate(lens_all)
    off_all = np.zeros(len(lens_all) + 1, np.int64)
    np.cumsum(lens_all, out=off_all[1:])
    return off_all, np.concatenate(vals_all), np.concatenate(keys_all), np.concatenate(extra_all), np.concatenate(owner), int(8 * cap * world)


def resolve_region_groups(ctx, off, val, keys, region_key, params):
    """bundle_group::resolve for a job whose SAMPLES are ingested on different ranks (meta/incubator.cc:461-471: a bundle group
    holds the bundles of ALL samples of one (chromosome, region, strand)).  Every rank contributes the splice signatures of its
    own bundles with their global keys (e.g. sample << 32 | bundle-in-sample) and region keys; after ONE all-gather every rank
    orders the bundles canonically -- (region key, key): the gset order must not depend on the rank count -- and runs the
    identical agpu_group_resolve_batch on its own GPU.  Returns (keys_sorted, region_sorted, cluster_of, owner_sorted,
    bytes_received): cluster_of[i] is the cluster of bundle keys_sorted[i] inside its region group."""
    from . import gpu as G
    off_all, val_all, keys_all, reg_all, owner, nbytes = gather_packed(off, val, keys, region_key)
    order = np.lexsort((keys_all, reg_all))
    loff, lval = G.reorder_lists(off_all, val_all, order)
    reg_s = reg_all[order]
    cuts = np.concatenate([[0], np.nonzero(np.diff(reg_s))[0] + 1, [len(reg_s)]]) if len(reg_s) else np.zeros(1, np.int64)
    group_off = np.asarray(cuts, np.int32)
    cl_of, _ = G.group_resolve_arrays(ctx, group_off, loff, lval, params)
    return keys_all[order], reg_s, cl_of[:len(order)], owner[order], nbytes


def bundle_region_keys(batch):
    """(chromosome, 1 Mb region, strand) key of every bundle of a packed host batch; for unstranded libraries the strand is the
    majority of the hits' XS tags (bundle_base::compute_strand, rnacore/bundle_base.cc:206-224)"""
    a = batch.a
    off = a["bundle_hit_off"]
    if batch.n_bundles == 0:
        return np.zeros(0, np.int64)
    first = np.minimum(off[:-1], max(batch.n_hits - 1, 0))
    strand = a["strand"][first].astype(np.int64)
    if batch.n_hits and np.any(strand == ord(".")):
        cp = np.concatenate([[0], np.cumsum(a["xs"] == ord("+"))])
        cm = np.concatenate([[0], np.cumsum(a["xs"] == ord("-"))])
        npl, nmi = cp[off[1:]] - cp[off[:-1]], cm[off[1:]] - cm[off[:-1]]
        strand = np.where(strand == ord("."), np.where(npl > nmi, ord("+"), np.where(npl < nmi, ord("-"), ord("."))), strand)
    return (a["bundle_tid"].astype(np.int64) << 40) | ((a["pos"][first].astype(np.int64) // REGION) << 8) | strand
