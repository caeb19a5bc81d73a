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
