#!/usr/bin/env python
"""small end-to-end run for compute-sanitizer: per-bundle pass, group pass, stage 5, fetches (synthetic + fuzz batches)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz  # noqa: E402
import parity  # noqa: E402
from aletsch_b200 import gpu as G, hostlib as H  # noqa: E402


def main():
    ctx = G.Context(0)
    batch, lt = parity.make_batch(H.SYNTH_PAIRED, 20000, samples=4)
    gp, _ = parity.params_pair(lt)
    for b in (batch, fuzz.random_batch(7, n_bundles=10, max_hits=400), fuzz.random_batch(8, n_bundles=6, max_hits=40, empty_every=3)):
        view = parity.lean_view(b) if b is batch else b.view()
        bt = ctx.upload(view, keepalive=b)
        bt.bridge_all(gp)
        ev = bt.fetch_evidence(b.a["bundle_hit_off"])
        bt.fetch_graph(); bt.fetch_fragments(); bt.fetch_clusters(ev); bt.fetch_bridge(bt.cluster_offsets())
        off, val = bt.fetch_splices()
        groups = parity.locus_groups(b) if b is batch else fuzz.strand_clusters(b, np.random.default_rng(1))
        if groups:
            bt.group_bridge(groups, gp)
            bt.fetch_group()
            bt.fetch_evidence(b.a["bundle_hit_off"])
        lists = [val[off[k]:off[k + 1]] for k in range(b.n_bundles)]
        G.group_resolve_batch(ctx, [lists[:40], lists[40:45], lists[45:]], gp)
        bt.reset()
        bt.bridge_all(gp)
        print("ok", b.n_bundles, bt.counts()["bridged"], flush=True)
        bt.free()
    ctx.close()


if __name__ == "__main__":
    main()
