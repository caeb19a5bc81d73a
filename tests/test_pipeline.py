"""The stream-pool pipeline (aletsch_b200/pipeline.py) on the kernel-logic build: sub-batches through several host threads, with
and without the double-buffered asynchronous upload, lean and compact inputs -- every variant must return the counters of the
whole batch processed in one piece.  The -m gpu tier repeats it on the CUDA path."""
import numpy as np
import pytest

import parity
from aletsch_b200 import gpu as G
from aletsch_b200 import hostlib as H
from aletsch_b200.pipeline import Pipeline


def sub_views(batch, parts, compact):
    views = []
    for ch in batch.split(parts):
        ch.a["bundle_strand"] = np.ascontiguousarray(ch.a["strand"][np.minimum(ch.a["bundle_hit_off"][:-1], max(ch.n_hits - 1, 0))])
        if compact:
            arr = ch.compact()
            views.append((H.compact_struct(arr, ch.n_cigar), (arr, ch)))
        else:
            v = ch.view()
            v.rpos, v.flag, v.strand = None, None, None
            v.bundle_strand = ch.a["bundle_strand"].ctypes.data
            views.append((v, ch))
    return views


def run_variants(lib_path, n_threads, device=0):
    batch, lt = parity.make_batch(H.SYNTH_PAIRED, 60000, seed=20260131)
    gp = G.default_params(library_type=lt)
    ctx = G.Context(device, lib_path=lib_path)
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.bridge_all(gp)
    whole = bt.counts()
    bt.free()
    ctx.close()
    keys = ("hits", "segments", "chains", "junctions", "vertices", "edges", "fragments", "clusters", "bridged", "piers")
    for compact in (False, True):
        views = sub_views(batch, 5, compact)
        for prefetch in (False, True):
            pipe = Pipeline(device, n_streams=n_threads, lib_path=lib_path, prefetch=prefetch)
            for rep in range(2):
                res = pipe.run(views * 2, gp, results=G.RESULT_ALL if rep else 0)
                for k in keys:
                    assert sum(r[k] for r in res) == 2 * whole[k], (compact, prefetch, rep, k)
                # rep 1 brings every structure bundle::bridge leaves behind back to the host: at least the fragments (12 B each)
                # and the hit -> chain handles (4 B per hit) have to show up in the byte count
                d2h = sum(r["d2h_bytes"] for r in res)
                assert (d2h >= 2 * (12 * whole["fragments"] + 4 * whole["hits"])) if rep else d2h == 0, (compact, prefetch, rep, d2h)
            pipe.close()


def test_pipeline_variants_agree(emu_lib):
    # the kernel-logic build keeps a kernel's "shared memory" in thread-local statics, so host threads can run side by side
    run_variants(emu_lib, 3)


@pytest.mark.gpu
def test_pipeline_variants_agree_gpu():
    run_variants(None, 3)
