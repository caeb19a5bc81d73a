"""Drop-in check of the second half of integration/adapter.cc (the binding a maintainer adds next to meta/bundle.cc /
meta/assembler.cc): bundle_base::{mmap, splices, hcst, frgs, fcst}, vector<pereads_cluster> and vector<bridge_path> rebuilt from the
views of ONE agpu_batch_results call, compared field by field -- inside the reference build, on the reference's own types -- with
what build_fragments + bundle::bridge (meta/bundle.cc:55-88) leave on a fresh bundle (oracle/ref_driver.cc:
ref_adapter_compare_bundle).  CPU tier: kernel-logic build; -m gpu: the CUDA path."""
import ctypes as C

import pytest

import parity
from aletsch_b200 import gpu as G, hostlib as H


def compare_all(ctx, chk, mode, templates, samples=1):
    L = chk.lib
    L.ref_adapter_compare_bundle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64]
    batch, lt = parity.make_batch(mode, templates, samples=samples)
    gp, op = parity.params_pair(lt)
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.bridge_all(gp)
    res = bt.results(G.RESULT_ALL)
    assert res.bytes > 0
    bad = []
    clusters = 0
    for k in range(batch.n_bundles):
        h = chk.new_bundle(batch.bundle(k), op)
        nd = L.ref_adapter_compare_bundle(h, C.byref(res.evidence), C.byref(res.fragments), C.byref(res.clusters), C.byref(res.bridges), k,
                                          int(batch.a["bundle_hit_off"][k]))
        chk.free_bundle(h)
        if nd != 0:
            bad.append((k, nd))
        clusters += int(res.clusters.clu_off[k + 1] - res.clusters.clu_off[k])
    bt.free()
    assert not bad, "%d bundles differ, first: %s" % (len(bad), bad[:5])
    return batch.n_bundles, clusters


@pytest.mark.parametrize("mode,templates", [(H.SYNTH_PAIRED, 30000), (H.SYNTH_SINGLE, 15000), (H.SYNTH_LONG, 2000)])
def test_adapter_bundle_structures(emu_lib, checkers, mode, templates):
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    ctx = G.Context(0, lib_path=emu_lib)
    nb, nc = compare_all(ctx, checkers["ref"], mode, templates)
    ctx.close()
    assert nb > 5 and (nc > 100 or mode != H.SYNTH_PAIRED)


@pytest.mark.gpu
def test_adapter_bundle_structures_gpu(checkers):
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    ctx = G.Context(0)
    nb, nc = compare_all(ctx, checkers["ref"], H.SYNTH_PAIRED, 60000, samples=2)
    ctx.close()
    assert nb > 5 and nc > 100
