"""Groundwork for the "next" row SURVEY 8f-3 (cross-sample support features, meta/assembler.cc:177-373): the reference-side
driver (oracle/ref_driver.cc: ref_group_support) runs the loops of assembler::assemble(vector<bundle*>) around the reference's own
member functions, and oracle/restate/support.cc restates them.  Nothing in the product implements this step yet.  The tests
pin down what an implementation has to respect: the per-edge bookkeeping invariants; that the step is ORDER DEPENDENT across
the members of a cluster (member k is assembled right after its own round); that the part of that dependence which comes
from extend_strands / group_start_boundaries / group_end_boundaries is restated exactly; and that the rest comes from scallop
decomposing member k's graph in place (it holds a reference), which no restatement short of scallop itself can follow."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
from aletsch_b200 import hostlib as H


def run_cluster(chk, batch, op, g):
    L = chk.lib
    L.ref_group_support.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    L.ref_bundle_set_sample.argtypes = [C.c_void_p, C.c_int]
    L.orc_bag_new.restype = C.c_void_p
    hs = []
    for k in g:
        h = chk.new_bundle(batch.bundle(k), op)
        chk.run(h, "fragments")
        chk.run(h, "bridge")
        L.ref_bundle_set_sample(h, int(batch.a["bundle_sample"][k]))
        hs.append(h)
    chk.group_bridge(hs)
    bag = L.orc_bag_new()
    arr = (C.c_void_p * len(hs))(*hs)
    rc = L.ref_group_support(arr, len(hs), bag)
    d = chk.bag_to_dict(bag)
    L.orc_bag_free(bag)
    for h in hs:
        chk.free_bundle(h)
    assert rc == 0
    return d


def test_reference_support_features(checkers):
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    chk = checkers["ref"]
    batch, lt = parity.make_batch(H.SYNTH_PAIRED, 40000, samples=4)
    _, op = parity.params_pair(lt)
    groups = parity.locus_groups(batch, max_groups=8)
    assert len(groups) >= 4
    multi = changed = 0
    for g in groups:
        os.environ.pop("ORC_SUPPORT_NO_ASSEMBLE", None)
        d = run_cluster(chk, batch, op, g)
        samples = sorted(set(int(batch.a["bundle_sample"][k]) for k in g))
        for pre in ["m%d_" % k for k in range(len(g))] + ["x_"]:
            e = d[pre + "sup_edge"].reshape(-1, 3)
            sizes = np.diff(d[pre + "sup_set_off"])
            assert np.array_equal(e[:, 2], sizes), pre                   # edge_info::count == |edge_info::samples|
            assert set(d[pre + "sup_set"].tolist()) <= set(samples + [-1]), pre
            assert np.all(d[pre + "sup_abd"] > 0), pre
            multi += int((e[:, 2] > 1).sum())
        os.environ["ORC_SUPPORT_NO_ASSEMBLE"] = "1"
        try:
            d2 = run_cluster(chk, batch, op, g)
        finally:
            os.environ.pop("ORC_SUPPORT_NO_ASSEMBLE", None)
        # member 0 is assembled last of nobody: its own round does not depend on the switch
        for n in ("m0_sup_edge", "m0_sup_abd", "m0_sup_loss"):
            assert d[n].shape == d2[n].shape and np.allclose(d[n], d2[n]), n
        if any(d[n].shape != d2[n].shape or not np.allclose(d[n], d2[n]) for n in d if n.startswith("m1_")):
            changed += 1
    assert multi > 20
    assert changed >= 1          # the order dependence is real: it is the rule on this data


def run_cluster_with(chk, prefix, batch, op, g):
    L = chk.lib
    fs = getattr(L, prefix + "_group_support")
    fs.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    ss = getattr(L, prefix + "_bundle_set_sample")
    ss.argtypes = [C.c_void_p, C.c_int]
    L.orc_bag_new.restype = C.c_void_p
    hs = []
    for k in g:
        h = chk.new_bundle(batch.bundle(k), op)
        chk.run(h, "fragments")
        chk.run(h, "bridge")
        ss(h, int(batch.a["bundle_sample"][k]))
        hs.append(h)
    chk.group_bridge(hs)
    bag = L.orc_bag_new()
    arr = (C.c_void_p * len(hs))(*hs)
    assert fs(arr, len(hs), bag) == 0
    d = chk.bag_to_dict(bag)
    L.orc_bag_free(bag)
    for h in hs:
        chk.free_bundle(h)
    return d


@pytest.mark.parametrize("mode,templates,samples", [(H.SYNTH_PAIRED, 60000, 4), (H.SYNTH_PAIRED, 60000, 6), (H.SYNTH_SINGLE, 30000, 4)])
def test_support_restatement_matches_reference_functions(checkers, mode, templates, samples):
    """every array of every member and of the combined graph, against the reference's own junction_support / start_end_support /
    non_splicing_support / boundary_extend / group_*_boundaries driven in the reference's order (scallop left out)"""
    if "ref" not in checkers or "orc" not in checkers:
        pytest.skip("needs both checkers")
    batch, lt = parity.make_batch(mode, templates, samples=samples)
    _, op = parity.params_pair(lt)
    groups = parity.locus_groups(batch, max_groups=30)
    assert len(groups) >= 5
    os.environ["ORC_SUPPORT_GROUP_ONLY"] = "1"
    try:
        edges = 0
        for g in groups:
            a = run_cluster_with(checkers["ref"], "ref", batch, op, g)
            b = run_cluster_with(checkers["orc"], "orc", batch, op, g)
            assert sorted(a) == sorted(b)
            for n in a:
                assert a[n].shape == b[n].shape, (g, n)
                if a[n].dtype == np.float64:
                    assert np.allclose(a[n], b[n], rtol=1e-9, atol=1e-12), (g, n)
                else:
                    assert np.array_equal(a[n], b[n]), (g, n)
            edges += len(a["x_sup_abd"])
        assert edges > 100
    finally:
        os.environ.pop("ORC_SUPPORT_GROUP_ONLY", None)
