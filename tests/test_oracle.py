"""CPU tier: pin the oracle.

The restatement (oracle/liboracle.so) is checked, array by array, against outputs of the REFERENCE ITSELF run here:
oracle/_ref/libaletsch_ref.so is the reference's own hot-path translation units compiled unchanged from
/root/reference against stand-in htslib / Boost.ICL headers (oracle/Makefile).  The reference ships no tests or
golden vectors for this path; the committed fixtures under tests/golden were generated
from the reference build by tests/golden/make_golden.py and pin the restatement where /root/reference is absent.
"""
import json
import os

import numpy as np
import pytest

import orclib
import parity
from aletsch_b200 import hostlib as H

ALL = ("bundle", "hits", "seg", "splices", "hcst_off", "hcst_val", "hcst_cnt", "hcst_grp", "hit_chain", "hit_chain_xs")
BRIDGE_INT = parity.INT_GRAPH + parity.INT_CLUSTER + parity.INT_BRIDGE + parity.INT_FCST + ("frgs", "frgs_clustered", "seg", "bridged")
BRIDGE_F64 = parity.F64_GRAPH + ("opt_score",)


def both(checkers):
    if "ref" not in checkers or "orc" not in checkers:
        pytest.skip("needs both oracle/_ref/libaletsch_ref.so and oracle/liboracle.so")
    return checkers["ref"], checkers["orc"]


@pytest.mark.parametrize("mode,templates,seed", [(H.SYNTH_PAIRED, 40000, 20260101), (H.SYNTH_PAIRED, 120000, 20260103),
                                                 (H.SYNTH_SINGLE, 20000, 20260104), (H.SYNTH_LONG, 3000, 20260105)])
def test_restatement_matches_reference_build(checkers, mode, templates, seed):
    ref, orc = both(checkers)
    batch, lt = parity.make_batch(mode, templates, seed=seed)
    _, op = parity.params_pair(lt)
    bad = []
    for k in range(batch.n_bundles):
        outs = []
        for chk in (ref, orc):
            h = chk.new_bundle(batch.bundle(k), op)
            _, ev = chk.run(h, "evidence")
            _, fr = chk.run(h, "fragments")
            cnt, br = chk.run(h, "bridge")
            chk.free_bundle(h)
            outs.append((ev, fr, br, cnt))
        (e0, f0, b0, c0), (e1, f1, b1, c1) = outs
        w = "bundle %d" % k
        for n in ALL:
            parity.cmp_int(n, e0[n], e1[n], w, bad)
        parity.cmp_int("frgs", f0["frgs"], f1["frgs"], w, bad)
        for n in BRIDGE_INT:
            parity.cmp_int(n, b0[n], b1[n], w, bad)
        for n in BRIDGE_F64:
            parity.cmp_f64(n, b0[n], b1[n], w, bad)
        assert c0 == c1
    assert not bad, bad[:5]


def test_group_bridge_and_resolve_match_reference_build(checkers):
    ref, orc = both(checkers)
    batch, lt = parity.make_batch(H.SYNTH_PAIRED, 30000, samples=4)
    _, op = parity.params_pair(lt, max_group_size=3)
    # bundles of different samples over the same locus: same side, overlapping extents
    a = batch.a
    loci = {}
    for k in range(batch.n_bundles):
        h0 = int(a["bundle_hit_off"][k])
        key = (int(a["bundle_side"][k]), int(a["pos"][h0]) // 50000)
        loci.setdefault(key, []).append(k)
    groups = [ks for ks in loci.values() if len(ks) >= 2][:12]
    assert groups
    bad = []
    for ks in groups:
        outs = []
        for chk in (ref, orc):
            hs = [chk.new_bundle(batch.bundle(k), op) for k in ks]
            for h in hs:
                chk.run(h, "fragments")
                chk.run(h, "bridge")
            tot, gb = chk.group_bridge(hs)
            _, gr = chk.group_resolve(hs, op)
            for h in hs:
                chk.free_bundle(h)
            outs.append((tot, gb, gr))
        (t0, g0, r0), (t1, g1, r1) = outs
        assert t0 == t1
        assert set(g0) == set(g1)
        for n in g0:
            if g0[n].dtype == np.float64:
                parity.cmp_f64(n, g0[n], g1[n], "group %s" % ks, bad)
            else:
                parity.cmp_int(n, g0[n], g1[n], "group %s" % ks, bad)
        for n in r0:
            parity.cmp_int(n, r0[n], r1[n], "group %s" % ks, bad)
    # cross-sample clustering of all bundles at once (bundle_group::resolve)
    outs = []
    for chk in (ref, orc):
        hs = [chk.new_bundle(batch.bundle(k), op) for k in range(min(batch.n_bundles, 120))]
        _, gr = chk.group_resolve(hs, op)
        for h in hs:
            chk.free_bundle(h)
        outs.append(gr)
    for n in outs[0]:
        parity.cmp_int(n, outs[0][n], outs[1][n], "resolve", bad)
    assert len(outs[0]["gvv_off"]) - 1 < min(batch.n_bundles, 120)      # something was grouped
    assert not bad, bad[:5]


def test_golden_fixtures_v2(checkers):
    """phase set and boundary revision fixtures (tests/golden/bundles_v2.json) pin whichever checker is present"""
    path = os.path.join(os.path.dirname(__file__), "golden", "bundles_v2.json")
    gold = json.load(open(path))
    batch, lt = parity.make_batch(gold["mode"], gold["templates"], seed=gold["seed"], chrom_len=gold["chrom_len"])
    _, op = parity.params_pair(lt, min_boundary_log_ratio=gold["min_boundary_log_ratio"])
    assert batch.n_bundles == gold["n_bundles"]
    for name, chk in checkers.items():
        for k, g in enumerate(gold["bundles"]):
            h = chk.new_bundle(batch.bundle(k), op)
            chk.run(h, "fragments")
            chk.run(h, "bridge")
            _, ph = chk.run(h, "phase")
            _, rv = chk.run(h, "revise")
            chk.free_bundle(h)
            for n, want in g["arrays"].items():
                got = ph[n] if n in ph else rv[n]
                if got.dtype == np.float64:
                    np.testing.assert_allclose(got, np.array(want), rtol=1e-9, atol=0, err_msg="%s bundle %d %s" % (name, k, n))
                else:
                    assert got.tolist() == want, "%s bundle %d %s" % (name, k, n)


def test_golden_fixtures(checkers):
    """fixtures generated from the reference build (tests/golden/make_golden.py) pin whichever checker is present"""
    path = os.path.join(os.path.dirname(__file__), "golden", "bundles_v1.json")
    if not os.path.exists(path):
        pytest.skip("no golden fixture")
    gold = json.load(open(path))
    batch, lt = parity.make_batch(gold["mode"], gold["templates"], seed=gold["seed"], chrom_len=gold["chrom_len"])
    _, op = parity.params_pair(lt)
    assert batch.n_bundles == gold["n_bundles"]
    for name, chk in checkers.items():
        for k, g in enumerate(gold["bundles"]):
            h = chk.new_bundle(batch.bundle(k), op)
            _, ev = chk.run(h, "evidence")
            chk.run(h, "fragments")
            cnt, br = chk.run(h, "bridge")
            chk.free_bundle(h)
            assert cnt == g["bridged"], (name, k)
            for n, want in g["arrays"].items():
                got = br[n] if n in br else ev[n]
                if got.dtype == np.float64:
                    np.testing.assert_allclose(got, np.array(want), rtol=1e-12, atol=0, err_msg="%s bundle %d %s" % (name, k, n))
                else:
                    assert got.tolist() == want, "%s bundle %d %s" % (name, k, n)
