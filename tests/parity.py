"""shared parity harness: run a packed batch through the C ABI (real CUDA library or, in the CPU
tier, the kernel-logic emulation build) and through a CPU checker, and compare every array."""
import numpy as np

from aletsch_b200 import gpu as G
from aletsch_b200 import hostlib as H
import orclib

INT_EVIDENCE = ("bundle", "seg", "splices", "hcst_off", "hcst_val", "hcst_cnt", "hcst_grp", "hit_chain")
INT_GRAPH = ("junc", "pexon", "vert", "edge", "graph")
F64_GRAPH = ("pexon_d", "vert_d", "edge_d")
REL_TOL = 1e-9      # BASELINE.json: double-valued scores within 1e-9 relative; integers bit-exact


def make_batch(mode=H.SYNTH_PAIRED, templates=20000, chrom_len=2_000_000, samples=1, seed=20260101, library_type=None, **cfgkw):
    cfg = H.default_config(mode, chrom_len=chrom_len, seed=seed, **cfgkw)
    s = H.Synth(cfg)
    recs = [s.sample(k, templates, threads=4) for k in range(samples)]
    if library_type is None:
        library_type = H.FR_FIRST if mode == H.SYNTH_PAIRED else H.UNSTRANDED
    return H.pack(recs, H.default_packer_params(library_type)), library_type


def params_pair(library_type, **kw):
    return G.default_params(library_type=library_type, **kw), orclib.default_params(library_type=library_type, **kw)


def cmp_int(name, a, b, where, out):
    if a.shape != b.shape or not np.array_equal(a, b):
        out.append("%s: %s differs (checker %s.. vs gpu %s.., len %d vs %d)" % (where, name, a[:8], b[:8], len(a), len(b)))


def cmp_f64(name, a, b, where, out):
    if a.shape != b.shape:
        out.append("%s: %s length differs (%d vs %d)" % (where, name, len(a), len(b)))
        return
    if len(a) == 0:
        return
    d = np.abs(a - b) / np.maximum(np.abs(a), 1e-300)
    if not np.all((d <= REL_TOL) | (a == b)):
        i = int(np.argmax(d))
        out.append("%s: %s differs at %d: %r vs %r" % (where, name, i, a[i], b[i]))


def compare_evidence_graph(ctx, batch, checker, gp, op):
    """stages 1-2 on the whole batch vs the checker bundle by bundle; returns list of mismatch strings"""
    bad = []
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.evidence(gp)
    ev = bt.fetch_evidence(batch.a["bundle_hit_off"])
    bt.graph(gp)
    gr = bt.fetch_graph()
    for k in range(batch.n_bundles):
        h = checker.new_bundle(batch.bundle(k), op)
        _, rev = checker.run(h, "evidence")
        _, rgr = checker.run(h, "graph")
        checker.free_bundle(h)
        w = "bundle %d" % k
        for n in INT_EVIDENCE:
            cmp_int(n, rev[n], ev[k][n], w, bad)
        for n in INT_GRAPH:
            cmp_int(n, rgr[n], gr[k][n], w, bad)
        for n in F64_GRAPH:
            cmp_f64(n, rgr[n], gr[k][n], w, bad)
    bt.free()
    return bad


INT_BRIDGE = ("opt", "opt_chain_off", "opt_chain_val", "opt_whole_off", "opt_whole_val")
INT_CLUSTER = ("clu_bounds", "clu_extend", "clu_count", "clu_c1_off", "clu_c1_val", "clu_c2_off", "clu_c2_val", "clu_fr_off", "clu_fr_val")
INT_FCST = ("fcst_off", "fcst_val", "fcst_cnt", "fcst_grp", "frg_chain")


def compare_full(ctx, batch, checker, gp, op, stats=None):
    """bundle::bridge (meta/bundle.cc:55-88) on the whole batch, every intermediate compared with the checker:
    evidence, fragments, graph, clusters, bridge choices, updated fragments / fcst / coverage, bridged count."""
    bad = []
    hit_off = batch.a["bundle_hit_off"]
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.evidence(gp)
    ev = bt.fetch_evidence(hit_off)
    bt.fragments()
    fr0 = bt.fetch_fragments()
    bt.graph(gp)
    gr = bt.fetch_graph()
    bt.cluster(gp)
    fr1 = bt.fetch_fragments()
    cl = bt.fetch_clusters(ev)
    bt.bridge(gp)
    br = bt.fetch_bridge(bt.cluster_offsets())
    bt.update()
    fr2 = bt.fetch_fragments()
    ev2 = bt.fetch_evidence(hit_off)
    if stats is not None:
        stats.update(bt.counts())
    for k in range(batch.n_bundles):
        h = checker.new_bundle(batch.bundle(k), op)
        _, rev = checker.run(h, "evidence")
        _, rfr = checker.run(h, "fragments")
        _, rbr = checker.run(h, "bridge")
        checker.free_bundle(h)
        w = "bundle %d" % k
        for n in INT_EVIDENCE:
            cmp_int(n, rev[n], ev[k][n], w, bad)
        cmp_int("frgs", rfr["frgs"], fr0[k]["frgs"], w + " build_fragments", bad)
        for n in INT_GRAPH:
            cmp_int(n, rbr[n], gr[k][n], w, bad)
        for n in F64_GRAPH:
            cmp_f64(n, rbr[n], gr[k][n], w, bad)
        cmp_int("frgs", rbr["frgs_clustered"], fr1[k]["frgs"], w + " group_pereads", bad)
        for n in INT_CLUSTER:
            cmp_int(n, rbr[n], cl[k][n], w, bad)
        for n in INT_BRIDGE:
            cmp_int(n, rbr[n], br[k][n], w, bad)
        cmp_f64("opt_score", rbr["opt_score"], br[k]["opt_score"], w, bad)
        cmp_int("frgs", rbr["frgs"], fr2[k]["frgs"], w + " update_bridges", bad)
        cmp_int("bridged", rbr["bridged"], fr2[k]["bridged"], w, bad)
        for n in INT_FCST:
            cmp_int(n, rbr[n], fr2[k][n], w, bad)
        cmp_int("seg", rbr["seg"], ev2[k]["seg"], w + " update_bridges", bad)
    bt.free()
    return bad


def compare_golden(ctx, golden_dir):
    """the device path against the COMMITTED fixtures (tests/golden, generated from the reference build by make_golden.py): no
    checker library involved.  v1: evidence, graph, clusters, bridging, update; v2: phase set and boundary revision."""
    import json
    import os
    g1 = json.load(open(os.path.join(golden_dir, "bundles_v1.json")))
    g2 = json.load(open(os.path.join(golden_dir, "bundles_v2.json")))
    batch, lt = make_batch(g1["mode"], g1["templates"], seed=g1["seed"], chrom_len=g1["chrom_len"])
    assert batch.n_bundles == g1["n_bundles"] == g2["n_bundles"]
    gp = G.default_params(library_type=lt)
    gp2 = G.default_params(library_type=lt, min_boundary_log_ratio=g2["min_boundary_log_ratio"])
    hit_off = batch.a["bundle_hit_off"]
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.evidence(gp)
    ev = bt.fetch_evidence(hit_off)
    bt.fragments()
    bt.graph(gp)
    gr = bt.fetch_graph()
    bt.cluster(gp)
    cl = bt.fetch_clusters(ev)
    bt.bridge(gp)
    br = bt.fetch_bridge(bt.cluster_offsets())
    bt.update()
    fr = bt.fetch_fragments()
    ev2 = bt.fetch_evidence(hit_off)
    bt.graph(gp2)
    ph = bt.phase_set()
    rv = bt.revise(gp2)
    bt.free()
    bad = []
    n_checked = 0
    for k in range(batch.n_bundles):
        w = "bundle %d" % k
        have = {}
        have.update(ev[k])
        have.update(gr[k])
        have.update(cl[k])
        have.update(br[k])
        have.update({n: fr[k][n] for n in ("frgs", "fcst_val", "fcst_cnt")})
        have["seg"] = ev2[k]["seg"]
        for n, want in g1["bundles"][k]["arrays"].items():
            want = np.array(want, np.float64 if have[n].dtype == np.float64 else np.int32)
            (cmp_f64 if have[n].dtype == np.float64 else cmp_int)(n, want, have[n], w, bad)
            n_checked += 1
        if int(fr[k]["bridged"][0]) != g1["bundles"][k]["bridged"]:
            bad.append("%s: bridged %d vs fixture %d" % (w, int(fr[k]["bridged"][0]), g1["bundles"][k]["bridged"]))
        have2 = dict(ph[k])
        have2.update(rv[k])
        for n, want in g2["bundles"][k]["arrays"].items():
            want = np.array(want, np.float64 if have2[n].dtype == np.float64 else np.int32)
            (cmp_f64 if have2[n].dtype == np.float64 else cmp_int)(n, want, have2[n], w, bad)
            n_checked += 1
    return bad, n_checked


def lean_view(batch):
    """agpu_batch_in without rpos / flag / per-hit strand (all optional): rpos is re-derived on the device, the strand goes
    up once per bundle"""
    a = batch.a
    if "bundle_strand" not in a:
        a["bundle_strand"] = np.ascontiguousarray(a["strand"][np.minimum(a["bundle_hit_off"][:-1], max(batch.n_hits - 1, 0))])
    v = batch.view()
    v.rpos = None
    v.flag = None
    v.strand = None
    v.bundle_strand = a["bundle_strand"].ctypes.data
    return v


def compare_lean_upload(ctx, batch, gp):
    """the lean upload must leave exactly the state of the full upload"""
    bad = []
    outs = []
    for view in (batch.view(), lean_view(batch)):
        bt = ctx.upload(view, keepalive=batch)
        bt.bridge_all(gp)
        ev = bt.fetch_evidence(batch.a["bundle_hit_off"])
        fr = bt.fetch_fragments()
        gr = bt.fetch_graph()
        outs.append((ev, fr, gr, bt.counts()))
        bt.free()
    (e0, f0, g0, c0), (e1, f1, g1, c1) = outs
    if c0 != c1:
        bad.append("counts differ: %s vs %s" % (c0, c1))
    for k in range(batch.n_bundles):
        for d0, d1 in ((e0[k], e1[k]), (f0[k], f1[k]), (g0[k], g1[k])):
            for n in d0:
                if not np.array_equal(d0[n], d1[n]):
                    bad.append("bundle %d: %s differs between full and lean upload" % (k, n))
    return bad


def compare_compact_upload(ctx, batch, gp, stats=None):
    """the compact upload (agpu_batch_packed, decoded on the device) must leave exactly the state of the full upload; also
    after a reset (the decoded arrays outlive the scratch the decoding used)"""
    bad = []
    arrays = batch.compact()
    cv = H.compact_struct(arrays, batch.n_cigar)
    if stats is not None:
        stats["bytes_full"] = sum(batch.a[f].nbytes for f in ("pos", "mpos", "isize", "xs", "qid", "cigar_off", "cigar"))
        stats["bytes_compact"] = sum(a.nbytes for n, a in arrays.items() if not n.startswith("bundle_"))
        stats["escapes"] = len(arrays["esc_pos_idx"]) + len(arrays["esc_mpos_idx"]) + len(arrays["esc_isize_idx"])
        stats["unit_escapes"] = len(arrays["esc_units_idx"])
        stats["default_cigars"] = int(((arrays["hit_meta"] & 63) == 62).sum())
        stats["long_ops"] = int(len(arrays["units"]) - (batch.n_cigar - stats["default_cigars"]))
    outs = []
    for view, keep, again in ((batch.view(), batch, False), (cv, (arrays, batch), False), (cv, (arrays, batch), True)):
        bt = ctx.upload(view, keepalive=keep)
        bt.bridge_all(gp)
        if again:
            bt.reset()
            bt.bridge_all(gp)
        ev = bt.fetch_evidence(batch.a["bundle_hit_off"])
        fr = bt.fetch_fragments()
        gr = bt.fetch_graph()
        outs.append((ev, fr, gr, bt.counts()))
        bt.free()
    for which, (e1, f1, g1, c1) in (("compact", outs[1]), ("compact after reset", outs[2])):
        e0, f0, g0, c0 = outs[0]
        if c0 != c1:
            bad.append("%s: counts differ: %s vs %s" % (which, c0, c1))
        for k in range(batch.n_bundles):
            for d0, d1 in ((e0[k], e1[k]), (f0[k], f1[k]), (g0[k], g1[k])):
                for n in d0:
                    if not np.array_equal(d0[n], d1[n]):
                        bad.append("bundle %d: %s differs between full and %s upload" % (k, n, which))
    return bad


def locus_groups(batch, width=50000, max_groups=12):
    """bundles of different samples over the same locus and side: candidate clusters for assembler::bridge"""
    a = batch.a
    loci = {}
    for k in range(batch.n_bundles):
        h0 = int(a["bundle_hit_off"][k])
        key = (int(a["bundle_side"][k]), int(a["pos"][h0]) // width)
        loci.setdefault(key, []).append(k)
    return [ks for ks in loci.values() if len(ks) >= 2][:max_groups]


CB_INT = ("combine_order", "cb_bundle", "cb_seg", "cb_hcst_off", "cb_hcst_val", "cb_hcst_cnt", "cb_hcst_grp", "cb_fcst_off", "cb_fcst_val",
          "cb_fcst_cnt", "cb_fcst_grp", "cb_junc", "cb_pexon", "cb_vert", "cb_edge")
CB_F64 = ("cb_pexon_d", "cb_vert_d", "cb_edge_d")


def compare_group_bridge(ctx, batch, checker, gp, op, groups, stats=None, first_round=True):
    """assembler::bridge (meta/assembler.cc:977-1018) on clusters of bundles: per-bundle bridge first (as assembler::resolve
    does), then the group pass; the combined bundles and every member's clusters / bridges / updated state are compared"""
    bad = []
    hit_off = batch.a["bundle_hit_off"]
    bt = ctx.upload(batch.view(), keepalive=batch)
    if first_round:
        bt.bridge_all(gp)
    else:
        bt.evidence(gp)            # no per-bundle bridging first: the group pass has every fragment to work on
        bt.fragments()
    before = bt.bundle_counts()[:, 3].copy()
    bt.group_bridge(groups, gp)
    cb = bt.fetch_group()
    ev = bt.fetch_evidence(hit_off)
    fr = bt.fetch_fragments()
    cl = bt.fetch_clusters(ev)
    br = bt.fetch_bridge(bt.cluster_offsets())
    after = bt.bundle_counts()[:, 3]
    if stats is not None:
        stats["group_bridged"] = int((after - before).sum())
    for gi, ks in enumerate(groups):
        hs = [checker.new_bundle(batch.bundle(k), op) for k in ks]
        for h in hs:
            checker.run(h, "fragments")
            if first_round:
                checker.run(h, "bridge")
        tot, ref = checker.group_bridge(hs)
        for h in hs:
            checker.free_bundle(h)
        w = "cluster %s" % ks
        for n in CB_INT:
            cmp_int(n, ref[n], cb[gi][n], w, bad)
        for n in CB_F64:
            cmp_f64(n, ref[n], cb[gi][n], w, bad)
        for j, k in enumerate(ks):
            pre = "b%d_" % j
            wk = "%s member %d (bundle %d)" % (w, j, k)
            for n in INT_CLUSTER:
                cmp_int(n, ref[pre + n], cl[k][n], wk, bad)
            for n in INT_BRIDGE:
                cmp_int(n, ref[pre + n], br[k][n], wk, bad)
            cmp_f64("opt_score", ref[pre + "opt_score"], br[k]["opt_score"], wk, bad)
            cmp_int("frgs", ref[pre + "frgs"], fr[k]["frgs"], wk, bad)
            for n in INT_FCST:
                cmp_int(n, ref[pre + n], fr[k][n], wk, bad)
            cmp_int("seg", ref[pre + "seg"], ev[k]["seg"], wk, bad)
            cmp_int("bridged", ref[pre + "bridged"], np.array([after[k] - before[k]], np.int32), wk, bad)
        if stats is not None:
            stats["ref_group_bridged"] = stats.get("ref_group_bridged", 0) + int(tot)
    bt.free()
    return bad


def compare_revise(ctx, batch, checker, gp, op, stats=None):
    """identify_boundaries + remove_false_boundaries (rnacore/graph_reviser.cc:1068-1377) after bundle::bridge, as
    assembler::assemble runs transform(bd, gr, true); also the whole revised graph = built edges + added ones"""
    bad = []
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.bridge_all(gp)
    bt.graph(gp)
    gr = bt.fetch_graph()
    rv = bt.revise(gp)
    bt.free()
    added = marked = 0
    for k in range(batch.n_bundles):
        h = checker.new_bundle(batch.bundle(k), op)
        checker.run(h, "fragments")
        checker.run(h, "bridge")
        n, ref = checker.run(h, "revise")
        checker.free_bundle(h)
        wk = "bundle %d" % k
        added += len(ref["rev_edge_d"])
        marked += int((ref["rev_vert"] > 0).sum())
        cmp_int("rev_edge", ref["rev_edge"], rv[k]["rev_edge"], wk, bad)
        cmp_f64("rev_edge_d", ref["rev_edge_d"], rv[k]["rev_edge_d"], wk, bad)
        cmp_int("rev_vert", ref["rev_vert"], rv[k]["rev_vert"], wk, bad)
        cmp_f64("rev_vert_d", ref["rev_vert_d"], rv[k]["rev_vert_d"], wk, bad)
        # the revised graph in out-edge order: alive built edges + added edges, sorted by (src, dst)
        st = np.concatenate([gr[k]["edge_ins"].reshape(-1, 3)[:, :2], rv[k]["rev_edge"].reshape(-1, 2)])
        w = np.concatenate([gr[k]["edge_ins_d"], rv[k]["rev_edge_d"]])
        o = np.lexsort((st[:, 1], st[:, 0]))
        cmp_int("rev_graph_edge", ref["rev_graph_edge"], st[o].reshape(-1).astype(np.int32), wk, bad)
        cmp_f64("rev_graph_edge_d", ref["rev_graph_edge_d"], w[o], wk, bad)
    if stats is not None:
        stats["rev_added"] = added
        stats["rev_marked"] = marked
    return bad


def compare_phase_set(ctx, batch, checker, gp, op, stats=None):
    """bundle_base::build_phase_set (rnacore/bundle_base.cc:338-418) after bundle::bridge, against the bundles' own splice
    graphs rebuilt from the updated evidence (transform(bd, gr, false))"""
    bad = []
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.bridge_all(gp)
    bt.graph(gp)
    ph = bt.phase_set()
    bt.free()
    total = 0
    for k in range(batch.n_bundles):
        h = checker.new_bundle(batch.bundle(k), op)
        checker.run(h, "fragments")
        checker.run(h, "bridge")
        n, ref = checker.run(h, "phase")
        checker.free_bundle(h)
        total += int(ref["phase_cnt"].sum())
        for name in ("phase_off", "phase_val", "phase_cnt"):
            cmp_int(name, ref[name], ph[k][name], "bundle %d" % k, bad)
    if stats is not None:
        stats["phase_count"] = total
    return bad


def checker_group_support(chk, batch, op, g):
    """the checker's side of the cross-sample support features for one cluster: per-bundle bridging, the group pass, then
    <prefix>_group_support (reference build: its own member functions in the reference's order, scallop left out --
    ORC_SUPPORT_GROUP_ONLY; restatement: oracle/restate/support.cc)"""
    import ctypes as C
    import os
    L = chk.lib
    fs = getattr(L, chk.prefix + "_group_support")
    fs.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    ss = getattr(L, chk.prefix + "_bundle_set_sample")
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
    old = os.environ.get("ORC_SUPPORT_GROUP_ONLY")
    os.environ["ORC_SUPPORT_GROUP_ONLY"] = "1"
    try:
        rc = fs(arr, len(hs), bag)
    finally:
        if old is None:
            os.environ.pop("ORC_SUPPORT_GROUP_ONLY", None)
        else:
            os.environ["ORC_SUPPORT_GROUP_ONLY"] = old
    d = chk.bag_to_dict(bag)
    L.orc_bag_free(bag)
    for h in hs:
        chk.free_bundle(h)
    assert rc == 0
    return d


def compare_group_support(ctx, batch, checker, gp, op, groups, stats=None):
    """agpu_batch_group_support (meta/assembler.cc:177-373: junction / start-end / non-splicing support, boundary_extend) on
    clusters of bundles against the checker: every member's and the combined graph's edge rows (source, target, count), abd,
    per-sample abd lists and the four boundary losses per vertex"""
    bad = []
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.bridge_all(gp)
    bt.group_bridge(groups, gp)
    got = bt.group_support(groups, gp)
    bt.free()
    edges = multi = 0
    for gi, g in enumerate(groups):
        ref = checker_group_support(checker, batch, op, g)
        d = got[gi]
        where = "cluster %d %s" % (gi, g)
        for name in sorted(ref):
            if name not in d:
                bad.append("%s: %s missing" % (where, name))
                continue
            a, b = ref[name], d[name]
            if a.dtype == np.float64:
                cmp_f64(name, a, np.asarray(b, np.float64), where, bad)
            else:
                cmp_int(name, a, np.asarray(b, np.int32), where, bad)
        e = ref["x_sup_edge"].reshape(-1, 3)
        edges += len(e)
        multi += int((e[:, 2] > 1).sum())
    if stats is not None:
        stats["support_edges"] = edges
        stats["support_multi"] = multi
    return bad
