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
