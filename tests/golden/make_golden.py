"""Generates tests/golden/bundles_v1.json (evidence, graph, clusters, bridging) and bundles_v2.json (phase set, boundary revision
at min_boundary_log_ratio 1.1 so that edges are added) from the REFERENCE BUILD (oracle/_ref/libaletsch_ref.so, i.e. the
reference's own translation units run in this container).  Run from the repo root: python tests/golden/make_golden.py
The fixture travels with the repository and pins the CPU checkers where /root/reference is not mounted."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import orclib   # noqa: E402
import parity   # noqa: E402
from aletsch_b200 import hostlib as H   # noqa: E402

MODE, TEMPLATES, SEED, CHROM = H.SYNTH_PAIRED, 6000, 20260199, 600_000
NAMES = ("bundle", "seg", "splices", "hcst_val", "hcst_cnt", "junc", "pexon", "pexon_d", "vert", "vert_d", "edge", "edge_d",
         "clu_bounds", "clu_count", "clu_fr_val", "opt", "opt_score", "opt_chain_val", "opt_whole_val", "frgs", "fcst_val", "fcst_cnt")


REV_RATIO = 1.1
V2_PHASE = ("phase_off", "phase_val", "phase_cnt")
V2_REVISE = ("rev_edge", "rev_edge_d", "rev_vert", "rev_vert_d")


def main():
    ref = orclib.Checker("ref")
    batch, lt = parity.make_batch(MODE, TEMPLATES, seed=SEED, chrom_len=CHROM)
    _, op = parity.params_pair(lt)
    out = {"mode": MODE, "templates": TEMPLATES, "seed": SEED, "chrom_len": CHROM, "n_bundles": batch.n_bundles, "bundles": [],
           "generated_by": "oracle/_ref/libaletsch_ref.so (reference TUs compiled unchanged from /root/reference)"}
    for k in range(batch.n_bundles):
        h = ref.new_bundle(batch.bundle(k), op)
        _, ev = ref.run(h, "evidence")
        ref.run(h, "fragments")
        cnt, br = ref.run(h, "bridge")
        ref.free_bundle(h)
        arrays = {}
        for n in NAMES:
            a = br[n] if n in br else ev[n]
            arrays[n] = [float(x) for x in a] if a.dtype.kind == "f" else [int(x) for x in a]
        out["bundles"].append({"bridged": int(cnt), "arrays": arrays})
    path = os.path.join(ROOT, "tests", "golden", "bundles_v1.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes,", batch.n_bundles, "bundles,", batch.n_hits, "hits")
    # v2: what follows the path -- build_phase_set and the boundary revision, on the bridged bundles
    _, op2 = parity.params_pair(lt, min_boundary_log_ratio=REV_RATIO)
    out2 = {"mode": MODE, "templates": TEMPLATES, "seed": SEED, "chrom_len": CHROM, "n_bundles": batch.n_bundles, "bundles": [],
            "min_boundary_log_ratio": REV_RATIO, "generated_by": out["generated_by"]}
    for k in range(batch.n_bundles):
        h = ref.new_bundle(batch.bundle(k), op2)
        ref.run(h, "fragments")
        ref.run(h, "bridge")
        _, ph = ref.run(h, "phase")
        _, rv = ref.run(h, "revise")
        ref.free_bundle(h)
        arrays = {}
        for src, names in ((ph, V2_PHASE), (rv, V2_REVISE)):
            for n in names:
                a = src[n]
                arrays[n] = [float(x) for x in a] if a.dtype.kind == "f" else [int(x) for x in a]
        out2["bundles"].append({"arrays": arrays})
    path = os.path.join(ROOT, "tests", "golden", "bundles_v2.json")
    json.dump(out2, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
