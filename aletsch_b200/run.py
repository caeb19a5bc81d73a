"""Command line: BAM files in, per-bundle evidence / splice graphs / bridged fragments / phasing paths on the GPU.

    python -m aletsch_b200.run [--library-type unstranded|first|second] [--device 0] [--clusters] a.bam [b.bam ...]

One sample per BAM.  The files are decoded on the host (host/bamio.cc), cut into bundles by the packer (the record loop of
meta/generator.cc:77-201), and the whole batch goes through bundle::bridge on the device; with --clusters the bundles are also
clustered across samples (bundle_group::resolve per 1 Mb region and strand) and every cluster is re-bridged against its combined
splice graph (assembler::bridge).  Prints one summary line per sample and the totals; the results themselves are available
through the fetch calls of aletsch_b200.gpu.Batch (this tool is a usage example, not a replacement for the assembler).
"""
import argparse
import json
import sys
import time

import numpy as np

from . import gpu as G
from . import hostlib as H

REGION = 1_000_000


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m aletsch_b200.run")
    ap.add_argument("bams", nargs="+")
    ap.add_argument("--library-type", default="first", choices=["unstranded", "first", "second", "auto"],
                    help="auto: previewer::infer_library_type over the first file's records")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--clusters", action="store_true", help="cross-sample clustering (-c 20 -s 0.2 style) + group-level re-bridge")
    ap.add_argument("--max-group-size", type=int, default=200)
    ap.add_argument("--min-grouping-similarity", type=float, default=0.10)
    ap.add_argument("--whole-file", action="store_true",
                    help="one record loop per file instead of the reference's region table (sample_profile::set_batch_boundaries), "
                         "which drops every region's first hit and the last region of the last chromosome")
    args = ap.parse_args(argv)
    lt = {"unstranded": H.UNSTRANDED, "first": H.FR_FIRST, "second": H.FR_SECOND, "auto": None}[args.library_type]
    t0 = time.time()
    recs, chrom_len = [], None
    for path in args.bams:
        r, cl = H.read_bam(path)
        if chrom_len is not None and not np.array_equal(cl, chrom_len):
            raise SystemExit("%s: reference dictionary differs from the first file's" % path)
        chrom_len = cl
        recs.append(r)
        print("%s: %d records" % (path, r["n"]), file=sys.stderr)
        if lt is None:
            pv = H.infer_library_type(r, H.default_packer_params(H.UNSTRANDED))
            lt = pv["library_type"]
            print("inferred library type %d from %s: %s" % (lt, path, pv), file=sys.stderr)
    batch = H.pack(recs, H.default_packer_params(lt), chrom_len=None if args.whole_file else chrom_len, region_length=REGION)
    t1 = time.time()
    gp = G.default_params(library_type=lt, max_group_size=args.max_group_size, min_grouping_similarity=args.min_grouping_similarity)
    ctx = G.Context(args.device)
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.bridge_all(gp)
    out = {"samples": len(recs), "bundles": batch.n_bundles, "decode_pack_s": round(t1 - t0, 3)}
    if args.clusters and batch.n_bundles:
        off, val = bt.fetch_splices()
        a = batch.a
        # bundle groups: (chromosome, 1 Mb region of the bundle's first hit, bundle strand).  The strand is the one
        # bundle_base::compute_strand leaves (for unstranded libraries the majority of the XS tags: '+', '-' and '.' bundles are
        # grouped separately, meta/incubator.cc:526-540).  The region is the NOMINAL one: the reference files a bundle under the
        # region-table entry its generator::resolve ran for, which reaches past the 1 Mb multiple until a gap wider than
        # min_bundle_gap (rnacore/sample_profile.cc:167-252), so a bundle starting just behind a boundary inside such a run joins
        # the previous group there and the next one here; and bundle_group::remove_duplicates (meta/bundle_group.cc:58-91: '+'
        # bundles that end at or before the previous region's end1 are emptied) is not applied.  This tool is a usage example of
        # the ABI, not a reference-equivalent scheduler: the incubator stays on the host (BASELINE.json: north_star).
        from .shard import bundle_region_keys
        key = bundle_region_keys(batch)
        order = np.lexsort((np.arange(batch.n_bundles), a["bundle_sample"], key))
        cuts = np.nonzero(np.diff(key[order]))[0] + 1
        groups = [g for g in np.split(order, cuts) if len(g)]
        goff = np.zeros(len(groups) + 1, np.int32)
        np.cumsum([len(g) for g in groups], out=goff[1:])
        loff, lval = G.reorder_lists(off, val, order)
        cl_of, ncl = G.group_resolve_arrays(ctx, goff, loff, lval, gp)
        clusters = []
        for gi in range(len(groups)):
            byc = {}
            for l in range(int(goff[gi]), int(goff[gi + 1])):
                byc.setdefault(int(cl_of[l]), []).append(int(order[l]))
            clusters.extend(v for _, v in sorted(byc.items()) if len(v) >= 2)
        if clusters:
            bt.group_bridge(clusters, gp)
        out["region_groups"] = len(groups)
        out["clusters_of_bundles"] = len(clusters)
    bt.graph(gp)
    phases = bt.phase_set()
    out.update(bt.counts())
    out["distinct_phases"] = int(sum(len(p["phase_cnt"]) for p in phases))
    out["device_s"] = round(time.time() - t1, 3)
    bt.free()
    ctx.close()
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
