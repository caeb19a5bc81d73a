"""CPU tier: the host-side SoA packer (aletsch_b200/host/packer.cc) against the REFERENCE'S OWN record loop, generator::resolve +
generator::generate (meta/generator.cc:51-227), compiled unchanged and fed the same records through the htslib stand-in:
record filters, the two dedupes, strand routing for every library type, bundle cuts, the single-exon skip."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib
from aletsch_b200 import hostlib as H


class RecordsIn(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_chrom", C.c_int32), ("chrom_len", C.c_void_p), ("tid", C.c_void_p), ("pos", C.c_void_p),
                ("mpos", C.c_void_p), ("isize", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p), ("xs", C.c_void_p),
                ("qid", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p)]


def records_in(rec, chrom_len):
    keep = []
    r = RecordsIn()
    r.n = rec["n"]
    cl = np.ascontiguousarray(chrom_len, np.int32)
    keep.append(cl)
    r.n_chrom, r.chrom_len = len(cl), cl.ctypes.data
    for k, dt in (("tid", np.int32), ("pos", np.int32), ("mpos", np.int32), ("isize", np.int32), ("flag", np.uint16), ("mapq", np.uint8),
                  ("xs", np.uint8), ("qid", np.uint64), ("cigar_off", np.uint32), ("cigar", np.uint32)):
        a = np.ascontiguousarray(rec[k], dt)
        keep.append(a)
        setattr(r, k, a.ctypes.data)
    return r, keep


def ref_generate(chk, rec, chrom_len, op, second):
    r, keep = records_in(rec, chrom_len)
    f = chk.lib.ref_generate
    f.argtypes = [C.POINTER(RecordsIn), C.POINTER(orclib.Params), C.c_int, C.c_void_p]
    chk.lib.orc_bag_new.restype = C.c_void_p
    bag = chk.lib.orc_bag_new()
    # generator::generate prints nothing, but set_tags / add_hit_intervals may: keep them off the test output
    nb = f(C.byref(r), C.byref(op), 1 if second else 0, bag)
    d = chk.bag_to_dict(bag)
    chk.lib.orc_bag_free(bag)
    return nb, d


@pytest.mark.parametrize("mode,lt,second", [(H.SYNTH_PAIRED, H.FR_FIRST, True), (H.SYNTH_PAIRED, H.FR_SECOND, True),
                                             (H.SYNTH_PAIRED, H.UNSTRANDED, False), (H.SYNTH_SINGLE, H.UNSTRANDED, True),
                                             (H.SYNTH_LONG, H.UNSTRANDED, True)])
def test_packer_matches_reference_generator(checkers, mode, lt, second):
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    chk = checkers["ref"]
    cfg = H.default_config(mode, chrom_len=2_000_000, seed=20260121, n_chrom=2)
    rec = H.Synth(cfg).sample(0, 25000 if mode != H.SYNTH_LONG else 3000, threads=4)
    pp = H.default_packer_params(lt, use_second_alignment=1 if second else 0)
    batch = H.pack([rec], pp)
    op = orclib.default_params(library_type=lt)
    nb, gen = ref_generate(chk, rec, [cfg.chrom_len] * cfg.n_chrom, op, second)
    assert nb == batch.n_bundles
    # fr-firststrand reads declared fr-secondstrand: every spliced hit contradicts its XS tag and is dropped, the rest form
    # single-exon bundles that generator::generate skips -- both sides must agree on "nothing"
    assert nb > 5 or lt == H.FR_SECOND
    assert np.array_equal(gen["gen_off"].astype(np.int64), batch.a["bundle_hit_off"])
    assert np.array_equal(gen["gen_bundle"].reshape(-1, 4)[:, 0], batch.a["bundle_tid"])
    for name, key in (("gen_pos", "pos"), ("gen_rpos", "rpos"), ("gen_mpos", "mpos"), ("gen_isize", "isize"), ("gen_flag", "flag"),
                      ("gen_strand", "strand"), ("gen_xs", "xs")):
        assert np.array_equal(gen[name], batch.a[key].astype(np.int32)), name


@pytest.mark.parametrize("mode,lt,region_length", [(H.SYNTH_PAIRED, H.FR_FIRST, 1_000_000), (H.SYNTH_PAIRED, H.FR_FIRST, 100_000),
                                                    (H.SYNTH_PAIRED, H.UNSTRANDED, 50_000), (H.SYNTH_LONG, H.UNSTRANDED, 200_000)])
def test_region_table_matches_reference(checkers, mode, lt, region_length):
    """sample_profile::set_batch_boundaries + one generator::resolve per region (the reference's end-to-end ingest, quirks
    included: a region's first hit is never seen, the last region of the last chromosome stays closed)"""
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    chk = checkers["ref"]
    cfg = H.default_config(mode, chrom_len=2_000_000, seed=20260122, n_chrom=2)
    rec = H.Synth(cfg).sample(0, 25000 if mode != H.SYNTH_LONG else 3000, threads=4)
    pp = H.default_packer_params(lt)
    chrom_len = [cfg.chrom_len] * cfg.n_chrom
    tables = []
    batch = H.pack([rec], pp, chrom_len=chrom_len, region_length=region_length, tables=tables)
    whole = H.pack([rec], pp)
    op = orclib.default_params(library_type=lt)
    r, keep = records_in(rec, chrom_len)
    f = chk.lib.ref_generate_regions
    f.argtypes = [C.POINTER(RecordsIn), C.POINTER(orclib.Params), C.c_int, C.c_int, C.c_void_p]
    chk.lib.orc_bag_new.restype = C.c_void_p
    bag = chk.lib.orc_bag_new()
    nb = f(C.byref(r), C.byref(op), 1, region_length, bag)
    gen = chk.bag_to_dict(bag)
    chk.lib.orc_bag_free(bag)
    t = tables[0]
    assert np.array_equal(gen["reg_off"].astype(np.int64), t["reg_off"])
    for a, b in (("reg_start1", "start1"), ("reg_start2", "start2"), ("reg_end1", "end1"), ("reg_start_off", "start_rec")):
        assert np.array_equal(gen[a].astype(np.int64), t[b].astype(np.int64)), a
    open_regions = int((t["start1"] < t["end1"]).sum())
    assert open_regions >= 2
    assert nb == batch.n_bundles and nb > 5
    # the quirks cost hits: fewer than the whole-file record loop admits
    assert batch.n_hits < whole.n_hits
    assert np.array_equal(gen["gen_off"].astype(np.int64), batch.a["bundle_hit_off"])
    for name, key in (("gen_pos", "pos"), ("gen_rpos", "rpos"), ("gen_mpos", "mpos"), ("gen_isize", "isize"), ("gen_flag", "flag"),
                      ("gen_strand", "strand"), ("gen_xs", "xs")):
        assert np.array_equal(gen[name], batch.a[key].astype(np.int32)), name


@pytest.mark.parametrize("mode,flip,caps", [(H.SYNTH_PAIRED, False, (2000000, 50000, 100)), (H.SYNTH_PAIRED, True, (2000000, 50000, 100)),
                                             (H.SYNTH_PAIRED, False, (5000, 200, 100)), (H.SYNTH_PAIRED, False, (2000000, 300, 100)),
                                             (H.SYNTH_SINGLE, False, (2000000, 50000, 100)), (H.SYNTH_LONG, False, (2000000, 50000, 100)),
                                             (H.SYNTH_PAIRED, False, (2000000, 50000, 1000000))])
def test_infer_library_type_matches_reference_previewer(checkers, mode, flip, caps):
    """previewer::infer_library_type (meta/previewer.cc:29-148), the first half of the reference's preview pass, on the host"""
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    chk = checkers["ref"]
    cfg = H.default_config(mode, chrom_len=2_000_000, seed=20260123, n_chrom=2)
    rec = H.Synth(cfg).sample(0, 25000 if mode != H.SYNTH_LONG else 3000, threads=4)
    if flip:                                  # the same reads as an fr-secondstrand library would tag them
        x = rec["xs"].copy()
        rec["xs"] = np.where(x == ord("+"), ord("-"), np.where(x == ord("-"), ord("+"), x)).astype(np.uint8)
    pp = H.default_packer_params(H.UNSTRANDED)
    ours = H.infer_library_type(rec, pp, caps[0], caps[1], caps[2], 0.8)
    r, keep = records_in(rec, [cfg.chrom_len] * cfg.n_chrom)
    f = chk.lib.ref_infer_library_type
    f.argtypes = [C.POINTER(RecordsIn), C.POINTER(orclib.Params), C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    chk.lib.orc_bag_new.restype = C.c_void_p
    bag = chk.lib.orc_bag_new()
    lt = f(C.byref(r), C.byref(orclib.default_params()), caps[0], caps[1], caps[2], 0.8, bag)
    ref = chk.bag_to_dict(bag)["preview"]
    chk.lib.orc_bag_free(bag)
    assert [ours["library_type"], ours["bam_with_xs"], ours["with_xs"], ours["used"]] == [int(x) for x in ref], (ours, ref)
    assert lt == ours["library_type"]
    if caps == (2000000, 50000, 100) and mode == H.SYNTH_PAIRED:
        assert ours["library_type"] == (H.FR_SECOND if flip else H.FR_FIRST), ours
    if caps[2] == 1000000:
        assert ours["library_type"] == H.UNSTRANDED
