"""CPU tier: host ingest (aletsch_b200/host/bamio.cc).  Synthetic records written as a real BGZF/BAM file and read back must
give the same decoded records (what htslib + hit::hit / hit::set_tags hand the reference, rnacore/hit.cc:52-141), the packer
must cut the same bundles from them, and the device path must not see a difference."""
import os

import numpy as np
import pytest

import parity
from aletsch_b200 import gpu as G
from aletsch_b200 import hostlib as H

SAME = ("tid", "pos", "rpos", "mpos", "isize", "flag", "mapq", "xs", "cigar_off", "cigar")


def qid_classes(q):
    """equality structure of the query-name keys: index of the first record with the same key"""
    first = {}
    return np.array([first.setdefault(int(x), i) for i, x in enumerate(q)], np.int64)


@pytest.mark.parametrize("mode,tag_mode", [(H.SYNTH_PAIRED, 0), (H.SYNTH_LONG, 1), (H.SYNTH_SINGLE, 0)])
def test_bam_round_trip(tmp_path, mode, tag_mode):
    cfg = H.default_config(mode, chrom_len=1_500_000, seed=20260111)
    rec = H.Synth(cfg).sample(0, 8000 if mode != H.SYNTH_LONG else 1500, threads=2)
    path = str(tmp_path / "s.bam")
    H.write_bam(path, rec, [cfg.chrom_len] * cfg.n_chrom, tag_mode=tag_mode)
    assert os.path.getsize(path) > 1000
    raw = open(path, "rb").read()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[-28:-24] == b"\x1f\x8b\x08\x04"      # BGZF blocks, EOF marker last
    # independent decoders: BGZF is a series of gzip members; the first alignment record parsed by hand (SAM/BAM spec 4.2)
    import gzip
    import struct
    plain = gzip.decompress(raw)
    assert plain[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", plain, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", plain, o)
    o += 4
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", plain, o)
        o += 4 + l_name + 4
    bs, tid, pos, lq, mapq, _bin, ncig, flag, lseq, mtid, mpos, isize = struct.unpack_from("<iiiBBHHHiiii", plain, o)
    assert (tid, pos, mapq, ncig, flag, mpos, isize) == (int(rec["tid"][0]), int(rec["pos"][0]), int(rec["mapq"][0]),
                                                         int(rec["cigar_off"][1]), int(rec["flag"][0]), int(rec["mpos"][0]), int(rec["isize"][0]))
    assert plain[o + 36:o + 36 + lq - 1] == ("q%x" % int(rec["qid"][0])).encode()
    back, chrom_len = H.read_bam(path)
    assert list(chrom_len) == [cfg.chrom_len] * cfg.n_chrom
    assert back["n"] == rec["n"]
    for k in SAME:
        assert np.array_equal(back[k], rec[k]), k
    assert np.array_equal(qid_classes(back["qid"]), qid_classes(rec["qid"]))


def test_bam_to_bundles_to_device(tmp_path, emu_lib):
    cfg = H.default_config(H.SYNTH_PAIRED, chrom_len=1_000_000, seed=20260112)
    syn = H.Synth(cfg)
    recs = [syn.sample(k, 6000, threads=2) for k in range(2)]
    back = []
    for k, r in enumerate(recs):
        path = str(tmp_path / ("s%d.bam" % k))
        H.write_bam(path, r, [cfg.chrom_len] * cfg.n_chrom)
        back.append(H.read_bam(path)[0])
    pp = H.default_packer_params(H.FR_FIRST)
    b0, b1 = H.pack(recs, pp), H.pack(back, pp)
    assert b0.n_bundles == b1.n_bundles and b0.n_hits == b1.n_hits
    for k in ("bundle_hit_off", "pos", "rpos", "mpos", "isize", "flag", "strand", "xs", "cigar_off", "cigar"):
        assert np.array_equal(b0.a[k], b1.a[k]), k
    gp, _ = parity.params_pair(H.FR_FIRST)
    ctx = G.Context(0, lib_path=emu_lib)
    outs = []
    for b in (b0, b1):
        bt = ctx.upload(b.view(), keepalive=b)
        bt.bridge_all(gp)
        outs.append((bt.counts(), bt.bundle_counts().copy()))
        bt.free()
    ctx.close()
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][0]["bridged"] > 0


def test_bam_reader_rejects_garbage(tmp_path):
    path = str(tmp_path / "bad.bam")
    open(path, "wb").write(b"not a bam file at all" * 10)
    with pytest.raises(RuntimeError):
        H.read_bam(path)


def test_qname_keys_survive_hash_collisions(tmp_path):
    """the ABI promises the device "equal qid <=> equal qname" inside a bundle (rnacore/bundle_base.cc:308 compares the names).
    With the name hashes truncated to 16 bits a few hundred of the ~6000 names collide with a different name nearby; the reader must
    still hand out keys with exactly the equality structure of the names (the synthetic names are "q<template id>")."""
    cfg = H.default_config(H.SYNTH_PAIRED, chrom_len=800_000, seed=20260113)
    rec = H.Synth(cfg).sample(0, 6000, threads=2)
    path = str(tmp_path / "c.bam")
    H.write_bam(path, rec, [cfg.chrom_len] * cfg.n_chrom)
    old = H.lib().bam_set_key_bits(16)
    try:
        back, _ = H.read_bam(path)
    finally:
        H.lib().bam_set_key_bits(old)
    # all records of this file lie within the 1.2 Mb window: the classes must agree over the whole file
    assert np.array_equal(qid_classes(back["qid"]), qid_classes(rec["qid"]))
    names = len(set(rec["qid"].tolist()))
    assert len(set(back["qid"].tolist())) == names and int(back["qid"].max()) < 65536
    # the birthday bound makes plain 16-bit truncation collide here (about names^2 / 2^17 pairs): without the re-keying the
    # classes could not have matched
    assert names * names // (1 << 17) > 50
    full, _ = H.read_bam(path)
    assert np.array_equal(qid_classes(full["qid"]), qid_classes(rec["qid"]))


def test_long_cigar_in_cg_tag(tmp_path):
    """a record with more than 65535 CIGAR operations keeps them in the CG:B,I tag behind a <read length>S<reference length>N
    placeholder (SAM specification 4.2.2); the reader swaps them in like htslib, so rpos and the splices are the real ones"""
    n_ops = 70001
    ops = np.empty(n_ops, np.uint32)
    ops[0::2] = (3 << 4) | 0          # 3M
    ops[1::2] = (2 << 4) | 2          # 2D
    ops[35001] = (500 << 4) | 3       # one intron in the middle
    ref_len = int(sum((int(c) >> 4) for c in ops if (int(c) & 0xf) in (0, 2, 3)))
    rec = {"n": 2, "tid": np.zeros(2, np.int32), "pos": np.array([1000, 1000 + ref_len + 50], np.int32),
           "rpos": np.array([1000 + ref_len, 1000 + ref_len + 150], np.int32), "mpos": np.zeros(2, np.int32), "isize": np.zeros(2, np.int32),
           "flag": np.zeros(2, np.uint16), "mapq": np.full(2, 60, np.uint8), "xs": np.array([ord("+"), ord(".")], np.uint8),
           "qid": np.array([7, 8], np.uint64), "cigar_off": np.array([0, n_ops, n_ops + 1], np.uint32),
           "cigar": np.concatenate([ops, np.array([(100 << 4) | 0], np.uint32)])}
    path = str(tmp_path / "l.bam")
    H.write_bam(path, rec, [2_000_000])
    back, _ = H.read_bam(path)
    assert back["n"] == 2
    for k in ("pos", "rpos", "cigar_off", "cigar", "xs"):
        assert np.array_equal(back[k], rec[k]), k


def test_dictionary_larger_than_buffer_is_an_error(tmp_path):
    cfg = H.default_config(H.SYNTH_PAIRED, chrom_len=300_000, n_chrom=3, seed=20260114)
    rec = H.Synth(cfg).sample(0, 500, threads=1)
    path = str(tmp_path / "d.bam")
    H.write_bam(path, rec, [cfg.chrom_len] * 3)
    import ctypes as C
    r = H.SynthRecords()
    n = C.c_int32(0)
    cl = np.zeros(2, np.int32)
    assert H.lib().bam_read_records(path.encode(), C.byref(r), C.byref(n), cl.ctypes.data, 2) == -4
