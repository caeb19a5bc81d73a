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
