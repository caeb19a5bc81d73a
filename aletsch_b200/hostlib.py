"""ctypes bindings of libaletsch_host.so: synthetic record generator (host/synth.h) and the
SoA packer (host/packer.h, the record loop of meta/generator.cc:77-201)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SYNTH_PAIRED, SYNTH_SINGLE, SYNTH_LONG = 0, 1, 2
UNSTRANDED, FR_FIRST, FR_SECOND = 0, 1, 2


class SynthConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mode", C.c_int32), ("n_chrom", C.c_int32), ("chrom_len", C.c_int32),
                ("gene_spacing", C.c_int32), ("read_len", C.c_int32), ("min_exons", C.c_int32), ("max_exons", C.c_int32),
                ("expressed_fraction", C.c_double), ("secondary_rate", C.c_double), ("indel_rate", C.c_double),
                ("clip_rate", C.c_double), ("odd_rate", C.c_double)]


class SynthRecords(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", C.POINTER(C.c_int32)), ("pos", C.POINTER(C.c_int32)), ("rpos", C.POINTER(C.c_int32)),
                ("mpos", C.POINTER(C.c_int32)), ("isize", C.POINTER(C.c_int32)), ("flag", C.POINTER(C.c_uint16)),
                ("mapq", C.POINTER(C.c_uint8)), ("xs", C.POINTER(C.c_uint8)), ("qid", C.POINTER(C.c_uint64)),
                ("cigar_off", C.POINTER(C.c_uint32)), ("cigar", C.POINTER(C.c_uint32)), ("n_cigar", C.c_int64)]


class PackerParams(C.Structure):
    _fields_ = [("library_type", C.c_int32), ("min_mapping_quality", C.c_int32), ("max_num_cigar", C.c_int32),
                ("max_read_span", C.c_int32), ("min_bundle_gap", C.c_int32), ("use_second_alignment", C.c_int32),
                ("skip_single_exon_transcripts", C.c_int32)]


class PackerRecords(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", C.c_void_p), ("pos", C.c_void_p), ("rpos", C.c_void_p), ("mpos", C.c_void_p),
                ("isize", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p), ("xs", C.c_void_p), ("qid", C.c_void_p),
                ("cigar_off", C.c_void_p), ("cigar", C.c_void_p)]


class BatchIn(C.Structure):
    """agpu_batch_in (include/aletsch_gpu.h)."""
    _fields_ = [("n_bundles", C.c_int32), ("n_hits", C.c_int64), ("n_cigar", C.c_int64),
                ("bundle_hit_off", C.c_void_p), ("bundle_tid", C.c_void_p), ("bundle_sample", C.c_void_p),
                ("pos", C.c_void_p), ("rpos", C.c_void_p), ("mpos", C.c_void_p), ("isize", C.c_void_p),
                ("flag", C.c_void_p), ("strand", C.c_void_p), ("xs", C.c_void_p), ("qid", C.c_void_p),
                ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("bundle_strand", C.c_void_p)]


class BatchPacked(C.Structure):
    """agpu_batch_packed (include/aletsch_gpu.h): the compact form of a batch for the host -> device link."""
    _fields_ = [("n_bundles", C.c_int32), ("n_hits", C.c_int64), ("n_cigar", C.c_int64), ("n_units", C.c_int64),
                ("default_unit", C.c_uint32), ("bundle_hit_off", C.c_void_p), ("bundle_tid", C.c_void_p), ("bundle_sample", C.c_void_p), ("bundle_strand", C.c_void_p),
                ("bundle_pos0", C.c_void_p), ("dpos", C.c_void_p), ("dmpos", C.c_void_p), ("isize16", C.c_void_p),
                ("qid", C.c_void_p), ("hit_meta", C.c_void_p), ("units", C.c_void_p),
                ("n_esc_pos", C.c_int64), ("n_esc_mpos", C.c_int64), ("n_esc_isize", C.c_int64), ("n_esc_units", C.c_int64),
                ("esc_pos_idx", C.c_void_p), ("esc_mpos_idx", C.c_void_p), ("esc_isize_idx", C.c_void_p), ("esc_units_idx", C.c_void_p),
                ("esc_pos_val", C.c_void_p), ("esc_mpos_val", C.c_void_p), ("esc_isize_val", C.c_void_p), ("esc_units_val", C.c_void_p)]


COMPACT_ARRAYS = [("bundle_hit_off", np.int64, "nb1"), ("bundle_tid", np.int32, "nb"), ("bundle_sample", np.int32, "nb"),
                  ("bundle_strand", np.uint8, "nb"), ("bundle_pos0", np.int32, "nb"), ("dpos", np.uint16, "nh"), ("dmpos", np.int16, "nh"),
                  ("isize16", np.int16, "nh"), ("qid", np.uint64, "nh"), ("hit_meta", np.uint8, "nh"),
                  ("units", np.uint16, "nu"), ("esc_pos_idx", np.int64, "ep"), ("esc_mpos_idx", np.int64, "em"), ("esc_isize_idx", np.int64, "ei"),
                  ("esc_units_idx", np.int64, "eu"), ("esc_pos_val", np.int32, "ep"), ("esc_mpos_val", np.int32, "em"),
                  ("esc_isize_val", np.int32, "ei"), ("esc_units_val", np.int32, "eu")]


def compact_struct(arrays, n_cigar, ptr=lambda a: a.ctypes.data, default_unit=None):
    """agpu_batch_packed over a dict of arrays (numpy, or anything `ptr` can turn into an address, e.g. pinned tensors)"""
    p = BatchPacked()
    p.n_bundles, p.n_hits = len(arrays["bundle_tid"]), len(arrays["dpos"])
    p.n_cigar, p.n_units = n_cigar, len(arrays["units"])
    p.n_esc_pos, p.n_esc_mpos, p.n_esc_isize = len(arrays["esc_pos_idx"]), len(arrays["esc_mpos_idx"]), len(arrays["esc_isize_idx"])
    p.n_esc_units = len(arrays["esc_units_idx"])
    p.default_unit = int(arrays["default_unit"][0]) if default_unit is None else default_unit
    for name, _, _ in COMPACT_ARRAYS:
        setattr(p, name, ptr(arrays[name]))
    return p


HIT_FIELDS = [("pos", np.int32), ("rpos", np.int32), ("mpos", np.int32), ("isize", np.int32), ("flag", np.uint16),
              ("strand", np.uint8), ("xs", np.uint8), ("qid", np.uint64)]


def build():
    """compile libaletsch_host.so in-tree (g++)."""
    import subprocess
    src = [os.path.join(_HERE, "host", f) for f in ("synth.cc", "packer.cc", "bamio.cc")]
    out = os.path.join(_HERE, "libaletsch_host.so")
    deps = src + [os.path.join(_HERE, "host", f) for f in ("synth.h", "packer.h", "bamio.h")] + \
        [os.path.join(_HERE, "..", "include", "aletsch_gpu.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", out] + src + ["-lpthread", "-lz"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libaletsch_host.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.synth_create.restype = C.c_void_p
        L.synth_create.argtypes = [C.POINTER(SynthConfig)]
        L.synth_destroy.argtypes = [C.c_void_p]
        L.synth_num_genes.argtypes = [C.c_void_p]
        L.synth_generate.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.POINTER(SynthRecords)]
        L.synth_default_config.argtypes = [C.POINTER(SynthConfig), C.c_int]
        L.synth_records_free.argtypes = [C.POINTER(SynthRecords)]
        L.bam_write_records.argtypes = [C.c_char_p, C.c_int32, C.c_void_p, C.POINTER(SynthRecords), C.c_int]
        L.bam_read_records.argtypes = [C.c_char_p, C.POINTER(SynthRecords), C.POINTER(C.c_int32), C.c_void_p, C.c_int32]
        L.bam_set_key_bits.argtypes = [C.c_int]
        L.packer_create.restype = C.c_void_p
        L.packer_destroy.argtypes = [C.c_void_p]
        L.packer_add_sample.argtypes = [C.c_void_p, C.POINTER(PackerRecords), C.POINTER(PackerParams), C.c_int32]
        L.packer_view.argtypes = [C.c_void_p, C.POINTER(BatchIn)]
        L.packer_infer_library_type.argtypes = [C.POINTER(PackerRecords), C.POINTER(PackerParams), C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                                C.c_void_p]
        L.packer_compact_create.restype = C.c_void_p
        L.packer_compact_create.argtypes = [C.POINTER(BatchIn)]
        L.packer_compact_view.restype = C.POINTER(BatchPacked)
        L.packer_compact_view.argtypes = [C.c_void_p]
        L.packer_compact_destroy.argtypes = [C.c_void_p]
        L.packer_region_table.restype = C.c_int64
        L.packer_region_table.argtypes = [C.c_void_p, C.POINTER(PackerRecords), C.c_int32, C.c_void_p, C.c_int32, C.POINTER(PackerParams)]
        L.packer_regions.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 5
        L.packer_add_sample_regions.argtypes = [C.c_void_p, C.POINTER(PackerRecords), C.POINTER(PackerParams), C.c_int32]
        L.packer_default_params.argtypes = [C.POINTER(PackerParams)]
        L.packer_preview_add.restype = C.c_int64
        L.packer_preview_add.argtypes = [C.c_void_p, C.POINTER(PackerRecords), C.POINTER(PackerParams), C.c_int32]
        L.packer_preview_view.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 2 + [C.POINTER(C.c_int64)] + [C.POINTER(C.c_void_p)] * 4
        L.packer_insertsize_profile.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.packer_reorder_lists.restype = C.c_int64
        L.packer_reorder_lists.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.packer_records_seen.restype = C.c_int64
        L.packer_records_seen.argtypes = [C.c_void_p]
        L.packer_bundle_side.restype = C.POINTER(C.c_uint8)
        L.packer_bundle_side.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _np(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(n, dtype=dtype)       # e.g. the one-entry cigar_off of a sample without records
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()


def default_config(mode=SYNTH_PAIRED, **kw):
    c = SynthConfig()
    lib().synth_default_config(C.byref(c), mode)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


class Synth:
    """synthetic transcriptome + read simulator "synth-v1"."""

    def __init__(self, cfg):
        self.cfg = cfg
        self._h = lib().synth_create(C.byref(cfg))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().synth_destroy(self._h)
            self._h = None

    @property
    def num_genes(self):
        return lib().synth_num_genes(self._h)

    def sample(self, sample, templates, threads=8):
        """coordinate-sorted records of one sample as a dict of numpy arrays."""
        r = SynthRecords()
        rc = lib().synth_generate(self._h, sample, templates, threads, C.byref(r))
        if rc != 0:
            raise RuntimeError("synth_generate failed")
        n, nc = r.n, r.n_cigar
        out = {"n": n,
               "tid": _np(r.tid, n, np.int32), "pos": _np(r.pos, n, np.int32), "rpos": _np(r.rpos, n, np.int32),
               "mpos": _np(r.mpos, n, np.int32), "isize": _np(r.isize, n, np.int32), "flag": _np(r.flag, n, np.uint16),
               "mapq": _np(r.mapq, n, np.uint8), "xs": _np(r.xs, n, np.uint8), "qid": _np(r.qid, n, np.uint64),
               "cigar_off": _np(r.cigar_off, n + 1, np.uint32), "cigar": _np(r.cigar, nc, np.uint32)}
        lib().synth_records_free(C.byref(r))
        return out


_REC_FIELDS = [("tid", np.int32), ("pos", np.int32), ("rpos", np.int32), ("mpos", np.int32), ("isize", np.int32), ("flag", np.uint16),
               ("mapq", np.uint8), ("xs", np.uint8), ("qid", np.uint64)]


def _records_struct(rec, keep):
    r = SynthRecords()
    r.n = rec["n"]
    r.n_cigar = len(rec["cigar"])
    for k, dt in _REC_FIELDS + [("cigar_off", np.uint32), ("cigar", np.uint32)]:
        a = np.ascontiguousarray(rec[k], dt)
        keep.append(a)
        setattr(r, k, a.ctypes.data_as(C.POINTER(np.ctypeslib.as_ctypes_type(dt))))
    return r


def write_bam(path, rec, chrom_len, tag_mode=0):
    """records of one sample (Synth.sample layout) -> coordinate-sorted BAM file (host/bamio.h); tag_mode 1 writes ts:A"""
    keep = []
    r = _records_struct(rec, keep)
    cl = np.ascontiguousarray(chrom_len, np.int32)
    rc = lib().bam_write_records(path.encode(), len(cl), cl.ctypes.data, C.byref(r), tag_mode)
    if rc != 0:
        raise RuntimeError("bam_write_records(%s) failed with %d" % (path, rc))


def read_bam(path):
    """BAM file -> (records in Synth.sample layout, chromosome lengths): the host ingest in front of the packer"""
    r = SynthRecords()
    n_chrom = C.c_int32(0)
    cl = np.zeros(4096, np.int32)
    rc = lib().bam_read_records(path.encode(), C.byref(r), C.byref(n_chrom), cl.ctypes.data, len(cl))
    if rc != 0:
        raise RuntimeError("bam_read_records(%s) failed with %d" % (path, rc))
    n, nc = r.n, r.n_cigar
    out = {"n": n}
    for k, dt in _REC_FIELDS:
        out[k] = _np(getattr(r, k), n, dt)
    out["cigar_off"] = _np(r.cigar_off, n + 1, np.uint32)
    out["cigar"] = _np(r.cigar, nc, np.uint32)
    lib().synth_records_free(C.byref(r))
    return out, cl[:min(n_chrom.value, len(cl))].copy()


def default_packer_params(library_type=FR_FIRST, **kw):
    p = PackerParams()
    lib().packer_default_params(C.byref(p))
    p.library_type = library_type
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class PackedBatch:
    """host copy of an agpu_batch_in: NB packed bundles (numpy arrays) + ctypes view."""

    def __init__(self, arrays):
        self.a = arrays
        self.n_bundles = len(arrays["bundle_tid"])
        self.n_hits = len(arrays["pos"])
        self.n_cigar = len(arrays["cigar"])

    def view(self):
        b = BatchIn()
        b.n_bundles, b.n_hits, b.n_cigar = self.n_bundles, self.n_hits, self.n_cigar
        for k in ("bundle_hit_off", "bundle_tid", "bundle_sample", "pos", "rpos", "mpos", "isize", "flag", "strand", "xs",
                  "qid", "cigar_off", "cigar"):
            setattr(b, k, self.a[k].ctypes.data)
        return b

    def compact(self):
        """the compact form (agpu_batch_packed) as a dict of numpy arrays; compact_struct() turns it into the C struct"""
        L = lib()
        v = self.view()
        if "bundle_strand" in self.a:
            v.bundle_strand = self.a["bundle_strand"].ctypes.data
        c = L.packer_compact_create(C.byref(v))
        if not c:
            raise ValueError("batch cannot be expressed in the compact form (pos decreasing in a bundle, or an oversized CIGAR)")
        try:
            p = L.packer_compact_view(c).contents
            n = {"nb": p.n_bundles, "nb1": p.n_bundles + 1, "nh": p.n_hits, "nu": p.n_units, "ep": p.n_esc_pos, "em": p.n_esc_mpos,
                 "ei": p.n_esc_isize, "eu": p.n_esc_units}
            out = {}
            for name, dt, size in COMPACT_ARRAYS:
                ptr = getattr(p, name)
                out[name] = _np(ptr, n[size], dt) if ptr else np.zeros(n[size], dt)
            out["default_unit"] = np.array([p.default_unit], np.int64)        # a scalar of the struct, carried with the arrays
            return out
        finally:
            L.packer_compact_destroy(c)

    def bundle(self, k):
        """arrays of bundle k alone (cigar offsets rebased)."""
        a = self.a
        h0, h1 = int(a["bundle_hit_off"][k]), int(a["bundle_hit_off"][k + 1])
        c0, c1 = int(a["cigar_off"][h0]), int(a["cigar_off"][h1])
        out = {f: np.ascontiguousarray(a[f][h0:h1]) for f, _ in HIT_FIELDS}
        out["cigar_off"] = np.ascontiguousarray(a["cigar_off"][h0:h1 + 1] - np.uint32(c0))
        out["cigar"] = np.ascontiguousarray(a["cigar"][c0:c1])
        out["tid"] = int(a["bundle_tid"][k])
        out["sample"] = int(a["bundle_sample"][k])
        return out

    def slice(self, b0, b1):
        """new PackedBatch holding the contiguous bundle range [b0, b1) (array slices, cigar offsets rebased)."""
        a = self.a
        h0, h1 = int(a["bundle_hit_off"][b0]), int(a["bundle_hit_off"][b1])
        c0, c1 = int(a["cigar_off"][h0]), int(a["cigar_off"][h1])
        arr = {f: np.ascontiguousarray(a[f][h0:h1]) for f, _ in HIT_FIELDS}
        arr["bundle_hit_off"] = np.ascontiguousarray(a["bundle_hit_off"][b0:b1 + 1] - h0)
        arr["cigar_off"] = np.ascontiguousarray(a["cigar_off"][h0:h1 + 1] - np.uint32(c0))
        arr["cigar"] = np.ascontiguousarray(a["cigar"][c0:c1])
        for f in ("bundle_tid", "bundle_sample", "bundle_side"):
            if f in a:
                arr[f] = np.ascontiguousarray(a[f][b0:b1])
        return PackedBatch(arr)

    def split(self, parts):
        """`parts` contiguous sub-batches of roughly equal hit counts (bundles stay whole)."""
        off = self.a["bundle_hit_off"]
        cuts = [0]
        for k in range(1, parts):
            cuts.append(max(cuts[-1], int(np.searchsorted(off, self.n_hits * k // parts))))
        cuts.append(self.n_bundles)
        return [self.slice(cuts[k], cuts[k + 1]) for k in range(parts) if cuts[k + 1] > cuts[k]]

    def select(self, ks):
        """new PackedBatch holding bundles ks in that order."""
        a = self.a
        parts = [self.bundle(k) for k in ks]
        arr = {f: (np.concatenate([p[f] for p in parts]) if parts else np.zeros(0, dt)) for f, dt in HIT_FIELDS}
        hit_off = np.zeros(len(parts) + 1, np.int64)
        cig_off = [np.zeros(1, np.uint32)]
        cbase = 0
        for i, p in enumerate(parts):
            hit_off[i + 1] = hit_off[i] + len(p["pos"])
            cig_off.append(p["cigar_off"][1:] + np.uint32(cbase))
            cbase += len(p["cigar"])
        arr["bundle_hit_off"] = hit_off
        arr["cigar_off"] = np.concatenate(cig_off).astype(np.uint32)
        arr["cigar"] = np.concatenate([p["cigar"] for p in parts]).astype(np.uint32) if parts else np.zeros(0, np.uint32)
        arr["bundle_tid"] = np.array([p["tid"] for p in parts], np.int32)
        arr["bundle_sample"] = np.array([p["sample"] for p in parts], np.int32)
        arr["bundle_side"] = np.array([a["bundle_side"][k] for k in ks], np.uint8) if "bundle_side" in a else np.zeros(len(parts), np.uint8)
        return PackedBatch(arr)


def infer_library_type(sample, params, max_preview_reads=2000000, max_preview_spliced_reads=50000, min_preview_spliced_reads=100,
                       preview_infer_ratio=0.8):
    """previewer::infer_library_type over one sample's records -> dict (library_type, bam_with_xs, reads, spliced, with_xs, used,
    first, second)"""
    r = PackerRecords()
    r.n = sample["n"]
    keep = []
    for k in ("tid", "pos", "rpos", "mpos", "isize", "flag", "mapq", "xs", "qid", "cigar_off", "cigar"):
        arr = np.ascontiguousarray(sample[k])
        keep.append(arr)
        setattr(r, k, arr.ctypes.data)
    out = np.zeros(8, np.int32)
    lib().packer_infer_library_type(C.byref(r), C.byref(params), max_preview_reads, max_preview_spliced_reads, min_preview_spliced_reads,
                                    preview_infer_ratio, out.ctypes.data)
    return dict(zip(("library_type", "bam_with_xs", "reads", "spliced", "with_xs", "used", "first", "second"), (int(x) for x in out)))


def _records_in(s, keep):
    r = PackerRecords()
    r.n = s["n"]
    for k in ("tid", "pos", "rpos", "mpos", "isize", "flag", "mapq", "xs", "qid", "cigar_off", "cigar"):
        arr = np.ascontiguousarray(s[k])
        keep.append(arr)
        setattr(r, k, arr.ctypes.data)
    return r


def _batch_of(L, pk):
    v = BatchIn()
    L.packer_view(pk, C.byref(v))
    nb, nh, nc = v.n_bundles, v.n_hits, v.n_cigar
    arr = {"bundle_hit_off": _np(v.bundle_hit_off, nb + 1, np.int64), "bundle_tid": _np(v.bundle_tid, nb, np.int32),
           "bundle_sample": _np(v.bundle_sample, nb, np.int32),
           "cigar_off": _np(v.cigar_off, nh + 1, np.uint32), "cigar": _np(v.cigar, nc, np.uint32)}
    for f, dt in HIT_FIELDS:
        arr[f] = _np(getattr(v, f), nh, dt)
    arr["bundle_side"] = _np(L.packer_bundle_side(pk), nb, np.uint8)
    return PackedBatch(arr)


def preview_pack(sample, params, min_num_hits_in_bundle=10):
    """the record loop of previewer::infer_insertsize (meta/previewer.cc:151-208) over one sample's records: the bundles
    previewer::process works on, packed, plus what the reference's never-flushed interval buffer does to their coverage maps.
    Returns (PackedBatch, event[NB], skip[H] uint16, extra = (bundle, l, r, count) int32 arrays)."""
    L = lib()
    pk = L.packer_create()
    try:
        keep = []
        r = _records_in(sample, keep)
        L.packer_preview_add(pk, C.byref(r), C.byref(params), min_num_hits_in_bundle)
        batch = _batch_of(L, pk)
        ev, sk, eb, el, er, ec = (C.c_void_p() for _ in range(6))
        ne = C.c_int64(0)
        L.packer_preview_view(pk, C.byref(ev), C.byref(sk), C.byref(ne), C.byref(eb), C.byref(el), C.byref(er), C.byref(ec))
        event = _np(ev.value, batch.n_bundles, np.int64)
        skip = _np(sk.value, batch.n_hits, np.uint16)
        extra = tuple(_np(x.value, ne.value, np.int32) for x in (eb, el, er, ec))
        return batch, event, skip, extra
    finally:
        L.packer_destroy(pk)


def insertsize_profile(clu_off, isize, event, max_preview_reads=2000000, min_preview_spliced_reads=100):
    """the tail of previewer::infer_insertsize (meta/previewer.cc:209-249) from the per-cluster fragment lengths of the device
    (agpu_preview_view): at most 1000 counted clusters per bundle (previewer::process), the break at max_preview_reads, then the
    percentiles.  Returns dict(insert_total, insertsize_low, insertsize_high, insertsize_median, insertsize_ave, insertsize_std);
    low / high / median / ave / std are None when fewer than min_preview_spliced_reads fragments were collected."""
    nb = len(clu_off) - 1
    valid = isize != np.iinfo(np.int32).min
    csum = np.concatenate([[0], np.cumsum(valid)])
    rank = csum[:-1] - np.repeat(csum[clu_off[:-1]], np.diff(clu_off))          # rank of a counted cluster inside its bundle
    keep = valid & (rank < 1000)
    kc = np.concatenate([[0], np.cumsum(keep)])
    d_off = np.ascontiguousarray(kc[clu_off], np.int64)
    d = np.ascontiguousarray(isize[keep], np.int32)
    ev = np.ascontiguousarray(event, np.int64)
    oi = np.full(4, -1, np.int32)
    od = np.zeros(2, np.float64)
    lib().packer_insertsize_profile(nb, d_off.ctypes.data, d.ctypes.data if len(d) else None, ev.ctypes.data if nb else None, max_preview_reads,
                                    min_preview_spliced_reads, oi.ctypes.data, od.ctypes.data)
    out = {"insert_total": int(oi[0])}
    ok = oi[0] >= min_preview_spliced_reads
    out.update({"insertsize_low": int(oi[1]) if ok else None, "insertsize_high": int(oi[2]) if ok else None,
                "insertsize_median": int(oi[3]) if ok else None, "insertsize_ave": float(od[0]) if ok else None,
                "insertsize_std": float(od[1]) if ok else None})
    return out


def pack(samples, params, sample_ids=None, chrom_len=None, region_length=1000000, tables=None):
    """run the record loop (meta/generator.cc:77-201) over each sample's records and pack all bundles.  With chrom_len the
    records go through the region table first (sample_profile::set_batch_boundaries) and every region gets its own record loop,
    as in the reference's end-to-end run; `tables` (a list) then receives each sample's table."""
    L = lib()
    pk = L.packer_create()
    try:
        for si, s in enumerate(samples):
            r = PackerRecords()
            r.n = s["n"]
            keep = []
            for k in ("tid", "pos", "rpos", "mpos", "isize", "flag", "mapq", "xs", "qid", "cigar_off", "cigar"):
                arr = np.ascontiguousarray(s[k])
                keep.append(arr)
                setattr(r, k, arr.ctypes.data)
            sid = si if sample_ids is None else sample_ids[si]
            if chrom_len is not None:
                cl = np.ascontiguousarray(chrom_len, np.int32)
                nr = L.packer_region_table(pk, C.byref(r), len(cl), cl.ctypes.data, region_length, C.byref(params))
                if nr < 0:
                    raise RuntimeError("packer_region_table: a record lies outside its chromosome")
                if tables is not None:
                    ptr = [C.c_void_p() for _ in range(5)]
                    L.packer_regions(pk, *[C.byref(x) for x in ptr])
                    tables.append({"reg_off": _np(ptr[0].value, len(cl) + 1, np.int64), "start1": _np(ptr[1].value, nr, np.int32),
                                   "start2": _np(ptr[2].value, nr, np.int32), "end1": _np(ptr[3].value, nr, np.int32),
                                   "start_rec": _np(ptr[4].value, nr, np.int64)})
                if L.packer_add_sample_regions(pk, C.byref(r), C.byref(params), sid) != 0:
                    raise RuntimeError("packer_add_sample_regions failed")
                continue
            if L.packer_add_sample(pk, C.byref(r), C.byref(params), sid) != 0:
                raise RuntimeError("packer_add_sample failed")
        v = BatchIn()
        L.packer_view(pk, C.byref(v))
        nb, nh, nc = v.n_bundles, v.n_hits, v.n_cigar
        arr = {"bundle_hit_off": _np(v.bundle_hit_off, nb + 1, np.int64), "bundle_tid": _np(v.bundle_tid, nb, np.int32),
               "bundle_sample": _np(v.bundle_sample, nb, np.int32),
               "cigar_off": _np(v.cigar_off, nh + 1, np.uint32), "cigar": _np(v.cigar, nc, np.uint32)}
        for f, dt in HIT_FIELDS:
            arr[f] = _np(getattr(v, f), nh, dt)
        arr["bundle_side"] = _np(L.packer_bundle_side(pk), nb, np.uint8)
        return PackedBatch(arr)
    finally:
        L.packer_destroy(pk)
