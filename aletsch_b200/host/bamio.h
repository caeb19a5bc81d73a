// Host ingest: coordinate-sorted BAM (BGZF, via zlib) -> decoded records in the SoA layout the packer takes
// (packer_records), i.e. what htslib + hit::hit / hit::set_tags hand the reference per record:
//   bam1_core_t fields                      rnacore/hit.cc:52-65   (pos, flag, mapq, mpos, isize, n_cigar; rpos = pos + cigar2rlen)
//   XS:A / ts:A -> xs                       rnacore/hit.cc:106-141 (ts is converted with the reverse-strand flag)
//   qname -> 64-bit key (equal <=> same qname, what build_fragments compares; rnacore/bundle_base.cc:308)
// plus a BAM writer used to turn synthetic records into real files for the round-trip tests and end-to-end runs.
// The region table of sample_profile::set_batch_boundaries (rnacore/sample_profile.cc:167-252) is not built here: a file
// is read front to back.
#ifndef ALETSCH_B200_HOST_BAMIO_H
#define ALETSCH_B200_HOST_BAMIO_H

#include <stdint.h>
#include "synth.h"

#ifdef __cplusplus
extern "C" {
#endif

// write records as a BAM file.  qnames are "q<qid in hex>"; spliced-strand tags: XS:A for tag_mode 0, ts:A (the minimap2
// convention, relative to the read) for tag_mode 1; NH:i:1 and HI:i:1 on every record.  Returns 0 on success.
int bam_write_records(const char *path, int32_t n_chrom, const int32_t *chrom_len, const synth_records *r, int tag_mode);

// read every record of a BAM file (free with synth_records_free).  n_chrom_out / chrom_len_out (optional, up to cap entries)
// receive the reference dictionary.  Returns 0 on success, < 0 on a malformed file, -4 when the file has more reference
// sequences than `cap`.
// qid: a 64-bit hash of the query name made collision-free where it matters -- every record's key is checked (with a second,
// independent hash) against the keys of all records within 1.2 Mb upstream, the only ones it can share a bundle with
// (max_read_span is 500000), and a different name under an equal key is given another key.  So inside a bundle "equal key <=>
// equal qname" holds, which is what the ABI promises the device (include/aletsch_gpu.h: agpu_batch_in::qid).
// Records with more than 65535 CIGAR operations carry their CIGAR in the CG:B,I tag: it is swapped in, as htslib does.
int bam_read_records(const char *path, synth_records *out, int32_t *n_chrom_out, int32_t *chrom_len_out, int32_t cap);

// test hook: truncate the query-name hashes to `bits` bits (64 = off) so that colliding keys occur on ordinary data and the
// re-keying above is exercised; returns the previous value
int bam_set_key_bits(int bits);

#ifdef __cplusplus
}
#endif
#endif
