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
// receive the reference dictionary.  Returns 0 on success, < 0 on a malformed file.
int bam_read_records(const char *path, synth_records *out, int32_t *n_chrom_out, int32_t *chrom_len_out, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif
