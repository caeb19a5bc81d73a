// Host-side SoA packer: the record loop of generator::resolve (meta/generator.cc:77-201) with
// bundle_base::add_hit's admission rules (rnacore/bundle_base.cc:73-104), emitting packed
// bundles in the layout of agpu_batch_in (include/aletsch_gpu.h) instead of bundle_base objects.
#ifndef ALETSCH_B200_HOST_PACKER_H
#define ALETSCH_B200_HOST_PACKER_H

#include <stdint.h>
#include "../../include/aletsch_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

// what the record loop reads from `parameters` (util/parameters.cc:53-62) and sample_profile
typedef struct packer_params
{
	int32_t library_type;
	int32_t min_mapping_quality;         // 1
	int32_t max_num_cigar;               // 10000
	int32_t max_read_span;               // 500000
	int32_t min_bundle_gap;              // 200
	int32_t use_second_alignment;        // 1
	int32_t skip_single_exon_transcripts;// 1
} packer_params;

// coordinate-sorted decoded records of one sample (what htslib hands the reference per record)
typedef struct packer_records
{
	int64_t n;
	const int32_t *tid, *pos, *rpos, *mpos, *isize;
	const uint16_t *flag;
	const uint8_t *mapq;
	const uint8_t *xs;
	const uint64_t *qid;
	const uint32_t *cigar_off;
	const uint32_t *cigar;
} packer_records;

void packer_default_params(packer_params *p);
void *packer_create(void);
void packer_destroy(void *pk);
// run the record loop over one sample and append its bundles to the batch under construction
int packer_add_sample(void *pk, const packer_records *r, const packer_params *p, int32_t sample);
// sample_profile::set_batch_boundaries (rnacore/sample_profile.cc:167-252) over the records of one sample: regions per
// chromosome (chrom_len / region_length + 1 each; region_partition_length is 1,000,000, util/parameters.cc:42) with the
// reference's start1 / start2 / end1 and, in place of the BGZF offset, the index of the record that FOLLOWS the region's
// first hit (bgzf_tell is taken after that hit was read).  Returns the number of regions, < 0 on a record outside its chromosome.
int64_t packer_region_table(void *pk, const packer_records *r, int32_t n_chrom, const int32_t *chrom_len, int32_t region_length,
		const packer_params *p);
// the table built last: reg_off[n_chrom + 1] and one entry per region; returns n_chrom
int packer_regions(void *pk, const int64_t **reg_off, const int32_t **start1, const int32_t **start2, const int32_t **end1, const int64_t **start_rec);
// one record loop per region of that table with start1 < end1 (meta/incubator.cc:355-380, meta/generator.cc:51-81)
int packer_add_sample_regions(void *pk, const packer_records *r, const packer_params *p, int32_t sample);
// previewer::infer_library_type (meta/previewer.cc:29-148) over decoded records; defaults of util/parameters.cc:65-68 are
// 2000000, 50000, 100, 0.8.  out[8]: library type, bam_with_xs, reads, spliced, with xs, used, first, second
int packer_infer_library_type(const packer_records *r, const packer_params *p, int32_t max_preview_reads, int32_t max_preview_spliced_reads,
		int32_t min_preview_spliced_reads, double preview_infer_ratio, int32_t *out);
// ---- previewer::infer_insertsize (meta/previewer.cc:151-304): the record loop of the insert-size preview.  Appends to the batch
// under construction the bundles previewer::process would work on (10 .. 20000 stored hits; both strand streams, in the order of
// the process() calls) and records what makes the reference's preview bundles differ from ordinary ones:
//   * previewer never calls add_buf_intervals and bundle_base::clear() does not reset interval_buf (rnacore/bundle_base.cc:106-204):
//     the last run of identical blocks in each of the ten buffer slots is missing from the bundle's coverage map when process()
//     looks at it -- skip[h] bit z set = the z-th BAM_CMATCH block of hit h adds no coverage;
//   * a run left in the buffer is flushed into whichever LATER bundle displaces it; on the same chromosome it lies left of that
//     bundle and is never looked at, after a chromosome change it can fall inside it -- extra (bundle, l, r, count) intervals.
// event[b] = index of the record whose arrival triggered process() on bundle b (both streams can share one): the host replays the
// `cnt >= max_preview_reads` break of the reference with it.  Returns the number of bundles appended.
int64_t packer_preview_add(void *pk, const packer_records *r, const packer_params *p, int32_t min_num_hits_in_bundle);
// arrays of the preview bundles appended so far: event[NB], skip[H] (aligned with the batch's hits), extra intervals
int packer_preview_view(void *pk, const int64_t **event, const uint16_t **skip, int64_t *n_extra, const int32_t **ex_bundle,
		const int32_t **ex_l, const int32_t **ex_r, const int32_t **ex_cnt);
// the insert-size profile from the fragment lengths the device found (meta/previewer.cc:214-249): d[] grouped by bundle
// (d_off[NB + 1], in cluster order, invalid clusters already removed and at most 1000 per bundle), the bundles' events, the
// break at max_preview_reads.  out_i[4] = insert_total, insertsize_low, insertsize_high, insertsize_median (low / high / median
// stay untouched, i.e. as passed in, when total < min_preview_spliced_reads); out_d[2] = insertsize_ave, insertsize_std
int packer_insertsize_profile(int64_t n_bundles, const int64_t *d_off, const int32_t *d, const int64_t *event, int32_t max_preview_reads,
		int32_t min_preview_spliced_reads, int32_t *out_i, double *out_d);
// the compact form of a batch for the host -> device link (agpu_batch_packed); the view points into the handle (and, for
// bundle_hit_off / bundle_tid / bundle_sample / xs / qid, into `in`).  NULL if `in` breaks the packing contract (pos
// decreasing inside a bundle) or holds an operation the units cannot express (length >= 2^24), or xs is not one of '+', '-', '.'.
void *packer_compact_create(const agpu_batch_in *in);
const agpu_batch_packed *packer_compact_view(void *c);
void packer_compact_destroy(void *c);
// packed int32 lists (off[n_lists + 1], val) re-packed in the order `order[n_out]` (e.g. the bundles' splice lists, fetched in
// bundle order, laid out bundle group by bundle group for agpu_group_resolve_batch).  out_off[n_out + 1]; out_val holds the sum of
// the chosen lists' lengths; returns that sum, or -1 when an index lies outside [0, n_lists)
int64_t packer_reorder_lists(int64_t n_lists, const int64_t *off, const int32_t *val, int64_t n_out, const int64_t *order, int64_t *out_off,
		int32_t *out_val);
// view of everything appended so far (pointers stay valid until the next add / destroy)
int packer_view(void *pk, agpu_batch_in *out);
int64_t packer_records_seen(void *pk);
// extra per-bundle info for tests: 0 = bb1 ('+' side), 1 = bb2 ('-' side)
const uint8_t *packer_bundle_side(void *pk);

#ifdef __cplusplus
}
#endif
#endif
