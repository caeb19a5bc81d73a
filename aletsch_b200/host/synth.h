// Synthetic sorted-BAM record generator "synth-v1" (SURVEY.md §8(d)).
//
// There is no network and the reference ships no data, so the benchmark and the parity tests
// run on simulated transcripts and reads.  The generator is a pure function of
// (seed, sample, record index): every draw comes from a counter-based hash, so any subset can
// be regenerated anywhere (GPU box, CPU baseline) bit-identically.
//
// Output is what htslib would hand the reference per record (bam1_core_t fields + CIGAR +
// XS/ts tag), in coordinate-sorted order per sample -- the input of meta/generator.cc:77.
#ifndef ALETSCH_B200_HOST_SYNTH_H
#define ALETSCH_B200_HOST_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SYNTH_PAIRED = 0, SYNTH_SINGLE = 1, SYNTH_LONG = 2 };

typedef struct synth_config
{
	uint64_t seed;
	int32_t mode;                 // SYNTH_PAIRED / SYNTH_SINGLE / SYNTH_LONG
	int32_t n_chrom;              // chromosomes (a 1 kb sentinel chromosome is NOT added here)
	int32_t chrom_len;            // bases per chromosome
	int32_t gene_spacing;         // one gene every ~gene_spacing bases
	int32_t read_len;             // short-read length
	int32_t min_exons, max_exons; // exons per gene, uniform
	double expressed_fraction;    // fraction of genes a sample expresses (single-cell: 0.1)
	double secondary_rate;        // extra secondary alignments per template
	double indel_rate;            // per-hit probability of one I or D op
	double clip_rate;             // per-hit probability of a soft clip
	double odd_rate;              // per-hit probability of =/X ops or a D-adjacent N
} synth_config;

// coordinate-sorted records of one sample (SoA, malloc'ed; free with synth_records_free)
typedef struct synth_records
{
	int64_t n;
	int32_t *tid;
	int32_t *pos;
	int32_t *rpos;               // pos + bam_cigar2rlen (rnacore/hit.cc:64)
	int32_t *mpos;
	int32_t *isize;
	uint16_t *flag;
	uint8_t *mapq;
	uint8_t *xs;                 // '+', '-', '.'  (hit::set_tags result, rnacore/hit.cc:106-141)
	uint64_t *qid;               // query-name key (equal <=> same template)
	uint32_t *cigar_off;         // [n+1]
	uint32_t *cigar;             // raw BAM ops
	int64_t n_cigar;
} synth_records;

void synth_default_config(synth_config *c, int mode);
void *synth_create(const synth_config *c);
void synth_destroy(void *s);
int32_t synth_num_genes(void *s);
// generate `templates` templates (pairs / single reads / long reads) of sample `sample`
int synth_generate(void *s, int32_t sample, int64_t templates, int32_t threads, synth_records *out);
void synth_records_free(synth_records *r);

#ifdef __cplusplus
}
#endif
#endif
