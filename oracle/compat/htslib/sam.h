/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the part of htslib's sam.h that the
 * reference's hot-path translation units touch.  htslib is not installed in the build
 * container.  Layouts and macros follow the published BAM record layout (SAM/BAM spec
 * section 4.2 and htslib >= 1.10 sam.h: bam1_core_t with l_extranul, 4-byte CIGAR ops
 * "len << 4 | op", BAM_CIGAR_TYPE = 0x3C1A7).
 *
 * Call sites (reference file:line):
 *   rnacore/hit.cc:52-65        bam_get_qname, bam_get_cigar, bam_cigar2rlen
 *   rnacore/hit.cc:77-104       bam_cigar_op / oplen / type, BAM_CREF_SKIP
 *   rnacore/hit.cc:106-141      bam_aux_get, bam_aux2A, bam_aux2i
 *   rnacore/bundle_base.cc:106-158  BAM_CMATCH / BAM_CINS / BAM_CDEL
 *   rnacore/essential.cc:491-700    bam_aux_append, bam_write1 (dead BAM writers)
 *   rnacore/sample_profile.cc, meta/generator.cc: file API (served from memory, hts_shim.cc)
 */
#ifndef ALETSCH_B200_ORACLE_COMPAT_HTSLIB_SAM_H
#define ALETSCH_B200_ORACLE_COMPAT_HTSLIB_SAM_H

#include <stdint.h>
#include <stddef.h>
#include <sys/types.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>

typedef struct bam1_core_t {
	int32_t tid;
	int32_t pos;
	uint16_t bin;
	uint8_t qual;
	uint8_t l_qname;
	uint16_t flag;
	uint8_t unused1;
	uint8_t l_extranul;
	uint32_t n_cigar;
	int32_t l_qseq;
	int32_t mtid;
	int32_t mpos;
	int32_t isize;
} bam1_core_t;

typedef struct bam1_t {
	bam1_core_t core;
	int l_data;
	uint32_t m_data;
	uint8_t *data;
	uint64_t id;
} bam1_t;

typedef struct bam_hdr_t {
	int32_t n_targets;
	uint32_t *target_len;
	char **target_name;
} bam_hdr_t;

typedef struct BGZF { int64_t pos; } BGZF;
typedef struct htsFile {
	union { BGZF *bgzf; void *other; } fp;
	int store;                 /* index into the in-memory file registry (hts_shim.cc) */
} htsFile;
typedef htsFile samFile;
typedef struct hts_idx_t { int unused; } hts_idx_t;
typedef struct hts_itr_t { int unused; } hts_itr_t;

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CBACK       9

#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define BAM_CIGAR_TYPE  0x3C1A7

#define bam_cigar_op(c) ((c)&BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c)>>BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE>>((o)<<1)&3)

#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (b)->core.l_qseq + (((b)->core.l_qseq + 1)>>1))
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar<<2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1)>>1))

static inline int64_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar)
{
	int64_t l = 0;
	for(int k = 0; k < n_cigar; ++k)
		if(bam_cigar_type(bam_cigar_op(cigar[k])) & 2) l += bam_cigar_oplen(cigar[k]);
	return l;
}

/* aux access: tag(2) type(1) value; integer types cCsSiI, 'A' printable char, 'Z' string */
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);
char bam_aux2A(const uint8_t *s);
int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data);

/* file API, served from an in-memory record store */
samFile *sam_open(const char *fn, const char *mode);
int sam_close(samFile *fp);
bam_hdr_t *sam_hdr_read(samFile *fp);
void bam_hdr_destroy(bam_hdr_t *h);
hts_idx_t *sam_index_load(samFile *fp, const char *fn);
void hts_idx_destroy(hts_idx_t *idx);
void hts_itr_destroy(hts_itr_t *it);
bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
int sam_read1(samFile *fp, bam_hdr_t *h, bam1_t *b);
int bam_write1(BGZF *fp, const bam1_t *b);

#endif
