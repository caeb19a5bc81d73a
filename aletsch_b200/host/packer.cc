// Host-side SoA packer; see packer.h.
#include "packer.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct side
{
	// bundle_base fields the record loop needs (rnacore/bundle_base.h:31-36)
	int32_t tid = -1;
	int32_t rpos = 0;
	bool spliced = false;
	int32_t last_pos = 0, last_rpos = 0;
	std::vector<int64_t> recs;          // record indices admitted so far
	std::vector<uint8_t> strand;
	void clear() { tid = -1; rpos = 0; spliced = false; recs.clear(); strand.clear(); }
};

struct packer
{
	std::vector<int64_t> hit_off{0};
	std::vector<int32_t> b_tid, b_sample;
	std::vector<uint8_t> b_side;
	std::vector<int32_t> pos, rpos, mpos, isize;
	std::vector<uint16_t> flag;
	std::vector<uint8_t> strand, xs;
	std::vector<uint64_t> qid;
	std::vector<uint32_t> cigar_off{0};
	std::vector<uint32_t> cigar;
	int64_t seen = 0;
	// insert-size preview (packer_preview_add): per bundle the triggering record, per hit the blocks without coverage, extra intervals
	std::vector<int64_t> pv_event;
	std::vector<uint16_t> pv_skip;
	std::vector<int32_t> pv_ex_bundle, pv_ex_l, pv_ex_r, pv_ex_cnt;
	// region table of the sample added last (sample_profile::start1 / start2 / end1 / start_off, flattened over tid, rid)
	std::vector<int64_t> reg_off{0}, reg_rec;
	std::vector<int32_t> reg_start1, reg_start2, reg_end1;
};

// hit::set_strand (rnacore/hit.cc:152-185)
inline char strand_of(uint16_t flag, int libtype)
{
	char s = '.';
	bool paired = (flag & 0x1) != 0, rev = (flag & 0x10) != 0, r1 = (flag & 0x40) != 0, r2 = (flag & 0x80) != 0;
	if(libtype == AGPU_FR_FIRST && paired)
	{
		if(!rev && r1 && !r2) s = '-';
		if(rev && r1 && !r2) s = '+';
		if(!rev && !r1 && r2) s = '+';
		if(rev && !r1 && r2) s = '-';
	}
	if(libtype == AGPU_FR_SECOND && paired)
	{
		if(!rev && r1 && !r2) s = '+';
		if(rev && r1 && !r2) s = '-';
		if(!rev && !r1 && r2) s = '-';
		if(rev && !r1 && r2) s = '+';
	}
	if(libtype == AGPU_FR_FIRST && !paired) s = rev ? '+' : '-';
	if(libtype == AGPU_FR_SECOND && !paired) s = rev ? '-' : '+';
	return s;
}

inline bool has_splice(const uint32_t *c, uint32_t n)
{
	for(uint32_t k = 0; k < n; k++) if((c[k] & 0xf) == 3) return true;      // hit::contain_splices, rnacore/hit.cc:67-75
	return false;
}

inline bool has_inner_splice(const uint32_t *c, uint32_t n)
{
	for(uint32_t k = 1; k + 1 < n; k++) if((c[k] & 0xf) == 3) return true;  // hit::extract_splices, rnacore/hit.cc:91-92
	return false;
}

// bundle_base::add_hit (rnacore/bundle_base.cc:73-104)
void admit(side &s, const packer_records &r, int64_t i, char strand)
{
	if(!s.recs.empty() && s.last_pos == r.pos[i] && s.last_rpos == r.rpos[i]) return;
	s.recs.push_back(i);
	s.strand.push_back((uint8_t)strand);
	s.last_pos = r.pos[i];
	s.last_rpos = r.rpos[i];
	int32_t p = r.rpos[i];
	if(r.mpos[i] > r.rpos[i] && r.mpos[i] <= r.rpos[i] + 500000) p = r.mpos[i];
	if(p > s.rpos) s.rpos = p;
	if(s.tid == -1) s.tid = r.tid[i];
	if(has_inner_splice(r.cigar + r.cigar_off[i], r.cigar_off[i + 1] - r.cigar_off[i])) s.spliced = true;
}

// generator::generate (meta/generator.cc:203-227)
void flush(packer &pk, side &s, const packer_records &r, const packer_params &p, int32_t sample, int which)
{
	if(s.tid < 0) return;
	if(p.skip_single_exon_transcripts && !s.spliced) return;
	for(size_t k = 0; k < s.recs.size(); k++)
	{
		int64_t i = s.recs[k];
		pk.pos.push_back(r.pos[i]); pk.rpos.push_back(r.rpos[i]); pk.mpos.push_back(r.mpos[i]); pk.isize.push_back(r.isize[i]);
		pk.flag.push_back(r.flag[i]); pk.strand.push_back(s.strand[k]); pk.xs.push_back(r.xs[i]); pk.qid.push_back(r.qid[i]);
		pk.cigar.insert(pk.cigar.end(), r.cigar + r.cigar_off[i], r.cigar + r.cigar_off[i + 1]);
		pk.cigar_off.push_back((uint32_t)pk.cigar.size());
	}
	pk.hit_off.push_back((int64_t)pk.pos.size());
	pk.b_tid.push_back(s.tid);
	pk.b_sample.push_back(sample);
	pk.b_side.push_back((uint8_t)which);
}

} // namespace

extern "C" {

void packer_default_params(packer_params *p)
{
	p->library_type = AGPU_FR_FIRST;
	p->min_mapping_quality = 1;
	p->max_num_cigar = 10000;
	p->max_read_span = 500000;
	p->min_bundle_gap = 200;
	p->use_second_alignment = 1;
	p->skip_single_exon_transcripts = 1;
}

void *packer_create(void) { return new packer; }
void packer_destroy(void *pk) { delete (packer*)pk; }
int64_t packer_records_seen(void *pk) { return ((packer*)pk)->seen; }
const uint8_t *packer_bundle_side(void *pk) { return ((packer*)pk)->b_side.data(); }

// the record loop of one generator::resolve call: records [begin, ...) until the file ends or, with a region (target >= 0),
// until a record lies at or behind end1 or on another chromosome (meta/generator.cc:80-81)
static void resolve_records(packer &pk, const packer_records &r, const packer_params &p, int32_t sample, int64_t begin, int32_t target, int32_t end1)
{
	side bb1, bb2;
	int32_t pre_lpos = -1, pre_rpos = -1;
	for(int64_t i = begin; i < r.n; i++)
	{
		if(target >= 0 && (r.pos[i] >= end1 || r.tid[i] != target)) break;
		pk.seen++;
		uint16_t fl = r.flag[i];
		uint32_t nc = r.cigar_off[i + 1] - r.cigar_off[i];
		// meta/generator.cc:87-97
		if((fl & 0x4) >= 1) continue;
		if((fl & 0x100) >= 1 && !p.use_second_alignment) continue;
		if((int64_t)nc > (int64_t)p.max_num_cigar) continue;
		if((int32_t)r.mapq[i] < p.min_mapping_quality) continue;
		if(nc < 1) continue;
		if(std::fabs((double)r.pos[i] - (double)r.rpos[i]) >= p.max_read_span) continue;
		if(((fl & 0x8) <= 0) && std::fabs((double)r.pos[i] - (double)r.mpos[i]) >= p.max_read_span) continue;
		if(r.pos[i] == pre_lpos && r.rpos[i] == pre_rpos) continue;
		pre_lpos = r.pos[i];
		pre_rpos = r.rpos[i];

		char xs = (char)r.xs[i];
		char st = strand_of(fl, p.library_type);

		// meta/generator.cc:107-135: close bundles left behind by more than min_bundle_gap
		if(!bb1.recs.empty() && (r.tid[i] != bb1.tid || r.pos[i] > bb1.rpos + p.min_bundle_gap)) { flush(pk, bb1, r, p, sample, 0); bb1.clear(); }
		if(!bb2.recs.empty() && (r.tid[i] != bb2.tid || r.pos[i] > bb2.rpos + p.min_bundle_gap)) { flush(pk, bb2, r, p, sample, 1); bb2.clear(); }

		// meta/generator.cc:155-179: strand routing
		if(p.library_type != AGPU_UNSTRANDED)
		{
			if(st == '+' && xs == '-') continue;
			if(st == '-' && xs == '+') continue;
			if(st == '.' && xs != '.') st = xs;
			if(st == '+') admit(bb1, r, i, st);
			if(st == '-') admit(bb2, r, i, st);
		}
		else
		{
			if(xs == '+') admit(bb1, r, i, st);
			if(xs == '-') admit(bb2, r, i, st);
			if(xs == '.' && !has_splice(r.cigar + r.cigar_off[i], nc)) { admit(bb1, r, i, st); admit(bb2, r, i, st); }
		}
	}
	flush(pk, bb1, r, p, sample, 0);
	flush(pk, bb2, r, p, sample, 1);
}

int packer_add_sample(void *pkp, const packer_records *rp, const packer_params *pp, int32_t sample)
{
	resolve_records(*(packer*)pkp, *rp, *pp, sample, 0, -1, 0);
	return 0;
}

// sample_profile::set_batch_boundaries (rnacore/sample_profile.cc:167-252): one pass over the mapped records cuts every
// chromosome into regions that start where a gap of more than min_bundle_gap crosses a multiple of region_length.  The
// table keeps the reference's quirks: the offset of a region is taken AFTER its first hit was read, so that hit is
// never seen by generator::resolve; end1 is written only when a later region or chromosome starts, so the last region of the
// last chromosome with records stays closed (start1 >= end1); regions that never start have start1 = end1 = 0.
int64_t packer_region_table(void *pkp, const packer_records *rp, int32_t n_chrom, const int32_t *chrom_len, int32_t region_length,
		const packer_params *pp)
{
	packer &pk = *(packer*)pkp;
	const packer_records &r = *rp;
	pk.reg_off.assign(1, 0);
	for(int t = 0; t < n_chrom; t++) pk.reg_off.push_back(pk.reg_off.back() + chrom_len[t] / region_length + 1);
	const int64_t nr = pk.reg_off.back();
	pk.reg_rec.assign(nr, 0); pk.reg_start1.assign(nr, 0); pk.reg_start2.assign(nr, 0); pk.reg_end1.assign(nr, 0);
	int32_t tid = -1, rid = 0, rpos = 0;
	for(int64_t i = 0; i < r.n; i++)
	{
		if((r.flag[i] & 0x4) >= 1) continue;
		if(std::fabs((double)r.pos[i] - (double)r.rpos[i]) >= pp->max_read_span) continue;
		if(r.tid[i] < 0 || r.tid[i] >= n_chrom) return -1;
		if(r.tid[i] != tid)
		{
			if(tid >= 0) pk.reg_end1[pk.reg_off[tid] + rid] = rpos;
			tid = r.tid[i];
			rid = 0;
			int64_t k = pk.reg_off[tid];
			pk.reg_start1[k] = r.pos[i]; pk.reg_start2[k] = r.rpos[i]; pk.reg_rec[k] = i + 1;
			rpos = r.rpos[i];
		}
		if(r.pos[i] > rpos + pp->min_bundle_gap && (int64_t)r.pos[i] >= (int64_t)region_length * (1 + rid))
		{
			pk.reg_end1[pk.reg_off[tid] + rid] = rpos;
			rid = r.pos[i] / region_length;
			if(pk.reg_off[tid] + rid >= pk.reg_off[tid + 1]) return -1;       // a record behind the end of its chromosome
			int64_t k = pk.reg_off[tid] + rid;
			pk.reg_start1[k] = r.pos[i]; pk.reg_start2[k] = r.rpos[i]; pk.reg_rec[k] = i + 1;
		}
		if(r.rpos[i] > rpos) rpos = r.rpos[i];
	}
	return nr;
}

int packer_regions(void *pkp, const int64_t **reg_off, const int32_t **start1, const int32_t **start2, const int32_t **end1, const int64_t **start_rec)
{
	packer &pk = *(packer*)pkp;
	*reg_off = pk.reg_off.data(); *start1 = pk.reg_start1.data(); *start2 = pk.reg_start2.data(); *end1 = pk.reg_end1.data();
	*start_rec = pk.reg_rec.data();
	return (int)pk.reg_off.size() - 1;
}

// what incubator::generate posts per sample (meta/incubator.cc:355-380): one generator::resolve per region of the table
// whose start1 < end1, chromosome by chromosome; the bundles of a region never see the records of another
int packer_add_sample_regions(void *pkp, const packer_records *rp, const packer_params *pp, int32_t sample)
{
	packer &pk = *(packer*)pkp;
	const int nt = (int)pk.reg_off.size() - 1;
	for(int t = 0; t < nt; t++)
		for(int64_t k = pk.reg_off[t]; k < pk.reg_off[t + 1]; k++)
		{
			if(pk.reg_start1[k] >= pk.reg_end1[k]) continue;
			resolve_records(pk, *rp, *pp, sample, pk.reg_rec[k], t, pk.reg_end1[k]);
		}
	return 0;
}

int packer_view(void *pkp, agpu_batch_in *out)
{
	packer &pk = *(packer*)pkp;
	out->n_bundles = (int32_t)pk.b_tid.size();
	out->n_hits = (int64_t)pk.pos.size();
	out->n_cigar = (int64_t)pk.cigar.size();
	out->bundle_hit_off = pk.hit_off.data();
	out->bundle_tid = pk.b_tid.data();
	out->bundle_sample = pk.b_sample.data();
	out->pos = pk.pos.data(); out->rpos = pk.rpos.data(); out->mpos = pk.mpos.data(); out->isize = pk.isize.data();
	out->flag = pk.flag.data(); out->strand = pk.strand.data(); out->xs = pk.xs.data(); out->qid = pk.qid.data();
	out->cigar_off = pk.cigar_off.data(); out->cigar = pk.cigar.data();
	out->bundle_strand = NULL;
	return 0;
}


// previewer::infer_library_type (meta/previewer.cc:29-148): a pass over the first records that are primary, mapped and carry an
// XS / ts strand; every spliced one votes "first" when the strand its flags predict under fr-firststrand equals the tag,
// "second" when it is the opposite.  out[0] = library type, out[1] = bam_with_xs, out[2] = reads looked at, out[3] = spliced,
// out[4] = with xs, out[5] = used (spn), out[6] = first, out[7] = second
int packer_infer_library_type(const packer_records *rp, const packer_params *pp, int32_t max_preview_reads, int32_t max_preview_spliced_reads,
		int32_t min_preview_spliced_reads, double preview_infer_ratio, int32_t *out)
{
	const packer_records &r = *rp;
	int total = 0, spliced = 0, num_xs = 0, first = 0, second = 0;
	int64_t n1 = 0, n2 = 0;
	for(int64_t i = 0; i < r.n; i++)
	{
		if(total >= max_preview_reads) break;
		if(n1 >= max_preview_spliced_reads && n2 >= max_preview_spliced_reads) break;
		const uint16_t fl = r.flag[i];
		const uint32_t nc = r.cigar_off[i + 1] - r.cigar_off[i];
		if((fl & 0x4) >= 1) continue;
		if((fl & 0x100) >= 1) continue;
		if((int64_t)nc > (int64_t)pp->max_num_cigar) continue;
		if((int32_t)r.mapq[i] < pp->min_mapping_quality) continue;
		if(nc < 1) continue;
		total++;
		if(!has_inner_splice(r.cigar + r.cigar_off[i], nc)) continue;
		spliced++;
		const char xs = (char)r.xs[i];
		if(xs == '.') continue;
		num_xs++;
		if(xs == '+' && n1 >= max_preview_spliced_reads) continue;
		if(xs == '-' && n2 >= max_preview_spliced_reads) continue;
		const bool paired = (fl & 0x1) != 0, rev = (fl & 0x10) != 0, mrev = (fl & 0x20) != 0, r1 = (fl & 0x40) != 0, r2 = (fl & 0x80) != 0;
		char p = '.';
		if(paired && !rev && mrev && r1 && !r2) p = '-';
		if(paired && rev && !mrev && !r1 && r2) p = '-';
		if(paired && rev && !mrev && r1 && !r2) p = '+';
		if(paired && !rev && mrev && !r1 && r2) p = '+';
		if(!paired && !rev) p = '-';
		if(!paired && rev) p = '+';
		if(p == '+') { n1++; if(p == xs) first++; else second++; }
		if(p == '-') { n2++; if(p == xs) first++; else second++; }
	}
	const int spn = (int)((n1 + n2) / 2.0);
	int lt = AGPU_UNSTRANDED;
	if(spn >= min_preview_spliced_reads && first > preview_infer_ratio * 2.0 * spn) lt = AGPU_FR_FIRST;
	if(spn >= min_preview_spliced_reads && second > preview_infer_ratio * 2.0 * spn) lt = AGPU_FR_SECOND;
	out[0] = lt;
	out[1] = (spliced > 0 && num_xs * 1.0 / spliced > preview_infer_ratio) ? 1 : 0;
	out[2] = total; out[3] = spliced; out[4] = num_xs; out[5] = spn; out[6] = first; out[7] = second;
	return 0;
}

// ---- insert-size preview: previewer::infer_insertsize (meta/previewer.cc:151-304) --------------------------------------------
namespace {

struct preview_side
{
	// bundle_base as the previewer uses it
	int32_t tid = -1, rpos = 0;
	std::vector<int64_t> recs;
	std::vector<uint8_t> strand;
	std::vector<uint8_t> nblk;             // BAM_CMATCH operations of every stored hit
	int32_t last_pos = 0, last_rpos = 0;
	// interval_buf / interval_cnt (rnacore/bundle_base.h:43-44): never flushed, never cleared by the previewer
	int32_t buf[10][2], cnt[10];
	int64_t run_owner[10];                 // serial number of the bundle the buffered run belongs to
	int run_first[10];                     // index (in recs) of the first hit of the run
	int64_t serial = 0;                    // serial number of the bundle being filled
	std::vector<int32_t> ex_l, ex_r, ex_cnt;   // runs of earlier bundles flushed into this one
	preview_side() { for(int z = 0; z < 10; z++) { buf[z][0] = buf[z][1] = -1; cnt[z] = 0; run_owner[z] = -1; run_first[z] = 0; } }
	void clear() { tid = -1; rpos = 0; recs.clear(); strand.clear(); nblk.clear(); ex_l.clear(); ex_r.clear(); ex_cnt.clear(); serial++; }
};

// bundle_base::add_hit_intervals as far as the preview needs it: add_hit + the buffer bookkeeping of add_intervals
void preview_admit(preview_side &s, const packer_records &r, int64_t i, char strand)
{
	if(!s.recs.empty() && s.last_pos == r.pos[i] && s.last_rpos == r.rpos[i]) return;
	s.recs.push_back(i);
	s.strand.push_back((uint8_t)strand);
	s.last_pos = r.pos[i]; s.last_rpos = r.rpos[i];
	int32_t q = r.rpos[i];
	if(r.mpos[i] > r.rpos[i] && r.mpos[i] <= r.rpos[i] + 500000) q = r.mpos[i];
	if(q > s.rpos) s.rpos = q;
	if(s.tid == -1) s.tid = r.tid[i];
	int32_t p = r.pos[i];
	int z = 0;
	for(uint32_t k = r.cigar_off[i]; k < r.cigar_off[i + 1]; k++)
	{
		const uint32_t c = r.cigar[k], op = c & 0xf, len = c >> 4;
		if((0x3C1A7 >> (op << 1)) & 2) p += (int32_t)len;
		if(op != 0) continue;
		const int32_t b0 = p - (int32_t)len;
		if(z < 10)
		{
			if(b0 == s.buf[z][0] && p == s.buf[z][1]) s.cnt[z]++;
			else
			{
				// the buffered run goes into the map of the bundle being filled NOW: its own blocks simply become coverage; a run of
				// an earlier bundle is a foreign interval
				if(s.buf[z][0] != -1 && s.buf[z][1] != -1 && s.run_owner[z] != s.serial)
				{ s.ex_l.push_back(s.buf[z][0]); s.ex_r.push_back(s.buf[z][1]); s.ex_cnt.push_back(s.cnt[z]); }
				s.buf[z][0] = b0; s.buf[z][1] = p; s.cnt[z] = 1;
				s.run_owner[z] = s.serial; s.run_first[z] = (int)s.recs.size() - 1;
			}
		}
		z++;
	}
	s.nblk.push_back((uint8_t)(z > 255 ? 255 : z));
}

// previewer::process as far as the host decides it: size limits; the bundle goes to the batch with its skip masks
void preview_flush(packer &pk, preview_side &s, const packer_records &r, int32_t min_hits, int64_t event, int which, int64_t &added)
{
	if((int64_t)s.recs.size() < (int64_t)min_hits || s.recs.size() > 20000 || s.tid < 0) return;
	const int32_t bundle = (int32_t)pk.b_tid.size();
	for(size_t k = 0; k < s.recs.size(); k++)
	{
		const int64_t i = s.recs[k];
		pk.pos.push_back(r.pos[i]); pk.rpos.push_back(r.rpos[i]); pk.mpos.push_back(r.mpos[i]); pk.isize.push_back(r.isize[i]);
		pk.flag.push_back(r.flag[i]); pk.strand.push_back(s.strand[k]); pk.xs.push_back(r.xs[i]); pk.qid.push_back(r.qid[i]);
		pk.cigar.insert(pk.cigar.end(), r.cigar + r.cigar_off[i], r.cigar + r.cigar_off[i + 1]);
		pk.cigar_off.push_back((uint32_t)pk.cigar.size());
		uint16_t mask = 0;
		for(int z = 0; z < 10; z++)
			if(s.run_owner[z] == s.serial && (int)k >= s.run_first[z] && s.nblk[k] > z) mask |= (uint16_t)(1u << z);   // still in the buffer
		pk.pv_skip.push_back(mask);
	}
	pk.hit_off.push_back((int64_t)pk.pos.size());
	pk.b_tid.push_back(s.tid);
	pk.b_sample.push_back(0);
	pk.b_side.push_back((uint8_t)which);
	pk.pv_event.push_back(event);
	for(size_t x = 0; x < s.ex_l.size(); x++)
	{ pk.pv_ex_bundle.push_back(bundle); pk.pv_ex_l.push_back(s.ex_l[x]); pk.pv_ex_r.push_back(s.ex_r[x]); pk.pv_ex_cnt.push_back(s.ex_cnt[x]); }
	added++;
}

} // namespace

int64_t packer_preview_add(void *pkp, const packer_records *rp, const packer_params *pp, int32_t min_num_hits_in_bundle)
{
	packer &pk = *(packer*)pkp;
	const packer_records &r = *rp;
	const packer_params &p = *pp;
	preview_side bb1, bb2;
	int64_t added = 0;
	// hits that were given no skip entry by earlier (ordinary) additions keep a zero mask
	pk.pv_skip.resize(pk.pos.size(), 0);
	pk.pv_event.resize(pk.b_tid.size(), -1);
	for(int64_t i = 0; i < r.n; i++)
	{
		const uint16_t fl = r.flag[i];
		const uint32_t nc = r.cigar_off[i + 1] - r.cigar_off[i];
		if((fl & 0x4) >= 1) continue;
		if((fl & 0x100) >= 1) continue;
		if((int64_t)nc > (int64_t)p.max_num_cigar) continue;
		if((int32_t)r.mapq[i] < p.min_mapping_quality) continue;
		if(nc < 1) continue;
		char strand = strand_of(fl, p.library_type);
		const char xs = (char)r.xs[i];
		// truncate (meta/previewer.cc:178-189): process() on a closed bundle, bb1 first
		if(r.tid[i] != bb1.tid || r.pos[i] > bb1.rpos + p.min_bundle_gap) { preview_flush(pk, bb1, r, min_num_hits_in_bundle, i, 0, added); bb1.clear(); }
		if(r.tid[i] != bb2.tid || r.pos[i] > bb2.rpos + p.min_bundle_gap) { preview_flush(pk, bb2, r, min_num_hits_in_bundle, i, 1, added); bb2.clear(); }
		// (the `cnt >= max_preview_reads` break needs the device's counts: packer_insertsize_profile replays it over the events)
		if(p.library_type != AGPU_UNSTRANDED && strand == '+' && xs == '-') continue;
		if(p.library_type != AGPU_UNSTRANDED && strand == '-' && xs == '+') continue;
		if(p.library_type != AGPU_UNSTRANDED && strand == '.' && xs != '.') strand = xs;
		if(p.library_type != AGPU_UNSTRANDED && strand == '+') preview_admit(bb1, r, i, strand);
		if(p.library_type != AGPU_UNSTRANDED && strand == '-') preview_admit(bb2, r, i, strand);
		if(p.library_type == AGPU_UNSTRANDED && xs == '.') { preview_admit(bb1, r, i, strand); preview_admit(bb2, r, i, strand); }
		if(p.library_type == AGPU_UNSTRANDED && xs == '+') preview_admit(bb1, r, i, strand);
		if(p.library_type == AGPU_UNSTRANDED && xs == '-') preview_admit(bb2, r, i, strand);
	}
	// the bundles still open when the file ends are never processed (meta/previewer.cc:206-208)
	return added;
}

int packer_preview_view(void *pkp, const int64_t **event, const uint16_t **skip, int64_t *n_extra, const int32_t **ex_bundle,
		const int32_t **ex_l, const int32_t **ex_r, const int32_t **ex_cnt)
{
	packer &pk = *(packer*)pkp;
	pk.pv_skip.resize(pk.pos.size(), 0);
	pk.pv_event.resize(pk.b_tid.size(), -1);
	*event = pk.pv_event.data(); *skip = pk.pv_skip.data();
	*n_extra = (int64_t)pk.pv_ex_l.size();
	*ex_bundle = pk.pv_ex_bundle.data(); *ex_l = pk.pv_ex_l.data(); *ex_r = pk.pv_ex_r.data(); *ex_cnt = pk.pv_ex_cnt.data();
	return 0;
}

int packer_insertsize_profile(int64_t n_bundles, const int64_t *d_off, const int32_t *d, const int64_t *event, int32_t max_preview_reads,
		int32_t min_preview_spliced_reads, int32_t *out_i, double *out_d)
{
	// the histogram m of previewer::infer_insertsize, filled bundle by bundle until the break (checked once per record, i.e.
	// after all the bundles one record closed)
	std::vector<std::pair<int32_t, int> > vv;
	{
		std::vector<int32_t> all;
		int64_t cnt = 0;
		for(int64_t b = 0; b < n_bundles; )
		{
			int64_t e = b;
			while(e < n_bundles && event[e] == event[b]) e++;
			for(int64_t x = d_off[b]; x < d_off[e]; x++) all.push_back(d[x]);
			cnt += d_off[e] - d_off[b];
			b = e;
			if(cnt >= max_preview_reads) break;
		}
		std::sort(all.begin(), all.end());
		for(size_t i = 0; i < all.size(); )
		{
			size_t j = i;
			while(j < all.size() && all[j] == all[i]) j++;
			vv.push_back(std::make_pair(all[i], (int)(j - i)));
			i = j;
		}
	}
	int total = 0;
	for(size_t k = 0; k < vv.size(); k++) total += vv[k].second;
	out_i[0] = total;
	if(total < min_preview_spliced_reads) return 0;
	int n = 0;
	double sx2 = 0, ave = 0;
	int low = -1, high = -1, median = -1;
	for(size_t k = 0; k < vv.size(); k++)
	{
		n += vv[k].second;
		if(n >= 0.5 * total && median < 0) median = vv[k].first;
		ave += vv[k].second * vv[k].first;
		sx2 += vv[k].second * vv[k].first * vv[k].first;
		if(low == -1 && n >= 0.005 * total) low = vv[k].first;
		if(high == -1 && n >= 0.990 * total) high = vv[k].first;
		if(n >= 0.998 * total) break;
	}
	ave = ave * 1.0 / n;
	out_i[1] = low; out_i[2] = high; out_i[3] = median;
	out_d[0] = ave;
	out_d[1] = sqrt((sx2 - n * ave * ave) * 1.0 / n);
	return 0;
}

// ---- compact form for the host -> device link (agpu_batch_packed, include/aletsch_gpu.h) --------------------------------
struct compact
{
	agpu_batch_packed v;
	std::vector<int32_t> pos0, ev_pos, ev_mpos, ev_isize, ev_units;
	std::vector<uint16_t> dpos, units;
	std::vector<int16_t> dmpos, is16;
	std::vector<int64_t> ei_pos, ei_mpos, ei_isize, ei_units;
	std::vector<uint8_t> bstrand, meta;
};

void *packer_compact_create(const agpu_batch_in *in)
{
	if(!in || (!in->strand && !in->bundle_strand)) return NULL;
	compact *c = new compact;
	const int nb = in->n_bundles;
	const int64_t nh = in->n_hits;
	c->pos0.assign(nb, 0); c->bstrand.assign(nb, (uint8_t)'.');
	c->dpos.resize(nh); c->dmpos.resize(nh); c->is16.resize(nh); c->meta.resize(nh);
	c->units.reserve((size_t)in->n_cigar + (size_t)in->n_cigar / 4);
	// the most common one-operation CIGAR of the batch travels once (default_unit) instead of once per hit
	uint32_t def_unit = 0;
	{
		std::vector<uint32_t> singles;
		for(int64_t i = 0; i < nh; i++)
			if(in->cigar_off[i + 1] - in->cigar_off[i] == 1 && (in->cigar[in->cigar_off[i]] >> 4) < 4096 && (in->cigar[in->cigar_off[i]] & 0xf) != 15)
				singles.push_back(in->cigar[in->cigar_off[i]]);
		std::sort(singles.begin(), singles.end());
		size_t best = 0;
		for(size_t x = 0; x < singles.size(); )
		{
			size_t y = x;
			while(y < singles.size() && singles[y] == singles[x]) y++;
			if(y - x > best) { best = y - x; def_unit = singles[x]; }
			x = y;
		}
		if(best == 0) def_unit = 0xFFFFFFFFu;      // no hit uses it
	}
	for(int b = 0; b < nb; b++)
	{
		const int64_t h0 = in->bundle_hit_off[b], h1 = in->bundle_hit_off[b + 1];
		if(in->bundle_strand) c->bstrand[b] = in->bundle_strand[b];
		else if(h1 > h0) c->bstrand[b] = in->strand[h0];
		if(h1 > h0) c->pos0[b] = in->pos[h0];
		for(int64_t i = h0; i < h1; i++)
		{
			int64_t d = i == h0 ? 0 : (int64_t)in->pos[i] - (int64_t)in->pos[i - 1];
			if(d < 0 || d > 0x7fffffff) { delete c; return NULL; }        // the packing contract: pos is non-decreasing in a bundle
			if(d >= 0xFFFF) { c->dpos[i] = 0xFFFF; c->ei_pos.push_back(i); c->ev_pos.push_back((int32_t)d); }
			else c->dpos[i] = (uint16_t)d;
			int64_t m = (int64_t)in->mpos[i] - (int64_t)in->pos[i];
			if(m <= -32768 || m > 32767) { c->dmpos[i] = (int16_t)-32768; c->ei_mpos.push_back(i); c->ev_mpos.push_back(in->mpos[i]); }
			else c->dmpos[i] = (int16_t)m;
			int32_t s = in->isize[i];
			if(s <= -32768 || s > 32767) { c->is16[i] = (int16_t)-32768; c->ei_isize.push_back(i); c->ev_isize.push_back(s); }
			else c->is16[i] = (int16_t)s;
			size_t u0 = c->units.size();
			const bool is_default = in->cigar_off[i + 1] - in->cigar_off[i] == 1 && in->cigar[in->cigar_off[i]] == def_unit;
			for(uint32_t k = in->cigar_off[i]; k < in->cigar_off[i + 1] && !is_default; k++)
			{
				uint32_t op = in->cigar[k] & 0xf, len = in->cigar[k] >> 4;
				if(op == 15 || len >= (1u << 24)) { delete c; return NULL; }
				if(len < 4096) c->units.push_back((uint16_t)(len << 4 | op));
				else { c->units.push_back((uint16_t)((len & 0xfff) << 4 | 15)); c->units.push_back((uint16_t)((len >> 12) << 4 | op)); }
			}
			const size_t nu = c->units.size() - u0;
			const uint8_t x = in->xs[i] == '+' ? 1 : (in->xs[i] == '-' ? 2 : (in->xs[i] == '.' ? 0 : 3));
			if(nu > 0x7fffffff || x == 3) { delete c; return NULL; }
			if(is_default) c->meta[i] = (uint8_t)(62 | x << 6);
			else if(nu >= 62) { c->meta[i] = (uint8_t)(63 | x << 6); c->ei_units.push_back(i); c->ev_units.push_back((int32_t)nu); }
			else c->meta[i] = (uint8_t)(nu | x << 6);
		}
	}
	agpu_batch_packed &v = c->v;
	memset(&v, 0, sizeof(v));
	v.n_bundles = nb; v.n_hits = nh; v.n_cigar = in->n_cigar; v.n_units = (int64_t)c->units.size();
	v.default_unit = def_unit == 0xFFFFFFFFu ? 0 : def_unit;
	v.bundle_hit_off = in->bundle_hit_off; v.bundle_tid = in->bundle_tid; v.bundle_sample = in->bundle_sample;
	v.bundle_strand = c->bstrand.data(); v.bundle_pos0 = c->pos0.data();
	v.dpos = c->dpos.data(); v.dmpos = c->dmpos.data(); v.isize16 = c->is16.data(); v.qid = in->qid;
	v.hit_meta = c->meta.data(); v.units = c->units.data();
	v.n_esc_pos = (int64_t)c->ei_pos.size(); v.n_esc_mpos = (int64_t)c->ei_mpos.size(); v.n_esc_isize = (int64_t)c->ei_isize.size();
	v.n_esc_units = (int64_t)c->ei_units.size(); v.esc_units_idx = c->ei_units.data(); v.esc_units_val = c->ev_units.data();
	v.esc_pos_idx = c->ei_pos.data(); v.esc_pos_val = c->ev_pos.data();
	v.esc_mpos_idx = c->ei_mpos.data(); v.esc_mpos_val = c->ev_mpos.data();
	v.esc_isize_idx = c->ei_isize.data(); v.esc_isize_val = c->ev_isize.data();
	return c;
}

const agpu_batch_packed *packer_compact_view(void *c) { return c ? &((compact*)c)->v : NULL; }
void packer_compact_destroy(void *c) { delete (compact*)c; }

int64_t packer_reorder_lists(int64_t n_lists, const int64_t *off, const int32_t *val, int64_t n_out, const int64_t *order, int64_t *out_off,
		int32_t *out_val)
{
	int64_t w = 0;
	out_off[0] = 0;
	for(int64_t k = 0; k < n_out; k++)
	{
		const int64_t l = order[k];
		if(l < 0 || l >= n_lists) return -1;
		const int64_t len = off[l + 1] - off[l];
		if(len > 0) memcpy(out_val + w, val + off[l], sizeof(int32_t) * (size_t)len);
		w += len;
		out_off[k + 1] = w;
	}
	return w;
}

}
