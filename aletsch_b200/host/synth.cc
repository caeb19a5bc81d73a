// Synthetic sorted-BAM record generator "synth-v1"; see synth.h.
#include "synth.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline uint64_t mix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

// counter-based stream: every draw is a hash of (key, counter)
struct rng
{
	uint64_t key;
	uint64_t ctr;
	rng(uint64_t a, uint64_t b, uint64_t c) : key(mix64(mix64(mix64(a) ^ b) ^ c)), ctr(0) {}
	uint64_t next() { return mix64(key ^ (0xD1B54A32D192ED03ULL * ++ctr)); }
	double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
	int32_t range(int32_t lo, int32_t hi) { return lo + (int32_t)(next() % (uint64_t)(hi - lo + 1)); }   // inclusive
	double normal()
	{
		double u1 = uniform(), u2 = uniform();
		if(u1 < 1e-300) u1 = 1e-300;
		return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
	}
};

struct isoform
{
	std::vector<int32_t> exons;   // exon indices of the gene, ascending
	std::vector<int32_t> cum;     // cumulative transcript length at the start of each exon (+ total at the end)
};

struct gene
{
	int32_t tid;
	char strand;
	std::vector<int32_t> el, er;  // exon [el, er)
	std::vector<isoform> iso;
};

struct model
{
	synth_config cfg;
	std::vector<gene> genes;
	std::vector<int32_t> iso_gene, iso_idx;   // flattened isoform table
};

void build_model(model &m)
{
	const synth_config &c = m.cfg;
	for(int32_t tid = 0; tid < c.n_chrom; tid++)
	{
		int32_t ng = c.chrom_len / c.gene_spacing;
		for(int32_t g = 0; g < ng; g++)
		{
			rng r(c.seed, 0x67656E65ULL + tid, g);
			gene ge;
			ge.tid = tid;
			ge.strand = (r.next() & 1) ? '+' : '-';
			int32_t p = 10000 + g * c.gene_spacing + r.range(0, c.gene_spacing / 5);
			int32_t ne = r.range(c.min_exons, c.max_exons);
			for(int32_t e = 0; e < ne; e++)
			{
				int32_t len = r.range(80, 400);
				ge.el.push_back(p);
				ge.er.push_back(p + len);
				p += len;
				// intron log-uniform in [100, 20000]
				double li = std::exp(std::log(100.0) + r.uniform() * (std::log(20000.0) - std::log(100.0)));
				p += (int32_t)li;
			}
			if(ge.er.back() + 1000 >= c.chrom_len) continue;
			int32_t ni = r.range(1, 5);
			for(int32_t k = 0; k < ni; k++)
			{
				isoform is;
				int32_t first = 0, last = ne - 1;
				if(k > 0 && ne >= 4 && r.uniform() < 0.2) first = 1;
				if(k > 0 && ne >= 4 && r.uniform() < 0.2) last = ne - 2;
				for(int32_t e = first; e <= last; e++)
				{
					bool skip = (k > 0 && e != first && e != last && r.uniform() < 0.3);
					if(!skip) is.exons.push_back(e);
				}
				if(is.exons.size() < 2) continue;
				bool dup = false;
				for(size_t j = 0; j < ge.iso.size(); j++) if(ge.iso[j].exons == is.exons) dup = true;
				if(dup) continue;
				int32_t t = 0;
				for(size_t j = 0; j < is.exons.size(); j++)
				{
					is.cum.push_back(t);
					t += ge.er[is.exons[j]] - ge.el[is.exons[j]];
				}
				is.cum.push_back(t);
				ge.iso.push_back(is);
			}
			m.genes.push_back(ge);
		}
	}
	for(size_t g = 0; g < m.genes.size(); g++)
		for(size_t k = 0; k < m.genes[g].iso.size(); k++)
		{
			m.iso_gene.push_back((int32_t)g);
			m.iso_idx.push_back((int32_t)k);
		}
}

struct hitrec
{
	int32_t tid, pos, rpos, mpos, isize;
	uint16_t flag;
	uint8_t xs;
	uint64_t qid;
	uint32_t ncig;
};

struct outbuf
{
	std::vector<hitrec> hits;
	std::vector<uint32_t> cigar;
};

inline uint32_t op(uint32_t len, uint32_t o) { return (len << 4) | o; }

// genomic blocks of transcript interval [a, b) of an isoform
void project(const gene &ge, const isoform &is, int32_t a, int32_t b, std::vector<int32_t> &bl, std::vector<int32_t> &br)
{
	bl.clear(); br.clear();
	size_t k = std::upper_bound(is.cum.begin(), is.cum.end(), a) - is.cum.begin() - 1;
	for(; k < is.exons.size() && is.cum[k] < b; k++)
	{
		int32_t e = is.exons[k];
		int32_t lo = std::max(a, is.cum[k]) - is.cum[k];
		int32_t hi = std::min(b, is.cum[k + 1]) - is.cum[k];
		if(lo >= hi) continue;
		int32_t gl = ge.el[e] + lo, gr = ge.el[e] + hi;
		if(!br.empty() && br.back() == gl) br.back() = gr;
		else { bl.push_back(gl); br.push_back(gr); }
	}
}

// CIGAR of an alignment whose matched blocks are bl/br; returns pos/rpos and whether spliced
void emit_cigar(const synth_config &c, rng &r, std::vector<int32_t> &bl, std::vector<int32_t> &br, bool allow_mods,
		std::vector<uint32_t> &cig, int32_t &pos, int32_t &rpos, bool &spliced)
{
	cig.clear();
	size_t nb = bl.size();
	spliced = nb >= 2;
	int mod = 0;            // 1 indel, 2 clip, 3 odd
	if(allow_mods)
	{
		double u = r.uniform();
		if(u < c.indel_rate) mod = 1;
		else if(u < c.indel_rate + c.clip_rate) mod = 2;
		else if(u < c.indel_rate + c.clip_rate + c.odd_rate) mod = 3;
	}
	size_t mb = nb ? (size_t)(r.next() % nb) : 0;
	bool coin = (r.next() & 1) != 0;
	int32_t k = r.range(1, 3);
	int32_t front_clip = 0, back_clip = 0;
	if(mod == 2)
	{
		int32_t s = r.range(1, 10);
		if(coin && br[0] - bl[0] > s + 5) { front_clip = s; bl[0] += s; }
		else if(!coin && br[nb - 1] - bl[nb - 1] > s + 5) { back_clip = s; br[nb - 1] -= s; }
	}
	pos = bl[0];
	rpos = br[nb - 1];
	if(front_clip) cig.push_back(op(front_clip, 4));
	for(size_t i = 0; i < nb; i++)
	{
		int32_t len = br[i] - bl[i];
		bool dn = false;
		if(i > 0)
		{
			int32_t gap = bl[i] - br[i - 1];
			// odd case B: a deletion directly before the intron (junction end falls on a D-op end)
			if(mod == 3 && !coin && i == std::max<size_t>(mb, 1) && gap > 10) { cig.push_back(op(2, 2)); cig.push_back(op(gap - 2, 3)); dn = true; }
			else cig.push_back(op(gap, 3));
		}
		(void)dn;
		if(mod == 1 && i == mb && len >= 20)
		{
			int32_t a = r.range(5, len - 8);
			if(coin) { cig.push_back(op(a, 0)); cig.push_back(op(k, 1)); cig.push_back(op(len - a, 0)); }
			else { cig.push_back(op(a, 0)); cig.push_back(op(k, 2)); cig.push_back(op(len - a - k, 0)); }
		}
		else if(mod == 3 && (coin || nb == 1) && i == mb && len >= 8)
		{
			// odd case A: '=' / 'X' ops advance the reference but add no coverage (rnacore/bundle_base.cc:119)
			int32_t a = r.range(2, len - 3);
			cig.push_back(op(a, 7)); cig.push_back(op(1, 8)); cig.push_back(op(len - a - 1, 7));
		}
		else cig.push_back(op(len, 0));
	}
	if(back_clip) cig.push_back(op(back_clip, 4));
}

void push_hit(outbuf &ob, int32_t tid, int32_t pos, int32_t rpos, int32_t mpos, int32_t isize, uint16_t flag, uint8_t xs, uint64_t qid,
		const std::vector<uint32_t> &cig)
{
	hitrec h;
	h.tid = tid; h.pos = pos; h.rpos = rpos; h.mpos = mpos; h.isize = isize; h.flag = flag; h.xs = xs; h.qid = qid;
	h.ncig = (uint32_t)cig.size();
	ob.hits.push_back(h);
	ob.cigar.insert(ob.cigar.end(), cig.begin(), cig.end());
}

void make_template(const model &m, int32_t iso_flat, rng &r, uint64_t qid, uint16_t extra_flag, outbuf &ob,
		std::vector<int32_t> &bl, std::vector<int32_t> &br, std::vector<uint32_t> &c1, std::vector<uint32_t> &c2)
{
	const synth_config &c = m.cfg;
	const gene &ge = m.genes[m.iso_gene[iso_flat]];
	const isoform &is = ge.iso[m.iso_idx[iso_flat]];
	int32_t tlen = is.cum.back();

	if(c.mode == SYNTH_LONG)
	{
		int32_t a = r.range(0, tlen / 20);
		int32_t b = tlen - r.range(0, tlen / 20);
		project(ge, is, a, b, bl, br);
		// 2% indel rate per block: applied as independent D ops inside blocks
		int32_t pos, rpos; bool sp;
		c1.clear();
		pos = bl[0]; rpos = br.back();
		for(size_t i = 0; i < bl.size(); i++)
		{
			int32_t len = br[i] - bl[i];
			if(i > 0) c1.push_back(op(bl[i] - br[i - 1], 3));
			double u = r.uniform();
			if(u < c.indel_rate && len >= 20)
			{
				int32_t x = r.range(5, len - 8);
				int32_t k = r.range(1, 3);
				if(r.next() & 1) { c1.push_back(op(x, 0)); c1.push_back(op(k, 1)); c1.push_back(op(len - x, 0)); }
				else { c1.push_back(op(x, 0)); c1.push_back(op(k, 2)); c1.push_back(op(len - x - k, 0)); }
			}
			else c1.push_back(op(len, 0));
		}
		sp = bl.size() >= 2;
		// unpaired records carry flag 0x8 and mpos -1: without 0x8 the reference's span filter
		// (meta/generator.cc:95) drops every unpaired hit beyond max_read_span
		uint16_t flag = (uint16_t)(((r.next() & 1) ? 0x10 : 0) | 0x8 | extra_flag);
		push_hit(ob, ge.tid, pos, rpos, -1, 0, flag, sp ? (uint8_t)ge.strand : (uint8_t)'.', qid, c1);
		return;
	}

	int32_t rl = std::min(c.read_len, tlen);
	if(c.mode == SYNTH_SINGLE)
	{
		int32_t s = r.range(0, tlen - rl);
		project(ge, is, s, s + rl, bl, br);
		int32_t pos, rpos; bool sp;
		emit_cigar(c, r, bl, br, true, c1, pos, rpos, sp);
		uint16_t flag = (uint16_t)(((r.next() & 1) ? 0x10 : 0) | 0x8 | extra_flag);
		push_hit(ob, ge.tid, pos, rpos, -1, 0, flag, sp ? (uint8_t)ge.strand : (uint8_t)'.', qid, c1);
		return;
	}

	// paired-end, fr-firststrand
	int32_t flen = (int32_t)std::lround(250.0 + 50.0 * r.normal());
	flen = std::max(120, std::min(500, flen));
	flen = std::min(flen, tlen);
	rl = std::min(rl, flen);
	int32_t s = r.range(0, tlen - flen);
	int32_t p1, q1, p2, q2; bool s1, s2;
	project(ge, is, s, s + rl, bl, br);
	emit_cigar(c, r, bl, br, true, c1, p1, q1, s1);
	project(ge, is, s + flen - rl, s + flen, bl, br);
	emit_cigar(c, r, bl, br, true, c2, p2, q2, s2);
	// left mate forward, right mate reverse; '+' gene: left = read2, right = read1 (rnacore/hit.cc:156-162)
	uint16_t fl, fr;
	if(ge.strand == '+') { fl = 0x1 | 0x2 | 0x20 | 0x80; fr = 0x1 | 0x2 | 0x10 | 0x40; }
	else { fl = 0x1 | 0x2 | 0x20 | 0x40; fr = 0x1 | 0x2 | 0x10 | 0x80; }
	int32_t span = std::max(q1, q2) - std::min(p1, p2);
	push_hit(ob, ge.tid, p1, q1, p2, span, (uint16_t)(fl | extra_flag), s1 ? (uint8_t)ge.strand : (uint8_t)'.', qid, c1);
	push_hit(ob, ge.tid, p2, q2, p1, -span, (uint16_t)(fr | extra_flag), s2 ? (uint8_t)ge.strand : (uint8_t)'.', qid, c2);
}

void generate_range(const model &m, int32_t sample, const std::vector<double> &cumw, int64_t t0, int64_t t1, int64_t total, outbuf &ob)
{
	std::vector<int32_t> bl, br;
	std::vector<uint32_t> c1, c2;
	double wtot = cumw.back();
	for(int64_t t = t0; t < t1; t++)
	{
		double u = ((double)t + 0.5) / (double)total * wtot;
		int32_t iso = (int32_t)(std::upper_bound(cumw.begin(), cumw.end(), u) - cumw.begin()) - 1;
		if(iso < 0) iso = 0;
		if(iso >= (int32_t)m.iso_gene.size()) iso = (int32_t)m.iso_gene.size() - 1;
		rng r(m.cfg.seed, ((uint64_t)sample << 40) ^ 0x72656164ULL, (uint64_t)t);
		uint64_t qid = ((uint64_t)t + 1) * 0x9E3779B97F4A7C15ULL + (uint64_t)sample;
		make_template(m, iso, r, qid, 0, ob, bl, br, c1, c2);
		if(r.uniform() < m.cfg.secondary_rate)
		{
			// secondary alignment of the same template: same isoform (another offset) or another isoform
			int32_t iso2 = iso;
			if(r.next() & 1)
			{
				double u2 = r.uniform() * wtot;
				iso2 = (int32_t)(std::upper_bound(cumw.begin(), cumw.end(), u2) - cumw.begin()) - 1;
				if(iso2 < 0) iso2 = 0;
				if(iso2 >= (int32_t)m.iso_gene.size()) iso2 = (int32_t)m.iso_gene.size() - 1;
			}
			make_template(m, iso2, r, qid, 0x100, ob, bl, br, c1, c2);
		}
	}
}

} // namespace

extern "C" {

void synth_default_config(synth_config *c, int mode)
{
	memset(c, 0, sizeof(*c));
	c->seed = 20260101;
	c->mode = mode;
	c->n_chrom = 1;
	c->chrom_len = 100000000;
	c->gene_spacing = 50000;
	c->read_len = 100;
	c->min_exons = (mode == SYNTH_LONG) ? 7 : 3;
	c->max_exons = (mode == SYNTH_LONG) ? 31 : 12;
	c->expressed_fraction = 1.0;
	c->secondary_rate = (mode == SYNTH_LONG) ? 0.0 : 0.03;
	c->indel_rate = (mode == SYNTH_LONG) ? 0.02 : 0.01;
	c->clip_rate = (mode == SYNTH_LONG) ? 0.0 : 0.02;
	c->odd_rate = (mode == SYNTH_LONG) ? 0.0 : 0.005;
}

void *synth_create(const synth_config *c)
{
	model *m = new model;
	m->cfg = *c;
	build_model(*m);
	return m;
}

void synth_destroy(void *s) { delete (model*)s; }

int32_t synth_num_genes(void *s) { return (int32_t)((model*)s)->genes.size(); }

int synth_generate(void *s, int32_t sample, int64_t templates, int32_t threads, synth_records *out)
{
	const model &m = *(model*)s;
	memset(out, 0, sizeof(*out));
	if(m.iso_gene.empty() || templates <= 0) return 0;
	if(threads < 1) threads = 1;

	// expression: log-normal(ln 50, 1.5) per gene, per-sample modulation, isoform shares, times length
	std::vector<double> cumw(m.iso_gene.size() + 1, 0.0);
	for(size_t i = 0; i < m.iso_gene.size(); i++)
	{
		int32_t g = m.iso_gene[i];
		rng rg(m.cfg.seed, 0x65787072ULL, (uint64_t)g);
		double base = std::exp(std::log(50.0) + 1.5 * rg.normal());
		rng rs(m.cfg.seed, 0x73616D70ULL + ((uint64_t)sample << 20), (uint64_t)g);
		bool expressed = rs.uniform() < m.cfg.expressed_fraction;
		double mod = std::exp(0.5 * rs.normal());
		rng ri(m.cfg.seed, 0x69736F66ULL, (uint64_t)i);
		double share = 0.1 + ri.uniform();
		double w = expressed ? base * mod * share * (double)m.genes[g].iso[m.iso_idx[i]].cum.back() : 0.0;
		cumw[i + 1] = cumw[i] + w;
	}
	if(cumw.back() <= 0) return 0;

	std::vector<outbuf> obs(threads);
	std::vector<std::thread> th;
	for(int32_t k = 0; k < threads; k++)
	{
		int64_t t0 = templates * k / threads, t1 = templates * (k + 1) / threads;
		th.emplace_back([&, k, t0, t1]() { generate_range(m, sample, cumw, t0, t1, templates, obs[k]); });
	}
	for(auto &t : th) t.join();

	int64_t n = 0, nc = 0;
	for(auto &o : obs) { n += (int64_t)o.hits.size(); nc += (int64_t)o.cigar.size(); }
	std::vector<hitrec> all;
	all.reserve(n);
	std::vector<uint32_t> coff;
	coff.reserve(n + 1);
	std::vector<uint32_t> cig;
	cig.reserve(nc);
	for(auto &o : obs)
	{
		size_t c0 = 0;
		for(size_t i = 0; i < o.hits.size(); i++)
		{
			coff.push_back((uint32_t)cig.size());
			cig.insert(cig.end(), o.cigar.begin() + c0, o.cigar.begin() + c0 + o.hits[i].ncig);
			c0 += o.hits[i].ncig;
			all.push_back(o.hits[i]);
		}
		outbuf().hits.swap(o.hits);
		std::vector<uint32_t>().swap(o.cigar);
	}
	coff.push_back((uint32_t)cig.size());

	// stable LSD radix sort of record indices by (tid, pos): coordinate-sorted BAM order
	std::vector<uint32_t> idx(n), tmp(n);
	for(int64_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
	uint64_t maxkey = 0;
	std::vector<uint64_t> key(n);
	for(int64_t i = 0; i < n; i++) { key[i] = ((uint64_t)(uint32_t)all[i].tid << 32) | (uint32_t)all[i].pos; maxkey = std::max(maxkey, key[i]); }
	for(int shift = 0; shift < 64 && (maxkey >> shift) != 0; shift += 11)
	{
		size_t cnt[2049];
		memset(cnt, 0, sizeof(cnt));
		for(int64_t i = 0; i < n; i++) cnt[((key[idx[i]] >> shift) & 2047) + 1]++;
		for(int d = 0; d < 2048; d++) cnt[d + 1] += cnt[d];
		for(int64_t i = 0; i < n; i++) tmp[cnt[(key[idx[i]] >> shift) & 2047]++] = idx[i];
		idx.swap(tmp);
	}

	out->n = n;
	out->n_cigar = nc;
	out->tid = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
	out->pos = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
	out->rpos = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
	out->mpos = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
	out->isize = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
	out->flag = (uint16_t*)malloc(sizeof(uint16_t) * (n + 1));
	out->mapq = (uint8_t*)malloc(n + 1);
	out->xs = (uint8_t*)malloc(n + 1);
	out->qid = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
	out->cigar_off = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
	out->cigar = (uint32_t*)malloc(sizeof(uint32_t) * (nc + 1));
	uint32_t w = 0;
	for(int64_t i = 0; i < n; i++)
	{
		const hitrec &h = all[idx[i]];
		out->tid[i] = h.tid; out->pos[i] = h.pos; out->rpos[i] = h.rpos; out->mpos[i] = h.mpos; out->isize[i] = h.isize;
		out->flag[i] = h.flag; out->mapq[i] = 60; out->xs[i] = h.xs; out->qid[i] = h.qid;
		out->cigar_off[i] = w;
		memcpy(out->cigar + w, cig.data() + coff[idx[i]], sizeof(uint32_t) * h.ncig);
		w += h.ncig;
	}
	out->cigar_off[n] = w;
	return 0;
}

void synth_records_free(synth_records *r)
{
	free(r->tid); free(r->pos); free(r->rpos); free(r->mpos); free(r->isize); free(r->flag); free(r->mapq);
	free(r->xs); free(r->qid); free(r->cigar_off); free(r->cigar);
	memset(r, 0, sizeof(*r));
}

}
