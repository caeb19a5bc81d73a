// BGZF / BAM reader and writer over zlib; see bamio.h.
#include "bamio.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------- BGZF (SAM/BAM specification, section 4.1)
const size_t BGZF_BLOCK = 0xff00;           // uncompressed bytes per block, as htslib uses

struct bgzf_writer
{
	FILE *fp = NULL;
	std::vector<uint8_t> buf;
	bool ok = true;

	void flush_block(const uint8_t *data, size_t n)
	{
		uint8_t out[0x10000 + 64];
		z_stream zs;
		memset(&zs, 0, sizeof(zs));
		if(deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; return; }
		zs.next_in = (Bytef*)data; zs.avail_in = (uInt)n;
		zs.next_out = out + 18; zs.avail_out = sizeof(out) - 18 - 8;
		if(deflate(&zs, Z_FINISH) != Z_STREAM_END) { ok = false; deflateEnd(&zs); return; }
		size_t clen = zs.total_out;
		deflateEnd(&zs);
		size_t bsize = 18 + clen + 8;
		const uint8_t hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
		memcpy(out, hdr, 12);
		out[12] = 'B'; out[13] = 'C'; out[14] = 2; out[15] = 0;
		out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
		uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data, (uInt)n);
		uint32_t isize = (uint32_t)n;
		memcpy(out + 18 + clen, &crc, 4);
		memcpy(out + 18 + clen + 4, &isize, 4);
		if(fwrite(out, 1, bsize, fp) != bsize) ok = false;
	}
	void write(const void *p, size_t n)
	{
		const uint8_t *b = (const uint8_t*)p;
		while(n > 0)
		{
			size_t room = BGZF_BLOCK - buf.size();
			size_t k = n < room ? n : room;
			buf.insert(buf.end(), b, b + k);
			b += k; n -= k;
			if(buf.size() == BGZF_BLOCK) { flush_block(buf.data(), buf.size()); buf.clear(); }
		}
	}
	void close()
	{
		if(!buf.empty()) { flush_block(buf.data(), buf.size()); buf.clear(); }
		flush_block(NULL, 0);                   // the empty end-of-file marker block
		if(fp) fclose(fp);
		fp = NULL;
	}
};

struct bgzf_reader
{
	FILE *fp = NULL;
	std::vector<uint8_t> cur;
	size_t at = 0;
	bool bad = false, eof = false;

	bool next_block()
	{
		uint8_t h[12];
		size_t got = fread(h, 1, 12, fp);
		if(got == 0) { eof = true; return false; }
		if(got != 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { bad = true; return false; }
		unsigned xlen = h[10] | (h[11] << 8);
		std::vector<uint8_t> extra(xlen);
		if(fread(extra.data(), 1, xlen, fp) != xlen) { bad = true; return false; }
		long bsize = -1;
		for(unsigned i = 0; i + 4 <= xlen; )
		{
			unsigned slen = extra[i + 2] | (extra[i + 3] << 8);
			if(extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= xlen) bsize = (extra[i + 4] | (extra[i + 5] << 8)) + 1;
			i += 4 + slen;
		}
		if(bsize < 0) { bad = true; return false; }
		long clen = bsize - 12 - (long)xlen - 8;
		if(clen < 0) { bad = true; return false; }
		std::vector<uint8_t> comp(clen + 8);
		if(fread(comp.data(), 1, clen + 8, fp) != (size_t)clen + 8) { bad = true; return false; }
		uint32_t crc, isize;
		memcpy(&crc, comp.data() + clen, 4);
		memcpy(&isize, comp.data() + clen + 4, 4);
		cur.resize(isize);
		at = 0;
		if(isize == 0) return true;
		z_stream zs;
		memset(&zs, 0, sizeof(zs));
		if(inflateInit2(&zs, -15) != Z_OK) { bad = true; return false; }
		zs.next_in = comp.data(); zs.avail_in = (uInt)clen;
		zs.next_out = cur.data(); zs.avail_out = isize;
		int rc = inflate(&zs, Z_FINISH);
		inflateEnd(&zs);
		if(rc != Z_STREAM_END || zs.total_out != isize) { bad = true; return false; }
		if((uint32_t)crc32(crc32(0L, Z_NULL, 0), cur.data(), isize) != crc) { bad = true; return false; }
		return true;
	}
	// read exactly n bytes; false at a clean end of file before the first byte, `bad` set on truncation
	bool read(void *p, size_t n)
	{
		uint8_t *o = (uint8_t*)p;
		size_t done = 0;
		while(done < n)
		{
			if(at == cur.size())
			{
				if(!next_block()) { if(done > 0) bad = true; return false; }
				continue;
			}
			size_t k = cur.size() - at;
			if(k > n - done) k = n - done;
			memcpy(o + done, cur.data() + at, k);
			at += k; done += k;
		}
		return true;
	}
};

// second, independent hash of a query name (different offset basis and finaliser): with hash64 a 128-bit identity
uint64_t hash64b(const char *s, size_t n)
{
	uint64_t h = 0x84222325cbf29ce4ULL;
	for(size_t i = 0; i < n; i++) { h = (h ^ (uint8_t)s[i]) * 0x100000001b3ULL; h ^= h >> 29; }
	h ^= h >> 31; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 32;
	return h;
}

// test hook (bam_set_key_bits): the 64-bit query-name keys truncated to this many bits, so that the collision handling below
// can be exercised on ordinary data
int g_key_bits = 64;

// The ABI's contract for qid is "equal <=> same qname" (include/aletsch_gpu.h); bundle_base::build_fragments compares the names
// themselves (rnacore/bundle_base.cc:308).  A 64-bit hash alone only makes a violation improbable.  This makes it (128-bit)
// certain where it matters: two records can only be paired inside one bundle, i.e. within max_read_span (500000) of each other,
// so every record's key is checked against the keys of the records of the last KEY_WINDOW positions; a different name under an
// equal key gets a new key derived from its second hash, until the key is free.  The same name always takes the same path, so
// mates still receive equal keys.
struct key_window
{
	enum { KEY_WINDOW = 1200000 };
	std::unordered_map<uint64_t, uint64_t> seen;             // key -> second hash of the name that owns it
	std::deque<std::pair<int64_t, uint64_t> > order;         // (position, key) in arrival order, for eviction
	int32_t tid = -1;
	int64_t collisions = 0;
	uint64_t resolve(int32_t t, int32_t pos, uint64_t key, uint64_t h2)
	{
		if(t != tid) { seen.clear(); order.clear(); tid = t; }
		while(!order.empty() && order.front().first + KEY_WINDOW < (int64_t)pos)
		{
			// (a key seen again later was re-inserted at the back: only drop it when this was its last sighting)
			std::unordered_map<uint64_t, uint64_t>::iterator it = seen.find(order.front().second);
			if(it != seen.end() && last_pos[order.front().second] == order.front().first) { seen.erase(it); last_pos.erase(order.front().second); }
			order.pop_front();
		}
		for(int round = 0; round < 64; round++)
		{
			std::unordered_map<uint64_t, uint64_t>::iterator it = seen.find(key);
			if(it == seen.end()) { seen[key] = h2; break; }
			if(it->second == h2) break;                          // the same name: its mate or another alignment of it
			collisions++;
			key = (key ^ h2) * 0x9E3779B97F4A7C15ULL + (uint64_t)round;     // another name owns the key: move on, deterministically
			if(g_key_bits < 64) key &= (1ULL << g_key_bits) - 1;
			if(key == 0xffffffffffffffffULL) key = 0x7fffffffffffffffULL;
		}
		last_pos[key] = pos;
		order.push_back(std::make_pair((int64_t)pos, key));
		return key;
	}
	std::unordered_map<uint64_t, int64_t> last_pos;
};

uint64_t hash64(const char *s, size_t n)
{
	uint64_t h = 0xcbf29ce484222325ULL;                          // FNV-1a, then a finaliser; never the reserved all-ones key
	for(size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 0x100000001b3ULL; }
	h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33;
	if(h == 0xffffffffffffffffULL) h = 0x7fffffffffffffffULL;
	return h;
}

int reg2bin(int64_t beg, int64_t end)                            // SAM/BAM specification, section 5.3
{
	--end;
	if(beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
	if(beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
	if(beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
	if(beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
	if(beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
	return 0;
}

template<typename T> void put(std::vector<uint8_t> &v, T x) { const uint8_t *p = (const uint8_t*)&x; v.insert(v.end(), p, p + sizeof(T)); }

template<typename T> T *grow(T *p, size_t n) { return (T*)realloc(p, (n ? n : 1) * sizeof(T)); }

} // namespace

extern "C" {

int bam_write_records(const char *path, int32_t n_chrom, const int32_t *chrom_len, const synth_records *r, int tag_mode)
{
	bgzf_writer w;
	w.fp = fopen(path, "wb");
	if(!w.fp) return -1;
	std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
	for(int k = 0; k < n_chrom; k++) text += "@SQ\tSN:chr" + std::to_string(k + 1) + "\tLN:" + std::to_string(chrom_len[k]) + "\n";
	std::vector<uint8_t> h;
	h.insert(h.end(), {'B', 'A', 'M', 1});
	put<int32_t>(h, (int32_t)text.size());
	h.insert(h.end(), text.begin(), text.end());
	put<int32_t>(h, n_chrom);
	for(int k = 0; k < n_chrom; k++)
	{
		std::string name = "chr" + std::to_string(k + 1);
		put<int32_t>(h, (int32_t)name.size() + 1);
		h.insert(h.end(), name.begin(), name.end());
		h.push_back(0);
		put<int32_t>(h, chrom_len[k]);
	}
	w.write(h.data(), h.size());
	std::vector<uint8_t> rec;
	for(int64_t i = 0; i < r->n; i++)
	{
		rec.clear();
		char qn[32];
		int ql = snprintf(qn, sizeof(qn), "q%llx", (unsigned long long)r->qid[i]) + 1;
		uint32_t nc = r->cigar_off[i + 1] - r->cigar_off[i];
		const uint32_t *cg = r->cigar + r->cigar_off[i];
		int32_t lseq = 0;
		for(uint32_t k = 0; k < nc; k++) { uint32_t op = cg[k] & 0xf; if(op == 0 || op == 1 || op == 4 || op == 7 || op == 8) lseq += (int32_t)(cg[k] >> 4); }
		put<int32_t>(rec, r->tid[i]);
		put<int32_t>(rec, r->pos[i]);
		rec.push_back((uint8_t)ql);
		rec.push_back(r->mapq[i]);
		put<uint16_t>(rec, (uint16_t)reg2bin(r->pos[i], r->rpos[i] > r->pos[i] ? r->rpos[i] : r->pos[i] + 1));
		const bool long_cigar = nc > 65535;            // SAM specification 4.2.2: placeholder CIGAR + CG:B,I tag
		put<uint16_t>(rec, (uint16_t)(long_cigar ? 2 : nc));
		put<uint16_t>(rec, r->flag[i]);
		put<int32_t>(rec, lseq);
		put<int32_t>(rec, (r->flag[i] & 0x1) ? r->tid[i] : -1);
		put<int32_t>(rec, r->mpos[i]);
		put<int32_t>(rec, r->isize[i]);
		rec.insert(rec.end(), qn, qn + ql);
		if(long_cigar)
		{
			put<uint32_t>(rec, ((uint32_t)lseq << 4) | 4u);
			put<uint32_t>(rec, ((uint32_t)(r->rpos[i] - r->pos[i]) << 4) | 3u);
		}
		else for(uint32_t k = 0; k < nc; k++) put<uint32_t>(rec, cg[k]);
		rec.insert(rec.end(), (size_t)(lseq + 1) / 2, (uint8_t)0x11);       // sequence: all 'A'
		rec.insert(rec.end(), (size_t)lseq, (uint8_t)0xff);                 // qualities absent
		if(long_cigar)
		{
			rec.insert(rec.end(), {'C', 'G', 'B', 'I'});
			put<int32_t>(rec, (int32_t)nc);
			for(uint32_t k = 0; k < nc; k++) put<uint32_t>(rec, cg[k]);
		}
		rec.insert(rec.end(), {'N', 'H', 'C', 1});
		rec.insert(rec.end(), {'H', 'I', 'C', 1});
		char xs = (char)r->xs[i];
		if(xs == '+' || xs == '-')
		{
			if(tag_mode == 0) rec.insert(rec.end(), {'X', 'S', 'A', (uint8_t)xs});
			else
			{
				// ts is relative to the read: hit::set_tags flips it for reverse-strand alignments (rnacore/hit.cc:116-123)
				char ts = (r->flag[i] & 0x10) ? (xs == '+' ? '-' : '+') : xs;
				rec.insert(rec.end(), {'t', 's', 'A', (uint8_t)ts});
			}
		}
		int32_t bs = (int32_t)rec.size();
		w.write(&bs, 4);
		w.write(rec.data(), rec.size());
	}
	w.close();
	return w.ok ? 0 : -2;
}

int bam_read_records(const char *path, synth_records *out, int32_t *n_chrom_out, int32_t *chrom_len_out, int32_t cap)
{
	memset(out, 0, sizeof(*out));
	bgzf_reader rd;
	rd.fp = fopen(path, "rb");
	if(!rd.fp) return -1;
	char magic[4];
	int32_t l_text = 0, n_ref = 0;
	if(!rd.read(magic, 4) || memcmp(magic, "BAM\1", 4) != 0 || !rd.read(&l_text, 4) || l_text < 0) { fclose(rd.fp); return -2; }
	std::vector<char> text(l_text);
	if(l_text > 0 && !rd.read(text.data(), l_text)) { fclose(rd.fp); return -2; }
	if(!rd.read(&n_ref, 4) || n_ref < 0) { fclose(rd.fp); return -2; }
	for(int k = 0; k < n_ref; k++)
	{
		int32_t l_name = 0, l_ref = 0;
		if(!rd.read(&l_name, 4) || l_name < 0) { fclose(rd.fp); return -2; }
		std::vector<char> name(l_name);
		if(l_name > 0 && !rd.read(name.data(), l_name)) { fclose(rd.fp); return -2; }
		if(!rd.read(&l_ref, 4)) { fclose(rd.fp); return -2; }
		if(chrom_len_out && k < cap) chrom_len_out[k] = l_ref;
	}
	if(n_chrom_out) *n_chrom_out = n_ref;
	if(chrom_len_out && n_ref > cap) { fclose(rd.fp); return -4; }        // the caller's dictionary buffer is too small: no silent truncation
	key_window keys;
	size_t capn = 0, capc = 0;
	int64_t n = 0, nc_tot = 0;
	std::vector<uint8_t> rec;
	while(true)
	{
		int32_t bs = 0;
		if(!rd.read(&bs, 4)) break;
		if(bs < 32) { rd.bad = true; break; }
		rec.resize(bs);
		if(!rd.read(rec.data(), bs)) { rd.bad = true; break; }
		int32_t tid, pos, lseq, mtid, mpos, isize;
		uint16_t ncig, flag;
		memcpy(&tid, &rec[0], 4); memcpy(&pos, &rec[4], 4);
		uint8_t lq = rec[8], mapq = rec[9];
		memcpy(&ncig, &rec[12], 2); memcpy(&flag, &rec[14], 2); memcpy(&lseq, &rec[16], 4);
		memcpy(&mtid, &rec[20], 4); memcpy(&mpos, &rec[24], 4); memcpy(&isize, &rec[28], 4);
		size_t o_q = 32, o_c = o_q + lq, o_s = o_c + 4 * (size_t)ncig, o_a = o_s + (size_t)(lseq + 1) / 2 + (size_t)lseq;
		if(lseq < 0 || o_a > (size_t)bs || lq < 1) { rd.bad = true; break; }
		if((size_t)n + 2 > capn)
		{
			capn = capn ? capn * 2 : 1 << 16;
			out->tid = grow(out->tid, capn); out->pos = grow(out->pos, capn); out->rpos = grow(out->rpos, capn);
			out->mpos = grow(out->mpos, capn); out->isize = grow(out->isize, capn); out->flag = grow(out->flag, capn);
			out->mapq = grow(out->mapq, capn); out->xs = grow(out->xs, capn); out->qid = grow(out->qid, capn);
			out->cigar_off = grow(out->cigar_off, capn + 1);
		}
		if((size_t)nc_tot + ncig + 1 > capc) { capc = (capc ? capc * 2 : 1 << 17) + ncig; out->cigar = grow(out->cigar, capc); }
		if(n == 0) out->cigar_off[0] = 0;
		int32_t rpos = pos;
		for(unsigned k = 0; k < ncig; k++)
		{
			uint32_t c;
			memcpy(&c, &rec[o_c + 4 * k], 4);
			out->cigar[nc_tot + k] = c;
			if((0x3C1A7 >> ((c & 0xf) << 1)) & 2) rpos += (int32_t)(c >> 4);       // bam_cigar2rlen
		}
		nc_tot += ncig;
		// aux fields: XS:A and ts:A (hit::set_tags, rnacore/hit.cc:106-123)
		char xs = '.', ts = '.';
		size_t a = o_a, cg_at = 0;
		int32_t cg_n = 0;
		while(a + 3 <= (size_t)bs)
		{
			char t0 = (char)rec[a], t1 = (char)rec[a + 1], ty = (char)rec[a + 2];
			a += 3;
			size_t len = 0;
			if(ty == 'A' || ty == 'c' || ty == 'C') len = 1;
			else if(ty == 's' || ty == 'S') len = 2;
			else if(ty == 'i' || ty == 'I' || ty == 'f') len = 4;
			else if(ty == 'Z' || ty == 'H') { while(a + len < (size_t)bs && rec[a + len] != 0) len++; len++; }
			else if(ty == 'B')
			{
				if(a + 5 > (size_t)bs) break;
				char st = (char)rec[a];
				int32_t cnt;
				memcpy(&cnt, &rec[a + 1], 4);
				size_t es = (st == 'c' || st == 'C') ? 1 : ((st == 's' || st == 'S') ? 2 : 4);
				if(cnt < 0 || (uint64_t)es * (uint64_t)cnt > (uint64_t)bs - a - 5) break;       // hostile count: would wrap / run past the record
				len = 5 + es * (size_t)cnt;
				if(t0 == 'C' && t1 == 'G' && st == 'I') { cg_at = a + 5; cg_n = cnt; }
			}
			else break;
			if(a + len > (size_t)bs) break;
			if(ty == 'A' && t0 == 'X' && t1 == 'S') xs = (char)rec[a];
			if(ty == 'A' && t0 == 't' && t1 == 's') ts = (char)rec[a];
			a += len;
		}
		// more than 65535 operations (long reads): the record holds the placeholder <read length>S <reference length>N and the real
		// CIGAR in the CG:B,I tag (SAM specification, section 4.2.2; htslib swaps it in when reading)
		if(cg_n > 0 && ncig == 2)
		{
			uint32_t c0, c1;
			memcpy(&c0, &rec[o_c], 4); memcpy(&c1, &rec[o_c + 4], 4);
			if((c0 & 0xf) == 4 && (int32_t)(c0 >> 4) == lseq && (c1 & 0xf) == 3)
			{
				nc_tot -= ncig;
				if((size_t)nc_tot + (size_t)cg_n + 1 > capc) { capc = capc * 2 + (size_t)cg_n; out->cigar = grow(out->cigar, capc); }
				rpos = pos;
				for(int32_t k = 0; k < cg_n; k++)
				{
					uint32_t c;
					memcpy(&c, &rec[cg_at + 4 * (size_t)k], 4);
					out->cigar[nc_tot + k] = c;
					if((0x3C1A7 >> ((c & 0xf) << 1)) & 2) rpos += (int32_t)(c >> 4);
				}
				nc_tot += cg_n;
			}
		}
		if(xs == '.' && ts != '.')
		{
			if((flag & 0x10) && ts == '+') xs = '-';
			if((flag & 0x10) && ts == '-') xs = '+';
			if(!(flag & 0x10) && ts == '+') xs = '+';
			if(!(flag & 0x10) && ts == '-') xs = '-';
		}
		out->tid[n] = tid; out->pos[n] = pos; out->rpos[n] = rpos; out->mpos[n] = mpos; out->isize[n] = isize;
		out->flag[n] = flag; out->mapq[n] = mapq; out->xs[n] = (uint8_t)xs;
		{
			uint64_t key = hash64((const char*)&rec[o_q], (size_t)lq - 1);
			if(g_key_bits < 64) key &= (1ULL << g_key_bits) - 1;
			out->qid[n] = keys.resolve(tid, pos, key, hash64b((const char*)&rec[o_q], (size_t)lq - 1));
		}
		out->cigar_off[n + 1] = (uint32_t)nc_tot;
		n++;
		(void)mtid;
	}
	fclose(rd.fp);
	out->n = n; out->n_cigar = nc_tot;
	if(n == 0)
	{
		out->cigar_off = grow(out->cigar_off, 1);
		out->cigar_off[0] = 0;
	}
	if(rd.bad) { synth_records_free(out); return -3; }
	return 0;
}

int bam_set_key_bits(int bits)
{
	int old = g_key_bits;
	g_key_bits = bits < 1 ? 1 : (bits > 64 ? 64 : bits);
	return old;
}

}
