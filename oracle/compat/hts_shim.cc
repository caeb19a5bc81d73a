// TEST INFRASTRUCTURE ONLY (oracle/): definitions behind oracle/compat/htslib/sam.h.
#include "hts_shim.h"
#include "htslib/bgzf.h"
#include <map>
#include <mutex>

namespace {
std::mutex g_lock;
std::map<std::string, hts_shim_file> g_files;
std::vector<const hts_shim_file*> g_open;
}

void hts_shim_register(const std::string &name, const hts_shim_file &f)
{
	std::lock_guard<std::mutex> g(g_lock);
	g_files[name] = f;
}

void hts_shim_clear()
{
	std::lock_guard<std::mutex> g(g_lock);
	g_files.clear();
}

hts_shim_record hts_shim_make_record(int32_t tid, int32_t pos, uint8_t mapq, uint16_t flag,
	int32_t mtid, int32_t mpos, int32_t isize, const std::string &qname,
	const uint32_t *cigar, uint32_t n_cigar, char xs, char ts, int nh, int hi, int nm)
{
	hts_shim_record r;
	memset(&r.core, 0, sizeof(r.core));
	size_t lq = qname.size() + 1;
	size_t extranul = (4 - (lq & 3)) & 3;
	r.core.tid = tid;
	r.core.pos = pos;
	r.core.qual = mapq;
	r.core.l_qname = (uint8_t)(lq + extranul);
	r.core.l_extranul = (uint8_t)extranul;
	r.core.flag = flag;
	r.core.n_cigar = n_cigar;
	r.core.l_qseq = 0;
	r.core.mtid = mtid;
	r.core.mpos = mpos;
	r.core.isize = isize;
	r.data.assign(qname.begin(), qname.end());
	for(size_t i = 0; i < 1 + extranul; i++) r.data.push_back(0);
	const uint8_t *c = (const uint8_t*)cigar;
	r.data.insert(r.data.end(), c, c + 4 * (size_t)n_cigar);
	if(xs != 0 && xs != '.') { r.data.push_back('X'); r.data.push_back('S'); r.data.push_back('A'); r.data.push_back((uint8_t)xs); }
	if(ts != 0 && ts != '.') { r.data.push_back('t'); r.data.push_back('s'); r.data.push_back('A'); r.data.push_back((uint8_t)ts); }
	int iv[3] = {nh, hi, nm};
	const char *it[3] = {"NH", "HI", "NM"};
	for(int k = 0; k < 3; k++)
	{
		if(iv[k] < 0) continue;
		r.data.push_back(it[k][0]); r.data.push_back(it[k][1]); r.data.push_back('i');
		int32_t v = iv[k];
		const uint8_t *p = (const uint8_t*)&v;
		r.data.insert(r.data.end(), p, p + 4);
	}
	return r;
}

void hts_shim_view(const hts_shim_record &rec, bam1_t *b)
{
	b->core = rec.core;
	b->data = const_cast<uint8_t*>(rec.data.data());
	b->l_data = (int)rec.data.size();
	b->m_data = (uint32_t)rec.data.size();
	b->id = 0;
}

static int aux_type_size(int t)
{
	switch(t)
	{
		case 'A': case 'c': case 'C': return 1;
		case 's': case 'S': return 2;
		case 'i': case 'I': case 'f': return 4;
		case 'd': return 8;
		default: return 0;
	}
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2])
{
	uint8_t *s = bam_get_aux(b);
	uint8_t *end = b->data + b->l_data;
	while(s + 3 <= end)
	{
		bool hit = (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]);
		uint8_t *val = s + 2;
		int t = val[0];
		uint8_t *nx;
		if(t == 'Z' || t == 'H') { nx = val + 1; while(nx < end && *nx) nx++; nx++; }
		else if(t == 'B')
		{
			int sz = aux_type_size(val[1]);
			uint32_t n; memcpy(&n, val + 2, 4);
			nx = val + 6 + (size_t)sz * n;
		}
		else { int sz = aux_type_size(t); if(sz == 0) return NULL; nx = val + 1 + sz; }
		if(hit) return val;
		s = nx;
	}
	return NULL;
}

int64_t bam_aux2i(const uint8_t *s)
{
	int t = *s++;
	if(t == 'c') return *(const int8_t*)s;
	if(t == 'C') return *s;
	if(t == 's') { int16_t v; memcpy(&v, s, 2); return v; }
	if(t == 'S') { uint16_t v; memcpy(&v, s, 2); return v; }
	if(t == 'i') { int32_t v; memcpy(&v, s, 4); return v; }
	if(t == 'I') { uint32_t v; memcpy(&v, s, 4); return v; }
	return 0;
}

char bam_aux2A(const uint8_t *s)
{
	if(*s == 'A') return (char)s[1];
	return 0;
}

// only used by the reference's BAM writers (rnacore/essential.cc:515-700), which the oracle never calls
int bam_aux_append(bam1_t *, const char *, char, int, const uint8_t *) { abort(); return -1; }
int bam_write1(BGZF *, const bam1_t *) { abort(); return -1; }

samFile *sam_open(const char *fn, const char *)
{
	std::lock_guard<std::mutex> g(g_lock);
	std::map<std::string, hts_shim_file>::const_iterator it = g_files.find(fn);
	if(it == g_files.end()) return NULL;
	samFile *f = new samFile;
	f->fp.bgzf = new BGZF;
	f->fp.bgzf->pos = 0;
	g_open.push_back(&it->second);
	f->store = (int)g_open.size() - 1;
	return f;
}

int sam_close(samFile *fp)
{
	if(fp == NULL) return 0;
	delete fp->fp.bgzf;
	delete fp;
	return 0;
}

bam_hdr_t *sam_hdr_read(samFile *fp)
{
	if(fp == NULL) return NULL;
	const hts_shim_file *f;
	{ std::lock_guard<std::mutex> g(g_lock); f = g_open[fp->store]; }
	bam_hdr_t *h = new bam_hdr_t;
	h->n_targets = (int32_t)f->target_name.size();
	h->target_len = new uint32_t[h->n_targets + 1];
	h->target_name = new char*[h->n_targets + 1];
	for(int i = 0; i < h->n_targets; i++)
	{
		h->target_len[i] = f->target_len[i];
		h->target_name[i] = strdup(f->target_name[i].c_str());
	}
	return h;
}

void bam_hdr_destroy(bam_hdr_t *h)
{
	if(h == NULL) return;
	for(int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
	delete[] h->target_name;
	delete[] h->target_len;
	delete h;
}

hts_idx_t *sam_index_load(samFile *, const char *) { return new hts_idx_t; }
void hts_idx_destroy(hts_idx_t *idx) { delete idx; }
void hts_itr_destroy(hts_itr_t *it) { delete it; }

bam1_t *bam_init1(void)
{
	bam1_t *b = new bam1_t;
	memset(b, 0, sizeof(*b));
	return b;
}

void bam_destroy1(bam1_t *b)
{
	if(b == NULL) return;
	free(b->data);
	delete b;
}

int sam_read1(samFile *fp, bam_hdr_t *, bam1_t *b)
{
	const hts_shim_file *f;
	{ std::lock_guard<std::mutex> g(g_lock); f = g_open[fp->store]; }
	int64_t k = fp->fp.bgzf->pos;
	if(k < 0 || k >= (int64_t)f->records.size()) return -1;
	const hts_shim_record &r = f->records[k];
	if(b->m_data < r.data.size())
	{
		b->data = (uint8_t*)realloc(b->data, r.data.size());
		b->m_data = (uint32_t)r.data.size();
	}
	memcpy(b->data, r.data.data(), r.data.size());
	b->l_data = (int)r.data.size();
	b->core = r.core;
	fp->fp.bgzf->pos = k + 1;
	return (int)r.data.size();
}

int64_t bgzf_seek(BGZF *fp, int64_t pos, int) { fp->pos = pos; return 0; }
int64_t bgzf_tell(BGZF *fp) { return fp->pos; }
BGZF *bgzf_open(const char *, const char *) { BGZF *b = new BGZF; b->pos = 0; return b; }
int bgzf_close(BGZF *fp) { delete fp; return 0; }
