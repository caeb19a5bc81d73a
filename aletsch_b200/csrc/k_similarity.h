// Stage 5 kernels: junction-set similarity of the bundles of one bundle_group
// (bundle_group::build_splice_similarity, meta/bundle_group.cc:190-231).
//
// The reference intersects two sorted splice lists with std::set_intersection for every pair, once per
// shared splice position.  Here every list becomes a bitset over the group's dictionary of splice
// positions and c = |A ∩ B| = popcount(A & B), computed once for all pairs in 32 x 32 tiles.
#ifndef ALETSCH_B200_CSRC_K_SIMILARITY_H
#define ALETSCH_B200_CSRC_K_SIMILARITY_H

#include "runtime.h"
#include "blockops.h"
#include "k_evidence.h"

namespace agpu {

#define SIM_TILE 32
#define SIM_CHUNK 32          // 64-bit words per shared-memory stage

KERNEL k_sim_keys(int64_t n, const int32_t *val, u64 *key)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	key[i] = (u64)(u32)val[i];
}

// single CTA: sort all positions and keep the distinct ones (the dictionary); *n_dict receives its size
KERNEL k_sim_dictionary(u64 *key, int n, int32_t *flag, int32_t *dict, int32_t *n_dict)
{
	block_sort_u64(key, n);
	for(int i = threadIdx.x; i < n; i += blockDim.x) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
	BLOCK_SYNC();
	int tot = block_excl_scan(flag, n);
	for(int i = threadIdx.x; i < n; i += blockDim.x) if(i == 0 || key[i] != key[i - 1]) dict[flag[i]] = (int32_t)(u32)key[i];
	if(threadIdx.x == 0) *n_dict = tot;
}

KERNEL k_sim_bitsets(int64_t n, int32_t n_lists, const int64_t *list_off, const int32_t *val, const int32_t *dict, const int32_t *n_dict,
		int words, u64 *bits)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	int g = find_segment(list_off, n_lists, i);
	int k = lower_bound_idx(dict, *n_dict, val[i]);
	atomicOr(&bits[(int64_t)g * words + (k >> 6)], (u64)1 << (k & 63));
}

// one CTA per 32 x 32 tile of pairs (only tiles on or above the diagonal do work)
KERNEL k_sim_tiles(int32_t n_lists, int words, const u64 *bits, int32_t *out_c)
{
	SHARED u64 sa[SIM_TILE][SIM_CHUNK + 1];
	SHARED u64 sb[SIM_TILE][SIM_CHUNK + 1];
	SHARED int sacc[SIM_TILE * SIM_TILE];
	int tiles = (n_lists + SIM_TILE - 1) / SIM_TILE;
	for(int64_t t = blockIdx.x; t < (int64_t)tiles * tiles; t += gridDim.x)
	{
		int ti = (int)(t / tiles), tj = (int)(t % tiles);
		if(tj < ti) continue;
		int nthr = blockDim.x;
		// every thread owns the pairs p = threadIdx.x, threadIdx.x + nthr, ... of the 1024 in the tile
		for(int p = threadIdx.x; p < SIM_TILE * SIM_TILE; p += nthr) sacc[p] = 0;
		for(int w0 = 0; w0 < words; w0 += SIM_CHUNK)
		{
			for(int x = threadIdx.x; x < SIM_TILE * SIM_CHUNK; x += nthr)
			{
				int r = x / SIM_CHUNK, c = x % SIM_CHUNK;
				int gi = ti * SIM_TILE + r, gj = tj * SIM_TILE + r;
				sa[r][c] = (gi < n_lists && w0 + c < words) ? bits[(int64_t)gi * words + w0 + c] : 0;
				sb[r][c] = (gj < n_lists && w0 + c < words) ? bits[(int64_t)gj * words + w0 + c] : 0;
			}
			BLOCK_SYNC();
			for(int p = threadIdx.x; p < SIM_TILE * SIM_TILE; p += nthr)
			{
				int a = p / SIM_TILE, b = p % SIM_TILE;
				int s = 0;
				for(int c = 0; c < SIM_CHUNK; c++) s += __popcll(sa[a][c] & sb[b][c]);
				sacc[p] += s;
			}
			BLOCK_SYNC();
		}
		for(int p = threadIdx.x; p < SIM_TILE * SIM_TILE; p += nthr)
		{
			int i = ti * SIM_TILE + p / SIM_TILE, j = tj * SIM_TILE + p % SIM_TILE;
			if(i < n_lists && j < n_lists && i < j) out_c[(int64_t)i * n_lists + j] = sacc[p];
		}
		BLOCK_SYNC();
	}
}

// ---- many small bundle groups at once: one CTA per group does dictionary, bitsets and all pairs
struct sim_batch
{
	int32_t n_groups;
	const int32_t *group_off;     // [n_groups + 1] lists of group g
	const int64_t *list_off;      // [n_lists + 1]
	const int32_t *val;
	const int64_t *bits_off;      // [n_groups] u64 words of scratch: G x words_ub, zeroed
	const int64_t *c_off;         // [n_groups] start of the group's dense G x G matrix in out_c
	const uint8_t *skip;          // [n_groups] 1: group is handled by the tiled kernels
	u64 *key;                     // [n_values] scratch
	int32_t *flag, *dict;         // [n_values] scratch
	u64 *bits;
	int32_t *out_c;
};

KERNEL k_sim_groups(sim_batch a)
{
	for(int g = blockIdx.x; g < a.n_groups; g += gridDim.x)
	{
		if(a.skip[g]) continue;
		const int l0 = a.group_off[g], G = a.group_off[g + 1] - l0;
		const int64_t v0 = a.list_off[l0];
		const int n = (int)(a.list_off[l0 + G] - v0);
		if(G < 2 || n <= 0) continue;
		u64 *key = a.key + v0;
		int32_t *flag = a.flag + v0, *dict = a.dict + v0;
		for(int i = threadIdx.x; i < n; i += blockDim.x) key[i] = (u64)(u32)a.val[v0 + i];
		BLOCK_SYNC();
		block_sort_u64(key, n);
		for(int i = threadIdx.x; i < n; i += blockDim.x) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
		BLOCK_SYNC();
		const int nd = block_excl_scan(flag, n);
		for(int i = threadIdx.x; i < n; i += blockDim.x) if(i == 0 || key[i] != key[i - 1]) dict[flag[i]] = (int32_t)(u32)key[i];
		BLOCK_SYNC();
		const int words = (n + 63) / 64;                      // row stride (upper bound of the dictionary size)
		const int used = (nd + 63) / 64;
		u64 *bits = a.bits + a.bits_off[g];
		for(int i = threadIdx.x; i < n; i += blockDim.x)
		{
			int li = find_segment(a.list_off + l0, G, v0 + i);
			int k = lower_bound_idx(dict, nd, a.val[v0 + i]);
			atomicOr(&bits[(int64_t)li * words + (k >> 6)], (u64)1 << (k & 63));
		}
		BLOCK_SYNC();
		int32_t *out = a.out_c + a.c_off[g];
		for(int p = threadIdx.x; p < G * G; p += blockDim.x)
		{
			int i = p / G, j = p % G;
			if(i >= j) continue;
			const u64 *x = bits + (int64_t)i * words, *y = bits + (int64_t)j * words;
			int s = 0;
			for(int w = 0; w < used; w++) s += __popcll(x[w] & y[w]);
			out[p] = s;
		}
		BLOCK_SYNC();
	}
}

// ---- all bundle groups of a call at once, any size, with the pairs filtered on the device ------------------------------------
// bundle_group::build_splice_similarity only ever keeps a pair (i, j) when both have at most max_num_junctions_to_combine
// junctions, c = |splices_i AND splices_j| > 0 and r = c / min(|i|, |j|) reaches the round's threshold (meta/bundle_group.cc:193-
// 221).  The thresholds of the two rounds are known up front, so the device hands back, per bundle, only the partners j > i that
// can pass the weaker one, with c, in ascending j: a CSR the host walks instead of a dense G x G matrix (1 GB of host memory and
// most of the time at 1000 cells per region).
//   dictionary : the group's splice positions as ranks in a per-group position bitmap (a region spans a few Mb: no sort)
//   bitsets    : one row of ceil(distinct / 64) words per bundle
//   tiles      : 32 x 32 pairs per CTA, AND + popcount over the rows staged in shared memory; pass 1 counts the qualifying pairs
//                per (bundle, column tile), a look-back scan places them, pass 2 writes (j, c)
struct sim_lists
{
	int32_t n_groups, n_lists;
	const int32_t *group_off;      // [n_groups + 1]
	const int32_t *list_group;     // [n_lists]
	const int64_t *list_off;       // [n_lists + 1]
	const int32_t *val;
	int32_t *gmin, *gmax;          // [n_groups]
	const int64_t *bm_off;         // [n_groups + 1] words of the position bitmaps
	u32 *bitmap;
	const u32 *wrank;              // global rank of every bitmap word
	const int64_t *row_off;        // [n_lists + 1] words of the bitset rows
	u64 *rows;
};

KERNEL k_simb_minmax(sim_lists a)
{
	int l = blockIdx.x * blockDim.x + threadIdx.x;
	if(l >= a.n_lists) return;
	const int64_t v0 = a.list_off[l], v1 = a.list_off[l + 1];
	if(v1 <= v0) return;
	const int g = a.list_group[l];
	atomicMin(&a.gmin[g], a.val[v0]);
	atomicMax(&a.gmax[g], a.val[v1 - 1]);
}

// list of every value (lists are short: a binary search over the offsets)
DEV int simb_list_of(const sim_lists &a, int64_t i) { return find_segment(a.list_off, a.n_lists, i); }

KERNEL k_simb_mark(sim_lists a, int64_t n_val)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n_val) return;
	const int g = a.list_group[simb_list_of(a, i)];
	const int64_t bit = (int64_t)a.val[i] - a.gmin[g];
	atomicOr(&a.bitmap[a.bm_off[g] + (bit >> 5)], 1u << (bit & 31));
}

KERNEL k_simb_rows(sim_lists a, int64_t n_val)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n_val) return;
	const int l = simb_list_of(a, i);
	const int g = a.list_group[l];
	const int64_t bit = (int64_t)a.val[i] - a.gmin[g];
	const int64_t w = a.bm_off[g] + (bit >> 5);
	const u32 k = a.wrank[w] - a.wrank[a.bm_off[g]] + (u32)__popc(a.bitmap[w] & ((1u << (bit & 31)) - 1u));
	atomicOr(&a.rows[a.row_off[l] + (k >> 6)], (u64)1 << (k & 63));
}

struct sim_tiles
{
	int64_t n_jobs;
	const int32_t *job_g, *job_ti, *job_tj;
	const int64_t *cnt_off;        // [n_groups] start of the group's (bundle, column tile) counters
	int max_junctions;
	double thr;                    // the weaker of the two rounds' thresholds
};

// one 32 x 32 tile: c of every pair into sacc; returns (to all threads) nothing -- helper shared by the two passes
DEV void simb_tile_counts(const sim_lists &a, int g, int ti, int tj, int G, int words, u64 (*sa)[SIM_CHUNK + 1], u64 (*sb)[SIM_CHUNK + 1], int *sacc)
{
	const int l0 = a.group_off[g];
	const int nthr = blockDim.x;
	for(int p = threadIdx.x; p < SIM_TILE * SIM_TILE; p += nthr) sacc[p] = 0;
	for(int w0 = 0; w0 < words; w0 += SIM_CHUNK)
	{
		BLOCK_SYNC();
		for(int x = threadIdx.x; x < SIM_TILE * SIM_CHUNK; x += nthr)
		{
			const int r = x / SIM_CHUNK, c = x % SIM_CHUNK;
			const int gi = ti * SIM_TILE + r, gj = tj * SIM_TILE + r;
			sa[r][c] = (gi < G && w0 + c < words) ? a.rows[a.row_off[l0 + gi] + w0 + c] : 0;
			sb[r][c] = (gj < G && w0 + c < words) ? a.rows[a.row_off[l0 + gj] + w0 + c] : 0;
		}
		BLOCK_SYNC();
		const int lim = words - w0 < SIM_CHUNK ? words - w0 : SIM_CHUNK;
		for(int p = threadIdx.x; p < SIM_TILE * SIM_TILE; p += nthr)
		{
			const int x = p / SIM_TILE, y = p % SIM_TILE;
			int s = 0;
			for(int c = 0; c < lim; c++) s += __popcll(sa[x][c] & sb[y][c]);
			sacc[p] += s;
		}
	}
	BLOCK_SYNC();
}

// does the pair (local bundles i < j of group g) pass the weaker round?  (meta/bundle_group.cc:196-221)
DEV bool simb_keep(const sim_lists &a, const sim_tiles &t, int l0, int i, int j, int G, int c)
{
	if(i >= G || j >= G || i >= j || c <= 0) return false;
	const int64_t ni = a.list_off[l0 + i + 1] - a.list_off[l0 + i], nj = a.list_off[l0 + j + 1] - a.list_off[l0 + j];
	if(ni / 2.0 > t.max_junctions || nj / 2.0 > t.max_junctions) return false;
	const int64_t small = ni < nj ? ni : nj;
	const double r = c * 1.0 / small;
	return !(r < t.thr);
}

KERNEL k_simb_count(sim_lists a, sim_tiles t, int32_t *count)
{
	SHARED u64 sa[SIM_TILE][SIM_CHUNK + 1];
	SHARED u64 sb[SIM_TILE][SIM_CHUNK + 1];
	SHARED int sacc[SIM_TILE * SIM_TILE];
	for(int64_t job = blockIdx.x; job < t.n_jobs; job += gridDim.x)
	{
		const int g = t.job_g[job], ti = t.job_ti[job], tj = t.job_tj[job];
		const int l0 = a.group_off[g], G = a.group_off[g + 1] - l0;
		const int T = (G + SIM_TILE - 1) / SIM_TILE;
		const int words = (int)(a.row_off[l0 + 1] - a.row_off[l0]);
		simb_tile_counts(a, g, ti, tj, G, words, sa, sb, sacc);
		for(int x = threadIdx.x; x < SIM_TILE; x += blockDim.x)
		{
			const int i = ti * SIM_TILE + x;
			if(i >= G) continue;
			int n = 0;
			for(int y = 0; y < SIM_TILE; y++) n += simb_keep(a, t, l0, i, tj * SIM_TILE + y, G, sacc[x * SIM_TILE + y]) ? 1 : 0;
			count[t.cnt_off[g] + (int64_t)i * T + tj] = n;
		}
		BLOCK_SYNC();
	}
}

KERNEL k_simb_fill(sim_lists a, sim_tiles t, const int64_t *place, int32_t *out_j, int32_t *out_c)
{
	SHARED u64 sa[SIM_TILE][SIM_CHUNK + 1];
	SHARED u64 sb[SIM_TILE][SIM_CHUNK + 1];
	SHARED int sacc[SIM_TILE * SIM_TILE];
	for(int64_t job = blockIdx.x; job < t.n_jobs; job += gridDim.x)
	{
		const int g = t.job_g[job], ti = t.job_ti[job], tj = t.job_tj[job];
		const int l0 = a.group_off[g], G = a.group_off[g + 1] - l0;
		const int T = (G + SIM_TILE - 1) / SIM_TILE;
		const int words = (int)(a.row_off[l0 + 1] - a.row_off[l0]);
		simb_tile_counts(a, g, ti, tj, G, words, sa, sb, sacc);
		for(int x = threadIdx.x; x < SIM_TILE; x += blockDim.x)
		{
			const int i = ti * SIM_TILE + x;
			if(i >= G) continue;
			int64_t w = place[t.cnt_off[g] + (int64_t)i * T + tj];
			for(int y = 0; y < SIM_TILE; y++)
			{
				const int j = tj * SIM_TILE + y, c = sacc[x * SIM_TILE + y];
				if(simb_keep(a, t, l0, i, j, G, c)) { out_j[w] = j; out_c[w] = c; w++; }
			}
		}
		BLOCK_SYNC();
	}
}

// row_ptr[l] = place of bundle l's first pair (its counters are contiguous); row_ptr[n_lists] = total
KERNEL k_simb_rowptr(sim_lists a, const int64_t *cnt_off, const int64_t *place, int64_t n_count, int64_t *row_ptr)
{
	int l = blockIdx.x * blockDim.x + threadIdx.x;
	if(l > a.n_lists) return;
	if(l == a.n_lists) { row_ptr[l] = place[n_count]; return; }
	const int g = a.list_group[l];
	const int l0 = a.group_off[g], G = a.group_off[g + 1] - l0;
	const int T = (G + SIM_TILE - 1) / SIM_TILE;
	row_ptr[l] = place[cnt_off[g] + (int64_t)(l - l0) * T];
}

} // namespace agpu


#endif
