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

} // namespace agpu

#endif
