// Stage 3a kernels: mate pairing (bundle_base::build_fragments, rnacore/bundle_base.cc:267-323).
//
// Reference semantics: for i ascending, an unpaired hit i takes the first unpaired u != i (in index
// order) with hits[u].pos == hits[i].mpos, isize sum 0 and the same qname; frgs is ordered by i.
// The bucket hash of the reference only accelerates the search, so the result is a function of
// the groups of equal qname alone: the greedy is run independently inside every qname group.
// Groups are found with a per-bundle open-addressing table keyed by the 64-bit qname key; members
// are chained through an atomicExch linked list and sorted by index by the group's first hit.
#ifndef ALETSCH_B200_CSRC_K_FRAGMENTS_H
#define ALETSCH_B200_CSRC_K_FRAGMENTS_H

#include "runtime.h"
#include "k_evidence.h"

namespace agpu {

#define QID_EMPTY 0xffffffffffffffffULL
#define SCAN_TILE 2048

// qname table slot: two 64-bit words, [0] = qname key (QID_EMPTY when free), [1] low half = head of the member list
// (bundle-local hit index, -1 when empty).  A memset with 0xff initialises both.
// Both kernels work on one WAVE of bundles at a time: the hits [hit_lo, hit_hi) of a run of consecutive bundles, whose table
// regions start at slot_base in the batch's region offsets.  Default: ONE wave, one table for the whole batch.  With
// AGPU_PAIR_WAVE_SLOTS=<n> (e.g. 4194304 = 64 MB of slots) the same table is cleared, filled and read wave after wave and stays
// in the 126 MB L2, so the 16-byte slots never travel to DRAM (one table for the whole batch, 750 MB at configs[1], crosses the
// DRAM bus three times).  Measured on B200 at configs[1]: 13 waves 0.78 + 0.61 ms, one wave 0.62 + 0.43 ms -- the launch tails of
// 26 short kernels cost more than the DRAM traffic they save, so the waves are off by default (profiles/r02_notes.md).
KERNEL k_qid_insert(hits_dev h, int64_t hit_lo, int64_t hit_hi, int64_t slot_base, const int32_t *hit_bundle, const int64_t *reg_off, u64 *slots,
		int64_t *hit_qslot, int32_t *next, int *err)
{
	int64_t i = hit_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= hit_hi) return;
	int b = hit_bundle[i];
	int64_t r0 = reg_off[b] - slot_base;
	u32 mask = (u32)(reg_off[b + 1] - reg_off[b]) - 1;
	u64 key = h.qid[i];
	if(key == QID_EMPTY) { atomicAdd(&err[ERR_QID], 1); hit_qslot[i] = -1; return; }
	u32 pos = (u32)(mix64(key) >> 11) & mask;
	int64_t sl = -1;
	for(u32 probe = 0; probe <= mask; probe++)
	{
		u64 cur = atomicCAS(&slots[2 * (r0 + pos)], (u64)QID_EMPTY, key);
		if(cur == QID_EMPTY || cur == key) { sl = r0 + pos; break; }
		pos = (pos + 1) & mask;
	}
	hit_qslot[i] = sl;
	if(sl < 0) { atomicAdd(&err[ERR_CAP], 1); return; }
	int32_t li = (int32_t)(i - h.bundle_hit_off[b]);
	next[i] = atomicExch((int32_t*)&slots[2 * sl + 1], li);
}

// the reference's greedy over one qname group whose members m[0..n) are in ascending hit index
DEV void pair_group(const hits_dev &h, int64_t h0, const int32_t *m, int n, int32_t *mate)
{
	for(int a = 0; a < n; a++)
	{
		int64_t ia = h0 + m[a];
		if(mate[ia] != -1) continue;
		for(int c = 0; c < n; c++)
		{
			if(c == a) continue;
			int64_t ic = h0 + m[c];
			if(mate[ic] != -1) continue;
			if(h.pos[ic] != h.mpos[ia]) continue;
			if(h.isize[ic] + h.isize[ia] != 0) continue;
			mate[ia] = m[c];                 // i discovered the pair
			mate[ic] = -2 - m[a];            // partner
			break;
		}
	}
}

// one thread per qname group (the hit at the head of the group's member list) runs the greedy; frgs order is by
// discoverer index, so which member runs it does not matter
#define PAIR_LOCAL 8
KERNEL k_pair(hits_dev h, int64_t hit_lo, int64_t hit_hi, const int32_t *hit_bundle, const int64_t *hit_qslot, const u64 *slots, const int32_t *next,
		int32_t *cursor, int32_t *members, int32_t *mate)
{
	int64_t i = hit_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= hit_hi) return;
	int64_t sl = hit_qslot[i];
	if(sl < 0) return;
	int b = hit_bundle[i];
	int64_t h0 = h.bundle_hit_off[b];
	const int32_t li = (int32_t)(i - h0);
	if((int32_t)(u32)(slots[2 * sl + 1] & 0xffffffffULL) != li) return;
	int32_t x1 = next[i];
	if(x1 < 0) return;                       // group of one
	int32_t x2 = next[h0 + x1];
	if(x2 < 0)
	{
		// the common case, a group of two: members (lo, hi) in index order
		int32_t lo = li < x1 ? li : x1, hi = li < x1 ? x1 : li;
		int64_t ia = h0 + lo, ic = h0 + hi;
		int32_t sum = h.isize[ia] + h.isize[ic];
		if(sum != 0) return;
		if(h.pos[ic] == h.mpos[ia]) { mate[ia] = hi; mate[ic] = -2 - lo; }
		else if(h.pos[ia] == h.mpos[ic]) { mate[ic] = lo; mate[ia] = -2 - hi; }
		return;
	}
	int32_t loc[PAIR_LOCAL];
	int n = 0;
	for(int32_t x = li; x >= 0; x = next[h0 + x]) { if(n < PAIR_LOCAL) loc[n] = x; n++; }
	int32_t *m = loc;
	if(n > PAIR_LOCAL)
	{
		m = members + h0 + atomicAdd(&cursor[b], n);
		int k = 0;
		for(int32_t x = li; x >= 0 && k < n; x = next[h0 + x]) m[k++] = x;
	}
	// ascending hit index (insertion sort; groups are tiny)
	for(int a = 1; a < n; a++)
	{
		int32_t v = m[a];
		int c = a - 1;
		while(c >= 0 && m[c] > v) { m[c + 1] = m[c]; c--; }
		m[c + 1] = v;
	}
	pair_group(h, h0, m, n, mate);
}

// ---- generic device-wide exclusive scan of 0/1 flags derived from an int array (flag = v[i] >= 0)
KERNEL k_flag_tile_count(const int32_t *v, int64_t n, int64_t n_tiles, int32_t *tile_cnt)
{
	SHARED int s;
	for(int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x)
	{
		if(threadIdx.x == 0) s = 0;
		BLOCK_SYNC();
		int acc = 0;
		for(int i = threadIdx.x; i < SCAN_TILE; i += blockDim.x)
		{
			int64_t g = t * SCAN_TILE + i;
			if(g < n && v[g] >= 0) acc++;
		}
		atomicAdd(&s, acc);
		BLOCK_SYNC();
		if(threadIdx.x == 0) tile_cnt[t] = s;
		BLOCK_SYNC();
	}
}

// rank[i] = number of flagged elements before i (written for every i, plus rank[n] = total)
KERNEL k_flag_tile_rank(const int32_t *v, int64_t n, int64_t n_tiles, const int64_t *tile_off, int64_t *rank)
{
	SHARED int f[SCAN_TILE];
	for(int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x)
	{
		for(int i = threadIdx.x; i < SCAN_TILE; i += blockDim.x)
		{
			int64_t g = t * SCAN_TILE + i;
			f[i] = (g < n && v[g] >= 0) ? 1 : 0;
		}
		BLOCK_SYNC();
		block_excl_scan(f, SCAN_TILE);
		for(int i = threadIdx.x; i < SCAN_TILE; i += blockDim.x)
		{
			int64_t g = t * SCAN_TILE + i;
			if(g <= n) rank[g] = tile_off[t] + f[i];
		}
		BLOCK_SYNC();
	}
}

KERNEL k_frag_emit(hits_dev h, const int32_t *hit_bundle, const int32_t *mate, const int64_t *rank, int32_t *f_h1, int32_t *f_h2, int32_t *f_type,
		int32_t *f_bundle)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= h.n_hits) return;
	if(mate[i] < 0) return;
	int64_t f = rank[i];
	f_bundle[f] = hit_bundle[i];
	f_h1[f] = (int32_t)(i - h.bundle_hit_off[hit_bundle[i]]);
	f_h2[f] = mate[i];
	f_type[f] = 0;
}

KERNEL k_gather_i64(int64_t n, const int64_t *idx, const int64_t *src, int64_t *out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	out[i] = src[idx[i]];
}

} // namespace agpu

struct fragments_state
{
	bool built = false;
	agpu::dbuf<agpu::u64> slots;
	agpu::dbuf<int32_t> next, cursor, members, mate, tile_cnt;
	agpu::dbuf<int64_t> hit_qslot, tile_off, rank, frg_off;
	agpu::dbuf<int32_t> f_h1, f_h2, f_type, f_bundle, bridged;
	int64_t n_frg = 0;
	std::vector<int64_t> frg_off_host;

	void release(agpu_ctx *ctx)
	{
		slots.release(ctx); next.release(ctx);
		cursor.release(ctx); members.release(ctx); mate.release(ctx); tile_cnt.release(ctx);
		hit_qslot.release(ctx); tile_off.release(ctx); rank.release(ctx); frg_off.release(ctx);
		f_h1.release(ctx); f_h2.release(ctx); f_type.release(ctx); f_bundle.release(ctx); bridged.release(ctx);
		built = false; n_frg = 0;
	}
};

#endif
