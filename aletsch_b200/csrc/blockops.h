// Block-cooperative building blocks used by the per-bundle kernels: every thread of the CTA
// calls them with the same arguments; data lives in global (L2-resident) scratch.
#ifndef ALETSCH_B200_CSRC_BLOCKOPS_H
#define ALETSCH_B200_CSRC_BLOCKOPS_H

#include "dev.h"

namespace agpu {

#define AGPU_MAX_BLOCK 256        // largest CTA of the block-cooperative per-bundle kernels

#define SORT_SMEM_CAP 512

// Normalised bitonic network (all compare-exchanges ascending; the first step of every merge
// pairs i with i ^ (k - 1)).  With this form, virtual +inf padding beyond n never moves, so
// any n works in place.  Keys must be distinct for a deterministic result (callers pack a
// unique rank into the low bits).  Up to SORT_SMEM_CAP keys are sorted in shared memory.
DEV void bitonic_u64(u64 *key, int n)
{
	for(int k = 2; (k >> 1) < n; k <<= 1)
	{
		for(int j = k >> 1, first = 1; j > 0; j >>= 1, first = 0)
		{
			for(int i = threadIdx.x; i < n; i += blockDim.x)
			{
				int p = first ? (i ^ (k - 1)) : (i ^ j);
				if(p > i && p < n)
				{
					u64 a = key[i], b = key[p];
					if(b < a) { key[i] = b; key[p] = a; }
				}
			}
			BLOCK_SYNC();
		}
	}
}

// one shared staging pool for both sorts: SORT_SMEM_CAP keys followed by SORT_SMEM_CAP 32-bit payloads
DEV u64 *sort_pool()
{
	SHARED u64 pool[SORT_SMEM_CAP + SORT_SMEM_CAP / 2];
	return pool;
}

DEV void block_sort_u64(u64 *key, int n)
{
	u64 *sk = sort_pool();
	if(n <= 1) return;
	if(n <= SORT_SMEM_CAP)
	{
		for(int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = key[i];
		BLOCK_SYNC();
		bitonic_u64(sk, n);
		for(int i = threadIdx.x; i < n; i += blockDim.x) key[i] = sk[i];
		BLOCK_SYNC();
	}
	else bitonic_u64(key, n);
}

DEV void bitonic_pairs(u64 *key, u32 *val, int n)
{
	for(int k = 2; (k >> 1) < n; k <<= 1)
	{
		for(int j = k >> 1, first = 1; j > 0; j >>= 1, first = 0)
		{
			for(int i = threadIdx.x; i < n; i += blockDim.x)
			{
				int p = first ? (i ^ (k - 1)) : (i ^ j);
				if(p > i && p < n)
				{
					u64 a = key[i], b = key[p];
					u32 va = val[i], vb = val[p];
					if(b < a || (b == a && vb < va)) { key[i] = b; key[p] = a; val[i] = vb; val[p] = va; }
				}
			}
			BLOCK_SYNC();
		}
	}
}

// same network over (key, val) pairs ordered by key then val
DEV void block_sort_pairs(u64 *key, u32 *val, int n)
{
	u64 *sk = sort_pool();
	u32 *sv = (u32*)(sk + SORT_SMEM_CAP);
	if(n <= 1) return;
	if(n <= SORT_SMEM_CAP)
	{
		for(int i = threadIdx.x; i < n; i += blockDim.x) { sk[i] = key[i]; sv[i] = val[i]; }
		BLOCK_SYNC();
		bitonic_pairs(sk, sv, n);
		for(int i = threadIdx.x; i < n; i += blockDim.x) { key[i] = sk[i]; val[i] = sv[i]; }
		BLOCK_SYNC();
	}
	else bitonic_pairs(key, val, n);
}

#ifndef AGPU_EMU
// exclusive prefix sum of a[0..n) in place; returns the total to every thread.  Coalesced sweeps of blockDim elements:
// warp-shuffle scan, warp totals through shared memory, running carry in a register.
DEV int block_excl_scan(int *a, int n)
{
	__shared__ int wsum[AGPU_MAX_BLOCK / 32];
	const int nt = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = (nt + 31) >> 5;
	if(n == nt * 8 && (nt & 31) == 0 && ((size_t)a & 15) == 0)
	{
		// a full tile of eight values per thread (the look-back kernels): every thread scans eight CONSECUTIVE values in registers
		// (two 128-bit shared-memory accesses each way), then one scan of the thread sums -- three barriers instead of seventeen
		__syncthreads();
		int4 *p = (int4*)(a + t * 8);
		int4 u = p[0], v = p[1];
		const int x0 = u.x, x1 = u.y, x2 = u.z, x3 = u.w, x4 = v.x, x5 = v.y, x6 = v.z, x7 = v.w;
		const int tsum = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
		int inc = tsum;
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, inc, o); if(lane >= o) inc += y; }
		if(lane == 31) wsum[w] = inc;
		__syncthreads();
		int pre = 0, tot = 0;
		for(int k = 0; k < nw; k++) { int s = wsum[k]; if(k < w) pre += s; tot += s; }
		int run = pre + inc - tsum;
		u.x = run; run += x0; u.y = run; run += x1; u.z = run; run += x2; u.w = run; run += x3;
		v.x = run; run += x4; v.y = run; run += x5; v.z = run; run += x6; v.w = run;
		p[0] = u; p[1] = v;
		__syncthreads();
		return tot;
	}
	int carry = 0;
	for(int base = 0; base < n; base += nt)
	{
		const int i = base + t;
		const int x = i < n ? a[i] : 0;
		int inc = x;
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, inc, o); if(lane >= o) inc += y; }
		if(nw > 1)
		{
			if(lane == 31) wsum[w] = inc;
			__syncthreads();
			int pre = 0, tot = 0;
			for(int k = 0; k < nw; k++) { int v = wsum[k]; if(k < w) pre += v; tot += v; }
			if(i < n) a[i] = carry + pre + inc - x;
			carry += tot;
			__syncthreads();
		}
		else
		{
			if(i < n) a[i] = carry + inc - x;
			carry += __shfl_sync(0xffffffffu, inc, 31);
		}
	}
	__syncthreads();
	return carry;
}

// inclusive running maximum of a[0..n) in place
DEV void block_incl_maxscan(int *a, int n)
{
	__shared__ int wmax[AGPU_MAX_BLOCK / 32];
	const int nt = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = (nt + 31) >> 5;
	const int NEG = -0x7fffffff;
	int carry = NEG;
	for(int base = 0; base < n; base += nt)
	{
		const int i = base + t;
		int inc = i < n ? a[i] : NEG;
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, inc, o); if(lane >= o && y > inc) inc = y; }
		if(nw > 1)
		{
			if(lane == 31) wmax[w] = inc;
			__syncthreads();
			int pre = NEG, tot = NEG;
			for(int k = 0; k < nw; k++) { int v = wmax[k]; if(k < w && v > pre) pre = v; if(v > tot) tot = v; }
			int r = inc > pre ? inc : pre;
			if(carry > r) r = carry;
			if(i < n) a[i] = r;
			if(tot > carry) carry = tot;
			__syncthreads();
		}
		else
		{
			int r = inc > carry ? inc : carry;
			if(i < n) a[i] = r;
			int tot = __shfl_sync(0xffffffffu, inc, 31);
			if(tot > carry) carry = tot;
		}
	}
	__syncthreads();
}
#else
// exclusive prefix sum of a[0..n) in place; returns the total to every thread
DEV int block_excl_scan(int *a, int n)
{
	SHARED int part[AGPU_MAX_BLOCK];
	SHARED int total;
	int nt = blockDim.x, t = threadIdx.x;
	int chunk = (n + nt - 1) / nt;
	int lo = t * chunk, hi = lo + chunk;
	if(lo > n) lo = n;
	if(hi > n) hi = n;
	int s = 0;
	for(int i = lo; i < hi; i++) s += a[i];
	part[t] = s;
	BLOCK_SYNC();
	if(t == 0)
	{
		int run = 0;
		for(int k = 0; k < nt; k++) { int v = part[k]; part[k] = run; run += v; }
		total = run;
	}
	BLOCK_SYNC();
	int run = part[t];
	for(int i = lo; i < hi; i++) { int v = a[i]; a[i] = run; run += v; }
	BLOCK_SYNC();
	int r = total;
	BLOCK_SYNC();
	return r;
}

// inclusive running maximum of a[0..n) in place
DEV void block_incl_maxscan(int *a, int n)
{
	SHARED int part[AGPU_MAX_BLOCK];
	int nt = blockDim.x, t = threadIdx.x;
	int chunk = (n + nt - 1) / nt;
	int lo = t * chunk, hi = lo + chunk;
	if(lo > n) lo = n;
	if(hi > n) hi = n;
	int s = -0x7fffffff;
	for(int i = lo; i < hi; i++) if(a[i] > s) s = a[i];
	part[t] = s;
	BLOCK_SYNC();
	if(t == 0)
	{
		int run = -0x7fffffff;
		for(int k = 0; k < nt; k++) { int v = part[k]; part[k] = run; if(v > run) run = v; }
	}
	BLOCK_SYNC();
	int run = part[t];
	for(int i = lo; i < hi; i++) { if(a[i] > run) run = a[i]; a[i] = run; }
	BLOCK_SYNC();
}

#endif

} // namespace agpu

#endif
