// Block-cooperative building blocks used by the per-bundle kernels: every thread of the CTA
// calls them with the same arguments; data lives in global (L2-resident) scratch.
#ifndef ALETSCH_B200_CSRC_BLOCKOPS_H
#define ALETSCH_B200_CSRC_BLOCKOPS_H

#include "dev.h"

namespace agpu {

#define AGPU_MAX_BLOCK 256        // largest CTA of the block-cooperative per-bundle kernels

#define SORT_SMEM_CAP 1024

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

} // namespace agpu

#endif
