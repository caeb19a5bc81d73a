// Thin execution layer under the kernels.
//
// Product build (nvcc, sm_100a): kernels are __global__ functions launched on the context's
// CUDA stream.  There is NO host execution path in the product library.
//
// Kernel-logic test build (g++ -DAGPU_EMU, tests/emu only, never linked into
// libaletsch_gpu.so): the same kernel bodies are compiled for the host and run with a serial
// thread model so that the index arithmetic, orderings and tie-breaks can be checked against
// the reference in the CPU-only test tier.  Per-thread kernels run every (block, thread)
// in turn; block-cooperative kernels (written as strided loops separated by block barriers)
// run with blockDim.x == 1, where every strided loop degenerates to a full loop and the
// barrier to a no-op.
#ifndef ALETSCH_B200_CSRC_DEV_H
#define ALETSCH_B200_CSRC_DEV_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>

#ifndef AGPU_EMU
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#define KERNEL __global__ void
// thread-per-item kernel of 256 threads whose register count is capped so that n CTAs fit an SM: these kernels wait on chains of
// dependent loads, the resident warps are their memory-level parallelism (measured at configs[1]: k_frag_group 0.40 -> 0.31 ms,
// k_frag_align 0.36 -> 0.29, k_update 0.37 -> 0.30, k_cluster_emit 0.34 -> 0.32, k_vote_type2 0.44 -> 0.41; the warp-per-item
// kernels k_graph_build, k_group_partition_warp and k_bridge_dp_warp got slower or stayed put under a cap and keep their registers)
#define KERNEL_OCC(n) __global__ void __launch_bounds__(256, n)
#define SHARED __shared__
#define SHARED16 __shared__ __align__(16)        // tiles block_excl_scan reads with 128-bit accesses
#define BLOCK_SYNC() __syncthreads()
#else
// ---------------------------------------------------------------- host emulation (tests only)
#define HD inline
#define DEV inline
#define KERNEL static void
#define KERNEL_OCC(n) static void
#define SHARED static thread_local
#define SHARED16 static thread_local
#define BLOCK_SYNC() do {} while(0)
struct agpu_emu_dim { unsigned x, y, z; };
extern thread_local agpu_emu_dim threadIdx, blockIdx, blockDim, gridDim;
typedef void *cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
template<typename T> inline T atomicAdd(T *p, T v) { T o = *p; *p = (T)(o + v); return o; }
template<typename T> inline T atomicMin(T *p, T v) { T o = *p; if(v < o) *p = v; return o; }
template<typename T> inline T atomicMax(T *p, T v) { T o = *p; if(v > o) *p = v; return o; }
template<typename T> inline T atomicOr(T *p, T v) { T o = *p; *p = (T)(o | v); return o; }
template<typename T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template<typename T> inline T atomicCAS(T *p, T c, T v) { T o = *p; if(o == c) *p = v; return o; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
#endif

namespace agpu {

typedef unsigned long long u64;
typedef unsigned int u32;

HD u64 mix64(u64 z)
{
	z += 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

HD u32 pow2_ceil(u32 x)
{
	u32 p = 1;
	while(p < x) p <<= 1;
	return p;
}

// first index in sorted a[0..n) with a[i] >= v
template<typename T> HD int lower_bound_idx(const T *a, int n, T v)
{
	int lo = 0, hi = n;
	while(lo < hi)
	{
		int m = (lo + hi) >> 1;
		if(a[m] < v) lo = m + 1; else hi = m;
	}
	return lo;
}

// first index with a[i] > v
template<typename T> HD int upper_bound_idx(const T *a, int n, T v)
{
	int lo = 0, hi = n;
	while(lo < hi)
	{
		int m = (lo + hi) >> 1;
		if(!(v < a[m])) lo = m + 1; else hi = m;
	}
	return lo;
}

// index b with off[b] <= i < off[b+1]
HD int find_segment(const int64_t *off, int nseg, int64_t i)
{
	int lo = 0, hi = nseg;
	while(lo < hi)
	{
		int m = (lo + hi) >> 1;
		if(off[m + 1] <= i) lo = m + 1; else hi = m;
	}
	return lo;
}

} // namespace agpu

#endif
