// Device-callable re-implementation of the PERMUTATION produced by libstdc++'s std::sort
// (g++ 13, bits/stl_algo.h: __sort / __introsort_loop / __unguarded_partition_pivot /
// __move_median_to_first / __final_insertion_sort; bits/stl_heap.h for the depth-limit
// fallback).  The reference sorts with comparators that are only partial orders
// (rnacore/graph_cluster.cc:183-186, bridge/bridge_solver.cc:525, :262, :272), and std::sort
// is unstable above 16 elements, so the order of tied elements -- which leaks into
// pereads_cluster::bounds, the K-cut of the bridging DP and the bridge chosen by vote -- is
// whatever this exact algorithm leaves.  The permutation depends only on the input order and
// on comparator outcomes, so running the same algorithm over an array of handles with the
// same comparator reproduces it.
#ifndef ALETSCH_B200_CSRC_STDSORT_H
#define ALETSCH_B200_CSRC_STDSORT_H

#include "dev.h"

namespace agpu {

// a[] holds handles (or small self-contained elements); less(x, y) compares the elements they stand for
template<typename T, typename Less>
struct std_sort_emul
{
	T *a;
	Less less;

	HD std_sort_emul(T *arr, Less l) : a(arr), less(l) {}

	HD void swap_at(int i, int j) { T t = a[i]; a[i] = a[j]; a[j] = t; }

	HD void unguarded_linear_insert(int last)
	{
		T val = a[last];
		int next = last - 1;
		while(less(val, a[next]))
		{
			a[last] = a[next];
			last = next;
			--next;
		}
		a[last] = val;
	}

	HD void insertion_sort(int first, int last)
	{
		if(first == last) return;
		for(int i = first + 1; i != last; ++i)
		{
			if(less(a[i], a[first]))
			{
				T val = a[i];
				for(int k = i; k > first; --k) a[k] = a[k - 1];
				a[first] = val;
			}
			else unguarded_linear_insert(i);
		}
	}

	HD void final_insertion_sort(int first, int last)
	{
		if(last - first > 16)
		{
			insertion_sort(first, first + 16);
			for(int i = first + 16; i != last; ++i) unguarded_linear_insert(i);
		}
		else insertion_sort(first, last);
	}

	HD void move_median_to_first(int result, int x, int y, int z)
	{
		if(less(a[x], a[y]))
		{
			if(less(a[y], a[z])) swap_at(result, y);
			else if(less(a[x], a[z])) swap_at(result, z);
			else swap_at(result, x);
		}
		else if(less(a[x], a[z])) swap_at(result, x);
		else if(less(a[y], a[z])) swap_at(result, z);
		else swap_at(result, y);
	}

	HD int unguarded_partition(int first, int last, int pivot)
	{
		while(true)
		{
			while(less(a[first], a[pivot])) ++first;
			--last;
			while(less(a[pivot], a[last])) --last;
			if(!(first < last)) return first;
			swap_at(first, last);
			++first;
		}
	}

	// ---- heap fallback (depth limit reached) ----
	HD void push_heap(int first, int hole, int top, T value)
	{
		int parent = (hole - 1) / 2;
		while(hole > top && less(a[first + parent], value))
		{
			a[first + hole] = a[first + parent];
			hole = parent;
			parent = (hole - 1) / 2;
		}
		a[first + hole] = value;
	}

	HD void adjust_heap(int first, int hole, int len, T value)
	{
		const int top = hole;
		int second = hole;
		while(second < (len - 1) / 2)
		{
			second = 2 * (second + 1);
			if(less(a[first + second], a[first + (second - 1)])) second--;
			a[first + hole] = a[first + second];
			hole = second;
		}
		if((len & 1) == 0 && second == (len - 2) / 2)
		{
			second = 2 * (second + 1);
			a[first + hole] = a[first + (second - 1)];
			hole = second - 1;
		}
		push_heap(first, hole, top, value);
	}

	HD void pop_heap(int first, int last, int result)
	{
		T value = a[result];
		a[result] = a[first];
		adjust_heap(first, 0, last - first, value);
	}

	HD void make_heap(int first, int last)
	{
		if(last - first < 2) return;
		const int len = last - first;
		int parent = (len - 2) / 2;
		while(true)
		{
			T value = a[first + parent];
			adjust_heap(first, parent, len, value);
			if(parent == 0) return;
			parent--;
		}
	}

	HD void partial_sort_all(int first, int last)
	{
		// __partial_sort(first, last, last): heap_select over an empty tail, then sort_heap
		make_heap(first, last);
		int l = last;
		while(l - first > 1)
		{
			--l;
			pop_heap(first, l, l);
		}
	}

	HD void sort(int first, int last)
	{
		if(first == last) return;
		int n = last - first;
		int lg = 0;
		while((n >> (lg + 1)) != 0) lg++;          // std::__lg
		// __introsort_loop with an explicit stack: the recursive call handles [cut, last) and the
		// loop continues on [first, cut); the two ranges are disjoint, so the order is immaterial
		int st_first[64], st_last[64], st_depth[64];
		int sp = 0;
		st_first[0] = first; st_last[0] = last; st_depth[0] = lg * 2; sp = 1;
		while(sp > 0)
		{
			--sp;
			int f = st_first[sp], l = st_last[sp], d = st_depth[sp];
			while(l - f > 16)
			{
				if(d == 0) { partial_sort_all(f, l); break; }
				--d;
				int mid = f + (l - f) / 2;
				move_median_to_first(f, f + 1, mid, l - 1);
				int cut = unguarded_partition(f + 1, l, f);
				st_first[sp] = cut; st_last[sp] = l; st_depth[sp] = d; sp++;
				l = cut;
			}
		}
		final_insertion_sort(first, last);
	}
};

template<typename T, typename Less>
HD void std_sort_handles(T *a, int n, Less less)
{
	std_sort_emul<T, Less> s(a, less);
	s.sort(0, n);
}

#ifndef AGPU_EMU
// Warp-cooperative version producing the SAME permutation as std_sort_emul::sort (hence as std::sort).
// All 32 lanes call it with identical arguments.  scratch holds 3 * n ints.
//
// Why the parallel steps are exact:
//  * __unguarded_partition swaps, for k = 1, 2, ..., the k-th element from the left that is not less than the
//    pivot with the k-th element from the right that the pivot is not less than, as long as the former lies left
//    of the latter.  Both position lists depend only on the array as it was when the partition started, so they
//    can be produced by two compactions and the swaps applied in parallel; the returned cut is
//    min(next left candidate, last right position swapped).
//  * After __introsort_loop every range of at most 16 elements is ordered relative to its neighbours, and
//    __final_insertion_sort never moves an element across such a boundary (it stops at the first element that is
//    not greater), so it equals an independent stable insertion sort of every such range.
template<typename T, typename Less>
__device__ void warp_std_sort(T *a, int n, Less less, int *scratch)
{
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	if(n <= 1) return;
	std_sort_emul<T, Less> seq(a, less);
	if(n <= 16) { if(lane == 0) seq.insertion_sort(0, n); __syncwarp(); return; }
	int *Lp = scratch, *Rp = scratch + n, *leaf = scratch + 2 * n;
	int nleaf = 0;
	int lg = 0;
	while((n >> (lg + 1)) != 0) lg++;
	int st_first[64], st_last[64], st_depth[64];
	int sp = 1;
	st_first[0] = 0; st_last[0] = n; st_depth[0] = lg * 2;
	while(sp > 0)
	{
		--sp;
		int f = st_first[sp], l = st_last[sp], d = st_depth[sp];
		while(l - f > 16)
		{
			if(d == 0)
			{
				if(lane == 0) seq.partial_sort_all(f, l);
				__syncwarp();
				f = l;                      // sorted: nothing left for the insertion pass
				break;
			}
			--d;
			if(lane == 0) seq.move_median_to_first(f, f + 1, f + (l - f) / 2, l - 1);
			__syncwarp();
			const T pv = a[f];
			int nL = 0, nR = 0;
			for(int base = f + 1; base < l; base += 32)
			{
				int i = base + lane;
				bool valid = i < l;
				T x = valid ? a[i] : pv;
				bool ge = valid && !less(x, pv);
				bool le = valid && !less(pv, x);
				unsigned mg = __ballot_sync(FULL, ge), ml = __ballot_sync(FULL, le);
				unsigned below = (1u << lane) - 1u;
				if(ge) Lp[nL + __popc(mg & below)] = i;
				if(le) Rp[nR + __popc(ml & below)] = i;
				nL += __popc(mg); nR += __popc(ml);
			}
			__syncwarp();
			// K = number of swaps: pairs (Lp[k], Rp[nR - 1 - k]) with the left one strictly left of the right one
			int m = nL < nR ? nL : nR;
			int cnt = 0;
			for(int k = lane; k < m; k += 32) if(Lp[k] < Rp[nR - 1 - k]) cnt++;
			for(int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
			const int K = cnt;
			for(int k = lane; k < K; k += 32)
			{
				int x = Lp[k], y = Rp[nR - 1 - k];
				T t = a[x]; a[x] = a[y]; a[y] = t;
			}
			int cut = l;
			if(K > 0) cut = Rp[nR - K];
			if(K < nL && Lp[K] < cut) cut = Lp[K];
			__syncwarp();
			st_first[sp] = cut; st_last[sp] = l; st_depth[sp] = d; sp++;
			l = cut;
		}
		if(l - f > 0) { if(lane == 0) leaf[nleaf] = (f << 5) | (l - f); nleaf++; }
	}
	__syncwarp();
	for(int k = lane; k < nleaf; k += 32)
	{
		int f = leaf[k] >> 5, len = leaf[k] & 31;
		// stable insertion sort of a[f .. f + len)
		for(int i = f + 1; i < f + len; i++)
		{
			T val = a[i];
			int j = i - 1;
			while(j >= f && less(val, a[j])) { a[j + 1] = a[j]; j--; }
			a[j + 1] = val;
		}
	}
	__syncwarp();
}

// The introsort loop of warp_std_sort alone, for arrays in shared memory: what it leaves is a sequence of leaves of at most 16
// elements whose final insertion sorts are independent STABLE sorts, so the caller can finish all leaves of all its ranges in
// one pass in which every element ranks itself inside its leaf.  seg[i] = (first element of i's leaf, relative to `base`) << 8 |
// leaf length; elements of heap-sorted stretches get a leaf of their own.  scratch holds 2 * n ints.
template<typename T, typename Less>
__device__ void warp_std_sort_loop(T *a, int n, Less less, int *scratch, int *seg, int base)
{
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	if(n <= 16) { for(int i = lane; i < n; i += 32) seg[i] = (base << 8) | n; __syncwarp(); return; }
	std_sort_emul<T, Less> seq(a, less);
	int *Lp = scratch, *Rp = scratch + n;
	int lg = 0;
	while((n >> (lg + 1)) != 0) lg++;
	int st_first[64], st_last[64], st_depth[64];
	int sp = 1;
	st_first[0] = 0; st_last[0] = n; st_depth[0] = lg * 2;
	while(sp > 0)
	{
		--sp;
		int f = st_first[sp], l = st_last[sp], d = st_depth[sp];
		while(l - f > 16)
		{
			if(d == 0)
			{
				if(lane == 0) seq.partial_sort_all(f, l);
				for(int i = f + lane; i < l; i += 32) seg[i] = ((base + i) << 8) | 1;
				__syncwarp();
				f = l;
				break;
			}
			--d;
			if(lane == 0) seq.move_median_to_first(f, f + 1, f + (l - f) / 2, l - 1);
			__syncwarp();
			const T pv = a[f];
			int nL = 0, nR = 0;
			for(int b0 = f + 1; b0 < l; b0 += 32)
			{
				int i = b0 + lane;
				bool valid = i < l;
				T x = valid ? a[i] : pv;
				bool ge = valid && !less(x, pv);
				bool le = valid && !less(pv, x);
				unsigned mg = __ballot_sync(FULL, ge), ml = __ballot_sync(FULL, le);
				unsigned below = (1u << lane) - 1u;
				if(ge) Lp[nL + __popc(mg & below)] = i;
				if(le) Rp[nR + __popc(ml & below)] = i;
				nL += __popc(mg); nR += __popc(ml);
			}
			__syncwarp();
			int m = nL < nR ? nL : nR;
			int cnt = 0;
			for(int k = lane; k < m; k += 32) if(Lp[k] < Rp[nR - 1 - k]) cnt++;
			for(int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
			const int K = cnt;
			for(int k = lane; k < K; k += 32)
			{
				int x = Lp[k], y = Rp[nR - 1 - k];
				T t = a[x]; a[x] = a[y]; a[y] = t;
			}
			int cut = l;
			if(K > 0) cut = Rp[nR - K];
			if(K < nL && Lp[K] < cut) cut = Lp[K];
			__syncwarp();
			st_first[sp] = cut; st_last[sp] = l; st_depth[sp] = d; sp++;
			l = cut;
		}
		if(l - f > 0 && lane < l - f) seg[f + lane] = ((base + f) << 8) | (l - f);
	}
	__syncwarp();
}
#endif

} // namespace agpu

#endif
