// Result packing for the agpu_*_fetch calls: the stages keep their arrays in upper-bound (per-bundle padded) column layouts;
// these kernels write the dense, interleaved layouts of the views in include/aletsch_gpu.h on the device, so that a fetch is
// a handful of device -> pinned-host copies of exactly the bytes the view holds and the host never loops over elements.
#ifndef ALETSCH_B200_CSRC_K_FETCH_H
#define ALETSCH_B200_CSRC_K_FETCH_H

#include "dev.h"

namespace agpu {

#define FETCH_MAX_COLS 9

template<typename T> struct fetch_cols { const T *c[FETCH_MAX_COLS]; };

// rows [src_off[k] + skew * k, + cnt[k] + extra) of every bundle k -> out rows [dst_off[k], ...), W values per row interleaved
// (out[(dst_off[k] + x) * W + f] = cols.c[f][src_off[k] + skew * k + x]).  One warp-sized CTA per bundle, strided.
// negcol: a column whose negative values are written as -1 (removed edges keep -1 - src on the device), or -1
template<typename T> DEV void rows_gather(int32_t nb, int W, const int64_t *src_off, int skew, const int32_t *cnt, int extra,
		const int64_t *dst_off, const fetch_cols<T> &cols, T *out, int negcol)
{
	for(int k = blockIdx.x; k < nb; k += gridDim.x)
	{
		const int64_t s0 = src_off[k] + (int64_t)skew * k, d0 = dst_off[k];
		const int n = (cnt[k] + extra) * W;
		for(int i = threadIdx.x; i < n; i += blockDim.x)
		{
			const int x = i / W, f = i - x * W;
			T v = cols.c[f][s0 + x];
			if(f == negcol && v < 0) v = (T)-1;
			out[d0 * W + i] = v;
		}
	}
}
KERNEL k_rows_gather_i(int32_t nb, int W, const int64_t *src_off, int skew, const int32_t *cnt, int extra, const int64_t *dst_off,
		fetch_cols<int32_t> cols, int32_t *out, int negcol)
{
	rows_gather<int32_t>(nb, W, src_off, skew, cnt, extra, dst_off, cols, out, negcol);
}
KERNEL k_rows_gather_d(int32_t nb, int W, const int64_t *src_off, int skew, const int32_t *cnt, int extra, const int64_t *dst_off,
		fetch_cols<double> cols, double *out, int negcol)
{
	rows_gather<double>(nb, W, src_off, skew, cnt, extra, dst_off, cols, out, negcol);
}

// partial exons, double part of the view: ave dev max pvalue
KERNEL k_pexon_d_pack(int32_t nb, const int64_t *src_off, const int32_t *cnt, const int64_t *dst_off, const double *ave, const double *dev,
		const double *mx, const int32_t *ptype, double *out)
{
	for(int k = blockIdx.x; k < nb; k += gridDim.x)
	{
		const int64_t s0 = src_off[k], d0 = dst_off[k];
		for(int x = threadIdx.x; x < cnt[k]; x += blockDim.x)
		{
			double *o = out + 4 * (d0 + x);
			o[0] = ave[s0 + x]; o[1] = dev[s0 + x]; o[2] = mx[s0 + x]; o[3] = ptype[s0 + x] ? 1.0 : 0.0;
		}
	}
}

// flat interleave of W columns of n rows
KERNEL k_cols_interleave_i(int64_t n, int W, fetch_cols<int32_t> cols, int32_t *out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n * W) return;
	const int64_t x = i / W;
	const int f = (int)(i - x * W);
	out[i] = cols.c[f][x];
}

// per-bundle counts plus a constant -> int32 (input of the offset scans of derived row sets, e.g. vertices = pexons + 2)
KERNEL k_count_plus(int64_t nb, const int32_t *cnt, int add, int32_t *out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= nb) return;
	out[i] = cnt[i] + add;
}

KERNEL k_i64_to_i32(int64_t n, const int64_t *in, int32_t *out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	out[i] = (int32_t)in[i];
}

// chain_set -> dense view, pass 1: length, AI3 counts and group index of every chain at its dense position
KERNEL k_chains_pack_head(int32_t nb, const int64_t *elem_off, const int32_t *n_chains, const int64_t *chain_base, const int32_t *c_rep,
		const int32_t *c_cnt, const int32_t *c_grp, const int32_t *elem_len, int32_t *out_len, int32_t *out_cnt, int32_t *out_grp)
{
	for(int k = blockIdx.x; k < nb; k += gridDim.x)
	{
		const int64_t e0 = elem_off[k], d0 = chain_base[k];
		for(int c = threadIdx.x; c < n_chains[k]; c += blockDim.x)
		{
			out_len[d0 + c] = elem_len[e0 + c_rep[e0 + c]];
			out_cnt[3 * (d0 + c)] = c_cnt[3 * (e0 + c)]; out_cnt[3 * (d0 + c) + 1] = c_cnt[3 * (e0 + c) + 1]; out_cnt[3 * (d0 + c) + 2] = c_cnt[3 * (e0 + c) + 2];
			out_grp[d0 + c] = c_grp[e0 + c];
		}
	}
}

// pass 2: coordinates of every chain at chain_off (one thread per chain; chains are a few coordinates long)
KERNEL k_chains_pack_val(int32_t nb, const int64_t *elem_off, const int32_t *n_chains, const int64_t *chain_base, const int32_t *c_rep,
		const int32_t *elem_len, const u32 *voff32, const int64_t *voff64, const int32_t *val, const int64_t *chain_off, int32_t *out_off32, int32_t *out_val)
{
	for(int k = blockIdx.x; k < nb; k += gridDim.x)
	{
		const int64_t e0 = elem_off[k], d0 = chain_base[k];
		for(int c = threadIdx.x; c < n_chains[k]; c += blockDim.x)
		{
			const int64_t e = e0 + c_rep[e0 + c];
			const int64_t vo = voff64 ? voff64[e] : (int64_t)voff32[e];
			const int64_t o = chain_off[d0 + c];
			out_off32[d0 + c] = (int32_t)o;
			for(int x = 0; x < elem_len[e]; x++) out_val[o + x] = val[vo + x];
		}
	}
}

} // namespace agpu

#endif
