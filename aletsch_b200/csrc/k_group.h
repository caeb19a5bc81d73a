// Kernels of the group-level re-bridge (assembler::bridge, meta/assembler.cc:977-1018): the bundles of one cluster (found by
// bundle_group::resolve) are merged into a combined bundle -- chain sets added, coverage maps added, bounds joined
// (bundle::combine, meta/bundle.cc:90-107; combine_bundles, meta/assembler.cc:152-175) -- whose splice graph every member
// is then clustered and bridged against.  The combined bundles of all clusters form an internal batch of their own, so
// graph_builder, the coverage scan and the chain-set ordering are the kernels of k_evidence.h / k_graph.h unchanged.
#ifndef ALETSCH_B200_CSRC_K_GROUP_H
#define ALETSCH_B200_CSRC_K_GROUP_H

#include "dev.h"
#include "k_evidence.h"
#include "k_graph.h"

namespace agpu {

// bounds of the combined bundles: copy_meta_information of gv[0] (meta/bundle.cc:45-53), then min / max over the members
KERNEL k_cb_bounds(int64_t ng, const int64_t *gm_off, const int32_t *gm, const int32_t *first_member, const int32_t *b_lpos,
		const int32_t *b_rpos, const int32_t *b_covhi, const uint8_t *b_strand, const int64_t *hit_off,
		int32_t *c_lpos, int32_t *c_rpos, int32_t *c_covhi, uint8_t *c_strand, int64_t *c_span, int *err)
{
	int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(g >= ng) return;
	int32_t lo = 1 << 30, hi = 0, cov = 0;
	bool any = false;
	const uint8_t st = b_strand[first_member[g]];
	for(int64_t j = gm_off[g]; j < gm_off[g + 1]; j++)
	{
		int m = gm[j];
		if(b_lpos[m] < lo) lo = b_lpos[m];
		if(b_rpos[m] > hi) hi = b_rpos[m];
		if(b_covhi[m] > cov) cov = b_covhi[m];
		if(hit_off[m + 1] > hit_off[m]) any = true;
		if(b_strand[m] != st) atomicAdd(&err[ERR_STRAND], 1);       // the reference asserts strand == bb.strand (meta/bundle.cc:93)
	}
	c_lpos[g] = lo; c_rpos[g] = hi; c_covhi[g] = cov; c_strand[g] = st;
	int64_t span = any ? ((int64_t)cov - (int64_t)lo + 1) : 0;
	c_span[g] = (span + COV_ALIGN - 1) / COV_ALIGN * COV_ALIGN;
}

// mmap += bb.mmap: every border of every member becomes a weighted point of the combined map (window position, difference)
// and a border bit of the combined bitmap.  One CTA per member (in combine order).
KERNEL k_cb_points(int64_t n_members, const int32_t *gm, const int32_t *m_group, const int64_t *pt_off, const int64_t *bord_off,
		const int32_t *posc, const int32_t *diffc, const int64_t *c_cov_base, const int32_t *c_lpos, u32 *c_border,
		int64_t *pt_g, int32_t *pt_d)
{
	for(int64_t j = blockIdx.x; j < n_members; j += gridDim.x)
	{
		const int m = gm[j], G = m_group[j];
		const int64_t i0 = bord_off[m];
		const int n = (int)(bord_off[m + 1] - i0);
		const int64_t base = c_cov_base[G] - (int64_t)c_lpos[G];
		for(int i = threadIdx.x; i < n; i += blockDim.x)
		{
			int64_t g = base + posc[i0 + i];
			pt_g[pt_off[j] + i] = g;
			pt_d[pt_off[j] + i] = diffc[i0 + i];
			atomicOr(&c_border[g >> 5], 1u << (g & 31));
		}
	}
}

// weighted points into the ranked difference array
KERNEL k_cov_add_points(int64_t n, const int64_t *pt_g, const int32_t *pt_d, const u32 *border, const u32 *wrank, int32_t *diffc)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	if(pt_d[i] != 0) atomicAdd(&diffc[border_rank(border, wrank, pt_g[i])], pt_d[i]);
}

// chain_set::add(const chain_set&) (rnacore/chain_set.cc): the elements of a combined chain set are the chains of the
// members in combine order, each with its AI3 counts
KERNEL k_cb_chain_count(int64_t n_members, const int32_t *gm, chains_view cv, int32_t *cnt)
{
	int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(j >= n_members) return;
	cnt[j] = cv.count(gm[j]);
}

KERNEL k_cb_chain_fill(int64_t n_members, const int32_t *gm, chains_view cv, const int64_t *el_off, int64_t *e_voff, int32_t *e_len,
		int32_t *e_cnt3)
{
	for(int64_t j = blockIdx.x; j < n_members; j += gridDim.x)
	{
		const int m = gm[j];
		const int n = cv.count(m);
		const int64_t e0 = cv.present ? cv.elem_off[m] : 0;
		for(int k = threadIdx.x; k < n; k += blockDim.x)
		{
			int64_t rep = e0 + cv.c_rep[e0 + k];
			int64_t o = el_off[j] + k;
			e_voff[o] = cv.voff64 ? cv.voff64[rep] : (int64_t)cv.voff32[rep];
			e_len[o] = cv.elem_len[rep];
			e_cnt3[3 * o] = cv.c_cnt[(e0 + k) * 3]; e_cnt3[3 * o + 1] = cv.c_cnt[(e0 + k) * 3 + 1]; e_cnt3[3 * o + 2] = cv.c_cnt[(e0 + k) * 3 + 2];
		}
	}
}

// the combined chain set keeps its own copy of the element coordinates (the members' arrays are replaced by later
// update rounds): element e moves from src[src_off[e] ..] to dst[dst_off[e] ..], and src_off[e] becomes dst_off[e]
KERNEL k_cb_chain_copy(int64_t n_elem, const int32_t *src, int64_t *e_voff, const int32_t *e_len, const int64_t *dst_off, int32_t *dst)
{
	int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(e >= n_elem) return;
	const int32_t *s = src + e_voff[e];
	int32_t *d = dst + dst_off[e];
	for(int i = 0; i < e_len[e]; i++) d[i] = s[i];
	e_voff[e] = dst_off[e];
}

KERNEL k_diff_i64_to_i32(int64_t n, const int64_t *off, int32_t *out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	out[i] = (int32_t)(off[i + 1] - off[i]);
}

// insertion of elements that carry counts (chain_set::add(v, AI3), rnacore/chain_set.cc:24-62)
KERNEL k_chain_insert_cnt(int64_t n_ent, int32_t n_groups, const int64_t *ent_goff, const int32_t *ent_cnt3, const int32_t *ent_len,
		const int64_t *ent_voff, const int32_t *val, const int64_t *reg_off, u64 *slot_word, int32_t *slot_first, int32_t *slot_cnt,
		int64_t *ent_slot, int *err)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n_ent) return;
	ent_slot[i] = -1;
	if(ent_len[i] <= 0) return;
	int b = find_segment(ent_goff, n_groups, i);
	chain_src s;
	s.val = val; s.off = ent_voff; s.off32 = NULL; s.len = ent_len;
	int64_t r0 = reg_off[b];
	u32 rs = (u32)(reg_off[b + 1] - r0);
	u64 hsh = chain_hash(val + ent_voff[i], ent_len[i]);
	int64_t sl = chain_table_insert(slot_word, r0, rs, s, i, hsh);
	ent_slot[i] = sl;
	if(sl < 0) { atomicAdd(&err[ERR_CAP], 1); return; }
	for(int x = 0; x < 3; x++) if(ent_cnt3[3 * i + x]) atomicAdd(&slot_cnt[sl * 3 + x], ent_cnt3[3 * i + x]);
	atomicMin(&slot_first[sl], (int32_t)(i - ent_goff[b]));
}

} // namespace agpu

#endif
