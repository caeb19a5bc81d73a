// Decoding of the compact upload format (agpu_batch_packed, include/aletsch_gpu.h) into the plain per-hit arrays every other
// kernel reads.  The format exists for the PCIe link only: positions as 16-bit deltas inside a bundle, mate position and
// insert size as 16-bit offsets, CIGAR operations as 16-bit units, each with an escape to a short side list for the values
// that do not fit.  Nothing here has a counterpart in the reference; the decoded arrays are bit-identical to what
// agpu_batch_upload would have been given (tests: every stage's output is compared between the two uploads).
#ifndef ALETSCH_B200_CSRC_K_UNPACK_H
#define ALETSCH_B200_CSRC_K_UNPACK_H

#include "dev.h"

namespace agpu {

#define PACK_ESC_U16 0xFFFFu
#define PACK_ESC_I16 (-32768)
#define PACK_LONG_OP 15u
#define PACK_ESC_UNITS 63u
#define PACK_DEFAULT_UNIT 62u

struct packed_dev
{
	int64_t n_hits;
	int32_t n_bundles;
	const int64_t *bundle_hit_off;
	const int32_t *bundle_pos0;
	const uint16_t *dpos;
	const int16_t *dmpos, *isize16;
	const uint8_t *hit_meta;
	const uint16_t *units;
	u32 default_unit;
	int64_t n_esc_pos, n_esc_mpos, n_esc_isize, n_esc_units;
	const int64_t *esc_pos_idx, *esc_mpos_idx, *esc_isize_idx, *esc_units_idx;
	const int32_t *esc_pos_val, *esc_mpos_val, *esc_isize_val, *esc_units_val;
	int64_t n_units, n_cigar;       // sizes of units[] and of the decoded cigar[]: every index below is clamped against them
};

DEV int32_t packed_escape(const int64_t *idx, const int32_t *val, int64_t n, int64_t i, int *err)
{
	int64_t lo = 0, hi = n;
	while(lo < hi)
	{
		int64_t m = (lo + hi) >> 1;
		if(idx[m] < i) lo = m + 1; else hi = m;
	}
	if(lo < n && idx[lo] == i) return val[lo];
	atomicAdd(err, 1);
	return 0;
}

// per hit: position delta and unit count as int32 (the inputs of the two device-wide scans), and xs
KERNEL k_unpack_widen(packed_dev p, int32_t *d32, int32_t *nun32, uint8_t *xs, int *err)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= p.n_hits) return;
	u32 d = p.dpos[i];
	d32[i] = d == PACK_ESC_U16 ? packed_escape(p.esc_pos_idx, p.esc_pos_val, p.n_esc_pos, i, err) : (int32_t)d;
	const u32 m = p.hit_meta[i];
	const u32 n = m & PACK_ESC_UNITS;
	int32_t nu = n == PACK_ESC_UNITS ? packed_escape(p.esc_units_idx, p.esc_units_val, p.n_esc_units, i, err)
			: (n == PACK_DEFAULT_UNIT ? 0 : (int32_t)n);
	if(nu < 0 || (int64_t)nu > p.n_units) { atomicAdd(err, 1); nu = 0; }     // a hostile escape value must not wrap the scan
	nun32[i] = nu;
	const u32 x = m >> 6;
	if(x == 3) atomicAdd(err, 1);
	xs[i] = x == 1 ? '+' : (x == 2 ? '-' : '.');
}

// pos = pos of the bundle's first hit + the deltas up to the hit; mpos, isize
KERNEL k_unpack_pos(packed_dev p, const int64_t *dsum, const int32_t *d32, int32_t *pos, int32_t *mpos, int32_t *isize, int *err)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= p.n_hits) return;
	const int b = find_segment(p.bundle_hit_off, p.n_bundles, i);
	const int64_t f = p.bundle_hit_off[b];
	const int32_t x = p.bundle_pos0[b] + (int32_t)(dsum[i] + d32[i] - dsum[f] - d32[f]);
	pos[i] = x;
	int m = p.dmpos[i];
	mpos[i] = m == PACK_ESC_I16 ? packed_escape(p.esc_mpos_idx, p.esc_mpos_val, p.n_esc_mpos, i, err) : x + m;
	int s = p.isize16[i];
	isize[i] = s == PACK_ESC_I16 ? packed_escape(p.esc_isize_idx, p.esc_isize_val, p.n_esc_isize, i, err) : s;
}

// CIGAR operations of a hit: a unit is op | len << 4 (len < 4096), or 15 | (len & 0xfff) << 4 followed by op | (len >> 12) << 4
// Inputs are untrusted: the unit range of a hit is clamped to units[0, n_units) and a long operation whose second unit
// would lie outside it counts as a violation (ERR_PACKED) instead of being read.
KERNEL k_unpack_count_ops(packed_dev p, const int64_t *unit_off, int32_t *nops, int *err)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= p.n_hits) return;
	int n = (p.hit_meta[i] & PACK_ESC_UNITS) == PACK_DEFAULT_UNIT ? 1 : 0;
	int64_t u1 = unit_off[i + 1];
	if(u1 > p.n_units) { u1 = p.n_units; atomicAdd(err, 1); }
	if(i + 1 == p.n_hits && unit_off[i + 1] != p.n_units) atomicAdd(err, 1);
	for(int64_t u = unit_off[i]; u < u1; u++, n++)
		if((p.units[u] & 0xf) == PACK_LONG_OP) { if(u + 1 >= u1) { atomicAdd(err, 1); break; } u++; }
	nops[i] = n;
}

// The scatter is a no-op for the whole batch unless the decoded operation total equals n_cigar (the size cigar[] was
// allocated with): a disagreeing batch is reported (k_unpack_check_total / the host check) and never written.
KERNEL k_unpack_cigar(packed_dev p, const int64_t *unit_off, const int64_t *op_off, u32 *cigar_off, u32 *cigar)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i > p.n_hits) return;
	const bool sane = op_off[p.n_hits] == p.n_cigar;
	cigar_off[i] = sane ? (u32)op_off[i] : 0u;
	if(i == p.n_hits || !sane) return;
	int64_t o = op_off[i];
	const int64_t o1 = op_off[i + 1];
	int64_t u1 = unit_off[i + 1];
	if(u1 > p.n_units) u1 = p.n_units;
	if((p.hit_meta[i] & PACK_ESC_UNITS) == PACK_DEFAULT_UNIT && o < o1) cigar[o++] = p.default_unit;
	for(int64_t u = unit_off[i]; u < u1 && o < o1; u++)
	{
		u32 a = p.units[u];
		if((a & 0xf) == PACK_LONG_OP)
		{
			if(u + 1 >= u1) break;
			u32 c = p.units[++u];
			cigar[o++] = (((c >> 4) << 12 | (a >> 4)) << 4) | (c & 0xf);
		}
		else cigar[o++] = a;           // len << 4 | op, the BAM encoding itself
	}
}

KERNEL k_unpack_check_total(const int64_t *total, int64_t expect, int *err)
{
	if(blockIdx.x == 0 && threadIdx.x == 0 && *total != expect) atomicAdd(err, 1);
}

} // namespace agpu

#endif
