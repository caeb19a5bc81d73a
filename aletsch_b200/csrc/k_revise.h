// Boundary revision of the bundles' splice graphs, the step transform(bd, gr, true) applies after the graph is built
// (meta/assembler.cc:930-944): identify_boundaries (rnacore/graph_reviser.cc:1068-1283) adds start edges 0 -> a and end edges
// b -> n where the coverage of a continuous run of vertices dwarfs what enters (leaves) it through junctions, and
// remove_false_boundaries (:1285-1377) annotates the vertices at which paired fragments that stayed unbridged leave / enter.
// The refine_splice_graph that follows is a no-op here: the graph was refined when it was built and edges are only added.
//
// The reference re-evaluates every vertex after every edge it adds.  What a vertex x contributes never changes -- its run
// [a(x), x] (left_continuous_extend), the largest vertex weight in the run and the weight entering the run from outside are
// functions of the built graph -- only whether the run already holds a start edge does.  So the candidates are evaluated
// once, and every round is an arg-max over the runs still free of a start edge (ties go to the larger x, like the
// reference's `if(r < bestr) continue`), until the best ratio falls below min_boundary_log_ratio.  Start and end rounds do
// not see each other's edges; the reference alternates them, which fixes the order in which the edges are added.
#ifndef ALETSCH_B200_CSRC_K_REVISE_H
#define ALETSCH_B200_CSRC_K_REVISE_H

#include "k_bridge.h"

namespace agpu {

#define REVISE_BLOCK 64

struct revise_dev
{
	// per vertex, compact: the vertices of bundle b start at voff[b] (scan of n_pex + 2)
	const int64_t *voff;
	int32_t *run;                    // a(x) of the start candidates, then reused for b(x) of the end candidates
	double *ratio, *weight;          // log(2 + maxcov) / log(2 + sum), maxcov - sum
	int32_t *open;                   // candidate still usable
	int32_t *has_s, *has_e;          // edge 0 -> x / x -> n present (built graph, then the added ones)
	int32_t *leave_cnt, *come_cnt;   // fb1 / fb2 of remove_false_boundaries (filled by k_revise_unbridged)
	double *leave_ratio, *come_ratio;
	// per bundle, up to 2 (V - 2) added edges at 2 * voff[b]: starts from the front, ends from the middle
	int32_t *add_v;                  // the inner vertex of the added edge
	double *add_w;
	int32_t *n_start, *n_end;
	double min_ratio;
};

// fb1 / fb2 (rnacore/graph_reviser.cc:1287-1320): one thread per fragment
KERNEL k_revise_unbridged(int64_t n_frg, hits_dev h, graph_dev g, const int32_t *f_bundle, const int32_t *f_h1, const int32_t *f_h2,
		const int32_t *f_type, revise_dev r)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(f >= n_frg || f_type[f] != 0) return;
	const int b = f_bundle[f];
	const int64_t h0 = h.bundle_hit_off[b];
	gview gv = graph_of(g, b);
	int u1 = locate_vertex(gv, h.rpos[h0 + f_h1[f]] - 1), u2 = locate_vertex(gv, h.pos[h0 + f_h2[f]]);
	if(u1 < 0 || u2 < 0 || u1 >= u2) return;
	const int64_t v0 = r.voff[b];
	atomicAdd(&r.leave_cnt[v0 + u1], 1);
	atomicAdd(&r.come_cnt[v0 + u2], 1);
}

KERNEL k_revise_nv(int n_bundles, const int32_t *n_pex, int32_t *nv)
{
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if(b < n_bundles) nv[b] = n_pex[b] + 2;
}

// is k a distant in-vertex of a vertex in (k, x] (left) -- i.e. has left_continuous_extend put k into its set by the time
// it gets there
DEV bool revise_blocked_left(const sgraph &g, int k, int x)
{
	for(int i = g.out_off[k]; i < g.out_off[k + 1]; i++)
	{
		int t = g.out_dst[i];
		if(t > x) break;
		if(g.gv.v_l[t] != g.gv.v_r[k]) return true;
	}
	return false;
}
DEV bool revise_blocked_right(const sgraph &g, int k, int x)
{
	for(int i = g.in_off[k + 1] - 1; i >= g.in_off[k]; i--)
	{
		int s = g.in_src[i];
		if(s < x) break;
		if(g.gv.v_r[s] != g.gv.v_l[k]) return true;
	}
	return false;
}

// block-wide arg-max of (ratio, x) over the open candidates; every thread returns the winner (-1: none)
DEV int revise_best(int lo, int hi, const double *ratio, const int32_t *open, double *s_r, int *s_x, double *best_r)
{
	double br = -1;
	int bx = -1;
	for(int x = lo + threadIdx.x; x < hi; x += blockDim.x)
		if(open[x] && ratio[x] >= br) { br = ratio[x]; bx = x; }
	s_r[threadIdx.x] = br; s_x[threadIdx.x] = bx;
	BLOCK_SYNC();
	if(threadIdx.x == 0)
	{
		for(int t = 1; t < (int)blockDim.x; t++)
			if(s_x[t] >= 0 && (s_x[0] < 0 || s_r[t] > s_r[0] || (s_r[t] == s_r[0] && s_x[t] > s_x[0]))) { s_r[0] = s_r[t]; s_x[0] = s_x[t]; }
	}
	BLOCK_SYNC();
	bx = s_x[0];
	*best_r = s_r[0];
	BLOCK_SYNC();
	return bx;
}

// one CTA per bundle
KERNEL k_revise(int n_bundles, graph_dev gd, const uint8_t *b_strand, revise_dev r)
{
	SHARED double s_r[REVISE_BLOCK];
	SHARED int s_x[REVISE_BLOCK];
	for(int b = blockIdx.x; b < n_bundles; b += gridDim.x)
	{
		sgraph g = sgraph_of(gd, b_strand, b);
		const int nv = g.gv.nv, n = nv - 1;
		const int64_t v0 = r.voff[b];
		const double *v_w = gd.v_w + vert_base(gd, b);
		int32_t *run = r.run + v0, *open = r.open + v0, *has_s = r.has_s + v0, *has_e = r.has_e + v0;
		double *ratio = r.ratio + v0, *weight = r.weight + v0;
		int32_t *add_v = r.add_v + 2 * v0;
		double *add_w = r.add_w + 2 * v0;
		for(int x = threadIdx.x; x < nv; x += blockDim.x)
		{
			has_s[x] = (x >= 1 && x < n && real_edge(g, 0, x) >= 0) ? 1 : 0;
			has_e[x] = (x >= 1 && x < n && real_edge(g, x, n) >= 0) ? 1 : 0;
		}
		BLOCK_SYNC();
		// ---- start boundaries: identify_start_boundary / determine_start_boundary
		for(int x = 1 + threadIdx.x; x < n; x += blockDim.x)
		{
			int a = x;
			for(int k = x; ; k--)
			{
				if(k != x && revise_blocked_left(g, k, x)) break;
				a = k;
				if(k - 1 <= 0) break;
				if(real_edge(g, k - 1, k) < 0) break;
				if(g.gv.v_r[k - 1] != g.gv.v_l[k]) break;
			}
			double maxcov = 0, sum = 0;
			bool ok = true;
			for(int k = a; k <= x && ok; k++)
			{
				if(has_s[k]) { ok = false; break; }
				if(maxcov < v_w[k]) maxcov = v_w[k];
				for(int i = g.in_off[k]; i < g.in_off[k + 1]; i++)
				{
					int v = g.in_src[i];
					if(v >= a && v <= x) continue;
					sum += g.e_w[g.in_eid[i]];
				}
			}
			run[x] = a; open[x] = ok ? 1 : 0;
			ratio[x] = log(2 + maxcov) / log(2 + sum);
			weight[x] = maxcov - sum;
		}
		BLOCK_SYNC();
		int ns = 0;
		while(true)
		{
			double br;
			int bx = revise_best(1, n, ratio, open, s_r, s_x, &br);
			if(bx < 0 || br < r.min_ratio) break;
			const int a = run[bx];
			if(threadIdx.x == 0) { add_v[ns] = a; add_w[ns] = weight[bx]; has_s[a] = 1; }
			ns++;
			for(int x = a + threadIdx.x; x < n; x += blockDim.x)
				if(run[x] <= a) open[x] = 0;
			BLOCK_SYNC();
		}
		BLOCK_SYNC();
		// ---- end boundaries: identify_end_boundary / determine_end_boundary
		for(int x = 1 + threadIdx.x; x < n; x += blockDim.x)
		{
			int e = x;
			for(int k = x; ; k++)
			{
				if(k != x && revise_blocked_right(g, k, x)) break;
				e = k;
				if(k + 1 >= n) break;
				if(real_edge(g, k, k + 1) < 0) break;
				if(g.gv.v_l[k + 1] != g.gv.v_r[k]) break;
			}
			double maxcov = 0, sum = 0;
			bool ok = true;
			for(int k = x; k <= e && ok; k++)
			{
				if(has_e[k]) { ok = false; break; }
				if(maxcov < v_w[k]) maxcov = v_w[k];
				for(int i = g.out_off[k]; i < g.out_off[k + 1]; i++)
				{
					int v = g.out_dst[i];
					if(v >= x && v <= e) continue;
					sum += g.e_w[g.out_eid[i]];
				}
			}
			run[x] = e; open[x] = ok ? 1 : 0;
			ratio[x] = log(2 + maxcov) / log(2 + sum);
			weight[x] = maxcov - sum;
		}
		BLOCK_SYNC();
		int ne = 0;
		const int mid = nv - 2 > 0 ? nv - 2 : 0;      // ends are kept behind the (at most V - 2) starts
		while(true)
		{
			double br;
			int bx = revise_best(1, n, ratio, open, s_r, s_x, &br);
			if(bx < 0 || br < r.min_ratio) break;
			const int e = run[bx];
			if(threadIdx.x == 0) { add_v[mid + ne] = e; add_w[mid + ne] = weight[bx]; has_e[e] = 1; }
			ne++;
			for(int x = 1 + threadIdx.x; x <= e; x += blockDim.x)
				if(run[x] >= e) open[x] = 0;
			BLOCK_SYNC();
		}
		BLOCK_SYNC();
		if(threadIdx.x == 0) { r.n_start[b] = ns; r.n_end[b] = ne; }
		// ---- remove_false_boundaries: annotate where the boundary edge exists (rnacore/graph_reviser.cc:1322-1375)
		for(int x = threadIdx.x; x < nv; x += blockDim.x)
		{
			const double w = v_w[x];
			int c1 = r.leave_cnt[v0 + x], c2 = r.come_cnt[v0 + x];
			bool k1 = c1 > 0 && has_e[x], k2 = c2 > 0 && has_s[x];
			r.leave_ratio[v0 + x] = k1 ? log(1 + c1 + w) - log(1 + w) : 0.0;
			r.come_ratio[v0 + x] = k2 ? log(1 + c2 + w) - log(1 + w) : 0.0;
			if(!k1) r.leave_cnt[v0 + x] = 0;
			if(!k2) r.come_cnt[v0 + x] = 0;
		}
		BLOCK_SYNC();
	}
}

} // namespace agpu

#endif
