// Cross-sample support features of assembler::assemble(vector<bundle*>) (meta/assembler.cc:177-373) for the clusters of bundles
// that bundle_group::resolve formed:
//   junction_support      meta/assembler.cc:375-417   every junction edge learns which samples of the cluster have that junction
//   non_splicing_support  :419-462                    vertices / adjacent edges of one sample's graph support adjacent edges of another
//   start_end_support     :678-779                    boundary edges support boundary edges (walking up to 200 bases inward)
//   boundary_extend       :781-880                    what a boundary of one graph would lose inside another (three position types)
// Member k of a cluster is handled right after its own round in the reference, and assemble(gr, ps, sid) then regroups the start /
// end boundaries of its graph (group_start_boundaries / group_end_boundaries, rnacore/graph_reviser.cc:916-1066) before the later
// members look at it.  The regrouping is a function of the member's own graph alone, so every member gets two versions of its
// graph up front -- A as transform(bd, gr, true) leaves it, B regrouped -- and round k reads B for the members before k and A for
// the others: all rounds of all clusters are then independent and run side by side (one thread per member / per cluster; the
// floating-point sums are taken in the reference's order).  What scallop does to a member's graph when it decomposes it in place
// is NOT part of this (see include/aletsch_gpu.h: agpu_batch_group_support).
#ifndef ALETSCH_B200_CSRC_K_SUPPORT_H
#define ALETSCH_B200_CSRC_K_SUPPORT_H

#include "k_revise.h"

namespace agpu {

// G graphs in one flat layout: the member bundles of all clusters in the caller's order, then one combined graph per cluster.
// Vertices of graph g: [voff[g], voff[g + 1]); edges in (source, target) order: [eoff[g], eoff[g] + ne[g]); per-vertex offset
// tables have nv + 2 entries per graph at voff[g] + 2 g.
struct sup_graphs
{
	int32_t n_graphs;
	const int64_t *voff, *eoff;
	int32_t *ne;
	int32_t *vl, *vr;
	double *vwA, *vwB;
	int32_t *out_off;          // out edges of vertex v: [out_off[v], out_off[v + 1]) (graph-local edge indices)
	int32_t *in_off;           // in edges of vertex v: in_eid[in_off[v + 1] .. in_off[v + 2]) ascending by source
	int32_t *in_eid;
	int32_t *es, *et;
	double *ewA, *ewB;
	uint8_t *aliveB;           // B: boundary edges the regrouping removed are 0
};

struct sgv                    // one graph, one version
{
	int nv, ne;
	const int32_t *vl, *vr, *out_off, *in_off, *in_eid, *es, *et;
	const double *vw, *ew;
	const uint8_t *alive;      // NULL: every edge
	HD bool live(int e) const { return alive == NULL || alive[e] != 0; }
};

HD sgv sup_view(const sup_graphs &s, int g, int version_b)
{
	sgv v;
	const int64_t v0 = s.voff[g], e0 = s.eoff[g], t0 = v0 + 2 * (int64_t)g;
	v.nv = (int)(s.voff[g + 1] - v0); v.ne = s.ne[g];
	v.vl = s.vl + v0; v.vr = s.vr + v0; v.out_off = s.out_off + t0; v.in_off = s.in_off + t0; v.in_eid = s.in_eid + e0;
	v.es = s.es + e0; v.et = s.et + e0;
	v.vw = (version_b ? s.vwB : s.vwA) + v0; v.ew = (version_b ? s.ewB : s.ewA) + e0;
	v.alive = version_b ? s.aliveB + e0 : NULL;
	return v;
}

HD int sup_find_edge(const sgv &g, int a, int b)
{
	if(a < 0 || b < 0 || a >= g.nv || b >= g.nv) return -1;
	for(int x = g.out_off[a]; x < g.out_off[a + 1]; x++) if(g.et[x] == b && g.live(x)) return x;
	return -1;
}

// splice_graph::locate_vertex (rnacore/splice_graph.cc:1166-1215)
HD int sup_locate(const sgv &g, int32_t p)
{
	int a = 1, b = g.nv - 1, m = -1;
	while(a < b)
	{
		int mid = (a + b) / 2;
		if(p >= g.vl[mid] && p < g.vr[mid]) { m = mid; break; }
		if(p < g.vl[mid]) b = mid; else a = mid + 1;
	}
	if(m < 0) m = b;
	if(m < 0 || m >= g.nv) return -1;
	if(p >= g.vl[m] && p < g.vr[m]) return m;
	return -1;
}

// the edge of g that is the junction (rpos, lpos): source vertex ends at rpos, target starts at lpos
HD int sup_find_junction(const sgv &g, int32_t rp, int32_t lp)
{
	const int n = g.nv - 1;
	int lo = 1, hi = n;                           // inner vertices 1 .. n - 1, vr ascending
	while(lo < hi) { int m = (lo + hi) >> 1; if(g.vr[m] < rp) lo = m + 1; else hi = m; }
	if(lo >= n || g.vr[lo] != rp) return -1;
	const int a = lo;
	lo = 1; hi = n;
	while(lo < hi) { int m = (lo + hi) >> 1; if(g.vl[m] < lp) lo = m + 1; else hi = m; }
	if(lo >= n || g.vl[lo] != lp) return -1;
	return sup_find_edge(g, a, lo);
}

// ---- S1: one thread per graph: vertices, edges of the revised graph in (s, t) order, in-lists
struct sup_src
{
	int32_t n_members, n_groups;
	const int32_t *members;          // [n_members] bundle of member m
	graph_dev gm;                    // the members' own graphs (agpu_batch_graph of the batch)
	graph_dev gc;                    // the combined graphs (batch of combined bundles)
	const int64_t *rv_voff;          // boundary revision of the members' graphs (k_revise.h)
	const int32_t *rv_nstart, *rv_nend, *rv_addv;
	const double *rv_addw;
};

// vertices and edge capacity of every graph (before the offsets exist)
KERNEL k_sup_sizes(sup_src s, int32_t *nv, int32_t *ecap)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if(g >= s.n_members + s.n_groups) return;
	if(g < s.n_members)
	{
		const int b = s.members[g];
		nv[g] = s.gm.n_pex[b] + 2;
		ecap[g] = s.gm.n_edge[b] + s.rv_nstart[b] + s.rv_nend[b];
	}
	else
	{
		const int c = g - s.n_members;
		nv[g] = s.gc.n_pex[c] + 2;
		ecap[g] = s.gc.n_edge[c];
	}
}

KERNEL k_sup_build(sup_src s, sup_graphs o)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if(g >= o.n_graphs) return;
	const bool member = g < s.n_members;
	const graph_dev &gd = member ? s.gm : s.gc;
	const int b = member ? s.members[g] : g - s.n_members;
	const int nv = gd.n_pex[b] + 2, n = nv - 1;
	const int64_t gv0 = vert_base(gd, b), go0 = voff_base(gd, b), ge0 = gd.edge_off[b];
	const int64_t v0 = o.voff[g], e0 = o.eoff[g], t0 = v0 + 2 * (int64_t)g;
	int ns = 0, nend = 0;
	const int32_t *addv = NULL;
	const double *addw = NULL;
	int mid = 0;
	if(member)
	{
		ns = s.rv_nstart[b]; nend = s.rv_nend[b];
		addv = s.rv_addv + 2 * s.rv_voff[b]; addw = s.rv_addw + 2 * s.rv_voff[b];
		mid = nv - 2 > 0 ? nv - 2 : 0;
	}
	for(int v = 0; v < nv; v++)
	{
		o.vl[v0 + v] = gd.v_l[gv0 + v]; o.vr[v0 + v] = gd.v_r[gv0 + v];
		o.vwA[v0 + v] = gd.v_w[gv0 + v]; o.vwB[v0 + v] = gd.v_w[gv0 + v];
	}
	int ne = 0;
	for(int v = 0; v < nv; v++)
	{
		o.out_off[t0 + v] = ne;
		const int x0 = gd.out_off[go0 + v], x1 = gd.out_off[go0 + v + 1];
		if(v == 0 && ns > 0)
		{
			// merge the built start edges (targets ascending) with the added ones (in the order they were added)
			int x = x0, last = -1;
			while(true)
			{
				int best = -1;                   // smallest added target above `last`
				for(int i = 0; i < ns; i++) if(addv[i] > last && (best < 0 || addv[i] < addv[best])) best = i;
				const int tb = x < x1 ? gd.out_dst[ge0 + x] : 0x7fffffff, ta = best >= 0 ? addv[best] : 0x7fffffff;
				if(tb == 0x7fffffff && ta == 0x7fffffff) break;
				if(tb <= ta) { o.es[e0 + ne] = 0; o.et[e0 + ne] = tb; o.ewA[e0 + ne] = gd.e_w[ge0 + gd.out_eid[ge0 + x]]; last = tb; x++; }
				else { o.es[e0 + ne] = 0; o.et[e0 + ne] = ta; o.ewA[e0 + ne] = addw[best]; last = ta; }
				ne++;
			}
			continue;
		}
		for(int x = x0; x < x1; x++) { o.es[e0 + ne] = v; o.et[e0 + ne] = gd.out_dst[ge0 + x]; o.ewA[e0 + ne] = gd.e_w[ge0 + gd.out_eid[ge0 + x]]; ne++; }
		// an added end edge v -> n is the vertex's largest target
		for(int i = 0; i < nend; i++) if(addv[mid + i] == v) { o.es[e0 + ne] = v; o.et[e0 + ne] = n; o.ewA[e0 + ne] = addw[mid + i]; ne++; }
	}
	o.out_off[t0 + nv] = ne;
	o.out_off[t0 + nv + 1] = ne;
	o.ne[g] = ne;
	for(int e = 0; e < ne; e++) { o.ewB[e0 + e] = o.ewA[e0 + e]; o.aliveB[e0 + e] = 1; }
	// in-lists: count, prefix, fill from the back (leaves the start of vertex t's list at in_off[t + 1])
	for(int v = 0; v < nv + 2; v++) o.in_off[t0 + v] = 0;
	for(int e = 0; e < ne; e++) o.in_off[t0 + o.et[e0 + e] + 1]++;
	for(int v = 0; v < nv + 1; v++) o.in_off[t0 + v + 1] += o.in_off[t0 + v];
	for(int e = ne - 1; e >= 0; e--) o.in_eid[e0 + --o.in_off[t0 + o.et[e0 + e] + 1]] = e;
	// now in_off[t + 1] = start of t, and in_off[nv + 1] must close the last list
	o.in_off[t0 + nv + 1] = ne;
}

// check_continuous_vertices (rnacore/essential.cc:436-446) on version B
DEV bool sup_continuous(const sgv &g, int x, int y)
{
	for(int i = x; i < y; i++)
	{
		if(sup_find_edge(g, i, i + 1) < 0) return false;
		if(g.vr[i] != g.vl[i + 1]) return false;
	}
	return true;
}

// ---- S2: one thread per member: version B = group_start_boundaries + group_end_boundaries (rnacore/graph_reviser.cc:916-1066)
// as far as they change the graph: boundaries of one continuous stretch within max_dist collapse onto the first (last) one
KERNEL k_sup_regroup(int32_t n_members, sup_graphs o, int32_t max_dist)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if(g >= n_members) return;
	const int64_t v0 = o.voff[g], e0 = o.eoff[g];
	sgv B = sup_view(o, g, 1);
	double *vw = o.vwB + v0, *ew = o.ewB + e0;
	uint8_t *alive = o.aliveB + e0;
	const int n = B.nv - 1;
	// start boundaries: out edges of 0, targets ascending
	{
		const int x0 = B.out_off[0], x1 = B.out_off[1];
		if(x1 - x0 > 1)
		{
			int32_t p2 = B.vl[B.et[x0]];
			int k1 = B.et[x0], k2 = B.et[x0];
			int ea = x0;
			double wa = ew[ea];
			for(int x = x0 + 1; x < x1; x++)
			{
				const int vi = B.et[x];
				const int32_t p = B.vl[vi];
				const int eb = x;
				const double wb = ew[eb];
				bool ok = sup_continuous(B, k2, vi);
				if(p - p2 > max_dist) ok = false;
				if(!ok) { p2 = p; k1 = vi; k2 = vi; ea = eb; wa = wb; continue; }
				for(int j = k1; j < vi; j++)
				{
					const int ec = sup_find_edge(B, j, j + 1);
					const double vc = vw[j], wc = ew[ec];
					vw[j] = vc + wb;
					ew[ec] = wc + wb;
				}
				wa += wb;
				ew[ea] = wa;
				alive[eb] = 0;
				k2 = vi;
				p2 = p;
			}
		}
	}
	// end boundaries: in edges of n, sources descending
	{
		const int i0 = B.in_off[n + 1], i1 = B.in_off[n + 2];
		if(i1 - i0 > 1)
		{
			int first = B.es[B.in_eid[i1 - 1]];
			int32_t p2 = B.vr[first];
			int k1 = first, k2 = first;
			int ea = B.in_eid[i1 - 1];
			double wa = ew[ea];
			for(int x = i1 - 2; x >= i0; x--)
			{
				const int eb = B.in_eid[x];
				const int vi = B.es[eb];
				const int32_t p = B.vr[vi];
				const double wb = ew[eb];
				bool ok = sup_continuous(B, vi, k2);
				if(p2 - p > max_dist) ok = false;
				if(!ok) { p2 = p; k1 = vi; k2 = vi; ea = eb; wa = wb; continue; }
				for(int j = vi; j < k1; j++)
				{
					const int ec = sup_find_edge(B, j, j + 1);
					const double wc = ew[ec];
					ew[ec] = wc + wb;
					vw[j + 1] = wc + wb;          // as the reference writes it: the edge's new weight, not the vertex's own plus wb
				}
				wa += wb;
				ew[ea] = wa;
				alive[eb] = 0;
				k2 = vi;
				p2 = p;
			}
		}
	}
}

// ---- the support state of one graph: dense per (edge, sample index of the cluster) tables
struct sup_state
{
	int S;                   // sample slots of the cluster: slot 0 = the combined graph (sample -1), then the distinct samples ascending
	uint8_t *has;            // [ne * S]
	double *val;             // [ne * S]
	double *abd;             // [ne]
	double *loss;            // [nv * 4]
};

DEV void sup_credit(sup_state &s, int e, int slot, double w)
{
	s.has[(int64_t)e * s.S + slot] = 1;
	s.val[(int64_t)e * s.S + slot] += w;
	s.abd[e] += w;
}

// assembler::start_end_support(sample, gr, gx): boundary edges of gr support boundary edges of gx
DEV void sup_start_end(int slot, const sgv &gr, const sgv &gx, sup_state &sx)
{
	const int nr = gr.nv - 1, nx = gx.nv - 1;
	for(int x = gr.out_off[0]; x < gr.out_off[1]; x++)
	{
		if(!gr.live(x)) continue;
		const int t = gr.et[x];
		const int32_t p = gr.vr[t];
		int k = sup_locate(gx, p - 1);
		if(k < 0) continue;
		int eb = sup_find_edge(gx, 0, k);
		bool cont = true;
		while(eb < 0)
		{
			k--;
			if(k == 0) { cont = false; break; }
			if(p - gx.vr[k] > 200) cont = false;
			if(gx.vl[k + 1] != gx.vr[k]) cont = false;
			if(sup_find_edge(gx, k, k + 1) < 0) cont = false;
			if(!cont) break;
			eb = sup_find_edge(gx, 0, k);
		}
		if(!cont) continue;
		sup_credit(sx, eb, slot, gr.ew[x]);
	}
	for(int y = gr.in_off[nr + 1]; y < gr.in_off[nr + 2]; y++)
	{
		const int x = gr.in_eid[y];
		if(!gr.live(x)) continue;
		const int s = gr.es[x];
		const int32_t p = gr.vl[s];
		int k = sup_locate(gx, p);
		if(k < 0) continue;
		int eb = sup_find_edge(gx, k, nx);
		bool cont = true;
		while(eb < 0)
		{
			k++;
			if(k == nx) { cont = false; break; }
			if(gx.vl[k] - p > 200) cont = false;
			if(gx.vr[k - 1] != gx.vl[k]) cont = false;
			if(sup_find_edge(gx, k - 1, k) < 0) cont = false;
			if(!cont) break;
			eb = sup_find_edge(gx, k, nx);
		}
		if(!cont) continue;
		sup_credit(sx, eb, slot, gr.ew[x]);
	}
}

// assembler::non_splicing_support(sample, gr, gx): vertices / edges of gr support the adjacent edges of gx
DEV void sup_non_splicing(int slot, const sgv &gr, const sgv &gx, sup_state &sx)
{
	const int n = gx.nv - 1;
	for(int e = 0; e < gx.ne; e++)
	{
		if(!gx.live(e)) continue;
		const int a = gx.es[e], b = gx.et[e];
		if(a == 0 || b == n) continue;
		if(gx.vr[a] != gx.vl[b]) continue;
		const int32_t p = gx.vl[b];
		const int k1 = sup_locate(gr, p - 1), k2 = sup_locate(gr, p);
		if(k1 < 0 || k2 < 0) continue;
		if(k1 == k2) sup_credit(sx, e, slot, gr.vw[k1]);
		else if(gr.vr[k1] == gr.vl[k2])
		{
			const int f = sup_find_edge(gr, k1, k2);
			if(f >= 0) sup_credit(sx, e, slot, gr.ew[f]);
		}
	}
}

DEV double sup_in_weights(const sgv &g, int k)
{
	double w = 0;
	for(int y = g.in_off[k + 1]; y < g.in_off[k + 2]; y++) if(g.live(g.in_eid[y])) w += g.ew[g.in_eid[y]];
	return w;
}
DEV double sup_out_weights(const sgv &g, int k)
{
	double w = 0;
	for(int x = g.out_off[k]; x < g.out_off[k + 1]; x++) if(g.live(x)) w += g.ew[x];
	return w;
}

// assembler::boundary_extend(sample, gr, gx, pos_type): what a boundary of gr would lose inside gx; slot 3 = the combined graph
DEV void sup_boundary_extend(int loss_slot, const sgv &gr, sup_state &sr, const sgv &gx, int pos_type)
{
	const int nr = gr.nv - 1, nx = gx.nv - 1;
	for(int x = gr.out_off[0]; x < gr.out_off[1]; x++)
	{
		if(!gr.live(x)) continue;
		const int t = gr.et[x];
		int k = -1;
		if(pos_type == 1) k = sup_locate(gx, gr.vl[t]);
		else if(pos_type == 2) k = sup_locate(gx, gr.vr[t] - 1);
		else if(pos_type == 3 && sup_find_edge(gr, t, t + 1) >= 0 && gr.vr[t] == gr.vl[t + 1] && t + 1 < nr) k = sup_locate(gx, gr.vr[t]);
		if(k <= 0 || sup_find_edge(gx, 0, k) >= 0) continue;
		double loss;
		const int f = sup_find_edge(gx, k - 1, k);
		if(f >= 0 && gx.vr[k - 1] == gx.vl[k]) loss = sup_in_weights(gx, k) - gx.ew[f];
		else loss = sup_in_weights(gx, k);
		sr.loss[4 * t + loss_slot] += loss;
	}
	for(int y = gr.in_off[nr + 1]; y < gr.in_off[nr + 2]; y++)
	{
		const int x = gr.in_eid[y];
		if(!gr.live(x)) continue;
		const int s = gr.es[x];
		int k = -1;
		if(pos_type == 1) k = sup_locate(gx, gr.vr[s] - 1);
		else if(pos_type == 2) k = sup_locate(gx, gr.vl[s]);
		else if(pos_type == 3 && s > 1 && sup_find_edge(gr, s - 1, s) >= 0 && gr.vr[s - 1] == gr.vl[s]) k = sup_locate(gx, gr.vl[s] - 1);
		if(k < 0 || k == nx || sup_find_edge(gx, k, nx) >= 0) continue;
		double loss;
		const int f = sup_find_edge(gx, k, k + 1);
		if(f >= 0 && gx.vr[k] == gx.vl[k + 1]) loss = sup_out_weights(gx, k) - gx.ew[f];
		else loss = sup_out_weights(gx, k);
		sr.loss[4 * s + loss_slot] += loss;
	}
}

// what the kernels need about the clusters
struct sup_clusters
{
	int32_t n_groups, n_members;
	const int32_t *member_off;     // [n_groups + 1] members of cluster c (the caller's order = the reference's gv)
	const int32_t *member_group;   // [n_members]
	const int32_t *slot;           // [n_members] sample slot of member m inside its cluster (1 ..)
	const int32_t *by_sample;      // [n_members] per cluster: its members ordered by (sample, position): the junction map keeps the
	                               // first entry of a sample (meta/assembler.cc:229-232)
	const int32_t *n_slots;        // [n_groups] S of the cluster
	const int64_t *tab_off;        // [n_graphs + 1] start of graph g's dense tables (units of ne * S entries)
};

DEV sup_state sup_state_of(const sup_graphs &o, const sup_clusters &c, int g, int S, uint8_t *has, double *val, double *abd, double *loss)
{
	sup_state s;
	s.S = S;
	s.has = has + c.tab_off[g]; s.val = val + c.tab_off[g];
	s.abd = abd + o.eoff[g];
	s.loss = loss + 4 * o.voff[g];
	return s;
}

// init (meta/assembler.cc:205-246) + junction_support (:375-417) of graph g (A version) against the cluster's junction map: the
// map holds, per junction, the samples that have it and the weight of the FIRST graph of every sample (combined graph first)
DEV void sup_init_junctions(const sup_graphs &o, const sup_clusters &c, int grp, int g, int own_slot, sup_state &s)
{
	const sgv G = sup_view(o, g, 0);
	const int n = G.nv - 1;
	const int m0 = c.member_off[grp], m1 = c.member_off[grp + 1];
	for(int e = 0; e < G.ne; e++)
	{
		const double w = G.ew[e];
		for(int z = 0; z < s.S; z++) { s.has[(int64_t)e * s.S + z] = 0; s.val[(int64_t)e * s.S + z] = 0; }
		s.has[(int64_t)e * s.S + own_slot] = 1;
		s.val[(int64_t)e * s.S + own_slot] = w;
		s.abd[e] = w;
	}
	for(int v = 0; v < 4 * G.nv; v++) s.loss[v] = 0;
	(void)n; (void)m0; (void)m1;
}

DEV void sup_junction_support(const sup_graphs &o, const sup_clusters &c, int grp, int g, sup_state &s)
{
	const sgv G = sup_view(o, g, 0);
	const int n = G.nv - 1;
	const int m0 = c.member_off[grp], m1 = c.member_off[grp + 1];
	const sgv X = sup_view(o, c.n_members + grp, 0);
	for(int e = 0; e < G.ne; e++)
	{
		const int a = G.es[e], b = G.et[e];
		if(a == 0 || b == n) continue;
		if(G.vr[a] == G.vl[b]) continue;
		const int32_t rp = G.vr[a], lp = G.vl[b];
		for(int z = 0; z < s.S; z++) { s.has[(int64_t)e * s.S + z] = 0; s.val[(int64_t)e * s.S + z] = 0; }
		// ascending sample: the combined graph (-1) first, then the members by (sample, position)
		const int fx = sup_find_junction(X, rp, lp);
		if(fx >= 0) { s.has[(int64_t)e * s.S] = 1; s.val[(int64_t)e * s.S] = X.ew[fx]; s.abd[e] += X.ew[fx]; }
		int done_slot = 0;
		for(int y = m0; y < m1; y++)
		{
			const int m = c.by_sample[y];
			const int sl = c.slot[m];
			if(sl == done_slot) continue;                 // this sample already has its (first) entry
			const sgv J = sup_view(o, m, 0);
			const int f = sup_find_junction(J, rp, lp);
			if(f < 0) continue;
			s.has[(int64_t)e * s.S + sl] = 1; s.val[(int64_t)e * s.S + sl] = J.ew[f]; s.abd[e] += J.ew[f];
			done_slot = sl;
		}
	}
}

// ---- S3: one thread per member: its own round (meta/assembler.cc:296-326)
KERNEL k_sup_member(sup_graphs o, sup_clusters c, uint8_t *has, double *val, double *abd, double *loss)
{
	int m = blockIdx.x * blockDim.x + threadIdx.x;
	if(m >= c.n_members) return;
	const int grp = c.member_group[m];
	const int m0 = c.member_off[grp], m1 = c.member_off[grp + 1];
	const int S = c.n_slots[grp];
	sup_state s = sup_state_of(o, c, m, S, has, val, abd, loss);
	sup_init_junctions(o, c, grp, m, c.slot[m], s);
	sup_junction_support(o, c, grp, m, s);
	const sgv K = sup_view(o, m, 0);
	for(int j = m0; j < m1; j++)
	{
		const sgv J = sup_view(o, j, j < m ? 1 : 0);       // the members before m have been assembled: their graphs are regrouped
		sup_start_end(c.slot[j], J, K, s);
		sup_non_splicing(c.slot[j], J, K, s);
		for(int t = 1; t <= 3; t++) sup_boundary_extend(t - 1, K, s, J, t);
	}
	const sgv X = sup_view(o, c.n_members + grp, 0);
	sup_boundary_extend(3, K, s, X, 1);
}

// ---- S4: one thread per cluster: the combined graph's support
KERNEL k_sup_combined(sup_graphs o, sup_clusters c, uint8_t *has, double *val, double *abd, double *loss)
{
	int grp = blockIdx.x * blockDim.x + threadIdx.x;
	if(grp >= c.n_groups) return;
	const int g = c.n_members + grp;
	const int S = c.n_slots[grp];
	sup_state s = sup_state_of(o, c, g, S, has, val, abd, loss);
	sup_init_junctions(o, c, grp, g, 0, s);
	const sgv X = sup_view(o, g, 0);
	for(int m = c.member_off[grp]; m < c.member_off[grp + 1]; m++)
	{
		const sgv K = sup_view(o, m, 0);
		sup_start_end(c.slot[m], K, X, s);
		sup_non_splicing(c.slot[m], K, X, s);
	}
	sup_junction_support(o, c, grp, g, s);
}

// ---- S5: dense tables -> sparse rows.  Pass 1: samples per edge; pass 2: (sample slot, weight) lists
KERNEL k_sup_count(int64_t n_edges_cap, sup_graphs o, sup_clusters c, const int32_t *edge_graph, const uint8_t *has, int32_t *count)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n_edges_cap) return;
	const int g = edge_graph[i];
	const int e = (int)(i - o.eoff[g]);
	if(g < 0 || e >= o.ne[g]) { count[i] = 0; return; }
	const int grp = g < c.n_members ? c.member_group[g] : g - c.n_members;
	const int S = c.n_slots[grp];
	const uint8_t *h = has + c.tab_off[g] + (int64_t)e * S;
	int n = 0;
	for(int z = 0; z < S; z++) n += h[z] ? 1 : 0;
	count[i] = n;
}

KERNEL k_sup_fill(int64_t n_edges_cap, sup_graphs o, sup_clusters c, const int32_t *edge_graph, const uint8_t *has, const double *val,
		const int64_t *row_off, int32_t *out_slot, double *out_val)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n_edges_cap) return;
	const int g = edge_graph[i];
	const int e = (int)(i - o.eoff[g]);
	if(g < 0 || e >= o.ne[g]) return;
	const int grp = g < c.n_members ? c.member_group[g] : g - c.n_members;
	const int S = c.n_slots[grp];
	const uint8_t *h = has + c.tab_off[g] + (int64_t)e * S;
	const double *v = val + c.tab_off[g] + (int64_t)e * S;
	int64_t w = row_off[i];
	for(int z = 0; z < S; z++) if(h[z]) { out_slot[w] = z; out_val[w] = v[z]; w++; }
}

// graph of every edge slot (edge arrays are laid out by capacity)
KERNEL k_sup_edge_graph(int32_t n_graphs, const int64_t *eoff, int32_t *edge_graph)
{
	for(int g = blockIdx.x; g < n_graphs; g += gridDim.x)
		for(int64_t e = eoff[g] + threadIdx.x; e < eoff[g + 1]; e += blockDim.x) edge_graph[e] = g;
}

} // namespace agpu

#endif
