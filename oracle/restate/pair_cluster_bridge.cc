// TEST INFRASTRUCTURE ONLY (oracle/): restatement of mate pairing, paired-read clustering, the bridge
// solver and update_bridges.
#include "restate.h"

#include <algorithm>
#include <cassert>
#include <cstdio>

namespace orc {

namespace {

// check_increasing_sequence (util/util.h:177-185): equal neighbours pass
bool increasing(const chain_t &x)
{
	for(size_t k = 0; k + 1 < x.size(); k++) if(x[k] > x[k + 1]) return false;
	return true;
}

// check_continuous_vertices (rnacore/essential.cc:436-446)
bool continuous(const graph &gr, int x, int y)
{
	for(int i = x; i < y; i++)
	{
		if(gr.find_edge(i, i + 1) < 0) return false;
		if(gr.vr[i] != gr.vl[i + 1]) return false;
	}
	return true;
}

// build_path_from_intron_coordinates (rnacore/essential.cc:368-403)
bool path_from_introns(const graph &gr, const chain_t &v, std::vector<int> &vv)
{
	vv.clear();
	if(v.empty()) return true;
	int n = (int)v.size() / 2;
	std::vector<std::pair<int, int> > pp(n);
	for(int k = 0; k < n; k++)
	{
		int32_t p = v[2 * k], q = v[2 * k + 1];
		if(p >= q) return false;
		std::map<int32_t, int>::const_iterator a = gr.rindex.find(p), b = gr.lindex.find(q);
		if(a == gr.rindex.end() || b == gr.lindex.end()) return false;
		pp[k] = std::make_pair(a->second, b->second);
	}
	vv.push_back(pp.front().first);
	for(int k = 0; k < n - 1; k++)
	{
		int a = pp[k].second, b = pp[k + 1].first;
		if(a > b) return false;
		if(!continuous(gr, a, b)) return false;
		for(int j = a; j <= b; j++) vv.push_back(j);
	}
	vv.push_back(pp.back().second);
	return true;
}

// align_hit_to_splice_graph (rnacore/essential.cc:461-472) -> build_path_from_mixed_coordinates (:405-434)
bool align_hit(const hit &h, const chain_t &chain, const graph &gr, std::vector<int> &vv)
{
	vv.clear();
	chain_t u;
	u.push_back(h.pos);
	u.insert(u.end(), chain.begin(), chain.end());
	u.push_back(h.rpos);
	if(!increasing(u)) return false;
	int u1 = gr.locate_vertex(u.front()), u2 = gr.locate_vertex(u.back() - 1);
	if(u1 < 0 || u2 < 0) return false;
	if(u1 > u2) return false;
	if(u.size() == 2)
	{
		for(int k = u1; k <= u2; k++) vv.push_back(k);
		return true;
	}
	std::vector<int> uu;
	if(!path_from_introns(gr, chain, uu)) return false;
	for(int i = u1; i < uu.front(); i++) vv.push_back(i);
	vv.insert(vv.end(), uu.begin(), uu.end());
	for(int i = uu.back() + 1; i <= u2; i++) vv.push_back(i);
	return true;
}

typedef std::array<int32_t, 5> frow;
template<int R> bool cmp_rank(const frow &x, const frow &y) { return x[R] < y[R]; }

// graph_cluster::partition (rnacore/graph_cluster.cc:170-203)
std::vector<std::vector<int> > partition(std::vector<frow> &fs, int r, int gap)
{
	std::vector<std::vector<int> > vv;
	if(fs.empty()) return vv;
	if(r >= 4)
	{
		std::vector<int> v;
		for(size_t k = 0; k < fs.size(); k++) v.push_back(fs[k][4]);
		vv.push_back(v);
		return vv;
	}
	if(r == 0) std::sort(fs.begin(), fs.end(), cmp_rank<0>);
	if(r == 1) std::sort(fs.begin(), fs.end(), cmp_rank<1>);
	if(r == 2) std::sort(fs.begin(), fs.end(), cmp_rank<2>);
	if(r == 3) std::sort(fs.begin(), fs.end(), cmp_rank<3>);
	size_t pre = 0;
	for(size_t k = 1; k <= fs.size(); k++)
	{
		if(k < fs.size() && fs[k][r] - fs[k - 1][r] <= gap) continue;
		std::vector<frow> fs1(fs.begin() + pre, fs.begin() + k);
		std::vector<std::vector<int> > vv1 = partition(fs1, r + 1, gap);
		vv.insert(vv.end(), vv1.begin(), vv1.end());
		pre = k;
	}
	return vv;
}

} // namespace

// bundle_base::build_fragments (rnacore/bundle_base.cc:267-323).  The reference's bucket index only narrows the
// search; the pairing it finds is: for i ascending, the first unpaired u != i with pos[u] == mpos[i],
// isize[u] + isize[i] == 0 and the same query name.
void build_fragments(bundle &bd)
{
	bd.frgs.clear();
	int n = (int)bd.hits.size();
	if(n == 0) return;
	std::vector<bool> paired(n, false);
	std::map<uint64_t, std::vector<int> > byname;
	for(int i = 0; i < n; i++) byname[bd.hits[i].qid].push_back(i);
	for(int i = 0; i < n; i++)
	{
		if(paired[i]) continue;
		const hit &h = bd.hits[i];
		const std::vector<int> &cand = byname[h.qid];
		int x = -1;
		for(size_t j = 0; j < cand.size(); j++)
		{
			int u = cand[j];
			if(u == i || paired[u]) continue;
			if(bd.hits[u].pos != h.mpos) continue;
			if(bd.hits[u].isize + h.isize != 0) continue;
			x = u;
			break;
		}
		if(x < 0) continue;
		bd.frgs.push_back(AI3({i, x, 0}));
		paired[i] = true;
		paired[x] = true;
	}
}

// graph_cluster::group_pereads + build_pereads_clusters (rnacore/graph_cluster.cc:28-168)
void cluster_fragments(graph &gr, bundle &bd, std::vector<cluster> &vc)
{
	typedef std::pair<std::vector<int>, std::vector<int> > PVV;
	std::map<PVV, int> findex;
	std::vector<std::vector<int> > groups;
	std::vector<int32_t> extend;
	static const chain_t empty;
	for(size_t i = 0; i < bd.frgs.size(); i++)
	{
		if(bd.frgs[i][2] != 0) continue;
		bd.frgs[i][2] = -1;
		int h1 = bd.frgs[i][0], h2 = bd.frgs[i][1];
		if(bd.hits[h1].pos > bd.hits[h2].pos) continue;
		if(bd.hits[h1].rpos > bd.hits[h2].rpos) continue;
		const chain_t *c1 = bd.hcst.get(h1), *c2 = bd.hcst.get(h2);
		std::vector<int> v1, v2;
		bool b1 = align_hit(bd.hits[h1], c1 ? *c1 : empty, gr, v1);
		bool b2 = align_hit(bd.hits[h2], c2 ? *c2 : empty, gr, v2);
		if(!b1 || !b2) continue;
		if(v1.empty() || v2.empty()) continue;
		bd.frgs[i][2] = 0;
		PVV pvv(v1, v2);
		std::map<PVV, int>::iterator it = findex.find(pvv);
		if(it == findex.end())
		{
			findex[pvv] = (int)groups.size();
			extend.push_back(gr.vl[v1.front()]); extend.push_back(gr.vr[v1.back()]);
			extend.push_back(gr.vl[v2.front()]); extend.push_back(gr.vr[v2.back()]);
			groups.push_back(std::vector<int>(1, (int)i));
		}
		else groups[it->second].push_back((int)i);
	}
	for(size_t g = 0; g < groups.size(); g++)
	{
		const std::vector<int> &fs = groups[g];
		std::vector<frow> rows;
		for(size_t i = 0; i < fs.size(); i++)
		{
			const hit &a = bd.hits[bd.frgs[fs[i]][0]], &b = bd.hits[bd.frgs[fs[i]][1]];
			rows.push_back(frow({a.pos, a.rpos, b.pos, b.rpos, (int32_t)i}));
		}
		std::vector<std::vector<int> > zz = partition(rows, 0, bd.prm.max_reads_partition_gap);
		for(size_t z = 0; z < zz.size(); z++)
		{
			if(zz[z].empty()) continue;
			int h1 = bd.frgs[fs[zz[z][0]]][0], h2 = bd.frgs[fs[zz[z][0]]][1];
			cluster pc;
			const chain_t *c1 = bd.hcst.get(h1), *c2 = bd.hcst.get(h2);
			if(c1) pc.chain1 = *c1;
			if(c2) pc.chain2 = *c2;
			int32_t base[4] = {bd.hits[h1].pos, bd.hits[h1].rpos, bd.hits[h2].pos, bd.hits[h2].rpos};
			pc.bounds.assign(4, 0);
			pc.count = 0;
			for(size_t k = 0; k < zz[z].size(); k++)
			{
				const hit &a = bd.hits[bd.frgs[fs[zz[z][k]]][0]], &b = bd.hits[bd.frgs[fs[zz[z][k]]][1]];
				pc.bounds[0] += a.pos - base[0]; pc.bounds[1] += a.rpos - base[1];
				pc.bounds[2] += b.pos - base[2]; pc.bounds[3] += b.rpos - base[3];
				pc.frlist.push_back(fs[zz[z][k]]);
				pc.count++;
			}
			for(int k = 0; k < 4; k++) pc.bounds[k] = pc.bounds[k] / pc.count + base[k];
			pc.extend.assign(extend.begin() + 4 * g, extend.begin() + 4 * g + 4);
			vc.push_back(pc);
		}
	}
}

namespace {

struct entry { std::vector<int> stack; int32_t length; int trace1, trace2; };

// entry_compare (bridge/bridge_solver.cc:21-30)
bool entry_compare(const entry &x, const entry &y)
{
	for(size_t i = 0; i < x.stack.size() && i < y.stack.size(); i++)
	{
		if(x.stack[i] > y.stack[i]) return true;
		if(x.stack[i] < y.stack[i]) return false;
	}
	return x.length < y.length;
}

// update_stack (:532-546)
std::vector<int> update_stack(const std::vector<int> &v, int s)
{
	std::vector<int> stack(v.size(), 0);
	for(size_t i = 0, j = 0; i < v.size() && j < v.size(); i++, j++)
	{
		if(i == j && v[i] > s)
		{
			stack[j] = s;
			j++;
			if(j >= stack.size()) break;
		}
		stack[j] = v[i];
	}
	return stack;
}

// compare_bridge_path_vertices / _stack (bridge/bridge_path.cc:58-84)
bool cmp_vertices(const bridge_path &a, const bridge_path &b)
{
	for(size_t k = 0; k < a.v.size() && k < b.v.size(); k++)
	{
		if(a.v[k] < b.v[k]) return true;
		if(a.v[k] > b.v[k]) return false;
	}
	return a.v.size() < b.v.size();
}
bool cmp_stack(const bridge_path &a, const bridge_path &b)
{
	for(size_t k = 0; k < a.stack.size() && k < b.stack.size(); k++)
	{
		if(a.stack[k] > b.stack[k]) return true;
		if(a.stack[k] < b.stack[k]) return false;
	}
	return a.stack.size() > b.stack.size();
}

enum { IDENTICAL = 0, FALL_RIGHT, FALL_LEFT, CONTAINED, CONTAINING, EXTEND_RIGHT, EXTEND_LEFT, NESTED, NESTING, CONFLICTING };

bool identical(const chain_t &x, int x1, int x2, const chain_t &y, int y1, int y2)
{
	if(x[x1] != y[y1] || x[x2] != y[y2] || x2 - x1 != y2 - y1) return false;
	for(int kx = x1, ky = y1; kx <= x2 && ky <= y2; kx++, ky++) if(x[kx] != y[ky]) return false;
	return true;
}

// compare_two_sorted_sequences (util/util.h:191-253)
int compare_sorted(const chain_t &ref, const chain_t &qry)
{
	if(ref.back() < qry.front()) return FALL_RIGHT;
	if(ref.front() > qry.back()) return FALL_LEFT;
	int nr = (int)ref.size(), nq = (int)qry.size();
	int kr1 = (int)(std::lower_bound(ref.begin(), ref.end(), qry.front()) - ref.begin());
	int kq1 = (int)(std::lower_bound(qry.begin(), qry.end(), ref.front()) - qry.begin());
	int kq2 = (int)(std::lower_bound(qry.begin(), qry.end(), ref.back()) - qry.begin());
	int kr2 = (int)(std::lower_bound(ref.begin(), ref.end(), qry.back()) - ref.begin());
	if(kr1 >= nr || kq1 >= nq) return CONFLICTING;
	bool r2end = kr2 >= nr, q2end = kq2 >= nq;
	if(qry[kq1] == ref.front() || ref[kr1] == qry.front())
	{
		if(!r2end && !q2end)
		{
			if(!identical(ref, kr1, kr2, qry, kq1, kq2)) return CONFLICTING;
			if(kr1 == 0 && kq1 == 0) return IDENTICAL;
			if(kr1 >= 1 && kq1 == 0) return CONTAINED;
			if(kr1 == 0 && kq1 >= 1) return CONTAINING;
			return CONFLICTING;
		}
		else if(!r2end && q2end)
		{
			if(!identical(ref, kr1, kr2, qry, kq1, nq - 1)) return CONFLICTING;
			return kq1 == 0 ? CONTAINED : EXTEND_LEFT;
		}
		else if(r2end && !q2end)
		{
			if(!identical(ref, kr1, nr - 1, qry, kq1, kq2)) return CONFLICTING;
			return kr1 == 0 ? CONTAINING : EXTEND_RIGHT;
		}
	}
	else if(ref[kr1] > qry.front() && kr2 == kr1 && ref[kr2] > qry.back()) return NESTED;
	else if(qry[kq1] > ref.front() && kq2 == kq1 && qry[kq2] > ref.back()) return NESTING;
	return CONFLICTING;
}

// merge_intron_chains (rnacore/essential.cc:474-483) over merge_two_sorted_sequences (util/util.h:255-299)
bool merge_intron_chains(const chain_t &x, const chain_t &y, chain_t &xy)
{
	xy.clear();
	if(!x.empty() && !y.empty() && x.front() > y.front()) return false;
	if(x.empty()) xy = y;
	else if(y.empty()) xy = x;
	else
	{
		int t = compare_sorted(x, y);
		if(t == CONFLICTING || t == NESTED || t == NESTING) return false;
		if(t == IDENTICAL || t == CONTAINED) xy = x;
		if(t == CONTAINING) xy = y;
		if(t == FALL_RIGHT) { xy = x; xy.insert(xy.end(), y.begin(), y.end()); }
		if(t == FALL_LEFT) { xy = y; xy.insert(xy.end(), x.begin(), x.end()); }
		if(t == EXTEND_LEFT)
		{
			chain_t::const_iterator q1 = std::lower_bound(y.begin(), y.end(), x.front());
			xy.insert(xy.end(), y.begin(), q1);
			xy.insert(xy.end(), x.begin(), x.end());
		}
		if(t == EXTEND_RIGHT)
		{
			chain_t::const_iterator q2 = std::lower_bound(y.begin(), y.end(), x.back());
			xy.insert(xy.end(), x.begin(), x.end());
			xy.insert(xy.end(), q2 + 1, y.end());
		}
	}
	int d = (int)(x.size() + y.size()) - (int)xy.size();
	return d % 2 == 0;
}

// check_strand_from_intron_coordinates (rnacore/essential.cc:164-200)
int check_strand(const graph &gr, const chain_t &v)
{
	if(v.empty()) return 0;
	bool b1 = false, b2 = false;
	for(size_t k = 0; k < v.size() / 2; k++)
	{
		int32_t p = v[2 * k], q = v[2 * k + 1];
		if(p >= q) return -1;
		std::map<int32_t, int>::const_iterator a = gr.rindex.find(p), b = gr.lindex.find(q);
		if(a == gr.rindex.end() || b == gr.lindex.end()) return -1;
		int e = gr.find_edge(a->second, b->second);
		if(e < 0) return -1;
		if(gr.edges[e].strand == 1) b1 = true;
		if(gr.edges[e].strand == 2) b2 = true;
	}
	if(b1 && b2) return -1;
	if(b1) return 1;
	if(b2) return 2;
	return 0;
}

} // namespace

// bridge_solver::bridge_solver (bridge/bridge_solver.cc:32-46)
void bridge_clusters(graph &gr, std::vector<cluster> &vc, const orc_params &prm, std::vector<bridge_path> &opt)
{
	int nv = gr.nv();
	// add_adjacent_edges + build_pseudo_introns (:71-108)
	std::vector<int> adj;
	std::set<std::pair<int32_t, int32_t> > pseudos;
	for(int i = 1; i < nv - 2; i++)
	{
		if(gr.find_edge(i, i + 1) >= 0) continue;
		adj.push_back(gr.add_edge(i, i + 1, 0.5, 0));
	}
	for(size_t i = 0; i < adj.size(); i++)
	{
		int32_t p1 = gr.vr[gr.edges[adj[i]].s], p2 = gr.vl[gr.edges[adj[i]].t];
		if(p1 < p2) pseudos.insert(std::make_pair(p1, p2));
	}

	// build_bridging_vertices (:53-69) with check_left/right_relaxing (:124-148)
	int n = nv - 1;
	std::vector<std::pair<int, int> > vpairs;
	for(size_t i = 0; i < vc.size(); i++)
	{
		const cluster &pc = vc[i];
		int v1 = gr.locate_vertex(pc.bounds[1] - 1), v2 = gr.locate_vertex(pc.bounds[2]);
		{
			int v = v1;
			bool ok = !(v <= 0 || v >= n) && !(v <= 1);
			if(ok && !continuous(gr, v - 1, v)) ok = false;
			if(ok && pc.bounds[1] - gr.vl[v] > prm.bridge_end_relaxing) ok = false;
			if(ok && !pc.chain1.empty() && pc.chain1.back() >= gr.vl[v]) ok = false;
			if(ok) v1--;
		}
		{
			int v = v2;
			bool ok = !(v <= 0 || v >= n) && !(v >= n - 1);
			if(ok && !continuous(gr, v, v + 1)) ok = false;
			if(ok && gr.vr[v] - pc.bounds[2] > prm.bridge_end_relaxing) ok = false;
			if(ok && !pc.chain2.empty() && pc.chain2.front() <= gr.vr[v]) ok = false;
			if(ok) v2++;
		}
		vpairs.push_back(std::make_pair(v1, v2));
	}

	// build_piers (:150-167) + build_bounds (:205-222)
	struct pier { int bs, bt; std::vector<bridge_path> bridges; };
	std::vector<pier> piers;
	{
		std::set<std::pair<int, int> > ss;
		for(size_t k = 0; k < vc.size(); k++)
		{
			std::pair<int, int> p = vpairs[k];
			if(p.first < 0 || p.second < 0 || p.first >= p.second) continue;
			if(!ss.insert(p).second) continue;
			pier pr;
			pr.bs = p.first; pr.bt = p.second;
			piers.push_back(pr);
		}
	}
	std::sort(piers.begin(), piers.end(), [](const pier &a, const pier &b) { return a.bs < b.bs || (a.bs == b.bs && a.bt < b.bt); });
	std::vector<int> bounds;
	if(!piers.empty())
	{
		bounds.push_back(0);
		for(size_t i = 1; i < piers.size(); i++)
			if(piers[i].bs != piers[i - 1].bs) { bounds.push_back((int)i - 1); bounds.push_back((int)i); }
		bounds.push_back((int)piers.size() - 1);
	}

	// nominate (:180-257) with dynamic_programming (:484-530) and trace_back (:548-568)
	std::vector<int> passes;
	if(gr.strand == '.') { passes.push_back(1); passes.push_back(2); }
	else if(gr.strand == '+') passes.push_back(1);
	else if(gr.strand == '-') passes.push_back(2);
	for(size_t ps = 0; ps < passes.size(); ps++)
	{
		int strand = passes[ps];
		for(size_t k = 0; k < bounds.size() / 2; k++)
		{
			int b1 = bounds[2 * k], b2 = bounds[2 * k + 1];
			int k1 = piers[b2].bs, k2 = piers[b2].bt;
			std::vector<std::vector<entry> > table(nv);
			table[k1].resize(1);
			table[k1][0].stack.assign(prm.bridge_dp_stack_size, 999999);
			table[k1][0].length = gr.vr[k1] - gr.vl[k1];
			table[k1][0].trace1 = table[k1][0].trace2 = -1;
			for(int kk = k1 + 1; kk <= k2; kk++)
			{
				std::vector<entry> v;
				int32_t len = gr.vr[kk] - gr.vl[kk];
				for(std::set<std::pair<int, int> >::const_iterator it = gr.in[kk].begin(); it != gr.in[kk].end(); ++it)
				{
					const edge &e = gr.edges[it->second];
					if(e.strand != 0 && e.strand != strand) continue;
					int j = e.s;
					int w = (int)e.w;
					if(j < k1) continue;
					for(size_t i = 0; i < table[j].size(); i++)
					{
						entry ne;
						ne.stack = update_stack(table[j][i].stack, w);
						ne.length = table[j][i].length + len;
						ne.trace1 = j;
						ne.trace2 = (int)i;
						v.push_back(ne);
					}
				}
				std::sort(v.begin(), v.end(), entry_compare);
				if((int)v.size() > prm.bridge_dp_solution_size) v.resize(prm.bridge_dp_solution_size);
				table[kk] = v;
			}
			for(int b = b1; b <= b2; b++)
			{
				int bt = piers[b].bt;
				for(size_t j = 0; j < table[bt].size(); j++)
				{
					bridge_path p;
					p.score = table[bt][j].stack.front();
					p.stack = table[bt][j].stack;
					int pp = bt, qq = (int)j;
					while(true)
					{
						p.v.push_back(pp);
						const entry &e = table[pp][qq];
						pp = e.trace1; qq = e.trace2;
						if(pp < 0) break;
					}
					std::reverse(p.v.begin(), p.v.end());
					// build_intron_coordinates_from_path (rnacore/essential.cc:148-162) + filter_pseudo_introns (:110-122)
					for(size_t i = 0; i + 1 < p.v.size(); i++)
					{
						int32_t a = gr.vr[p.v[i]], c = gr.vl[p.v[i + 1]];
						if(a == c) continue;
						if(pseudos.count(std::make_pair(a, c))) continue;
						p.chain.push_back(a); p.chain.push_back(c);
					}
					piers[b].bridges.push_back(p);
				}
			}
		}
	}
	// refine_pier (:259-274)
	for(size_t i = 0; i < piers.size(); i++)
	{
		std::vector<bridge_path> &br = piers[i].bridges;
		if(br.empty()) continue;
		std::sort(br.begin(), br.end(), cmp_vertices);
		std::vector<bridge_path> v(1, br[0]);
		for(size_t k = 1; k < br.size(); k++) if(br[k].v != br[k - 1].v) v.push_back(br[k]);
		br = v;
		std::sort(br.begin(), br.end(), cmp_stack);
	}

	// vote (:276-385)
	std::map<std::pair<int, int>, int> pindex;
	for(size_t k = 0; k < piers.size(); k++) pindex[std::make_pair(piers[k].bs, piers[k].bt)] = (int)k;
	opt.assign(vc.size(), bridge_path());
	for(size_t r = 0; r < vc.size(); r++)
	{
		bridge_path &bbp = opt[r];
		bbp.type = -1;
		int ss = vpairs[r].first, tt = vpairs[r].second;
		if(ss < 0 || tt < 0) continue;
		const cluster &pc = vc[r];
		int type = 0;
		std::vector<chain_t> chains, wholes;
		std::vector<int> scores, strands;
		if(ss >= tt)
		{
			chain_t w;
			if(!merge_intron_chains(pc.chain1, pc.chain2, w)) continue;
			if(!increasing(w)) continue;
			int s = check_strand(gr, w);
			if(s < 0) continue;
			type = 1;
			chains.push_back(chain_t()); wholes.push_back(w); scores.push_back(10); strands.push_back(s);
		}
		else if(pindex.count(std::make_pair(ss, tt)))
		{
			type = 2;
			std::vector<bridge_path> &pb = piers[pindex[std::make_pair(ss, tt)]].bridges;
			for(size_t e = 0; e < pb.size(); e++)
			{
				chain_t w = pc.chain1;
				w.insert(w.end(), pb[e].chain.begin(), pb[e].chain.end());
				w.insert(w.end(), pc.chain2.begin(), pc.chain2.end());
				if(!increasing(w)) continue;
				int s = check_strand(gr, w);
				if(s < 0) continue;
				wholes.push_back(w); chains.push_back(pb[e].chain); scores.push_back((int)pb[e].score); strands.push_back(s);
			}
		}
		int be = -1, choices = 0;
		for(size_t e = 0; e < chains.size(); e++)
		{
			if(!wholes[e].empty() && wholes[e].front() <= pc.bounds[0]) continue;
			if(!wholes[e].empty() && wholes[e].back() >= pc.bounds[3]) continue;
			int32_t intron = 0;
			for(size_t k = 0; k < wholes[e].size() / 2; k++) intron += wholes[e][2 * k + 1] - wholes[e][2 * k];
			int32_t length = pc.bounds[3] - pc.bounds[0] - intron;
			if(length < prm.insertsize_low || length > prm.insertsize_high) continue;
			if(be < 0) be = (int)e;
			choices++;
		}
		if(be < 0) continue;
		bbp.type = type; bbp.score = scores[be]; bbp.chain = chains[be]; bbp.whole = wholes[be]; bbp.strand = strands[be]; bbp.choices = choices;
	}
	// remove_adjacent_edges (:88-95)
	for(size_t i = 0; i < adj.size(); i++) gr.remove_edge(adj[i]);
}

// bundle_base::update_bridges (rnacore/bundle_base.cc:420-507)
int update_bridges(bundle &bd, const std::vector<int> &frlist, const chain_t &chain, int strand)
{
	int cnt = 0;
	for(size_t i = 0; i < frlist.size(); i++)
	{
		int k = frlist[i];
		const hit &h1 = bd.hits[bd.frgs[k][0]], &h2 = bd.hits[bd.frgs[k][1]];
		chain_t v1;
		v1.push_back(h1.rpos);
		v1.insert(v1.end(), chain.begin(), chain.end());
		v1.push_back(h2.pos);
		if(h1.rpos < h2.pos && !increasing(v1)) continue;
		cnt++;
		if(chain.empty()) bd.frgs[k][2] = 1;
		else
		{
			char s = '.';
			if(h1.xs != '.') s = h1.xs;
			if(h2.xs != '.') s = h2.xs;
			if(h1.xs != '.' && h2.xs != '.' && h1.xs != h2.xs) s = '.';
			char ss = '.';
			if(strand == 1) ss = '+';
			if(strand == 2) ss = '-';
			bd.frgs[k][2] = 2;
			if(s == ss) bd.fcst.add(chain, k, ss);
			else if(s != '.' && ss == '.') bd.fcst.add(chain, k, s);
			else if(ss != '.' && s == '.') bd.fcst.add(chain, k, ss);
			else bd.fcst.add(chain, k, '.');
		}
		for(size_t j = 0; j < v1.size() / 2; j++)
		{
			if(v1[2 * j] >= v1[2 * j + 1]) continue;
			bd.mmap.add(v1[2 * j], v1[2 * j + 1], 1);
		}
	}
	return cnt;
}


// bundle_base::build_phase_set (rnacore/bundle_base.cc:338-418): phasing paths (exon coordinate lists) with multiplicities,
// keyed and ordered like phase_set::pmap (rnacore/phase_set.h:23-27)
void build_phase_set(const bundle &bd, const graph &gr, std::map<chain_t, int> &ps)
{
	static const chain_t none;
	std::vector<int> fb(bd.hits.size(), -1);
	for(size_t i = 0; i < bd.frgs.size(); i++)
	{
		if(bd.frgs[i][2] <= -1) continue;
		int h1 = bd.frgs[i][0], h2 = bd.frgs[i][1];
		if(bd.frgs[i][2] == 0) { fb[h1] = 0; fb[h2] = 0; continue; }
		int u1 = gr.locate_vertex(bd.hits[h1].pos), u2 = gr.locate_vertex(bd.hits[h2].rpos - 1);
		if(u1 < 0 || u2 < 0) continue;
		int32_t p1 = gr.vl[u1], p2 = gr.vr[u2];
		const chain_t *c1 = bd.hcst.get(h1), *c2 = bd.hcst.get(h2);
		const chain_t &v1 = c1 ? *c1 : none, &v2 = c2 ? *c2 : none;
		chain_t xy;
		if(bd.frgs[i][2] == 1)
		{
			if(!merge_intron_chains(v1, v2, xy)) continue;
		}
		if(bd.frgs[i][2] >= 2)
		{
			const chain_t *cv = bd.fcst.get((int)i);
			xy.insert(xy.end(), v1.begin(), v1.end());
			if(cv) xy.insert(xy.end(), cv->begin(), cv->end());
			xy.insert(xy.end(), v2.begin(), v2.end());
		}
		xy.insert(xy.begin(), p1);
		xy.insert(xy.end(), p2);
		if(!increasing(xy)) continue;
		fb[h1] = 1; fb[h2] = 1;
		ps[xy] += 1;
	}
	for(size_t i = 0; i < bd.hits.size(); i++)
	{
		if(fb[i] >= 0) continue;
		int u1 = gr.locate_vertex(bd.hits[i].pos), u2 = gr.locate_vertex(bd.hits[i].rpos - 1);
		if(u1 < 0 || u2 < 0) continue;
		const chain_t *c = bd.hcst.get((int)i);
		chain_t xy = c ? *c : none;
		xy.insert(xy.begin(), gr.vl[u1]);
		xy.insert(xy.end(), gr.vr[u2]);
		if(!increasing(xy)) continue;
		ps[xy] += 1;
	}
}

} // namespace orc
