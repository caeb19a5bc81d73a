// TEST INFRASTRUCTURE ONLY (oracle/).  Restatement of the boundary revision that assembler::transform(bd, gr, true) applies
// (meta/assembler.cc:930-944): identify_boundaries (rnacore/graph_reviser.cc:1068-1283) and remove_false_boundaries
// (:1285-1377), written the way the reference runs them -- every vertex is re-examined after every edge that is added.
#include "restate.h"

#include <cmath>

namespace orc {

// left_continuous_extend (rnacore/graph_reviser.cc:1210-1230)
static int extend_left(const graph &gr, int x)
{
	int z = -1;
	std::set<int> far;
	for(int k = x; k > 0; k--)
	{
		if(far.count(k)) break;
		z = k;
		for(auto &pr : gr.in[k]) if(gr.vr[pr.first] != gr.vl[k]) far.insert(pr.first);       // add_distant_in_vertices
		if(k - 1 <= 0) break;
		if(gr.find_edge(k - 1, k) < 0) break;
		if(gr.vr[k - 1] != gr.vl[k]) break;
	}
	return z;
}

// right_continuous_extend (rnacore/graph_reviser.cc:1232-1253)
static int extend_right(const graph &gr, int x)
{
	int z = -1;
	const int n = gr.nv() - 1;
	std::set<int> far;
	for(int k = x; k < n; k++)
	{
		if(far.count(k)) break;
		z = k;
		for(auto &pr : gr.out[k]) if(gr.vl[pr.first] != gr.vr[k]) far.insert(pr.first);      // add_distant_out_vertices
		if(k + 1 >= n) break;
		if(gr.find_edge(k, k + 1) < 0) break;
		if(gr.vl[k + 1] != gr.vr[k]) break;
	}
	return z;
}

// identify_start_boundary + determine_start_boundary (rnacore/graph_reviser.cc:1079-1115, 1155-1181)
static bool start_round(graph &gr, double min_ratio, std::vector<std::array<int, 2> > &added, std::vector<double> &added_w)
{
	int besta = -1;
	double bestr = 0, bestw = 0;
	for(int x = 1; x < gr.nv() - 1; x++)
	{
		int a = extend_left(gr, x);
		if(a < 0 || a > x) continue;
		double maxcov = 0, sum = 0;
		bool skip = false;
		for(int k = a; k <= x && !skip; k++)
		{
			if(gr.find_edge(0, k) >= 0) { skip = true; break; }
			if(maxcov < gr.vw[k]) maxcov = gr.vw[k];
			for(auto &pr : gr.in[k])
			{
				if(pr.first >= a && pr.first <= x) continue;
				sum += gr.edges[pr.second].w;
			}
		}
		if(skip) continue;
		double r = std::log(2 + maxcov) / std::log(2 + sum);
		if(r < bestr) continue;
		bestr = r; besta = a; bestw = maxcov - sum;
	}
	if(besta < 0 || bestr < min_ratio) return false;
	gr.add_edge(0, besta, bestw, 0);
	added.push_back({0, besta}); added_w.push_back(bestw);
	return true;
}

// identify_end_boundary + determine_end_boundary (rnacore/graph_reviser.cc:1117-1153, 1183-1208)
static bool end_round(graph &gr, double min_ratio, std::vector<std::array<int, 2> > &added, std::vector<double> &added_w)
{
	const int n = gr.nv() - 1;
	int bestb = -1;
	double bestr = 0, bestw = 0;
	for(int x = 1; x < n; x++)
	{
		int b = extend_right(gr, x);
		if(b < 0 || x > b) continue;
		double maxcov = 0, sum = 0;
		bool skip = false;
		for(int k = x; k <= b && !skip; k++)
		{
			if(gr.find_edge(k, n) >= 0) { skip = true; break; }
			if(maxcov < gr.vw[k]) maxcov = gr.vw[k];
			for(auto &pr : gr.out[k])
			{
				if(pr.first >= x && pr.first <= b) continue;
				sum += gr.edges[pr.second].w;
			}
		}
		if(skip) continue;
		double r = std::log(2 + maxcov) / std::log(2 + sum);
		if(r < bestr) continue;
		bestr = r; bestb = b; bestw = maxcov - sum;
	}
	if(bestb < 0 || bestr < min_ratio) return false;
	gr.add_edge(bestb, n, bestw, 0);
	added.push_back({bestb, n}); added_w.push_back(bestw);
	return true;
}

void revise_graph(const bundle &bd, graph &gr, revision &rv)
{
	rv.added.clear(); rv.added_w.clear();
	while(true)
	{
		bool b1 = start_round(gr, bd.prm.min_boundary_log_ratio, rv.added, rv.added_w);
		bool b2 = end_round(gr, bd.prm.min_boundary_log_ratio, rv.added, rv.added_w);
		if(!b1 && !b2) break;
	}
	// remove_false_boundaries (rnacore/graph_reviser.cc:1285-1377)
	const int nv = gr.nv();
	std::map<int, int> fb1, fb2;
	for(size_t i = 0; i < bd.frgs.size(); i++)
	{
		if(bd.frgs[i][2] != 0) continue;
		const hit &h1 = bd.hits[bd.frgs[i][0]], &h2 = bd.hits[bd.frgs[i][1]];
		int u1 = gr.locate_vertex(h1.rpos - 1), u2 = gr.locate_vertex(h2.pos);
		if(u1 < 0 || u2 < 0 || u1 >= u2) continue;
		fb1[u1]++; fb2[u2]++;
	}
	rv.leave_cnt.assign(nv, 0); rv.come_cnt.assign(nv, 0); rv.leave_ratio.assign(nv, 0.0); rv.come_ratio.assign(nv, 0.0);
	for(auto &x : fb1)
	{
		if(gr.find_edge(x.first, nv - 1) < 0) continue;
		double w = gr.vw[x.first];
		rv.leave_cnt[x.first] = x.second;
		rv.leave_ratio[x.first] = std::log(1 + x.second + w) - std::log(1 + w);
	}
	for(auto &x : fb2)
	{
		if(gr.find_edge(0, x.first) < 0) continue;
		double w = gr.vw[x.first];
		rv.come_cnt[x.first] = x.second;
		rv.come_ratio[x.first] = std::log(1 + x.second + w) - std::log(1 + w);
	}
}

} // namespace orc
