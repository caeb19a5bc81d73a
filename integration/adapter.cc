// see adapter.h
#include "adapter.h"

int agpu_adapter_graph(const agpu_graph_view *g, const agpu_revise_view *r, int b, const std::string &chrm, const std::string &gid,
		splice_graph &gr)
{
	gr.clear();
	gr.strand = (char)g->strand[b];
	gr.chrm = chrm;
	gr.gid = gid;
	const int v0 = g->vert_off[b], nv = g->vert_off[b + 1] - v0;
	if(nv < 2) return -1;
	// rnacore/graph_builder.cc:305-342
	for(int x = 0; x < nv; x++)
	{
		const int32_t *vi5 = g->vert + 5 * (size_t)(v0 + x);
		const double *vd3 = g->vert_d + 3 * (size_t)(v0 + x);
		gr.add_vertex();
		vertex_info vi;
		vi.lpos = vi5[0];
		vi.rpos = vi5[1];
		vi.type = 0;
		if(x >= 1 && x < nv - 1)
		{
			vi.length = vi5[2];
			vi.type = vi5[3];
			vi.regional = vi5[4] != 0;
			vi.stddev = vd3[1];
			if(vd3[2] >= 0) vi.maxcov = vd3[2];          // -1 marks the 1-bp stubs whose maximum the reference never sets
		}
		gr.set_vertex_weight(x, vd3[0]);
		gr.set_vertex_info(x, vi);
	}
	// :344-425, minus what refine_splice_graph removed
	for(int k = g->edge_off[b]; k < g->edge_off[b + 1]; k++)
	{
		const int32_t *e3 = g->edge + 3 * (size_t)k;
		if(e3[0] < 0) continue;
		edge_descriptor p = gr.add_edge(e3[0], e3[1]);
		edge_info ei;
		ei.weight = g->edge_d[k];
		ei.strand = e3[2];
		gr.set_edge_info(p, ei);
		gr.set_edge_weight(p, g->edge_d[k]);
	}
	if(r)
	{
		// identify_boundaries (rnacore/graph_reviser.cc:1109-1111, 1147-1149)
		for(int64_t k = r->edge_off[b]; k < r->edge_off[b + 1]; k++)
		{
			edge_descriptor p = gr.add_edge(r->edge[2 * k], r->edge[2 * k + 1]);
			gr.set_edge_weight(p, r->edge_w[k]);
			gr.set_edge_info(p, edge_info());
		}
		// remove_false_boundaries (:1346-1347, 1372-1373)
		const int64_t w0 = r->vert_off[b];
		if(r->vert_off[b + 1] - w0 != nv) return -2;
		for(int x = 0; x < nv; x++)
		{
			if(r->unbridge[2 * (w0 + x)] == 0 && r->unbridge[2 * (w0 + x) + 1] == 0) continue;
			vertex_info &vi = gr.get_editable_vertex_info(x);
			vi.unbridge_leaving_count = r->unbridge[2 * (w0 + x)];
			vi.unbridge_leaving_ratio = r->unbridge_ratio[2 * (w0 + x)];
			vi.unbridge_coming_count = r->unbridge[2 * (w0 + x) + 1];
			vi.unbridge_coming_ratio = r->unbridge_ratio[2 * (w0 + x) + 1];
		}
	}
	gr.build_vertex_index();
	return 0;
}

int agpu_adapter_phase_set(const agpu_phase_view *p, int b, phase_set &ps)
{
	ps.pmap.clear();
	for(int64_t k = p->phase_off[b]; k < p->phase_off[b + 1]; k++)
	{
		std::vector<int32_t> v(p->coords + p->coord_off[k], p->coords + p->coord_off[k + 1]);
		ps.add(v, p->count[k]);
	}
	return 0;
}


int agpu_adapter_chain_set(const agpu_chainset_view *v, int b, int64_t handle_off, int64_t n_handles, const uint8_t *handle_xs, chain_set &cs)
{
	cs.clear();
	const int c0 = v->bundle_chain_off[b], c1 = v->bundle_chain_off[b + 1];
	// chains arrive in insertion order: group index chain_grp, position inside the group = order of appearance
	std::vector<PI> where(c1 - c0);
	for(int c = c0; c < c1; c++)
	{
		const int gi = v->chain_grp[c];
		if(gi < 0) return -1;
		if(gi >= (int)cs.chains.size()) cs.chains.resize(gi + 1);
		std::vector<int32_t> coords(v->chain_val + v->chain_off[c], v->chain_val + v->chain_off[c + 1]);
		AI3 a = {v->chain_cnt[3 * c], v->chain_cnt[3 * c + 1], v->chain_cnt[3 * c + 2]};
		if(cs.chains[gi].empty() && !coords.empty()) cs.pmap.insert(PI(coords[0], gi));
		cs.chains[gi].push_back(PVI3(coords, a));
		where[c - c0] = PI(gi, (int)cs.chains[gi].size() - 1);
	}
	for(int64_t h = 0; h < n_handles; h++)
	{
		const int c = v->handle_chain[handle_off + h];
		if(c < 0) continue;
		if(c >= c1 - c0) return -2;
		const int xs = handle_xs ? (handle_xs[h] == '+' ? 1 : (handle_xs[h] == '-' ? 2 : (handle_xs[h] <= 2 ? handle_xs[h] : 0))) : 0;
		cs.hmap.insert(std::make_pair((int)h, AI3({where[c].first, where[c].second, xs})));
	}
	return 0;
}

int agpu_adapter_bundle(const agpu_evidence_view *ev, const agpu_fragments_view *fr, int b, int64_t hit_off, const uint8_t *hit_xs, bundle_base &bb)
{
	bb.lpos = ev->lpos[b];
	bb.rpos = ev->rpos[b];
	bb.strand = (char)ev->strand[b];
	// mmap: the segments in order (split_interval_map keeps every border: inserting them one by one reproduces the map)
	bb.mmap.clear();
	for(int64_t i = ev->seg_off[b]; i < ev->seg_off[b + 1]; i++)
		bb.mmap += std::make_pair(ROI(ev->seg[3 * i], ev->seg[3 * i + 1]), ev->seg[3 * i + 2]);
	bb.splices.assign(ev->splices + ev->splice_off[b], ev->splices + ev->splice_off[b + 1]);
	const int64_t nh = (int64_t)bb.hits.size();
	int rc = agpu_adapter_chain_set(&ev->hcst, b, hit_off, nh, hit_xs, bb.hcst);
	if(rc != 0) return rc;
	if(!fr) return 0;
	const int64_t f0 = fr->frg_off[b], f1 = fr->frg_off[b + 1];
	bb.frgs.clear();
	for(int64_t f = f0; f < f1; f++) bb.frgs.push_back(AI3({fr->frgs[3 * f], fr->frgs[3 * f + 1], fr->frgs[3 * f + 2]}));
	// fcst handles: the strand class update_bridges added the fragment with (rnacore/bundle_base.cc:472-495) is not part of the
	// view; it is the class of the chain's only non-zero count when there is just one (else 0: the consumers of fcst read the
	// chains and counts, hmap's third entry is only used by chain_set::remove)
	std::vector<uint8_t> fxs((size_t)(f1 - f0), 0);
	for(int64_t f = f0; f < f1; f++)
	{
		const int c = fr->fcst.handle_chain[f];
		if(c < 0) continue;
		const int32_t *cnt = fr->fcst.chain_cnt + 3 * (size_t)(fr->fcst.bundle_chain_off[b] + c);
		const int nz = (cnt[0] > 0) + (cnt[1] > 0) + (cnt[2] > 0);
		if(nz == 1) fxs[f - f0] = (uint8_t)(cnt[1] > 0 ? 1 : (cnt[2] > 0 ? 2 : 0));
	}
	return agpu_adapter_chain_set(&fr->fcst, b, f0, f1 - f0, fxs.data(), bb.fcst);
}

static void adapter_chain_coords(const agpu_chainset_view *hcst, int b, int chain, std::vector<int32_t> &out)
{
	out.clear();
	if(chain < 0) return;
	const int c = hcst->bundle_chain_off[b] + chain;
	out.assign(hcst->chain_val + hcst->chain_off[c], hcst->chain_val + hcst->chain_off[c + 1]);
}

int agpu_adapter_clusters(const agpu_cluster_view *cv, const agpu_chainset_view *hcst, int b, std::vector<pereads_cluster> &vc)
{
	vc.clear();
	for(int64_t c = cv->clu_off[b]; c < cv->clu_off[b + 1]; c++)
	{
		pereads_cluster pc;
		adapter_chain_coords(hcst, b, cv->chain1[c], pc.chain1);
		adapter_chain_coords(hcst, b, cv->chain2[c], pc.chain2);
		pc.bounds.assign(cv->bounds + 4 * c, cv->bounds + 4 * c + 4);
		pc.extend.assign(cv->extend + 4 * c, cv->extend + 4 * c + 4);
		pc.frlist.assign(cv->frlist + cv->frlist_off[c], cv->frlist + cv->frlist_off[c + 1]);
		pc.count = cv->count[c];
		vc.push_back(pc);
	}
	return 0;
}

int agpu_adapter_bridges(const agpu_bridge_view *bv, const agpu_cluster_view *cv, int b, std::vector<bridge_path> &opt)
{
	opt.clear();
	for(int64_t c = cv->clu_off[b]; c < cv->clu_off[b + 1]; c++)
	{
		bridge_path p;
		p.type = bv->type[c];
		p.strand = bv->strand[c];
		p.choices = bv->choices[c];
		p.score = bv->score[c];
		p.chain.assign(bv->chain + bv->chain_off[c], bv->chain + bv->chain_off[c + 1]);
		p.whole.assign(bv->whole + bv->whole_off[c], bv->whole + bv->whole_off[c + 1]);
		opt.push_back(p);
	}
	return 0;
}
