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
