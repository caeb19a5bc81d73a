// Reference-side adapter: the code a maintainer of the reference adds next to meta/assembler.cc to consume the C ABI
// (include/aletsch_gpu.h) in place of assembler::transform(bd, gr, true) and bundle_base::build_phase_set(ps, gr).  It is
// compiled against the REFERENCE'S headers (splice_graph.h, phase_set.h), so in this repository it is built only into the
// test-side reference build (oracle/_ref, see oracle/Makefile) where tests/test_adapter_transcripts.py runs the reference's
// own scallop on graphs rebuilt from the views and compares the transcripts with the reference's end-to-end result.
#ifndef ALETSCH_B200_INTEGRATION_ADAPTER_H
#define ALETSCH_B200_INTEGRATION_ADAPTER_H

#include "../include/aletsch_gpu.h"
#include "splice_graph.h"
#include "phase_set.h"
#include "chain_set.h"
#include "bundle_base.h"
#include "pereads_cluster.h"
#include "bridge_path.h"

// splice graph of bundle b as assembler::transform(bd, gr, true) leaves it (meta/assembler.cc:930-944): vertices and alive
// edges of agpu_graph_view in insertion order, then the boundary edges and vertex annotations of agpu_revise_view (pass NULL
// for transform(bd, gr, false)), then the vertex index
int agpu_adapter_graph(const agpu_graph_view *g, const agpu_revise_view *r, int b, const std::string &chrm, const std::string &gid,
		splice_graph &gr);

// phase set of bundle b as bundle_base::build_phase_set fills it (rnacore/bundle_base.cc:338-418)
int agpu_adapter_phase_set(const agpu_phase_view *p, int b, phase_set &ps);

// chain_set of bundle b (hcst from agpu_evidence_view, fcst from agpu_fragments_view) as chain_set::add leaves it
// (rnacore/chain_set.cc:64-123): chains[i][j] in insertion order with their AI3 counts, pmap, and hmap for the handles (hit /
// fragment indices local to the bundle: handle_off = bundle_hit_off[b] resp. frg_off[b]); handle_xs[h] gives the third hmap
// entry (0 '.', 1 '+', 2 '-': the XS class the handle was added with)
int agpu_adapter_chain_set(const agpu_chainset_view *v, int b, int64_t handle_off, int64_t n_handles, const uint8_t *handle_xs, chain_set &cs);

// what bundle::bridge (meta/bundle.cc:55-88) leaves in bundle_base, rebuilt from the evidence and fragments views of bundle b:
// lpos / rpos / strand, mmap, splices, hcst, frgs, fcst.  `bb.hits` must already hold the bundle's hits (the host has them: the
// device never changes a hit); handle_xs for hcst = xs of the hits, for fcst = the strand class update_bridges used (recovered
// from the fcst counts: a chain's handles share one class whenever the chain has a single non-zero count)
int agpu_adapter_bundle(const agpu_evidence_view *ev, const agpu_fragments_view *fr, int b, int64_t hit_off, const uint8_t *hit_xs, bundle_base &bb);

// vector<pereads_cluster> of bundle b as graph_cluster::build_pereads_clusters leaves it (rnacore/graph_cluster.cc:93-168):
// chain1 / chain2 expanded from the hcst chain indices of the cluster view
int agpu_adapter_clusters(const agpu_cluster_view *cv, const agpu_chainset_view *hcst, int b, std::vector<pereads_cluster> &vc);

// bridge_solver::opt of bundle b (bridge/bridge_solver.cc:276-385): type, strand, choices, score, chain, whole per cluster
int agpu_adapter_bridges(const agpu_bridge_view *bv, const agpu_cluster_view *cv, int b, std::vector<bridge_path> &opt);

#endif
