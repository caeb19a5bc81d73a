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

// splice graph of bundle b as assembler::transform(bd, gr, true) leaves it (meta/assembler.cc:930-944): vertices and alive
// edges of agpu_graph_view in insertion order, then the boundary edges and vertex annotations of agpu_revise_view (pass NULL
// for transform(bd, gr, false)), then the vertex index
int agpu_adapter_graph(const agpu_graph_view *g, const agpu_revise_view *r, int b, const std::string &chrm, const std::string &gid,
		splice_graph &gr);

// phase set of bundle b as bundle_base::build_phase_set fills it (rnacore/bundle_base.cc:338-418)
int agpu_adapter_phase_set(const agpu_phase_view *p, int b, phase_set &ps);

#endif
