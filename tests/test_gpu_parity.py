"""-m gpu: parity of the CUDA path (through the C ABI) against the CPU checkers on a B200."""
import pytest

import parity
from aletsch_b200 import gpu as G
from aletsch_b200 import hostlib as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = G.Context(0)          # raises if libaletsch_gpu.so is missing or no device: no CPU fallback
    yield c
    c.close()


@pytest.mark.parametrize("mode,templates", [(H.SYNTH_PAIRED, 30000), (H.SYNTH_SINGLE, 20000), (H.SYNTH_LONG, 3000)])
def test_evidence_graph_parity(ctx, checkers, mode, templates):
    assert checkers, "no CPU checker library found (oracle/liboracle.so or oracle/_ref/libaletsch_ref.so)"
    batch, lt = parity.make_batch(mode, templates)
    assert batch.n_bundles > 0
    gp, op = parity.params_pair(lt)
    for name, chk in checkers.items():
        bad = parity.compare_evidence_graph(ctx, batch, chk, gp, op)
        assert not bad, "%s: %d mismatches, first: %s" % (name, len(bad), bad[:3])
