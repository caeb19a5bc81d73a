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


@pytest.mark.parametrize("mode,templates,seed", [(H.SYNTH_PAIRED, 60000, 20260101), (H.SYNTH_PAIRED, 150000, 20260102),
                                                 (H.SYNTH_SINGLE, 20000, 20260104), (H.SYNTH_LONG, 3000, 20260105)])
def test_bundle_bridge_parity(ctx, checkers, mode, templates, seed):
    """the whole per-bundle path (bundle::bridge) with every intermediate compared"""
    assert checkers
    batch, lt = parity.make_batch(mode, templates, seed=seed)
    gp, op = parity.params_pair(lt)
    for name, chk in checkers.items():
        stats = {}
        bad = parity.compare_full(ctx, batch, chk, gp, op, stats)
        assert not bad, "%s: %d mismatches, first: %s" % (name, len(bad), bad[:3])
        if mode == H.SYNTH_PAIRED:
            assert stats["bridged"] > 0 and stats["clusters"] > 0
