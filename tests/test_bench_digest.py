"""bench.py's full-size cross-check: the per-bundle digests it forms from the views of agpu_batch_results (mmap segments, frgs,
splices) equal the digests oracle/ref_driver.cc forms from the reference's own bundle objects after bundle::bridge()
(ref_timing_digest) -- and a corrupted view is noticed.  CPU tier: kernel-logic build; -m gpu: the CUDA path."""
import argparse
import os
import sys

import numpy as np
import pytest

import parity
from aletsch_b200 import gpu as G, hostlib as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_check(ctx, mode, templates):
    import bench
    batch, lt = parity.make_batch(mode, templates)
    gp, _ = parity.params_pair(lt)
    cfg = {"library_type": lt, "min_junction_support": 1}
    bt = ctx.upload(batch.view(), keepalive=batch)
    bt.bridge_all(gp)
    got = bench.view_digests(bt.results(G.RESULT_EVIDENCE | G.RESULT_FRAGMENTS), batch.n_bundles)
    bridged = bt.bundle_counts()[:, 3]
    bt.free()
    rt = bench.RefTimer(batch, cfg, 2, 1e9, calibrate=False)          # every bundle
    assert rt.step == 1 and len(rt.sample) == batch.n_bundles
    want = rt.digests()
    rt.run()
    assert np.array_equal(rt.per[:batch.n_bundles], bridged)
    rt.close()
    assert got.shape == want.shape and batch.n_bundles > 5
    assert np.array_equal(got, want), np.nonzero((got != want).any(axis=1))[0][:5]
    assert (got[:, 0] != 0).any() and (got[:, 1] != 0).any() == (mode == H.SYNTH_PAIRED)
    return batch.n_bundles


@pytest.mark.parametrize("mode,templates", [(H.SYNTH_PAIRED, 30000), (H.SYNTH_LONG, 2000)])
def test_view_digests_match_reference(emu_lib, checkers, mode, templates):
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    ctx = G.Context(0, lib_path=emu_lib)
    run_check(ctx, mode, templates)
    ctx.close()


def test_digest_is_order_and_value_sensitive():
    import bench
    a = np.array([5, 7, 9], np.int32)
    j = np.arange(3)
    z = np.zeros(3, np.int32)
    base = bench._mix_rows(a, z, z, j).sum(dtype=np.uint64)
    assert base != bench._mix_rows(a[::-1], z, z, j).sum(dtype=np.uint64)
    assert base != bench._mix_rows(a + np.array([0, 1, 0], np.int32), z, z, j).sum(dtype=np.uint64)
    off = np.array([0, 2, 2, 3], np.int64)
    s = bench._segmented_sum(bench._mix_rows(a, z, z, j), off)
    assert s[1] == 0 and s[0] + s[2] == base


@pytest.mark.gpu
def test_view_digests_match_reference_gpu(checkers):
    if "ref" not in checkers:
        pytest.skip("needs oracle/_ref/libaletsch_ref.so")
    ctx = G.Context(0)          # raises if libaletsch_gpu.so is missing or no device: no CPU fallback
    run_check(ctx, H.SYNTH_PAIRED, 60000)
    ctx.close()
