"""Adversarial random bundles (tests/fuzz.py) through the whole per-bundle path against BOTH CPU checkers: ragged CIGARs with
every op (M I D N S = X), position ties, qname groups of 1-4 hits, inconsistent mates, empty bundles (H = 0),
single-hit bundles, an empty batch.  CPU tier: kernel-logic build; the -m gpu tier repeats it on the CUDA path."""
import numpy as np
import pytest

import fuzz
import parity
from aletsch_b200 import gpu as G
from aletsch_b200 import hostlib as H


def run_seeds(ctx, checkers, seeds, big=False):
    assert checkers
    for seed in seeds:
        lt = H.FR_FIRST if seed % 4 else H.UNSTRANDED
        # the reference asserts strand != '.' for an empty bundle of a stranded library (rnacore/bundle_base.cc:208)
        batch = fuzz.random_batch(seed, n_bundles=8, max_hits=(1200 if big else (70 if seed % 3 else 250)),
                                  empty_every=(4 if lt == H.UNSTRANDED else 0), exon_grid=(seed % 2 == 0))
        gp, op = parity.params_pair(lt)
        for name, chk in checkers.items():
            bad = parity.compare_full(ctx, batch, chk, gp, op, {})
            assert not bad, "seed %d vs %s: %d mismatches, first: %s" % (seed, name, len(bad), bad[:3])
            if lt != H.UNSTRANDED:
                bad = parity.compare_phase_set(ctx, batch, chk, gp, op)
                assert not bad, "seed %d vs %s (phase set): %d mismatches, first: %s" % (seed, name, len(bad), bad[:3])
                gr, orr = parity.params_pair(lt, min_boundary_log_ratio=(2.0 if seed % 2 else 0.9))
                bad = parity.compare_revise(ctx, batch, chk, gr, orr)
                assert not bad, "seed %d vs %s (revision): %d mismatches, first: %s" % (seed, name, len(bad), bad[:3])


def run_pair_seeds(ctx, checkers, seeds):
    """mate pairing under candidate relations that are not isolated pairs (bundle_base::build_fragments, rnacore/bundle_base.cc:
    267-323): shared query names, several hits per position, self-pointing and one-sided mates -- the join on the sorted
    positions must hand exactly these to the greedy of the reference and pair the rest as it would"""
    for seed in seeds:
        batch = fuzz.random_batch(500 + seed, n_bundles=6, max_hits=(3000 if seed % 4 == 0 else 300), exon_grid=False, pair_heavy=True)
        gp, op = parity.params_pair(H.FR_FIRST)
        for name, chk in checkers.items():
            bad = parity.compare_full(ctx, batch, chk, gp, op, {})
            assert not bad, "seed %d vs %s: %d mismatches, first: %s" % (seed, name, len(bad), bad[:3])


def run_group_seeds(ctx, checkers, seeds):
    """assembler::bridge on random clusters of random bundles, with and without a per-bundle bridging round before"""
    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        batch = fuzz.random_batch(seed, n_bundles=9, max_hits=(60 if seed % 3 else 200), exon_grid=True)
        groups = fuzz.strand_clusters(batch, rng)
        gp, op = parity.params_pair(H.FR_FIRST)
        for name, chk in checkers.items():
            for first_round in (True, False):
                bad = parity.compare_group_bridge(ctx, batch, chk, gp, op, groups, {}, first_round=first_round)
                assert not bad, "seed %d vs %s (first_round=%s): %d mismatches, first: %s" % (seed, name, first_round, len(bad), bad[:3])


def run_degenerate(ctx, checkers):
    gp, op = parity.params_pair(H.UNSTRANDED)
    # an empty batch: every stage must accept it
    empty = fuzz.random_batch(1, n_bundles=0)
    bt = ctx.upload(empty.view(), keepalive=empty)
    bt.bridge_all(gp)
    c = bt.counts()
    assert c["hits"] == 0 and c["segments"] == 0 and c["fragments"] == 0 and c["clusters"] == 0
    bt.free()
    # only empty bundles; single-hit bundles
    for batch in (fuzz.random_batch(2, n_bundles=3, max_hits=1, empty_every=1), fuzz.random_batch(3, n_bundles=6, max_hits=1)):
        for name, chk in checkers.items():
            bad = parity.compare_full(ctx, batch, chk, gp, op, {})
            assert not bad, (name, bad[:3])


@pytest.fixture(scope="module")
def ctx(emu_lib):
    c = G.Context(0, lib_path=emu_lib)
    yield c
    c.close()


def test_fuzz_bundles(ctx, checkers):
    run_seeds(ctx, checkers, range(14))


def test_fuzz_pairing(ctx, checkers):
    run_pair_seeds(ctx, checkers, range(8))


def test_fuzz_group_bridge(ctx, checkers):
    run_group_seeds(ctx, checkers, range(8))


def test_degenerate_batches(ctx, checkers):
    run_degenerate(ctx, checkers)
