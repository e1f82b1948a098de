"""Pins the CPU oracle (oracle/) to the reference: its golden files, outputs of the reference
itself (tests/golden/make_golden.py) and Random123 known-answer vectors.  CPU only."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import ALL_TAGS, load_golden, oracle_params
from oracle import (closed_form_site_probability, mil_inference, mod_ratio, noisy_or_site_probability,
                    philox4x32_10, read_probabilities, sample_indices, sample_indices_many, sample_indices_mt19937,
                    block_layout)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))
        assert [int(v) for v in got] == want


def test_sample_indices_range_and_determinism():
    a = sample_indices(5, 2**33 + 3, 37, 64, 20)
    b = sample_indices(5, 2**33 + 3, 37, 64, 20)
    c = sample_indices(5, 2**33 + 4, 37, 64, 20)
    assert a.shape == (64, 20) and a.min() >= 0 and a.max() < 37
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # iteration i does not depend on n_iters while the block layout is the same (n_iters <= 2048)
    assert np.array_equal(sample_indices(5, 9, 37, 8, 20), sample_indices(5, 9, 37, 2048, 20)[:8])
    assert block_layout(1000) == (8, 4) and block_layout(2048) == (8, 8) and block_layout(10000) == (8, 40)
    assert block_layout(16384) == (8, 64) and block_layout(16385) == (9, 57) and block_layout(1) == (8, 1)
    # the first round of every lane is the first n_samples draws of its generator (both regimes)
    assert np.array_equal(sample_indices(5, 9, 37, 32, 7), sample_indices(5, 9, 37, 32, 8)[:, :7])
    assert np.array_equal(sample_indices(5, 9, 300, 32, 7), sample_indices(5, 9, 300, 32, 8)[:, :7])
    # paired regime (n_reads <= 256): even draws are the single-regime draws of every second word
    a = sample_indices(5, 9, 256, 32, 20)
    assert a.min() >= 0 and a.max() < 256 and len(np.unique(a)) > 200
    # vectorised == per-site
    many = sample_indices_many(5, [9, 2**33 + 3], [37, 41], 300, 20)
    assert np.array_equal(many[0], sample_indices(5, 9, 37, 300, 20))
    assert np.array_equal(many[1], sample_indices(5, 2**33 + 3, 41, 300, 20))


def test_mwc64x_recurrence_matches_reference_formulation():
    """The uint64 form t = A*x + c used by the oracle equals D. B. Thomas' 32-bit formulation
    (hi = mul_hi(x, A); x = x*A + c; c = hi + (x < c)) for random states."""
    rng = np.random.default_rng(0)
    A = 4294883355
    for _ in range(2000):
        x, c = int(rng.integers(0, 2**32)), int(rng.integers(0, A))
        t = A * x + c
        hi = (x * A) >> 32
        x2 = (x * A + c) & 0xFFFFFFFF
        c2 = hi + (1 if x2 < c else 0)
        assert (t & 0xFFFFFFFF, t >> 32) == (x2, c2)


def test_sample_indices_lanes_and_rounds_are_uncorrelated():
    idx = sample_indices(3, 77, 1000, 2048, 20).astype(np.float64)
    per_iter = idx.mean(axis=1)
    lanes = per_iter.reshape(-1, 32)                   # [round, lane]
    cc = np.corrcoef(lanes.T)                            # lane x lane over rounds
    off = cc[~np.eye(32, dtype=bool)]
    assert np.abs(off).max() < 0.5 and abs(off.mean()) < 0.05
    x = idx.ravel()
    r1 = np.corrcoef(x[:-1], x[1:])[0, 1]                # serial correlation of consecutive draws
    assert abs(r1) < 0.02


def test_sample_indices_uniform():
    for n, dof_bound in ((50, 100.0), (256, 370.0), (300, 420.0)):     # paired regime (first and second halves) and single
        idx = sample_indices(0, 11, n, 20000, 20)
        for part in (idx[:, 0::2].ravel(), idx[:, 1::2].ravel()):
            counts = np.bincount(part, minlength=n)
            expected = len(part) / n
            chi2 = ((counts - expected) ** 2 / expected).sum()
            assert chi2 < dof_bound, (n, chi2)                          # P(chi2_{n-1} > bound) ~ 1e-5
    # the two indices drawn from one word are uncorrelated
    idx = sample_indices(1, 3, 64, 4000, 20).astype(np.float64)
    assert abs(np.corrcoef(idx[:, 0::2].ravel(), idx[:, 1::2].ravel())[0, 1]) < 0.01


def test_mt19937_replay_matches_numpy_choice():
    p = np.linspace(0, 1, 41, dtype=np.float32)
    np.random.seed(3)
    drawn = np.random.choice(p, 5 * 20, replace=True).reshape(5, 20)
    idx = sample_indices_mt19937(3, 41, 5, 20)
    assert np.array_equal(drawn, p[idx])


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_read_probabilities_match_reference_model(tag, synthetic_inputs):
    g = load_golden(tag)
    params = oracle_params(tag)
    n_reads = np.diff(synthetic_inputs["read_off"])
    kmer_rows = np.repeat(synthetic_inputs["kmer_idx"], n_reads, axis=0)
    p = read_probabilities(params, synthetic_inputs["feats"], kmer_rows)
    # reference = torch CPU forward (fixtures); the restatement differs only by summation order
    assert np.max(np.abs(p - g["read_prob"])) <= 2e-6
    np.testing.assert_allclose(p, g["read_prob"], rtol=2e-4, atol=1e-30)


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_mil_inference_matches_literal_forward(tag, synthetic_inputs):
    """Factored oracle (p once per read, table look-ups) == the reference's literal MIL forward
    on the gathered bags, on the shared Philox index stream."""
    g = load_golden(tag)
    params = oracle_params(tag)
    rp, sp, mc = mil_inference(params, synthetic_inputs["feats"], synthetic_inputs["read_off"], synthetic_inputs["kmer_idx"],
                               n_iters=int(g["n_iters"]), seed=int(g["seed"]), site_id_base=int(g["site_id_base"]),
                               n_samples=int(g["n_samples"]), read_threshold=float(g["threshold"]))
    assert np.max(np.abs(sp - g["site_prob"])) <= 1e-5
    assert np.array_equal(mc, g["mod_count"])
    assert np.allclose(mc / np.diff(synthetic_inputs["read_off"]), g["mod_ratio"], rtol=0, atol=0)


def test_bundled_read_probs_match_reference_goldens(bundled_flat, golden_dir):
    """Reference golden data.indiv_proba.csv.gz with the reference test's own tolerance
    (m6anet/tests/test_inference.py:32: np.allclose default rtol 1e-5, atol 1e-8)."""
    params = oracle_params("HCT116_RNA002")
    b = bundled_flat
    n_reads = np.diff(b["read_off"])
    kmer_rows = np.repeat(b["kmer_idx"], n_reads, axis=0)
    p = read_probabilities(params, b["feats"], kmer_rows)
    assert np.max(np.abs(p - b["read_prob"])) <= 1e-6
    gold = pd.read_csv(os.path.join(golden_dir, "bundled", "data.indiv_proba.csv.gz"))
    mine = pd.DataFrame({"transcript_id": np.repeat(b["tx_id"], n_reads), "transcript_position": np.repeat(b["tx_pos"], n_reads),
                         "read_index": b["read_id"].astype(np.int64), "probability_modified": p.astype(np.float64)})
    keys = ["transcript_id", "transcript_position", "read_index"]
    gold = gold.sort_values(keys).reset_index(drop=True)
    mine = mine.sort_values(keys).reset_index(drop=True)
    assert np.all(gold["transcript_id"] == mine["transcript_id"])
    assert np.all(gold["transcript_position"] == mine["transcript_position"])
    assert np.all(gold["read_index"] == mine["read_index"])
    assert np.allclose(gold["probability_modified"], mine["probability_modified"])


def test_bundled_site_outputs_match_reference_goldens(bundled_flat, golden_dir):
    """Reference golden data.site_proba.csv.gz: mod_ratio allclose, site probability atol 1e-2
    (m6anet/tests/test_inference.py:36-37; the reference tests at 10000 iterations)."""
    params = oracle_params("HCT116_RNA002")
    b = bundled_flat
    rp, sp, mc = mil_inference(params, b["feats"], b["read_off"], b["kmer_idx"], n_iters=10000, seed=0,
                               read_threshold=0.033379376)
    gold = pd.read_csv(os.path.join(golden_dir, "bundled", "data.site_proba.csv.gz"))
    mine = pd.DataFrame({"transcript_id": b["tx_id"], "transcript_position": b["tx_pos"],
                         "probability_modified": sp.astype(np.float64), "mod_ratio": mc / np.diff(b["read_off"])})
    keys = ["transcript_id", "transcript_position"]
    gold = gold.sort_values(keys).reset_index(drop=True)
    mine = mine.sort_values(keys).reset_index(drop=True)
    assert len(gold) == len(mine) == 101
    assert np.all(gold["transcript_id"] == mine["transcript_id"])
    assert np.allclose(gold["mod_ratio"], mine["mod_ratio"])
    assert np.allclose(gold["probability_modified"], mine["probability_modified"], atol=1e-2)
    # RNG sanity against the closed form E = 1-(1-mean p)^20
    for s in range(len(sp)):
        e = closed_form_site_probability(rp[b["read_off"][s]:b["read_off"][s + 1]])
        assert abs(sp[s] - e) < 3 * 0.5 / np.sqrt(10000) + 1e-6


def test_mt19937_replay_matches_reference_function(bundled_flat):
    """oracle noisy-OR on the replayed MT19937 indices == the reference's `_calculate_site_proba`
    itself (fixture computed by the reference right after np.random.seed(0))."""
    b = bundled_flat
    for s, want in zip(b["replay_sites"], b["replay_site_prob"]):
        p = b["read_prob"][b["read_off"][s]:b["read_off"][s + 1]]
        idx = sample_indices_mt19937(int(b["replay_seed"]), len(p), int(b["replay_n_iters"]), 20)
        got = noisy_or_site_probability(p, idx)
        assert got.dtype == np.float32
        assert got == want  # same library calls, bit-identical


def test_mod_ratio_threshold_is_float32():
    thr = 0.033379376
    p = np.array([np.float32(thr), np.nextafter(np.float32(thr), np.float32(0))], dtype=np.float32)
    assert mod_ratio(p, thr) == 0.5


def test_empty_site_and_single_read():
    params = oracle_params("HCT116_RNA002")
    feats = np.zeros((1, 9), np.float32)
    rp, sp, mc = mil_inference(params, feats, np.array([0, 0, 1]), np.array([[0, 1, 2], [3, 4, 5]]), n_iters=4)
    assert np.isnan(sp[0]) and mc[0] == 0
    assert np.isclose(sp[1], 1 - (1 - rp[0]) ** 20, atol=1e-6)


def test_torch_port_used_for_cpu_timing_equals_numpy_restatement(synthetic_inputs):
    from oracle.cpu_baseline import read_probabilities_torch
    si = synthetic_inputs
    kmer_rows = np.repeat(si["kmer_idx"], np.diff(si["read_off"]), axis=0)
    for tag in ("HCT116_RNA002", "signal_only"):
        params = oracle_params(tag)
        a = read_probabilities(params, si["feats"], kmer_rows)
        b = read_probabilities_torch(params, si["feats"], kmer_rows)
        assert np.max(np.abs(a - b)) <= 2e-6
        assert np.array_equal(b, load_golden(tag)["read_prob"])   # identical calls to the reference => identical bits


def test_c_restatement_equals_numpy_oracle(synthetic_inputs):
    """oracle/c (used to check every site of full-size jobs) against the NumPy restatement and the reference goldens."""
    from oracle import c_oracle
    si = synthetic_inputs
    for seed, site, n in [(0, 0, 20), (1234, 7_000_000_123, 50), (2**63 + 5, 2**40 + 1, 4001)]:
        assert np.array_equal(c_oracle.sample_indices(seed, site, n, 2100, 20).astype(np.int64), sample_indices(seed, site, n, 2100, 20))
    for tag in ("HCT116_RNA002", "signal_only"):
        g = load_golden(tag)
        kw = dict(n_iters=int(g["n_iters"]), seed=int(g["seed"]), site_id_base=int(g["site_id_base"]), n_samples=20,
                  read_threshold=float(g["threshold"]))
        rp, sp, mc = c_oracle.mil_inference(oracle_params(tag), si["feats"], si["read_off"], si["kmer_idx"], **kw)
        assert np.max(np.abs(rp - g["read_prob"])) <= 2e-6
        assert np.max(np.abs(sp - g["site_prob"])) <= 1e-5          # reference literal MIL forward on the shared stream
        assert np.array_equal(mc, g["mod_count"])
