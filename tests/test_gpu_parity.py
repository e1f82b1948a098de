"""GPU parity tests: the CUDA kernel (through the C ABI) against the oracle and the committed
reference goldens, on the same inputs and the same index stream.

Tolerances (BASELINE.json north_star): probability_modified <= 1e-4 absolute on shared indices;
per-read probability held to 2e-6 absolute / the reference test's np.allclose; integer outputs
(indices, mod_count away from the threshold) bit-exact."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import ALL_TAGS, ASSETS, MODEL_FILES, load_golden, oracle_params

pytestmark = pytest.mark.gpu

SITE_ATOL = 1e-4
READ_ATOL = 2e-6


def product_weights(tag):
    from m6anet_b200 import weights as W
    if tag == "signal_only":
        g = load_golden(tag)
        return W.from_raw({k[2:]: g[k] for k in g if k.startswith("w_")})
    return W.from_npz(os.path.join(ASSETS, "model_states", MODEL_FILES[tag]))


_ENGINES = {}
_ENCODER = "tc"


@pytest.fixture(autouse=True, params=["tc", "ffma"])
def _both_encoders(request):
    """Every test of this module runs with the tensor-core read encoder (the default) AND with the CUDA-core FFMA2 encoder:
    same tolerances, same index streams.  (Entry points the tensor-core kernel does not serve -- explicit indices, other bag
    sizes -- run the FFMA kernel under either setting.)"""
    global _ENCODER
    _ENCODER = request.param
    for e in _ENGINES.values():
        e.set_encoder(_ENCODER)
    yield request.param


def engine(tag):
    from m6anet_b200.engine import MilEngine
    if tag not in _ENGINES:
        _ENGINES[tag] = MilEngine(product_weights(tag), "cuda:0")
    _ENGINES[tag].set_encoder(_ENCODER)
    return _ENGINES[tag]


def run_device(eng, feats, read_off, kmer_idx, n_iters, **kw):
    import torch
    dev = eng.device
    f = torch.from_numpy(np.ascontiguousarray(feats, np.float32)).to(dev)
    o = torch.from_numpy(np.ascontiguousarray(read_off, np.int64)).to(dev)
    k = None if kmer_idx is None else torch.from_numpy(np.ascontiguousarray(kmer_idx, np.int32)).to(dev)
    if kw.get("sample_idx") is not None:
        kw["sample_idx"] = torch.from_numpy(np.ascontiguousarray(kw["sample_idx"], np.uint16)).to(dev)
    rp, sp, mc = eng.infer_device(f, o, k, n_iters, **kw)
    torch.cuda.synchronize()
    return rp.cpu().numpy(), sp.cpu().numpy(), mc.cpu().numpy()


def assert_mod_count(mc, read_prob_ref, read_off, thr):
    """exact wherever no read sits within 1e-6 of the threshold"""
    thr32 = np.float32(thr)
    for s in range(len(mc)):
        p = read_prob_ref[read_off[s]:read_off[s + 1]]
        want = int(np.count_nonzero(p >= thr32))
        slack = int(np.count_nonzero(np.abs(p.astype(np.float64) - float(thr32)) < 1e-6))
        assert abs(int(mc[s]) - want) <= slack, (s, mc[s], want, slack)


def test_library_is_loaded_and_native():
    from m6anet_b200 import _cabi
    L = _cabi.lib()
    assert L.m6a_version() == 100
    with open("/proc/self/maps") as f:
        assert "libm6anet_b200.so" in f.read()


@pytest.mark.parametrize("seed,site,n", [(0, 0, 20), (1234, 7_000_000_123, 50), (2**63 + 5, 2**40 + 1, 4001), (7, 3, 1),
                                         (7, 3, 256), (7, 3, 257)])      # 256 | 257: paired / single index regimes
@pytest.mark.parametrize("n_samples", [20, 7])
def test_device_index_stream_matches_oracle(seed, site, n, n_samples):
    from oracle import sample_indices
    got = engine("HCT116_RNA002").sample_indices(seed, site, n, 257, n_samples).cpu().numpy()
    want = sample_indices(seed, site, n, 257, n_samples)
    assert np.array_equal(got.astype(np.int64), want)


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_synthetic_parity_vs_reference_goldens(tag, synthetic_inputs):
    """Kernel vs the unmodified reference (torch CPU read encoder + literal MILModel.forward on the
    gathered bags) on the shared Philox stream; all 66 k-mer ids, ragged n_reads 20..700."""
    g = load_golden(tag)
    si = synthetic_inputs
    rp, sp, mc = run_device(engine(tag), si["feats"], si["read_off"], si["kmer_idx"], int(g["n_iters"]),
                            seed=int(g["seed"]), site_id_base=int(g["site_id_base"]), n_samples=int(g["n_samples"]),
                            read_threshold=float(g["threshold"]))
    assert np.max(np.abs(rp - g["read_prob"])) <= READ_ATOL
    np.testing.assert_allclose(rp, g["read_prob"], rtol=3e-4, atol=1e-30)
    assert np.max(np.abs(sp - g["site_prob"])) <= SITE_ATOL
    assert_mod_count(mc, g["read_prob"], si["read_off"], float(g["threshold"]))


@pytest.mark.parametrize("tag", ["HCT116_RNA002", "signal_only"])
def test_synthetic_parity_vs_oracle(tag, synthetic_inputs):
    from oracle import mil_inference
    si = synthetic_inputs
    kw = dict(n_iters=1000, seed=99, site_id_base=123456789012, n_samples=20, read_threshold=0.033379376)
    rp, sp, mc = run_device(engine(tag), si["feats"], si["read_off"], si["kmer_idx"], kw.pop("n_iters"), **kw)
    orp, osp, omc = mil_inference(oracle_params(tag), si["feats"], si["read_off"], si["kmer_idx"], n_iters=1000, **kw)
    assert np.max(np.abs(rp - orp)) <= READ_ATOL
    assert np.max(np.abs(sp - osp)) <= SITE_ATOL
    assert_mod_count(mc, orp, si["read_off"], kw["read_threshold"])


def test_bundled_data_vs_reference_golden_csvs(bundled_flat, golden_dir):
    """The reference's own acceptance test (m6anet/tests/test_inference.py:10-37) on flat buffers."""
    b = bundled_flat
    rp, sp, mc = run_device(engine("HCT116_RNA002"), b["feats"], b["read_off"], b["kmer_idx"], 10000, seed=0,
                            read_threshold=0.033379376)
    n_reads = np.diff(b["read_off"])
    gold = pd.read_csv(os.path.join(golden_dir, "bundled", "data.indiv_proba.csv.gz"))
    mine = pd.DataFrame({"transcript_id": np.repeat(b["tx_id"], n_reads), "transcript_position": np.repeat(b["tx_pos"], n_reads),
                         "read_index": b["read_id"].astype(np.int64), "probability_modified": rp.astype(np.float64)})
    keys = ["transcript_id", "transcript_position", "read_index"]
    gold = gold.sort_values(keys).reset_index(drop=True)
    mine = mine.sort_values(keys).reset_index(drop=True)
    assert np.all(gold["read_index"] == mine["read_index"])
    assert np.allclose(gold["probability_modified"], mine["probability_modified"])
    gs = pd.read_csv(os.path.join(golden_dir, "bundled", "data.site_proba.csv.gz"))
    ms = pd.DataFrame({"transcript_id": b["tx_id"], "transcript_position": b["tx_pos"],
                       "probability_modified": sp.astype(np.float64), "mod_ratio": mc / n_reads})
    keys = ["transcript_id", "transcript_position"]
    gs = gs.sort_values(keys).reset_index(drop=True)
    ms = ms.sort_values(keys).reset_index(drop=True)
    assert np.allclose(gs["mod_ratio"], ms["mod_ratio"])
    assert np.allclose(gs["probability_modified"], ms["probability_modified"], atol=1e-2)


def test_explicit_indices_replay_reference_mt19937(bundled_flat):
    """Explicit-index mode fed with the reference's own MT19937 draw reproduces the value the
    reference function `_calculate_site_proba` returned (fixture made by the reference)."""
    from oracle import sample_indices_mt19937
    b = bundled_flat
    sites = [int(s) for s in b["replay_sites"]]
    n_iters = int(b["replay_n_iters"])
    ro = b["read_off"]
    rows = np.concatenate([np.arange(ro[s], ro[s + 1]) for s in sites])
    off = np.concatenate([[0], np.cumsum([ro[s + 1] - ro[s] for s in sites])])
    idx = np.stack([sample_indices_mt19937(int(b["replay_seed"]), int(ro[s + 1] - ro[s]), n_iters, 20) for s in sites])
    rp, sp, mc = run_device(engine("HCT116_RNA002"), b["feats"][rows], off, b["kmer_idx"][sites], n_iters,
                            sample_idx=idx.astype(np.uint16))
    assert np.max(np.abs(sp - b["replay_site_prob"])) <= 2e-6


def _random_case(rng, n_reads_list, n_kmer=66):
    n_reads = np.asarray(n_reads_list, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(n_reads)])
    feats = (rng.standard_normal((int(off[-1]), 9)) * 1.2).astype(np.float32)
    kmer = rng.integers(0, n_kmer, size=(len(n_reads), 3)).astype(np.int32)
    return feats, off, kmer


@pytest.mark.parametrize("n_reads_list,n_iters,n_samples", [
    ([1, 1, 2, 3, 0, 5, 0], 100, 20),                 # tiny and empty sites
    ([20] * 7 + [0] + [33] * 5, 1, 20),               # single iteration
    ([25] * 40, 33, 20),                              # slab tail
    ([5000, 20, 4097, 21], 64, 20),                   # tile larger than the shared q table (global fallback)
    ([40] * 9, 5000, 20),                             # more than 2048 iterations: several iterations per lane
    ([30] * 11, 200, 7),                              # generic n_samples
    ([30] * 11, 200, 1),
    ([600] * 3 + [20] * 100, 128, 20),                # multi-chunk tiles
    ([256, 257, 255, 300, 20], 96, 20),               # paired / single index regimes side by side
    ([256, 257, 40], 96, 5),                          # odd n_samples in the paired regime
])
def test_edge_cases_vs_oracle(n_reads_list, n_iters, n_samples):
    from oracle import mil_inference
    rng = np.random.default_rng(len(n_reads_list) * 1000 + n_iters)
    feats, off, kmer = _random_case(rng, n_reads_list)
    kw = dict(seed=5, site_id_base=2**35, n_samples=n_samples, read_threshold=0.05)
    rp, sp, mc = run_device(engine("arabidopsis_RNA002"), feats, off, kmer, n_iters, **kw)
    orp, osp, omc = mil_inference(oracle_params("arabidopsis_RNA002"), feats, off, kmer, n_iters=n_iters, **kw)
    assert np.max(np.abs(rp - orp), initial=0) <= READ_ATOL
    empty = np.diff(off) == 0
    assert np.all(np.isnan(sp[empty])) and not np.any(np.isnan(sp[~empty]))
    assert np.max(np.abs(sp[~empty] - osp[~empty]), initial=0) <= SITE_ATOL
    assert_mod_count(mc, orp, off, 0.05)


@pytest.mark.parametrize("n_sites,n_reads", [(148, 50), (300, 50), (1000, 50), (5000, 50), (2000, 23), (777, 128), (40, 400)])
def test_short_jobs_with_few_tiles_per_sm(n_sites, n_reads):
    """Jobs of a few MMA tiles per SM: the tile size is balanced over the SMs, so a CTA gets one tile of 1-4 MMA tiles and its
    two staging groups send the stop right after their last tile (the hand-over that once completed a barrier phase too
    early).  Repeated, because the failure was a race; every site against the oracle once, bit-identical afterwards."""
    from oracle import c_oracle
    rng = np.random.default_rng(n_sites * 131 + n_reads)
    feats = rng.standard_normal((n_sites * n_reads, 9), dtype=np.float32)
    off = np.arange(n_sites + 1, dtype=np.int64) * n_reads
    kmer = rng.integers(0, 66, size=(n_sites, 3)).astype(np.int32)
    eng = engine("HCT116_RNA002")
    kw = dict(seed=2, site_id_base=5 * 10**9, read_threshold=0.033379376)
    first = run_device(eng, feats, off, kmer, 1000, **kw)
    orp, osp, omc = c_oracle.mil_inference(oracle_params("HCT116_RNA002"), feats, off, kmer, 1000, n_samples=20, **kw)
    assert np.max(np.abs(first[0] - orp)) <= READ_ATOL
    assert np.max(np.abs(first[1] - osp)) <= SITE_ATOL
    assert_mod_count(first[2], orp, off, 0.033379376)
    for _ in range(20):
        again = run_device(eng, feats, off, kmer, 100, **kw)
        assert np.array_equal(again[0], first[0])


def test_random_job_shapes_both_encoders_agree(_both_encoders):
    """Stress of the role hand-overs of the warp-specialised kernel: 120 random job shapes (site counts 1..4000, uniform,
    ragged, with empty sites and with a site larger than the q table, random tile sizes) -- the two encoders are independent
    implementations of the same expression, so they must agree to float32 round-off everywhere (and a lost arrival would trap)."""
    if _both_encoders != "tc":
        pytest.skip("one pass compares both encoders")
    import torch
    eng = engine("HCT116_RNA002")
    rng = np.random.default_rng(2024)
    for case in range(120):
        kind = case % 4
        n_sites = int(rng.integers(1, 4000))
        if kind == 0:
            n_reads = np.full(n_sites, int(rng.integers(1, 120)), dtype=np.int64)
        elif kind == 1:
            n_reads = np.clip(np.round(np.exp(rng.normal(np.log(30), 0.9, size=n_sites))), 1, 900).astype(np.int64)
        elif kind == 2:
            n_reads = rng.integers(0, 60, size=n_sites).astype(np.int64)
            n_reads[rng.random(n_sites) < 0.3] = 0
        else:
            n_reads = rng.integers(20, 40, size=min(n_sites, 300)).astype(np.int64)
            n_reads[int(rng.integers(0, len(n_reads)))] = int(rng.integers(6145, 9000))
        feats, off, kmer = _random_case(rng, n_reads)
        tile = 0 if case % 3 else int(rng.integers(1, 33)) * 128      # 0 = automatic
        n_iters = int(rng.choice([1, 50, 300, 1000]))
        out = {}
        for enc in ("tc", "ffma"):
            eng.set_encoder(enc)
            f = torch.from_numpy(feats).to(eng.device)
            o = torch.from_numpy(off).to(eng.device)
            k = torch.from_numpy(kmer).to(eng.device)
            eng.set_tile_reads(tile if enc == "tc" else 0)
            rp, sp, mc = eng.infer_device(f, o, k, n_iters, seed=case, site_id_base=case * 10**6)
            torch.cuda.synchronize()
            eng.set_tile_reads(0)
            out[enc] = (rp.cpu().numpy(), sp.cpu().numpy(), mc.cpu().numpy())
        eng.set_encoder("tc")
        live = n_reads > 0
        assert np.max(np.abs(out["tc"][0] - out["ffma"][0]), initial=0) <= 3e-6, case
        assert np.all(np.isnan(out["tc"][1][~live])) and np.all(out["tc"][2][~live] == 0), case
        assert np.max(np.abs(out["tc"][1][live] - out["ffma"][1][live]), initial=0) <= SITE_ATOL, case


def test_sharding_and_tiling_do_not_change_results(synthetic_inputs):
    """Counter-based RNG keyed by global site id + tiling-independent summation order: any split
    of the site list gives bit-identical outputs."""
    si = synthetic_inputs
    eng = engine("HEK293T_RNA004")
    ro = si["read_off"]
    full = run_device(eng, si["feats"], ro, si["kmer_idx"], 777, seed=3, site_id_base=10**12)
    for cuts in ([0, 100, 288], [0, 1, 2, 150, 287, 288]):
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            parts.append(run_device(eng, si["feats"][ro[a]:ro[b]], ro[a:b + 1] - ro[a], si["kmer_idx"][a:b], 777,
                                    seed=3, site_id_base=10**12 + a))
        for i in range(3):
            assert np.array_equal(np.concatenate([p[i] for p in parts]), full[i])


def test_unaligned_feature_pointer_uses_plain_loads(synthetic_inputs):
    import torch
    si = synthetic_inputs
    eng = engine("HCT116_RNA002")
    want = run_device(eng, si["feats"], si["read_off"], si["kmer_idx"], 50, seed=1)
    dev = eng.device
    buf = torch.zeros(si["feats"].size + 1, dtype=torch.float32, device=dev)
    buf[1:] = torch.from_numpy(si["feats"].reshape(-1)).to(dev)
    f = buf[1:].view(-1, 9)           # base pointer is 4-byte but not 16-byte aligned
    assert f.data_ptr() % 16 != 0
    o = torch.from_numpy(si["read_off"]).to(dev)
    k = torch.from_numpy(si["kmer_idx"]).to(dev)
    rp, sp, mc = eng.infer_device(f, o, k, 50, seed=1)
    torch.cuda.synchronize()
    assert np.array_equal(rp.cpu().numpy(), want[0]) and np.array_equal(sp.cpu().numpy(), want[1])


def test_host_pipeline_equals_device_call(synthetic_inputs):
    si = synthetic_inputs
    eng = engine("HCT116_RNA002")
    want = run_device(eng, si["feats"], si["read_off"], si["kmer_idx"], 300, seed=11, site_id_base=42)
    for n_chunks in (1, 3, 7):
        got = eng.infer_host(si["feats"], si["read_off"], si["kmer_idx"], 300, seed=11, site_id_base=42, n_chunks=n_chunks)
        for i in range(3):
            assert np.array_equal(got[i], want[i])


def test_invalid_arguments_are_rejected(synthetic_inputs):
    from m6anet_b200._cabi import M6AError
    from m6anet_b200 import weights as W
    from m6anet_b200.engine import MilEngine
    si = synthetic_inputs
    eng = engine("HCT116_RNA002")
    with pytest.raises(M6AError):
        run_device(eng, si["feats"], si["read_off"], si["kmer_idx"], 0)
    with pytest.raises(M6AError):
        run_device(eng, si["feats"], si["read_off"], si["kmer_idx"], 10, n_samples=65)
    w = product_weights("HCT116_RNA002")
    bad = W.EncoderWeights(w.emb, w.w1, w.b1, w.w2[:16], w.b2[:16], w.w3[:16], w.b3)
    with pytest.raises(M6AError):
        MilEngine(bad, "cuda:0")
    # the caller-owned workspace is validated: missing or too small -> M6A_EINVAL, nothing is launched
    import torch
    from m6anet_b200 import _cabi
    L = _cabi.lib()
    f = torch.zeros((40, 9), device="cuda:0")
    off = torch.tensor([0, 20, 40], dtype=torch.int64, device="cuda:0")
    k = torch.zeros((2, 3), dtype=torch.int32, device="cuda:0")
    rp, sp, mc = torch.empty(40, device="cuda:0"), torch.empty(2, device="cuda:0"), torch.empty(2, dtype=torch.int32, device="cuda:0")
    ws = torch.empty(4, dtype=torch.int64, device="cuda:0")
    args = lambda wsp, wsb: (eng._handle, f.data_ptr(), off.data_ptr(), k.data_ptr(), 2, 40, 0, 20, 10, 0, None, 0.5,
                             rp.data_ptr(), sp.data_ptr(), mc.data_ptr(), wsp, wsb, None)
    assert L.m6a_mil_workspace_bytes(40) == 24 and L.m6a_mil_workspace_bytes(10**6) >= (10**6 // 1000 + 2) * 8
    assert L.m6a_mil_infer_f32(*args(None, 0)) == -1
    assert L.m6a_mil_infer_f32(*args(ws.data_ptr(), 8)) == -1
    assert L.m6a_mil_infer_f32(*args(ws.data_ptr(), 32)) == 0
    torch.cuda.synchronize()


from oracle import read_probabilities_float64 as _read_probs_float64   # noqa: E402


def _full_size_check(tag, n_sites, n_reads, thr, seed=0, site_id_base=0, pooled_from=1):
    """Run a BASELINE-size config and check EVERY site against the C restatement of the oracle on the shared index
    stream, plus size-independent properties (determinism, closed-form expectation)."""
    from oracle import c_oracle
    eng = engine(tag)
    rng = np.random.default_rng(n_sites + n_reads)
    feats = rng.standard_normal((n_sites * n_reads, 9), dtype=np.float32)
    off = np.arange(n_sites + 1, dtype=np.int64) * n_reads
    kmer = rng.integers(0, 66, size=(n_sites, 3)).astype(np.int32)
    kw = dict(seed=seed, site_id_base=site_id_base, read_threshold=thr)
    rp, sp, mc = run_device(eng, feats, off, kmer, 1000, **kw)
    rp2, sp2, mc2 = run_device(eng, feats, off, kmer, 1000, **kw)
    assert np.array_equal(rp, rp2) and np.array_equal(sp, sp2) and np.array_equal(mc, mc2)      # idempotent
    orp, osp, omc = c_oracle.mil_inference(oracle_params(tag), feats, off, kmer, 1000, n_samples=20, **kw)
    # Over millions of reads the float32 evaluations (the reference's torch calls included) each sit up to ~1.9e-6 from
    # the exact value, so two of them can differ by more than READ_ATOL; the kernel is held to the float64 truth instead.
    truth = _read_probs_float64(oracle_params(tag), feats, np.repeat(kmer, n_reads, axis=0))
    assert np.max(np.abs(rp - truth)) <= 3e-6
    assert np.max(np.abs(rp - orp)) <= 5e-6
    assert np.max(np.abs(sp - osp)) <= SITE_ATOL                                               # every site
    near = np.abs(orp.astype(np.float64) - float(np.float32(thr))) < 1e-6
    slack = np.add.reduceat(near.astype(np.int64), off[:-1])
    assert np.all(np.abs(mc.astype(np.int64) - omc) <= slack)
    pm = rp.reshape(n_sites, n_reads).astype(np.float64)
    closed = 1.0 - (1.0 - pm.mean(axis=1)) ** 20
    assert np.max(np.abs(sp - closed)) < 4.5 * 0.5 / np.sqrt(1000)
    assert abs(float(np.mean(sp - closed))) < 2e-4            # no systematic bias from the device index stream
    return rp, sp, mc


def test_full_size_config2_every_site():
    """BASELINE config 2: 100k sites x 20 reads, HCT116_RNA002, 1000 iterations."""
    _full_size_check("HCT116_RNA002", 100_000, 20, 0.033379376)


def test_full_size_config4_every_site(_both_encoders):
    """BASELINE config 4 at its full size: HEK293T_RNA004 weights, 500k sites x 30 reads (the CUDA-core encoder, whose
    every-site check at this size would double the oracle's time, is held to 200k sites)."""
    n_sites = 500_000 if _both_encoders == "tc" else 200_000
    _full_size_check("HEK293T_RNA004", n_sites, 30, 0.033379376, seed=4, site_id_base=3_000_000_000)


def test_headline_size_config3_sampled_sites():
    """BASELINE config 3 (the job every bench number is quoted on): 1M sites x 50 reads, 1000 iterations -- 60 runs of 1000
    sites spread over the whole job (first, last, strided) against the C oracle, like bench.py's in-run parity check."""
    if _ENCODER != "tc":
        pytest.skip("the headline job is checked with the default encoder")
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(ASSETS), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    eng = engine("HCT116_RNA002")
    feats, off, kmer = bench.synth_shard(0, 1_000_000, 50, 1, total_sites=1_000_000)
    rp, sp, mc = eng.infer_host(feats, off, kmer, 1000, seed=0, read_threshold=0.033379376)

    class A:
        model, iters = "HCT116_RNA002", 1000
    par = bench.parity_check(A, 0, feats, off, kmer, rp, sp, mc, 0.033379376, 60_000)
    assert par["sites"] >= 50_000
    assert par["max_abs_site"] <= SITE_ATOL and par["max_abs_read"] <= 3e-6 and par["mod_count_mismatch"] == 0, par


def test_config5_replicate_pooling_shape():
    """BASELINE config 5 shape: arabidopsis_RNA002, sites pooled from 4 input directories x 20 reads = 80 reads/site."""
    _full_size_check("arabidopsis_RNA002", 50_000, 80, 0.0032978046219796, seed=5)
