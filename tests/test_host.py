"""CPU tests of the host-side mirror of the reference interface: registry, model-config plugin,
dataset ingest (bit-exact against the reference's NanopolishDS output), CSV row formats, sharding."""
import gzip
import os
import shutil

import numpy as np
import pytest

from conftest import ASSETS, GOLDEN


@pytest.fixture(scope="session")
def bundled_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("bundled")
    with gzip.open(os.path.join(GOLDEN, "bundled", "data.json.gz"), "rb") as f, open(d / "data.json", "wb") as g:
        shutil.copyfileobj(f, g)
    shutil.copyfile(os.path.join(GOLDEN, "bundled", "data.info"), d / "data.info")
    return str(d)


def test_registry_mirrors_reference():
    from m6anet_b200 import constants as C
    assert C.DEFAULT_PRETRAINED_MODELS == ['HCT116_RNA002', 'arabidopsis_RNA002', 'HEK293T_RNA004']
    assert set(C.PRETRAINED_CONFIGS) == set(C.DEFAULT_PRETRAINED_MODELS) | {'HEK293T_RNA004_M6ACE'}
    assert C.PRETRAINED_CONFIGS['arabidopsis_RNA002'][1] == 0.0032978046219796
    assert C.PRETRAINED_CONFIGS['HEK293T_RNA004'][1:] == (C.DEFAULT_READ_THRESHOLD, C.DEFAULT_NORM_PATH)
    for w, _, n in C.PRETRAINED_CONFIGS.values():
        assert os.path.exists(w) and os.path.exists(n)
    assert len(C.ALL_KMERS) == 66 and list(C.ALL_KMERS) == sorted(C.ALL_KMERS) and len(C.M6A_KMERS) == 18
    assert {C.KMER_TO_INT[k] for k in ("GGGAC", "GGACT", "GACTT")} == {45, 50, 51}


def test_model_config_plugin_and_state_dict_keys():
    from m6anet_b200 import constants as C
    from m6anet_b200.model import MILModel, load_model_config
    m = MILModel(load_model_config(C.DEFAULT_MODEL_CONFIG))
    assert (m.n_sig, m.emb_dim, m.n_kmer, m.h1, m.h2, m.bn1, m.bn2, m.n_reads_per_site) == (9, 2, 66, 150, 32, True, False, 20)
    keys = m.expected_keys()
    assert keys["read_level_encoder.1.embedding_layer.weight"] == (66, 2)
    assert keys["read_level_encoder.3.layers.1.running_var"] == (150,)
    assert keys["read_level_encoder.4.layers.0.weight"] == (32, 150) and "read_level_encoder.4.layers.1.weight" not in keys
    m.load_weights(C.DEFAULT_MODEL_WEIGHTS)
    w = m.encoder_weights()
    from m6anet_b200 import weights as W
    w2 = W.from_npz(C.DEFAULT_MODEL_WEIGHTS)
    for k in ("emb", "w1", "b1", "w2", "b2", "w3", "b3"):
        assert np.array_equal(getattr(w, k), getattr(w2, k))
    s = MILModel(load_model_config(os.path.join(ASSETS, "model_configs", "prod_pooling_signal.toml")))
    assert (s.emb_dim, s.in1) == (0, 9) and "read_level_encoder.2.layers.0.weight" in s.expected_keys()
    with pytest.raises(RuntimeError, match="missing keys"):
        s.load_state_dict({})


def test_unsupported_blocks_are_rejected_loudly():
    from m6anet_b200.model import MILModel
    base = [{"block_type": "DeaggregateNanopolish", "num_neighboring_features": 1}, {"block_type": "ExtractSignal"},
            {"block_type": "Linear", "input_channel": 9, "output_channel": 150, "activation": "relu", "batch_norm": True},
            {"block_type": "Linear", "input_channel": 150, "output_channel": 32, "activation": "relu", "batch_norm": False}]
    with pytest.raises(NotImplementedError, match="probability_layer"):
        MILModel({"block": base + [{"block_type": "Attention", "input_channel": 32, "hidden_layers_1": 8}]})
    with pytest.raises(AttributeError):
        MILModel({"block": base + [{"block_type": "NoSuchBlock"}]})
    with pytest.raises(NotImplementedError):
        MILModel({"block": base})
    ok = MILModel({"block": base + [{"block_type": "SigmoidMeanPooling", "input_channel": 32}]})
    assert ok.pooling_filter.block_type == "SigmoidMeanPooling"


def test_dataset_matches_reference_ingest_bit_for_bit(bundled_dir, bundled_flat):
    """features/k-mer ids/read ids equal what the reference's NanopolishDS + inference_collate produced."""
    from m6anet_b200 import constants as C
    from m6anet_b200.data import NanopolishDS, inference_collate
    ds = NanopolishDS(bundled_dir, 20, C.DEFAULT_NORM_PATH, mode='Inference')
    assert len(ds) == 101 and ds.total_neighboring_features == 1 and list(ds.indices) == list(range(9))
    b = ds.load_sites(0, len(ds))
    g = bundled_flat
    assert np.array_equal(b.read_off, g["read_off"]) and np.array_equal(b.kmer_idx, g["kmer_idx"])
    assert np.array_equal(b.feats, g["feats"])                      # float64 normalisation, one rounding: identical bits
    assert np.array_equal(b.read_ids, g["read_id"].astype(np.int64))
    assert np.array_equal(b.tx_ids, g["tx_id"]) and np.array_equal(b.tx_pos, g["tx_pos"])
    assert set(b.kmers) == {"GGACT"}
    feats, kmers, n_reads, tx, pos, rid = inference_collate([ds[0], ds[1]])
    assert feats.shape == (662 + 441, 9) and kmers.shape == (662 + 441, 3) and list(n_reads) == [662, 441]
    assert np.array_equal(kmers[0], g["kmer_idx"][0]) and tx[0] == "ENST00000393394.5" and pos[-1] == 234
    # a reference .joblib-free path: norm factors load from npz with the same values
    assert len(ds.norm_dict) == 1024 and ds.norm_dict["AAACG"][0].dtype == np.float64
    ds.close()


def test_replicate_dataset_pools_sites(bundled_dir, tmp_path, bundled_flat):
    from m6anet_b200 import constants as C
    from m6anet_b200.data import NanopolishDS, NanopolishReplicateDS
    rep = tmp_path / "rep1"
    shutil.copytree(bundled_dir, rep)
    # second replicate lacks the first site and has a private one
    lines = open(rep / "data.info").read().splitlines()
    with open(rep / "data.info", "w") as f:
        f.write("\n".join([lines[0]] + lines[2:]) + "\n")
    single = NanopolishDS(bundled_dir, 20, C.DEFAULT_NORM_PATH)
    ds = NanopolishReplicateDS([bundled_dir, str(rep)], 20, C.DEFAULT_NORM_PATH)
    # sites with >= 20 pooled reads: all 248 keys of dir 0, pooled counts double except the first
    import pandas as pd
    info = pd.read_csv(os.path.join(bundled_dir, "data.info"))
    pooled = info["n_reads"].to_numpy() * 2
    pooled[0] = info["n_reads"][0]
    assert len(ds) == int((pooled >= 20).sum())
    assert ds.n_reads[0] == 662 and ds.n_reads[1] == 882
    b = ds.load_sites(0, 2)
    assert list(np.diff(b.read_off)) == [662, 882]
    assert b.read_rep[0] == 0 and b.read_rep[662 + 441] == 1 and b.read_ids[662 + 441] == b.read_ids[662]
    s1 = single.load_sites(1, 2)
    assert np.array_equal(b.feats[662:662 + 441], s1.feats) and np.array_equal(b.feats[662 + 441:], s1.feats)
    # reference-shaped item: "{read_id}_{replicate}" strings, same features as the flat path
    feats, kmers, tx, pos, rid = ds[1]
    assert np.array_equal(feats, b.feats[662:]) and rid[0] == f"{b.read_ids[662]}_0" and rid[441] == f"{b.read_ids[662]}_1"


def test_csv_row_formats_match_reference(bundled_flat):
    """'%s,%d,%s,%.16f,%s,%.16f' and '%s,%d,%s,%.16f' (reference utils/inference_utils.py:59-67)."""
    from m6anet_b200.data import SiteBatch
    from m6anet_b200.inference import write_indiv_rows, write_site_rows
    batch = SiteBatch(np.zeros((3, 9), np.float32), np.array([0, 2, 3]), np.zeros((2, 3), np.int32), np.array([966210, 7, 8]),
                      np.array(["ENST1.5", "ENST2"]), np.array([130, 7]), np.array(["GGACT", "AAACA"]))
    import tempfile
    with tempfile.TemporaryFile("w+b") as f, tempfile.TemporaryFile("w+b") as g, tempfile.TemporaryFile("w+b") as h:
        write_site_rows(f, batch, np.array([0.25, np.float32(0.1)], np.float32), np.array([1, 0], np.int32))
        write_indiv_rows(g, batch, np.array([0.5, 0.125, np.float32(1e-20)], np.float32))
        batch.read_rep = np.array([0, 1, 1], np.int32)
        write_indiv_rows(h, batch, np.array([0.5, 0.125, 1.0], np.float32))
        f.seek(0), g.seek(0), h.seek(0)
        fl, gl, hl = f.read().decode().splitlines(), g.read().decode().splitlines(), h.read().decode().splitlines()
    assert fl[0] == "ENST1.5,130,2,0.2500000000000000,GGACT,0.5000000000000000"
    assert fl[1] == "ENST2,7,1,%.16f,AAACA,0.0000000000000000" % float(np.float32(0.1))
    assert gl == ["ENST1.5,130,966210,0.5000000000000000", "ENST1.5,130,7,0.1250000000000000", "ENST2,7,8,0.0000000000000000"]
    assert hl == ["ENST1.5,130,966210_0,0.5000000000000000", "ENST1.5,130,7_1,0.1250000000000000", "ENST2,7,8_1,1.0000000000000000"]
    # many rows, random values: identical to Python's own '%' formatting (what the reference executes)
    rng = np.random.default_rng(0)
    S = 2000
    n = rng.integers(1, 40, S)
    off = np.concatenate([[0], np.cumsum(n)])
    big = SiteBatch(np.zeros((off[-1], 9), np.float32), off, np.zeros((S, 3), np.int32), rng.integers(0, 10**7, off[-1]),
                    np.array([f"ENST{rng.integers(10**10)}.{i % 9}" for i in range(S)]), rng.integers(0, 10**5, S),
                    rng.choice(["GGACT", "AAACA", "TGACC"], S))
    rp = np.exp(rng.uniform(-50, 0, off[-1])).astype(np.float32)
    sp = rng.random(S).astype(np.float32)
    mc = np.minimum(rng.integers(0, 40, S), n).astype(np.int32)
    with tempfile.TemporaryFile("w+b") as f, tempfile.TemporaryFile("w+b") as g:
        write_site_rows(f, big, sp, mc, n_threads=4)
        write_indiv_rows(g, big, rp, n_threads=4)
        f.seek(0), g.seek(0)
        fl, gl = f.read().decode(), g.read().decode()
    want_f = "".join('%s,%d,%s,%.16f,%s,%.16f\n' % (t, p, k, float(a), km, m / k) for t, p, k, a, km, m in
                     zip(big.tx_ids, big.tx_pos, n, sp, big.kmers, mc.astype(np.float64)))
    want_g = "".join('%s,%d,%s,%.16f\n' % (t, p, r, float(a)) for t, p, r, a in
                     zip(np.repeat(big.tx_ids, n), np.repeat(big.tx_pos, n), big.read_ids, rp))
    assert fl == want_f and gl == want_g


def test_csv_probability_formatting_is_exact():
    """The writers format '%.16f' with an exact 128-bit integer path for values in [0, 1] (round-half-even on the binary
    value, like printf / Python) and fall back to printf otherwise.  Checked against Python's '%' on the cases that break
    approximate formatters: exact ties at the 17th decimal (odd m / 2^17), powers of two, float32 denormals, values just
    below 1, every ratio count / n, NaN, negative and > 1 values, negative positions and ids."""
    import tempfile
    from m6anet_b200.data import SiteBatch
    from m6anet_b200.inference import write_indiv_rows, write_site_rows
    rng = np.random.default_rng(4)
    ties = (np.arange(1, 4001, 2, dtype=np.float64) / 2.0 ** 17).astype(np.float32)               # x * 1e16 ends in .5 exactly
    pow2 = (2.0 ** -np.arange(0, 150, dtype=np.float64)).astype(np.float32)                       # down into the denormals
    near1 = np.nextafter(np.float32(1), np.float32(0)) - np.arange(50, dtype=np.float32) * np.float32(6e-8)
    bits = rng.integers(0, 0x3F800000, 20000, dtype=np.int64).astype(np.uint32).view(np.float32)  # uniform over bit patterns
    special = np.array([0.0, 1.0, np.nan, -0.25, 1.5, np.inf, -np.inf, 1e-45, 3.4e38, -0.0], dtype=np.float32)
    rp = np.concatenate([ties, pow2, near1, bits, special]).astype(np.float32)
    R = len(rp)
    n = np.full(R // 10, 10)
    n[-1] += R - n.sum()
    S = len(n)
    off = np.concatenate([[0], np.cumsum(n)])
    ids = rng.integers(-10**6, 10**12, R)
    batch = SiteBatch(np.zeros((R, 9), np.float32), off, np.zeros((S, 3), np.int32), ids,
                      np.array([f"tx{i}" for i in range(S)]), rng.integers(-5, 10**9, S), np.array(["GGACT"] * S))
    with tempfile.TemporaryFile("w+b") as g:
        write_indiv_rows(g, batch, rp, n_threads=3)
        g.seek(0)
        got = g.read().decode()
    want = "".join('%s,%d,%s,%.16f\n' % (t, p, r, float(a)) for t, p, r, a in
                   zip(np.repeat(batch.tx_ids, n), np.repeat(batch.tx_pos, n), ids, rp))
    assert got == want
    # mod_ratio: every count / n up to 120 reads (doubles with full 53-bit mantissas), site probabilities incl. NaN
    pairs = [(c, k) for k in range(1, 121) for c in range(k + 1)]
    S = len(pairs)
    k_arr = np.array([k for _, k in pairs])
    off = np.concatenate([[0], np.cumsum(k_arr)])
    sp = np.resize(np.concatenate([rp[:5000], special]), S).astype(np.float32)
    batch = SiteBatch(np.zeros((off[-1], 9), np.float32), off, np.zeros((S, 3), np.int32), np.zeros(off[-1], np.int64),
                      np.array([f"tx{i}" for i in range(S)]), np.arange(S), np.array(["GGACT"] * S))
    with tempfile.TemporaryFile("w+b") as f:
        write_site_rows(f, batch, sp, np.array([c for c, _ in pairs], np.int32), n_threads=2)
        f.seek(0)
        got = f.read().decode()
    want = "".join('%s,%d,%s,%.16f,%s,%.16f\n' % (f"tx{i}", i, k, float(a), "GGACT", c / k)
                   for i, ((c, k), a) in enumerate(zip(pairs, sp)))
    assert got == want


def test_argparser_keeps_reference_flags():
    from m6anet_b200 import constants as C
    from m6anet_b200.inference import argparser
    a = argparser().parse_args(["--input_dir", "x", "y", "--out_dir", "o"])
    assert a.input_dir == ["x", "y"] and a.pretrained_model == "HCT116_RNA002" and a.num_iterations == 1000
    assert (a.batch_size, a.save_per_batch, a.n_processes, a.seed) == (16, 2, 25, 0)
    assert a.read_proba_threshold == C.DEFAULT_READ_THRESHOLD and a.model_state_dict is None and a.device == "cuda"


def test_main_rejects_unknown_pretrained_model_and_cpu_device(bundled_dir, tmp_path):
    from m6anet_b200 import inference
    a = inference.argparser().parse_args(["--input_dir", bundled_dir, "--out_dir", str(tmp_path), "--pretrained_model", "nope"])
    with pytest.raises(ValueError, match="Invalid pretrained model"):
        inference.main(a)
    a = inference.argparser().parse_args(["--input_dir", bundled_dir, "--out_dir", str(tmp_path), "--device", "cpu"])
    with pytest.raises(RuntimeError, match="no CPU"):
        inference.main(a)


def test_shard_bounds_balance_reads():
    from m6anet_b200.dist import shard_bounds
    from m6anet_b200.inference import plan_batches
    rng = np.random.default_rng(0)
    n = rng.integers(20, 1000, size=5000)
    for w in (1, 2, 3, 8):
        b = shard_bounds(n, w)
        assert b[0] == 0 and b[-1] == 5000 and len(b) == w + 1 and all(x <= y for x, y in zip(b[:-1], b[1:]))
        tot = [n[b[i]:b[i + 1]].sum() for i in range(w)]
        assert max(tot) - min(tot) <= 2 * n.max()
    assert shard_bounds(np.array([], dtype=np.int64), 4) == [0, 0, 0, 0, 0]
    assert shard_bounds(np.full(8, 50), 8) == list(range(9))
    spans = plan_batches(n, 100, 4000, 50_000)
    assert spans[0][0] == 100 and spans[-1][1] == 4000 and all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
    assert all(n[a:b].sum() < 50_000 + 1000 for a, b in spans)


def test_native_ingest_equals_python_path_and_handles_number_formats(tmp_path):
    """m6a_ingest_parts vs the per-site Python/json path on a hand-written data.json: exponents, negatives,
    integers, 9-mer context (two flanks in the file, one wanted), every DRACH k-mer id."""
    import json
    from m6anet_b200 import constants as C
    from m6anet_b200.data import NanopolishDS
    rng = np.random.default_rng(5)
    lines, info, off = [], ["transcript_id,transcript_position,start,end,n_reads"], 0
    seven = ["GGGACTT", "AAAACAA", "TTGACTG", "CAGACCA"]
    fmt = [lambda v: repr(float(v)), lambda v: "%.3e" % v, lambda v: "%.12f" % v, lambda v: str(int(v * 100))]
    for i, k in enumerate(seven * 3):
        n = int(rng.integers(20, 60))
        rows = []
        for r in range(n):
            vals = rng.normal([0.01, 3, 110] * 3, [0.005, 1.5, 8] * 3)
            vals[0] = -abs(vals[0]) if r == 0 else vals[0]
            txt = ",".join(fmt[(r + c) % 4](v) for c, v in enumerate(vals)) + "," + repr(float(rng.integers(1, 10**6)))
            rows.append("[" + txt + "]")
        line = '{"tx%d":{"%d":{"%s":[%s]}}}\n' % (i, 100 + i, k, ",".join(rows))
        json.loads(line)
        info.append(f"tx{i},{100 + i},{off},{off + len(line)},{n}")
        off += len(line)
        lines.append(line)
    (tmp_path / "data.json").write_text("".join(lines))
    (tmp_path / "data.info").write_text("\n".join(info) + "\n")
    ds = NanopolishDS(str(tmp_path), 20, C.DEFAULT_NORM_PATH)
    flat = ds.load_sites(0, len(ds), n_threads=3)
    for i in range(len(ds)):
        feats, kmers, tx, pos, rid = ds[i]                       # Python json path (reference-shaped)
        sl = slice(flat.read_off[i], flat.read_off[i + 1])
        assert np.array_equal(flat.feats[sl], feats) and np.array_equal(flat.read_ids[sl], rid)
        assert np.array_equal(flat.kmer_idx[i], kmers[0]) and flat.kmers[i] == seven[i % 4][1:6]
    # partial ranges and an empty range
    part = ds.load_sites(3, 7)
    assert np.array_equal(part.feats, flat.feats[flat.read_off[3]:flat.read_off[7]]) and part.n_sites == 4
    assert ds.load_sites(5, 5).n_sites == 0


def test_native_ingest_reports_bad_input(tmp_path):
    from m6anet_b200 import constants as C
    from m6anet_b200._cabi import M6AError
    from m6anet_b200.data import NanopolishDS
    row = "[" + ",".join(["1.0"] * 9) + ",7.0]"
    line = '{"t":{"5":{"GGGACTT":[%s]}}}\n' % ",".join([row] * 20)
    (tmp_path / "data.json").write_text(line)
    (tmp_path / "data.info").write_text(f"transcript_id,transcript_position,start,end,n_reads\nt,5,0,{len(line)},21\n")
    ds = NanopolishDS(str(tmp_path), 20, C.DEFAULT_NORM_PATH)
    with pytest.raises(M6AError, match="t:5"):                   # data.info says 21 reads, the line has 20
        ds.load_sites(0, 1)
    (tmp_path / "data.info").write_text(f"transcript_id,transcript_position,start,end,n_reads\nt,5,0,{len(line)},20\n")
    assert NanopolishDS(str(tmp_path), 20, C.DEFAULT_NORM_PATH).load_sites(0, 1).feats.shape == (20, 9)
    with pytest.raises(M6AError):                                # arabidopsis norm factors lack most 5-mers of HCT116 data? use a non-DRACH k-mer
        bad = line.replace("GGGACTT", "GGGTCTT")
        (tmp_path / "data.json").write_text(bad)
        NanopolishDS(str(tmp_path), 20, C.DEFAULT_NORM_PATH).load_sites(0, 1)


def test_replicate_join_matches_reference_order_and_content(bundled_dir, tmp_path):
    """Site order, pooled read counts, features and "{id}_{rep}" read ids equal what the reference's
    NanopolishReplicateDS produced for two directories with missing and reordered rows (fixture made by the reference)."""
    import importlib.util
    from m6anet_b200 import constants as C
    from m6anet_b200.data import NanopolishReplicateDS
    spec = importlib.util.spec_from_file_location("make_golden_dirs", os.path.join(GOLDEN, "make_golden.py"))
    src = open(spec.origin).read()
    ns = {"os": os, "shutil": shutil}
    start = src.index("def make_replicate_dirs")
    exec(src[start:src.index("def replicates()")], ns)          # only the pure directory builder, no reference import
    dirs = ns["make_replicate_dirs"](bundled_dir, str(tmp_path))
    g = np.load(os.path.join(GOLDEN, "replicate_golden.npz"))
    ds = NanopolishReplicateDS(dirs, 20, C.DEFAULT_NORM_PATH)
    assert list(ds._tx) == list(g["tx_id"]) and np.array_equal(ds._pos, g["tx_pos"]) and np.array_equal(ds.n_reads, g["n_reads"])
    flat = ds.load_sites(0, len(ds))
    for j, i in enumerate(g["pick"]):
        feats, kmers, tx, pos, rid = ds[int(i)]
        assert np.array_equal(feats, g[f"feats_{j}"]) and np.array_equal(kmers[0], g[f"kmer_{j}"])
        assert [str(r) for r in rid] == list(g[f"read_id_{j}"])
        sl = slice(flat.read_off[i], flat.read_off[i + 1])
        assert np.array_equal(flat.feats[sl], g[f"feats_{j}"])
        assert [f"{a}_{b}" for a, b in zip(flat.read_ids[sl], flat.read_rep[sl])] == list(g[f"read_id_{j}"])


def test_native_parser_number_formats_property(tmp_path):
    """Property test: for arbitrary finite doubles written in any of Python's float spellings, the native parser
    (std::from_chars) yields the same float32 features as json.loads + float64 normalisation."""
    import json
    from hypothesis import HealthCheck, given, settings, strategies as st
    from m6anet_b200 import constants as C
    from m6anet_b200.data import NanopolishDS

    spell = [repr, lambda v: "%.17g" % v, lambda v: "%.6e" % v, lambda v: "%.3f" % v, lambda v: "%d" % int(v) if abs(v) < 1e15 else repr(v)]
    finite = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False) | \
        st.floats(min_value=-1e-300, max_value=1e-300, allow_nan=False) | st.sampled_from([0.0, -0.0, 5e-324, 1e22, 123456789.125])

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(st.tuples(st.lists(finite, min_size=9, max_size=9), st.lists(st.integers(0, 4), min_size=9, max_size=9),
                              st.integers(0, 10**9)), min_size=20, max_size=24))
    def check(rows):
        txt = []
        for vals, how, rid in rows:
            txt.append("[" + ",".join(spell[h](v) for v, h in zip(vals, how)) + "," + repr(float(rid)) + "]")
        line = '{"tx":{"42":{"AGGACTG":[%s]}}}\n' % ",".join(txt)
        parsed = json.loads(line)["tx"]["42"]["AGGACTG"]
        (tmp_path / "data.json").write_text(line)
        (tmp_path / "data.info").write_text(f"transcript_id,transcript_position,start,end,n_reads\ntx,42,0,{len(line)},{len(rows)}\n")
        ds = NanopolishDS(str(tmp_path), 20, C.DEFAULT_NORM_PATH)
        flat = ds.load_sites(0, 1, n_threads=1)
        raw = np.array(parsed, dtype=np.float64)
        mean, std = ds.get_norm_factor(["AGGAC", "GGACT", "GACTG"])
        want = ((raw[:, :9] - mean) / std).astype(np.float32)
        assert np.array_equal(flat.feats, want, equal_nan=True)
        assert np.array_equal(flat.read_ids, raw[:, 9].astype(np.int64))
        ds.close()

    check()


def test_native_info_reader_matches_pandas(bundled_dir, tmp_path):
    import pandas as pd
    from m6anet_b200._cabi import M6AError
    from m6anet_b200.data import read_info
    want = pd.read_csv(os.path.join(bundled_dir, "data.info"))
    tx, pos, start, end, n = read_info(bundled_dir)
    assert list(tx.astype(str)) == list(want["transcript_id"]) and np.array_equal(pos, want["transcript_position"])
    assert np.array_equal(start, want["start"]) and np.array_equal(end, want["end"]) and np.array_equal(n, want["n_reads"])
    # column order from the header, extra columns, CRLF, float-formatted integers, no trailing newline
    d = tmp_path / "odd"
    d.mkdir()
    (d / "data.info").write_text("n_reads,extra,end,transcript_id,start,transcript_position\r\n"
                                 "25,x,120,ENST0001.1,0,17\r\n20.0,y,260,T2,120,3\r\n\r\n7,z,300,a_very_long_transcript_name.12,260,99")
    tx, pos, start, end, n = read_info(str(d))
    assert list(tx.astype(str)) == ["ENST0001.1", "T2", "a_very_long_transcript_name.12"]
    assert list(pos) == [17, 3, 99] and list(start) == [0, 120, 260] and list(end) == [120, 260, 300] and list(n) == [25, 20, 7]
    (d / "data.info").write_text("transcript_id,transcript_position,start,end\nt,1,0,5\n")
    with pytest.raises(M6AError):
        read_info(str(d))
    (d / "data.info").write_text("transcript_id,transcript_position,start,end,n_reads\n")
    assert len(read_info(str(d))[0]) == 0
    with pytest.raises(M6AError):
        read_info(str(tmp_path / "missing"))


def test_native_number_parser_is_correctly_rounded_and_strict(tmp_path):
    """m6a_ingest_parts parses numbers either on a Clinger fast path (mantissa <= 2^53, <= 22 fraction digits: one exact
    division) or with std::from_chars.  Both must give Python's float() to the last bit: the norm tables are set to
    mean = float(text), std = ulp, so any difference shows up as a non-zero feature.  Malformed numbers and malformed
    lines (anything json.loads rejects) must be refused with M6A_EPARSE."""
    import ctypes as C
    import json
    from m6anet_b200 import _cabi
    L = _cabi.lib()
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    enc = lambda k: sum(code[ch] * 4 ** (4 - i) for i, ch in enumerate(k))
    codes = [enc(k) for k in ("AGGAC", "GGACT", "GACTG")]
    kid = np.full(1024, -1, dtype=np.int32)
    kid[codes] = [0, 1, 2]
    path = tmp_path / "n.json"
    paths = (C.c_char_p * 1)(os.fsencode(str(path)))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def ingest(line: str, n_rows=1, mean=None, std=None):
        path.write_text(line)
        parts = np.zeros(1, dtype=_cabi.PART_DTYPE)
        parts[0] = (0, 0, 0, len(line.encode()), 0, n_rows, 0, 1, 0)
        m = np.zeros((1024, 3)) if mean is None else mean
        s = np.ones((1024, 3)) if std is None else std
        feats = np.full((max(n_rows, 1), 9), 99.0, dtype=np.float32)
        ids = np.zeros(max(n_rows, 1), dtype=np.int64)
        kmer = np.zeros((1, 3), dtype=np.int32)
        bad = C.c_int64(-1)
        rc = L.m6a_ingest_parts(paths, 1, vp(parts), 1, 1, vp(np.ascontiguousarray(m)), vp(np.ascontiguousarray(s)), vp(kid),
                                vp(feats), vp(ids), vp(kmer), 1, C.byref(bad))
        return rc, feats, ids

    rng = np.random.default_rng(11)

    def random_decimal():
        kind = rng.integers(0, 6)
        if kind == 0:        # fast path: short mantissa, a few fraction digits (what dataprep writes after rounding)
            return "%d.%0*d" % (rng.integers(0, 10**4), int(rng.integers(1, 8)), rng.integers(0, 10**6) % 10 ** int(rng.integers(1, 7)))
        if kind == 1:        # repr of a double: 15-17 significant digits, both sides of the 2^53 mantissa limit
            return repr(float(rng.normal(0, 1) * 10.0 ** int(rng.integers(-8, 8))))
        if kind == 2:        # mantissa around 2^53 with a random number of fraction digits
            w = (1 << 53) + int(rng.integers(-3, 4))
            t = str(w)
            k = int(rng.integers(0, 16))
            return t[:len(t) - k] + "." + t[len(t) - k:] if k else t
        if kind == 3:        # many leading zeros in the fraction (up to and beyond 22 fraction digits)
            return "0." + "0" * int(rng.integers(0, 24)) + str(rng.integers(1, 10**5))
        if kind == 4:        # exponents
            return "%.*e" % (int(rng.integers(0, 17)), rng.normal(0, 1) * 10.0 ** int(rng.integers(-30, 30)))
        return str(int(rng.integers(-10**15, 10**15)))            # integers

    for _ in range(300):
        texts = [("-" if rng.random() < 0.3 and not t.startswith("-") else "") + t for t in (random_decimal() for _ in range(9))]
        want = np.array([float(t) for t in texts])
        mean = np.zeros((1024, 3))
        std = np.ones((1024, 3))
        for j, c in enumerate(codes):
            mean[c] = want[3 * j:3 * j + 3]
            std[c] = np.maximum(np.spacing(np.abs(want[3 * j:3 * j + 3])), 5e-324)
        line = '{"t":{"1":{"AGGACTG":[[%s,12345.0]]}}}\n' % ",".join(texts)
        assert json.loads(line)
        rc, feats, ids = ingest(line, 1, mean, std)
        assert rc == 0, (rc, line)
        assert np.array_equal(feats[0], np.zeros(9, dtype=np.float32)), (texts, feats[0])    # 0 ulp from float(text)
        assert ids[0] == 12345

    good_row = "[" + ",".join(["1.5"] * 9) + ",7.0]"
    ok = '{"t":{"1":{"AGGACTG":[%s]}}}\n' % good_row
    assert ingest(ok)[0] == 0
    assert ingest(' { "t" : { "1" : { "AGGACTG" : [ [ %s , 7.0 ] ] } } } \n' % " , ".join(["1.5"] * 9))[0] == 0   # JSON whitespace
    for bad_number in ("01.5", "1.", ".5", "+1.5", "1.5.2", "1e", "1e+", "0x10", "nan", "inf", "1,5", "--1", "1e400", ""):
        line = ok.replace("1.5", bad_number, 1)
        with pytest.raises(ValueError):
            v = json.loads(line)                 # json.loads rejects all of these, or (1e400) gives inf ...
            if bad_number in ("1e400", "1,5"):   # ... or a different row width: refused as well
                raise ValueError
        assert ingest(line)[0] == -6, bad_number
    for broken in (ok[:-2] + "\n",                       # a closing brace missing
                   ok.rstrip("\n")[:-3],                # truncated after the list
                   ok.replace("]}}}", "}}}"),           # outer list not closed
                   ok.replace("[[", "["),               # rows not nested
                   ok.replace('"AGGACTG":', '"AGGACTG"'),
                   ok.replace("7.0]", "7.0"),           # row not closed (cut inside the byte range)
                   ok.rstrip("\n") + "x\n",            # trailing garbage inside the byte range
                   ok.replace("1.5,", "1.5,,", 1),
                   ok.replace('"t"', '"t\\"x"')):       # escapes never occur in transcript ids
        assert ingest(broken)[0] == -6, broken
    assert ingest(ok, n_rows=2)[0] == -6 and ingest(ok, n_rows=0)[0] == -6           # row count must match data.info
    # transcript ids: any valid UTF-8 without escapes, exactly when json.loads can decode the line
    pool = [0x41, 0x7A, 0xC2, 0xA9, 0xE2, 0x82, 0xAC, 0xF0, 0x9F, 0x98, 0x80, 0xED, 0xA0, 0x80, 0xC0, 0xAF, 0xFF, 0xE0, 0x9F, 0xF4, 0x90]
    for _ in range(400):
        name = bytes(rng.choice(pool, int(rng.integers(1, 6))).tolist())
        line_b = b'{"' + name + b'":{"1":{"AGGACTG":[' + good_row.encode() + b']}}}\n'
        try:
            json.loads(line_b)
            valid = True
        except ValueError:
            valid = False
        path.write_bytes(line_b)
        parts = np.zeros(1, dtype=_cabi.PART_DTYPE)
        parts[0] = (0, 0, 0, len(line_b), 0, 1, 0, 1, 0)
        f_, i_, k_ = np.zeros((1, 9), np.float32), np.zeros(1, np.int64), np.zeros((1, 3), np.int32)
        z_, o_ = np.zeros((1024, 3)), np.ones((1024, 3))
        rc = L.m6a_ingest_parts(paths, 1, vp(parts), 1, 1, vp(z_), vp(o_), vp(kid), vp(f_), vp(i_), vp(k_), 1, None)
        assert (rc == 0) == valid, (name, rc, valid)
    # byte range past the end of the file: data.info does not belong to this data.json
    path.write_text(ok)
    parts = np.zeros(1, dtype=_cabi.PART_DTYPE)
    parts[0] = (0, 0, 0, len(ok) + 10, 0, 1, 0, 1, 0)
    feats, ids, kmer = np.zeros((1, 9), np.float32), np.zeros(1, np.int64), np.zeros((1, 3), np.int32)
    bad = C.c_int64(-1)
    z, o = np.zeros((1024, 3)), np.ones((1024, 3))
    assert L.m6a_ingest_parts(paths, 1, vp(parts), 1, 1, vp(z), vp(o), vp(kid), vp(feats), vp(ids), vp(kmer), 1, C.byref(bad)) == -6
    assert bad.value == 0


def test_ingest_fuzz_accepts_only_what_json_accepts():
    """tools/fuzz_ingest.py: corrupted site lines / part tables are either refused with a status or parsed to exactly
    what json.loads gives; never a crash or a write outside the output window (guard bands)."""
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("fuzz_ingest", os.path.join(ROOT, "tools", "fuzz_ingest.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    stats = mod.run(250, seed=3, verbose=False)
    assert stats["ok"] + stats["rejected"] == 250 and stats["rejected"] > 100 and stats["ok"] > 0


def test_csv_writers_on_ingested_batches_equal_python_formatting(bundled_dir):
    """The path run_inference takes: load_sites() batches (transcript ids as fixed-width bytes) -> native writers."""
    import tempfile
    from m6anet_b200 import constants as C
    from m6anet_b200.data import NanopolishDS
    from m6anet_b200.inference import write_indiv_rows, write_site_rows
    ds = NanopolishDS(bundled_dir, 20, C.DEFAULT_NORM_PATH)
    b = ds.load_sites(3, 60, n_threads=2)
    assert b.tx_bytes is not None and b.tx_bytes.dtype.kind == "S" and list(b.tx_bytes.astype(str)) == list(b.tx_ids)
    rng = np.random.default_rng(0)
    n = np.diff(b.read_off)
    rp, sp = rng.random(int(n.sum())).astype(np.float32), rng.random(b.n_sites).astype(np.float32)
    mc = np.minimum(rng.integers(0, 700, b.n_sites), n).astype(np.int32)
    with tempfile.TemporaryFile() as f, tempfile.TemporaryFile() as g:
        write_site_rows(f, b, sp, mc, 3)
        write_indiv_rows(g, b, rp, 3)
        f.seek(0), g.seek(0)
        fl, gl = f.read().decode(), g.read().decode()
    assert fl == "".join('%s,%d,%s,%.16f,%s,%.16f\n' % (t, p, k, float(a), km, m / k) for t, p, k, a, km, m in
                         zip(b.tx_ids, b.tx_pos, n, sp, b.kmers, mc.astype(np.float64)))
    assert gl == "".join('%s,%d,%s,%.16f\n' % (t, p, r, float(a)) for t, p, r, a in
                         zip(np.repeat(b.tx_ids, n), np.repeat(b.tx_pos, n), b.read_ids, rp))


def test_bind_host_to_device_is_best_effort(monkeypatch):
    """No NVML / implausible answers change nothing; a proper GPU-local CPU mask is applied (NVML bit words -> CPU list,
    CUDA_VISIBLE_DEVICES remapping honoured)."""
    import sys
    import types
    from m6anet_b200 import dist as D
    before = os.sched_getaffinity(0)
    monkeypatch.setitem(sys.modules, "pynvml", None)            # import fails
    assert D.bind_host_to_device(0) == [] and os.sched_getaffinity(0) == before
    allowed = sorted(before)
    if len(allowed) < 8:
        pytest.skip("needs >= 8 CPUs to exercise the binding")
    local = allowed[: len(allowed) // 2]
    seen = {}
    fake = types.SimpleNamespace(
        nvmlInit=lambda: None, nvmlShutdown=lambda: None,
        nvmlDeviceGetHandleByIndex=lambda i: seen.setdefault("index", i),
        nvmlDeviceGetCpuAffinity=lambda h, n: [sum(1 << (c - 64 * w) for c in local if c // 64 == w) for w in range(n)])
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    applied = []
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: applied.append(list(cpus)))
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,5")
    assert D.bind_host_to_device(1) == local and seen["index"] == 5 and applied == [local]
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "GPU-deadbeef")
    assert D.bind_host_to_device(0) == []                        # UUID form: not resolved, nothing changed
    fake.nvmlDeviceGetCpuAffinity = lambda h, n: [sum(1 << (c % 64) for c in allowed if c // 64 == w) for w in range(n)]
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES")
    assert D.bind_host_to_device(0) == []                        # every CPU is local: single socket, nothing to do


def test_append_file_concatenates_shards(tmp_path):
    """Rank 0 concatenates the per-rank CSV shards behind the headers: in-kernel copy when the descriptor allows it
    ('r+b'), buffered copy for an O_APPEND descriptor; later writes land behind the appended bytes either way."""
    from m6anet_b200.inference import append_file
    blob = os.urandom(3_000_001)
    (tmp_path / "shard").write_bytes(blob)
    (tmp_path / "empty").write_bytes(b"")
    for mode in ("ab", "r+b"):
        out_path = tmp_path / f"out_{mode}"
        out_path.write_bytes(b"header\n")
        with open(out_path, mode) as out:
            out.seek(0, os.SEEK_END)
            append_file(out, str(tmp_path / "shard"))
            append_file(out, str(tmp_path / "empty"))
            append_file(out, str(tmp_path / "shard"))
            out.write(b"tail")
        assert out_path.read_bytes() == b"header\n" + blob + blob + b"tail"


def test_m6anet_console_entry():
    """`m6anet inference ...` of the reference (m6anet/__init__.py:11-30, setup.py:47-54) is served by bin/m6anet and by the
    `m6anet` script of pyproject.toml; the sub-commands outside the hot path say so instead of failing obscurely."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "bin", "m6anet")
    r = subprocess.run([sys.executable, exe, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and "m6anet-b200" in r.stdout
    r = subprocess.run([sys.executable, exe, "inference", "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--pretrained_model" in r.stdout and "--num_iterations" in r.stdout
    r = subprocess.run([sys.executable, exe, "dataprep"], capture_output=True, text=True)
    assert r.returncode != 0 and "outside the m6anet_b200 scope" in r.stderr
    text = open(os.path.join(root, "pyproject.toml")).read()
    assert 'm6anet = "m6anet_b200.__main__:main"' in text


def test_native_ingest_checks_keys_against_data_info(tmp_path):
    """A data.info whose byte ranges point at ANOTHER site's line (same read count) must not be scored under the wrong
    transcript / position: the reference indexes json.loads(line)[tx_id][str(tx_pos)] and raises KeyError
    (utils/data_utils.py:185); here the native parser refuses the part (M6A_EPARSE)."""
    from m6anet_b200 import _cabi
    from m6anet_b200.constants import PRETRAINED_CONFIGS
    from m6anet_b200.data import NanopolishDS
    sys_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    import sys
    sys.path.insert(0, sys_path)
    from make_synthetic_dataset import write_dataset
    d = tmp_path / "d"
    write_dataset(str(d), 6, 20, seed=3)
    norm = PRETRAINED_CONFIGS["HCT116_RNA002"][2]
    assert len(NanopolishDS(str(d), 20, norm, mode="Inference").load_sites(0, 6).read_off) == 7       # intact: fine
    lines = open(d / "data.info").read().splitlines()
    a, b = lines[2].split(","), lines[3].split(",")
    a[2:4], b[2:4] = b[2:4], a[2:4]                                  # swap the byte ranges of sites 1 and 2
    lines[2], lines[3] = ",".join(a), ",".join(b)
    open(d / "data.info", "w").write("\n".join(lines) + "\n")
    ds = NanopolishDS(str(d), 20, norm, mode="Inference")
    with pytest.raises(_cabi.M6AError) as e:
        ds.load_sites(0, 6)
    assert e.value.status == -6                                       # M6A_EPARSE
