"""validate()-style literal MIL forward: bags WITHOUT replacement, pooled by the model's
pooling block, passes averaged, scored against labels -- reference utils/training_utils.py:213-268.

CPU part (`-m "not gpu"`): the Floyd bag stream (oracle and the host build of the SHIPPED device generator), the
oracle against the outputs of the reference's own validate() (tests/golden/validate_golden.npz, made by
tests/golden/make_golden.py), the metrics against sklearn, the labelled dataset modes, argument errors.
GPU part (`-m gpu`): the kernel through the C ABI against those goldens (explicit bags = replay of the reference's
MT19937 draws) and against the oracle on the device stream.

Tolerances: bag / site probabilities <= 2e-6 absolute (float32, same bags; north_star allows 1e-4); bag indices bit-exact;
roc_auc / pr_auc exact unless two averaged predictions sit within 2e-6 of each other (none do in the fixtures)."""
import ctypes as C
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ASSETS, GOLDEN, MODEL_FILES, ROOT, oracle_params

BAG_ATOL = 2e-6
POOLS = ("prod", "mean", "max")
POOL_BLOCK = {"prod": "SigmoidProdPooling", "mean": "SigmoidMeanPooling", "max": "SigmoidMaxPooling"}


@pytest.fixture(scope="module")
def vgold():
    z = np.load(os.path.join(GOLDEN, "validate_golden.npz"))
    return {k: z[k] for k in z.files}


def golden_sites(vgold, bundled_flat, mode):
    """flat buffers of the reference's `mode` set (rows of bundled_flat.npz), its bags and labels"""
    si = vgold[f"{mode}_site_index"]
    ro = bundled_flat["read_off"]
    n = np.diff(ro)[si]
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    feats = np.concatenate([bundled_flat["feats"][ro[s]:ro[s + 1]] for s in si])
    return feats, off, bundled_flat["kmer_idx"][si], vgold[f"{mode}_bags"]


@pytest.fixture(scope="module")
def labelled_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("labelled")
    src = os.path.join(GOLDEN, "bundled")
    for fn in ("data.info", "data.info.labelled"):
        shutil.copyfile(os.path.join(src, fn), d / fn)
    with gzip.open(os.path.join(src, "data.json.gz"), "rb") as f, open(d / "data.json", "wb") as g:
        shutil.copyfileobj(f, g)
    return str(d)


NORM = os.path.join(ASSETS, "norm_factors", "rna002_hct116.npz")


def product_model(pool="prod"):
    import copy
    from m6anet_b200.constants import DEFAULT_MODEL_CONFIG
    from m6anet_b200.model import MILModel, load_model_config
    cfg = copy.deepcopy(load_model_config(DEFAULT_MODEL_CONFIG))
    assert cfg["block"][-1]["block_type"] == "SigmoidProdPooling"
    cfg["block"][-1]["block_type"] = POOL_BLOCK[pool]
    model = MILModel(cfg)
    model.load_weights(os.path.join(ASSETS, "model_states", MODEL_FILES["HCT116_RNA002"]))
    return model


# ===================================================== CPU ==============================================================
@pytest.fixture(scope="module")
def rng_emul(tmp_path_factory):
    """g++ build of the shipped device generator source (m6anet_b200/csrc/m6a_rng.cuh) -- tests/native/rng_emul.cpp"""
    out = tmp_path_factory.mktemp("native") / "rng_emul.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", str(out),
                    os.path.join(ROOT, "tests", "native", "rng_emul.cpp")], check=True)
    return C.CDLL(str(out))


def emul_sample(lib, seed, site, n_reads, n_iters, n_samples, without_replacement):
    from oracle import block_layout
    ipl, n_blocks = block_layout(n_iters)
    out = np.full((n_iters, n_samples), -1, dtype=np.int32)
    lib.emul_sample(C.c_uint64(seed), C.c_uint64(site), C.c_uint32(n_reads), n_iters, n_samples, n_blocks, ipl,
                    int(without_replacement), out.ctypes.data_as(C.c_void_p))
    return out.astype(np.int64)


@pytest.mark.parametrize("seed,site,n,n_iters,k", [(0, 0, 20, 5, 20), (7, 2**33 + 5, 57, 300, 20), (3, 9, 700, 1000, 20),
                                                   (1, 4, 33, 70, 7), (2, 5, 64, 40, 64), (2**40 + 3, 1, 300, 2100, 20),
                                                   (5, 6, 256, 64, 20), (5, 6, 257, 64, 20)])
def test_shipped_generator_source_matches_oracle_on_host(rng_emul, seed, site, n, n_iters, k):
    """The device generator header compiles for the host; both of its streams (inference draws with replacement, Floyd
    bags without) equal the NumPy specification bit for bit -- on the CPU, before any GPU run."""
    from oracle import bag_indices, sample_indices
    assert np.array_equal(emul_sample(rng_emul, seed, site, n, n_iters, k, True), bag_indices(seed, site, n, n_iters, k))
    assert np.array_equal(emul_sample(rng_emul, seed, site, n, n_iters, k, False), sample_indices(seed, site, n, n_iters, k))


def test_floyd_bags_are_distinct_in_range_and_uniform():
    from oracle import bag_indices_many
    for n, k in ((20, 20), (21, 20), (33, 7), (700, 20), (64, 64)):
        b = bag_indices_many(3, np.arange(5) + 2**35, np.full(5, n), 96, k)
        assert b.shape == (5, 96, k) and b.min() >= 0 and b.max() < n
        assert all(len(set(row)) == k for site in b for row in site)
    # every read is in a bag with probability k/n, every pair with k(k-1)/(n(n-1))
    n, k, S, it = 30, 20, 200, 1000
    b = bag_indices_many(1, np.arange(S), np.full(S, n), it, k)
    freq = np.bincount(b.reshape(-1), minlength=n) / (S * it)
    assert np.abs(freq - k / n).max() < 4 * np.sqrt((k / n) * (1 - k / n) / (S * it))
    member = np.zeros((S * it, n), dtype=bool)
    np.put_along_axis(member, b.reshape(-1, k), True, axis=1)
    pair = (member[:, 0] & member[:, 1]).mean()
    want = k * (k - 1) / (n * (n - 1))
    assert abs(pair - want) < 4 * np.sqrt(want * (1 - want) / (S * it))
    # streams of different sites / seeds differ, same arguments reproduce
    assert not np.array_equal(b[0], b[1])
    assert np.array_equal(b[:3], bag_indices_many(1, np.arange(3), np.full(3, n), it, k))
    with pytest.raises(ValueError):
        bag_indices_many(0, [0], [19], 4, 20)


def test_mt19937_bag_replay_equals_numpy_choice():
    """np.random.choice(n, 20, replace=False) of the reference dataset (utils/data_utils.py:214) is replayed exactly."""
    from oracle import bag_indices_mt19937
    np.random.seed(5)
    rs = np.random.RandomState(5)
    for n in (20, 37, 662, 21):
        assert np.array_equal(np.random.choice(n, 20, replace=False), bag_indices_mt19937(rs, n, 20))


@pytest.mark.parametrize("pool", POOLS)
@pytest.mark.parametrize("mode", ("Val", "Test"))
def test_oracle_validate_matches_reference_validate(vgold, bundled_flat, pool, mode):
    """PIN: the restatement reproduces the per-pass predictions of the reference's own validate() on the same bags."""
    from oracle import mil_validate
    feats, off, kmer, bags = golden_sites(vgold, bundled_flat, mode)
    n_iters = int(vgold["n_iters"])
    _, bag_prob, site_mean, _ = mil_validate(oracle_params("HCT116_RNA002"), feats, off, kmer, n_iters, pool=pool,
                                             sample_idx=bags.astype(np.int64))
    y_pred = vgold[f"{pool}_{mode}_y_pred"]                       # [n_iters, n_sites]
    assert np.abs(bag_prob.T - y_pred).max() <= 3e-7
    assert np.abs(site_mean - np.mean([list(r) for r in y_pred], axis=0)).max() <= 3e-7


def test_oracle_validate_on_device_stream_is_consistent(synthetic_inputs):
    from oracle import bag_indices, mil_validate, pool_bags
    si = synthetic_inputs
    sel = slice(0, 12)
    off = si["read_off"][:13]
    feats = si["feats"][:off[-1]]
    rp, bag_prob, mean, mc = mil_validate(oracle_params("HCT116_RNA002"), feats, off, si["kmer_idx"][sel], 9, seed=4,
                                          site_id_base=100, n_samples=20)
    for s in (0, 5, 11):
        idx = bag_indices(4, 100 + s, int(off[s + 1] - off[s]), 9, 20)
        assert np.array_equal(bag_prob[s], pool_bags(rp[off[s]:off[s + 1]], idx, "prod"))
    assert np.allclose(mean, bag_prob.mean(axis=1), atol=1e-6)
    # a bag that holds every read once is the deterministic noisy-OR of the site (n_reads == n_samples)
    full = pool_bags(rp[:20], np.arange(20)[None, :], "prod")[0]
    assert abs(full - (1 - np.prod(1 - rp[:20].astype(np.float64)))) < 1e-6


def test_c_oracle_validate_equals_numpy_oracle(synthetic_inputs):
    """oracle/c (used to check EVERY site of full-size jobs) against the NumPy restatement: bags bit-exact, pooled values
    to float32 rounding of the read encoder."""
    from oracle import bag_indices, c_oracle, mil_validate
    si = synthetic_inputs
    feats, off, kmer = si["feats"], si["read_off"], si["kmer_idx"]
    P = oracle_params("HCT116_RNA002")
    for seed, site, n, it, k in ((3, 7, 57, 300, 20), (2**63 + 1, 2**40 + 5, 700, 70, 20), (1, 2, 64, 33, 64), (9, 9, 21, 9, 7)):
        assert np.array_equal(c_oracle.sample_bags(seed, site, n, it, k), bag_indices(seed, site, n, it, k))
    for pool in POOLS:
        rp, bag, _, _ = mil_validate(P, feats, off, kmer, 12, seed=5, site_id_base=2**33, pool=pool)
        crp, cbag = c_oracle.mil_validate(P, feats, off, kmer, 12, seed=5, site_id_base=2**33, pool=pool)
        assert np.abs(rp - crp).max() <= 1e-6 and np.abs(bag - cbag).max() <= 1e-6


def test_metrics_match_sklearn():
    from sklearn.metrics import auc, precision_recall_curve, roc_curve
    from m6anet_b200.validation import get_accuracy, get_pr_auc, get_roc_auc
    rng = np.random.default_rng(0)
    for t in range(200):
        n = int(rng.integers(2, 300))
        y = rng.integers(0, 2, n)
        if y.min() == y.max():
            y[0] = 1 - y[0]
        p = rng.random(n).astype(np.float32)
        if t % 3 == 0:
            p = np.round(p, 1)        # ties
        fpr, tpr, _ = roc_curve(y, p)
        prec, rec, _ = precision_recall_curve(y, p, pos_label=1)
        assert abs(get_roc_auc(y, p) - auc(fpr, tpr)) < 1e-12
        assert abs(get_pr_auc(y, p) - auc(rec, prec)) < 1e-12
    assert np.isnan(get_roc_auc(np.ones(5, dtype=int), rng.random(5)))
    assert get_accuracy([0, 1, 1, 0], [0, 1, 0, 0]) == 0.75


def test_losses_match_torch():
    import torch
    from m6anet_b200.validation import (binary_cross_entropy_loss, build_loss_function,
                                        weighted_binary_cross_entropy_loss)
    rng = np.random.default_rng(1)
    y = rng.integers(0, 2, 400)
    p = rng.random(400).astype(np.float32)
    p[:3] = (0.0, 1.0, 1e-30)       # BCELoss clamps the logs at -100
    pt, yt = torch.from_numpy(p), torch.from_numpy(y)
    want = torch.nn.BCELoss()(pt.flatten(), yt.float()).item()
    assert abs(binary_cross_entropy_loss(p, y) - want) <= 1e-5 * abs(want)
    # the reference's weighting (loss_functions.py:35-39)
    _, counts = torch.unique(yt, return_counts=True)
    pos_weight, neg_weight = counts
    want_w = (torch.nn.BCELoss(reduction="none")(pt, yt.float()) * torch.where(yt == 0, neg_weight, pos_weight)).mean().item()
    assert abs(weighted_binary_cross_entropy_loss(p, y) - want_w) <= 1e-5 * abs(want_w)
    assert build_loss_function({"loss_function_type": "binary_cross_entropy_loss"}) is binary_cross_entropy_loss
    with pytest.raises(ValueError):
        build_loss_function({})


@pytest.mark.parametrize("pool", POOLS)
@pytest.mark.parametrize("mode", ("Val", "Test"))
def test_metrics_reproduce_reference_validate_results(vgold, pool, mode):
    from m6anet_b200.validation import binary_cross_entropy_loss, get_pr_auc, get_roc_auc
    y_pred, y_true = vgold[f"{pool}_{mode}_y_pred"], vgold[f"{pool}_{mode}_y_true"]
    avg = np.mean([list(r) for r in y_pred], axis=0)
    roc, pr, loss = vgold[f"{pool}_{mode}_metrics"]
    assert abs(get_roc_auc(y_true, avg) - roc) < 1e-12
    assert abs(get_pr_auc(y_true, avg) - pr) < 1e-12
    assert abs(binary_cross_entropy_loss(avg, y_true) - loss) <= 1e-6 * loss


def test_labelled_dataset_modes_match_reference(vgold, bundled_flat, labelled_dir):
    """NanopolishDS(mode='Val'/'Test'): data.info.labelled filtered by set_type and min_reads, labels from
    modification_status (reference utils/data_utils.py:102-103,124-129)."""
    from m6anet_b200.data import NanopolishDS
    key = {(str(t), int(p)): i for i, (t, p) in enumerate(zip(bundled_flat["tx_id"], bundled_flat["tx_pos"]))}
    for mode in ("Val", "Test"):
        ds = NanopolishDS(labelled_dir, 20, NORM, mode=mode)
        assert np.array_equal(ds.labels, vgold[f"prod_{mode}_y_true"])
        si = np.array([key[(t, int(p))] for t, p in zip(ds._tx, ds._pos)])
        assert np.array_equal(si, vgold[f"{mode}_site_index"])
        df = ds.data_info
        assert list(df["modification_status"]) == list(ds.labels) and set(df["set_type"]) == {mode}
        # flat ingest of the labelled subset == the same rows of the inference ingest
        feats, off, kmer, _ = golden_sites(vgold, bundled_flat, mode)
        b = ds.load_sites(0, len(ds))
        assert np.array_equal(b.feats, feats) and np.array_equal(b.read_off, off) and np.array_equal(b.kmer_idx, kmer)
        # reference-shaped item: a bag of min_reads rows drawn without replacement from the NumPy stream + the label
        np.random.seed(3)
        x, k, y = ds[0]
        np.random.seed(3)
        pick = np.random.choice(int(ds.n_reads[0]), 20, replace=False)
        assert x.shape == (20, 9) and k.shape == (20, 3) and y == ds.labels[0]
        assert np.array_equal(x, feats[off[0]:off[1]][pick])
    train = NanopolishDS(labelled_dir, 20, NORM, mode="Train")
    assert len(train) == 57 and int(train.labels.sum()) == 15
    with pytest.raises(ValueError):
        NanopolishDS(labelled_dir, 20, NORM, mode="Validation")


def test_labelled_replicate_dataset_pools_directories(labelled_dir, tmp_path):
    """NanopolishReplicateDS in a labelled mode joins on (transcript, position, modification_status, set_type)
    (reference utils/data_utils.py:345-373)."""
    from m6anet_b200.data import NanopolishDS, NanopolishReplicateDS
    other = tmp_path / "rep1"
    shutil.copytree(labelled_dir, other)
    single = NanopolishDS(labelled_dir, 20, NORM, mode="Val")
    ds = NanopolishReplicateDS([labelled_dir, str(other)], 40, NORM, mode="Val")
    assert len(ds) == len(single)
    assert np.array_equal(ds.n_reads, 2 * single.n_reads) and np.array_equal(ds.labels, single.labels)
    assert list(ds._tx) == list(single._tx)
    # min_reads applies to the pooled count: sites of 10..19 reads per directory qualify when pooled
    pooled = NanopolishReplicateDS([labelled_dir, str(other)], 20, NORM, mode="Val")
    assert len(pooled) > len(single)


def test_validate_argument_errors(labelled_dir):
    from m6anet_b200 import _cabi
    from m6anet_b200.data import NanopolishDS
    from m6anet_b200.validation import binary_cross_entropy_loss, validate
    L = _cabi.lib()
    assert L.m6a_mil_validate_f32(None, None, None, None, 0, 0, 0, 20, 5, 0, None, 0, 0, 0.5, None, None, None, None, None,
                                  0, None) == -1                                   # M6A_EINVAL: no model
    assert L.m6a_mil_validate_host_f32(None, None, None, None, 1, 0, 20, 5, 0, 7, 0, 0.5, None, None, None, None, 0) == -1
    assert L.m6a_sample_bags(0, 0, 19, 4, 20, C.c_void_p(8), None) == -4           # M6A_ERANGE: fewer reads than the bag
    assert L.m6a_sample_bags(0, 0, 100, 4, 65, C.c_void_p(8), None) == -1
    model = product_model()
    with pytest.raises(ValueError):         # unlabelled dataset
        validate(model, NanopolishDS(labelled_dir, 20, NORM, mode="Inference"), "cuda", binary_cross_entropy_loss, 2)
    with pytest.raises(RuntimeError):       # no CPU path
        validate(model, NanopolishDS(labelled_dir, 20, NORM, mode="Val"), "cpu", binary_cross_entropy_loss, 2)


# ===================================================== GPU ==============================================================
_ENG = {}


def engine():
    from m6anet_b200 import weights as W
    from m6anet_b200.engine import MilEngine
    if "e" not in _ENG:
        _ENG["e"] = MilEngine(W.from_npz(os.path.join(ASSETS, "model_states", MODEL_FILES["HCT116_RNA002"])), "cuda:0")
    return _ENG["e"]


def run_validate_device(feats, off, kmer, n_iters, **kw):
    import torch
    eng = engine()
    dev = eng.device
    f = torch.from_numpy(np.ascontiguousarray(feats, np.float32)).to(dev)
    o = torch.from_numpy(np.ascontiguousarray(off, np.int64)).to(dev)
    k = torch.from_numpy(np.ascontiguousarray(kmer, np.int32)).to(dev)
    if kw.get("sample_idx") is not None:
        kw["sample_idx"] = torch.from_numpy(np.ascontiguousarray(kw["sample_idx"], np.uint16)).to(dev)
    out = eng.validate_device(f, o, k, n_iters, **kw)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in out]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,site,n,k", [(0, 0, 20, 20), (1234, 7_000_000_123, 50, 20), (2**63 + 5, 2**40 + 1, 4001, 20),
                                           (7, 3, 21, 20), (7, 3, 300, 7), (7, 3, 64, 64)])
def test_device_bag_stream_matches_oracle(seed, site, n, k):
    from oracle import bag_indices
    got = engine().sample_bags(seed, site, n, 257, k).cpu().numpy()
    assert np.array_equal(got.astype(np.int64), bag_indices(seed, site, n, 257, k))


@pytest.mark.gpu
@pytest.mark.parametrize("pool", POOLS)
@pytest.mark.parametrize("mode", ("Val", "Test"))
def test_kernel_reproduces_reference_validate_on_replayed_bags(vgold, bundled_flat, pool, mode):
    """Explicit bags = the draws of the reference's MT19937 stream: the kernel's per-pass predictions, their average and
    the metrics equal what the reference's own validate() returned."""
    from m6anet_b200.validation import binary_cross_entropy_loss, get_pr_auc, get_roc_auc
    feats, off, kmer, bags = golden_sites(vgold, bundled_flat, mode)
    n_iters = int(vgold["n_iters"])
    rp, bag_prob, site_mean, _ = run_validate_device(feats, off, kmer, n_iters, pooling=pool, sample_idx=bags)
    y_pred, y_true = vgold[f"{pool}_{mode}_y_pred"], vgold[f"{pool}_{mode}_y_true"]
    assert np.abs(bag_prob.T - y_pred).max() <= BAG_ATOL
    avg_ref = np.mean([list(r) for r in y_pred], axis=0)
    assert np.abs(site_mean - avg_ref).max() <= BAG_ATOL
    ro = bundled_flat["read_off"]
    want_rp = np.concatenate([bundled_flat["read_prob"][ro[s]:ro[s + 1]] for s in vgold[f"{mode}_site_index"]])
    assert np.abs(rp - want_rp).max() <= BAG_ATOL
    avg = bag_prob.T[0].copy()
    for row in bag_prob.T[1:]:
        avg = avg + row
    avg = avg / np.float32(n_iters)
    roc, pr, loss = vgold[f"{pool}_{mode}_metrics"]
    assert abs(get_roc_auc(y_true, avg) - roc) < 1e-9 and abs(get_pr_auc(y_true, avg) - pr) < 1e-9
    assert abs(binary_cross_entropy_loss(avg, y_true) - loss) <= 1e-4 * loss


@pytest.mark.gpu
@pytest.mark.parametrize("pool", POOLS)
@pytest.mark.parametrize("n_samples,n_iters,replace", [(20, 5, False), (20, 300, False), (7, 40, False), (20, 64, True),
                                                       (33, 9, True)])
def test_kernel_matches_oracle_on_device_bag_stream(synthetic_inputs, pool, n_samples, n_iters, replace):
    """288 ragged sites (20..700 reads): bags drawn on the device (Floyd, or the inference stream with replace=True)."""
    from oracle import mil_validate, pool_bags, sample_indices_many
    si = synthetic_inputs
    feats, off, kmer = si["feats"], si["read_off"], si["kmer_idx"]
    base = 7_000_000_000
    rp, bag_prob, site_mean, mc = run_validate_device(feats, off, kmer, n_iters, seed=11, site_id_base=base,
                                                      n_samples=n_samples, pooling=pool, replace=replace)
    P = oracle_params("HCT116_RNA002")
    if replace:
        idx = sample_indices_many(11, base + np.arange(len(off) - 1), np.diff(off), n_iters, n_samples)
        orp, obag, omean, omc = mil_validate(P, feats, off, kmer, n_iters, n_samples=n_samples, pool=pool, sample_idx=idx)
    else:
        orp, obag, omean, omc = mil_validate(P, feats, off, kmer, n_iters, seed=11, site_id_base=base,
                                             n_samples=n_samples, pool=pool)
    assert np.abs(rp - orp).max() <= BAG_ATOL
    assert np.abs(bag_prob - obag).max() <= BAG_ATOL
    # the device averages the passes pairwise (butterfly over lanes, then blocks): held to the float64 mean of the bags;
    # the reference-order float32 average (np.mean(all_y_pred, axis=0)) is formed on the host from bag_prob
    assert np.abs(site_mean - obag.astype(np.float64).mean(axis=1)).max() <= BAG_ATOL
    assert np.abs(omean - obag.astype(np.float64).mean(axis=1)).max() <= 1e-5
    assert np.abs(mc - omc).max() <= 1


@pytest.mark.gpu
def test_validate_host_path_equals_device_path_and_ignores_sharding(synthetic_inputs):
    si = synthetic_inputs
    feats, off, kmer = si["feats"], si["read_off"], si["kmer_idx"]
    S = len(off) - 1
    rp, bag, mean, mc = run_validate_device(feats, off, kmer, 12, seed=5, pooling="prod")
    eng = engine()
    hrp, hbag, hmean, hmc = eng.validate_host(feats, off, kmer, 12, seed=5, pooling="prod", n_chunks=5)
    assert np.array_equal(hrp, rp) and np.array_equal(hbag, bag) and np.array_equal(hmean, mean) and np.array_equal(hmc, mc)
    # two shards with their own site_id_base give the same bags
    cut = S // 3
    a = eng.validate_host(feats[:off[cut]], off[:cut + 1], kmer[:cut], 12, seed=5)
    b = eng.validate_host(feats[off[cut]:], off[cut:] - off[cut], kmer[cut:], 12, seed=5, site_id_base=cut)
    assert np.array_equal(np.concatenate([a[1], b[1]]), bag)
    # the inference entry point with the same (CUDA-core) read encoder gives the same per-read probabilities bit for bit;
    # with the tensor-core encoder they agree to float32 round-off
    eng.set_encoder("ffma")
    irp, isp, imc = eng.infer_host(feats, off, kmer, 100, seed=5)
    assert np.array_equal(irp, rp) and np.array_equal(imc, mc) and np.isfinite(isp).all()
    eng.set_encoder("tc")
    irp, isp, imc = eng.infer_host(feats, off, kmer, 100, seed=5)
    assert np.max(np.abs(irp - rp)) <= 2e-6 and np.isfinite(isp).all()


@pytest.mark.gpu
def test_sites_with_fewer_reads_than_the_bag_have_no_bag(synthetic_inputs):
    """np.random.choice(n, 20, replace=False) raises for n < 20 in the reference (its datasets drop such sites); the
    kernel marks them NaN and leaves every other site untouched.  Empty sites are NaN in every mode."""
    si = synthetic_inputs
    rng = np.random.default_rng(2)
    n_reads = np.array([25, 19, 0, 20, 3, 40])
    off = np.concatenate([[0], np.cumsum(n_reads)]).astype(np.int64)
    feats = rng.standard_normal((int(off[-1]), 9), dtype=np.float32)
    kmer = si["kmer_idx"][:6]
    _, bag, mean, mc = run_validate_device(feats, off, kmer, 6, seed=1)
    ok = n_reads >= 20
    assert np.isfinite(bag[ok]).all() and np.isnan(bag[~ok]).all()
    assert np.isfinite(mean[ok]).all() and np.isnan(mean[~ok]).all()
    _, bag_r, mean_r, _ = run_validate_device(feats, off, kmer, 6, seed=1, replace=True)
    assert np.isnan(bag_r[2]).all() and np.isfinite(np.delete(bag_r, 2, axis=0)).all()
    # a site with exactly n_samples reads: every bag is the whole site => the deterministic noisy-OR
    rp, bag3, _, _ = run_validate_device(feats, off, kmer, 6, seed=9)
    p = rp[off[3]:off[4]].astype(np.float64)
    assert np.abs(bag3[3] - (1 - np.prod(1 - p))).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("pool", POOLS)
def test_validate_end_to_end_on_the_labelled_bundled_data(vgold, bundled_flat, labelled_dir, pool):
    """The drop-in call validate(model, val_dl, device, criterion, n_iterations): data.info.labelled + data.json ->
    the reference's result dictionary, checked against the oracle on the same device bag stream."""
    from m6anet_b200.data import NanopolishDS
    from m6anet_b200.validation import binary_cross_entropy_loss, get_pr_auc, get_roc_auc, validate
    from oracle import mil_validate
    model = product_model(pool)
    ds = NanopolishDS(labelled_dir, 20, NORM, mode="Test")

    class Loader:           # the reference hands validate() a DataLoader; only `.dataset` is used
        dataset = ds

    res = validate(model, Loader, "cuda", binary_cross_entropy_loss, 5, seed=3)
    assert set(res) == {"y_pred", "y_true", "compute_time", "roc_auc", "pr_auc", "avg_loss"}
    feats, off, kmer, _ = golden_sites(vgold, bundled_flat, "Test")
    _, obag, omean, _ = mil_validate(oracle_params("HCT116_RNA002"), feats, off, kmer, 5, seed=3, pool=pool)
    y_pred = np.asarray(res["y_pred"], dtype=np.float32)
    assert y_pred.shape == (5, len(ds)) and np.abs(y_pred - obag.T).max() <= BAG_ATOL
    assert np.array_equal(res["y_true"], vgold[f"{pool}_Test_y_true"])
    assert abs(res["roc_auc"] - get_roc_auc(res["y_true"], omean)) < 1e-9
    assert abs(res["pr_auc"] - get_pr_auc(res["y_true"], omean)) < 1e-9
    assert abs(res["avg_loss"] - binary_cross_entropy_loss(omean, res["y_true"])) < 1e-4
    # the model separates the bundled labels about as well as in the reference's own run (different bags, same data)
    assert abs(res["roc_auc"] - vgold[f"{pool}_Test_metrics"][0]) < 0.1
    # reproducible, and independent of the ingest batching
    from m6anet_b200.validation import predict_bags
    again = predict_bags(model, ds, "cuda", 5, seed=3, reads_per_batch=300)
    assert np.array_equal(again, y_pred)
    import torch

    def torch_criterion(y_pred, y_true):            # a reference-style torch criterion is accepted as well
        return torch.nn.BCELoss()(y_pred.flatten(), y_true.float())

    res_t = validate(model, ds, "cuda:0", torch_criterion, 5, seed=3)
    assert abs(res_t["avg_loss"] - res["avg_loss"]) < 1e-5


@pytest.mark.gpu
def test_validate_properties_at_scale(synthetic_inputs):
    """100 000 ragged sites (2.9 M reads), 64 passes: size-independent properties of bags drawn WITHOUT replacement.
      * mean pooling is an unbiased estimator of the site's mean read probability (each read is in a bag w.p. k/n);
      * max pooling never exceeds the site maximum and noisy-OR is never below max pooling (1 - prod(1-p) >= max p);
      * a site with exactly k reads gives the same value in every pass;
      * the result does not depend on how the sites are cut into shards / chunks."""
    rng = np.random.default_rng(5)
    S, k, passes = 100_000, 20, 64
    n = np.minimum(20 + rng.geometric(0.12, S) - 1, 400)
    n[::97] = 20
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    feats = rng.standard_normal((int(off[-1]), 9), dtype=np.float32)
    kmer = synthetic_inputs["kmer_idx"][rng.integers(0, len(synthetic_inputs["kmer_idx"]), S)]
    eng = engine()
    rp, bag_mean, _, _ = eng.validate_host(feats, off, kmer, passes, seed=21, pooling="mean")
    _, bag_max, _, _ = eng.validate_host(feats, off, kmer, passes, seed=21, pooling="max", n_chunks=7)
    _, bag_prod, site_prod, _ = eng.validate_host(feats, off, kmer, passes, seed=21, pooling="prod", n_chunks=3)
    assert np.isfinite(bag_mean).all() and np.isfinite(bag_max).all() and np.isfinite(bag_prod).all()
    p64 = rp.astype(np.float64)
    site_mean = np.add.reduceat(p64, off[:-1]) / n
    site_sq = np.add.reduceat(p64 * p64, off[:-1]) / n
    var_read = np.maximum(site_sq - site_mean ** 2, 0)
    # variance of the mean of k draws without replacement: var/k * (n-k)/(n-1); over `passes` independent bags
    sd = np.sqrt(var_read / k * (n - k) / np.maximum(n - 1, 1) / passes)
    zscore = (bag_mean.astype(np.float64).mean(axis=1) - site_mean) / np.maximum(sd, 1e-9)
    full = n == k
    assert np.abs(zscore[~full]).max() < 8 and abs(zscore[~full].mean()) < 0.03          # unbiased, calibrated
    assert np.abs(bag_mean[full] - site_mean[full, None]).max() < 1e-6                    # the bag is the whole site
    assert (np.ptp(bag_prod[full], axis=1) == 0).all() and (np.ptp(bag_max[full], axis=1) == 0).all()
    site_max = np.maximum.reduceat(rp, off[:-1])
    assert (bag_max <= site_max[:, None]).all() and (bag_max > 0).all()
    assert (bag_prod >= bag_max - 1e-6).all() and (bag_prod <= 1).all()                   # 1 - prod(1-p) >= max p
    assert np.abs(site_prod - bag_prod.astype(np.float64).mean(axis=1)).max() < 2e-6
    # sharding independence at scale: the second half scored alone with its site_id_base
    h = S // 2
    _, bag_b, _, _ = eng.validate_host(feats[off[h]:], off[h:] - off[h], kmer[h:], passes, seed=21, pooling="prod",
                                       site_id_base=h)
    assert np.array_equal(bag_b, bag_prod[h:])


@pytest.mark.gpu
@pytest.mark.parametrize("pool", POOLS)
def test_validate_every_site_of_a_large_job_against_the_c_oracle(synthetic_inputs, pool):
    """100 000 ragged sites (about 4 M reads), 5 passes: every per-read probability and every bag of every site against
    the C restatement (oracle/c) on the same device bag stream."""
    from oracle import c_oracle
    rng = np.random.default_rng(8)
    S = 100_000
    n = rng.integers(20, 61, S)
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    feats = rng.standard_normal((int(off[-1]), 9), dtype=np.float32)
    kmer = synthetic_inputs["kmer_idx"][rng.integers(0, len(synthetic_inputs["kmer_idx"]), S)]
    rp, bag, _, _ = engine().validate_host(feats, off, kmer, 5, seed=77, site_id_base=3_000_000_000, pooling=pool)
    crp, cbag = c_oracle.mil_validate(oracle_params("HCT116_RNA002"), feats, off, kmer, 5, seed=77, site_id_base=3_000_000_000,
                                      pool=pool)
    assert np.abs(rp - crp).max() <= 3e-6          # float32 evaluations of 4 M reads sit up to ~2e-6 apart
    assert np.abs(bag - cbag).max() <= 3e-6
