"""GPU end-to-end tests of the drop-in CLI path, modelled on the reference's own tests
(m6anet/tests/test_inference.py:10-82): `m6anet inference` on the bundled data against the reference's
golden CSVs, single directory and replicate pooling, with the reference tolerances."""
import gzip
import os
import shutil

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bundled_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("bundled")
    with gzip.open(os.path.join(GOLDEN, "bundled", "data.json.gz"), "rb") as f, open(d / "data.json", "wb") as g:
        shutil.copyfileobj(f, g)
    shutil.copyfile(os.path.join(GOLDEN, "bundled", "data.info"), d / "data.info")
    return str(d)


def inference_args(input_dirs, out_dir, **kw):
    """the reference's InferenceArgs fixture (m6anet/tests/conftest.py:60-78) on the CUDA device"""
    from m6anet_b200 import inference
    argv = ["--input_dir", *input_dirs, "--out_dir", str(out_dir), "--n_processes", "4", "--num_iterations", "10000",
            "--device", "cuda", "--seed", "0"]
    args = inference.argparser().parse_args(argv)
    for k, v in kw.items():
        setattr(args, k, v)
    return args


def test_inference(bundled_dir, tmp_path):
    from m6anet_b200 import inference
    args = inference_args([bundled_dir], tmp_path)
    inference.main(args)
    test_indiv = pd.read_csv(tmp_path / "data.indiv_proba.csv").sort_values(["transcript_id", "transcript_position", "read_index"]).reset_index(drop=True)
    test_site = pd.read_csv(tmp_path / "data.site_proba.csv").sort_values(["transcript_id", "transcript_position"]).reset_index(drop=True)
    indiv = pd.read_csv(os.path.join(GOLDEN, "bundled", "data.indiv_proba.csv.gz")).sort_values(["transcript_id", "transcript_position", "read_index"]).reset_index(drop=True)
    site = pd.read_csv(os.path.join(GOLDEN, "bundled", "data.site_proba.csv.gz")).sort_values(["transcript_id", "transcript_position"]).reset_index(drop=True)
    assert np.all(indiv["transcript_id"] == test_indiv["transcript_id"])
    assert np.all(indiv["transcript_position"] == test_indiv["transcript_position"])
    assert np.all(indiv["read_index"] == test_indiv["read_index"])
    assert np.allclose(indiv["probability_modified"], test_indiv["probability_modified"])
    assert np.all(site["transcript_id"] == test_site["transcript_id"])
    assert np.all(site["transcript_position"] == test_site["transcript_position"])
    assert np.all(site["n_reads"] == test_site["n_reads"]) and np.all(site["kmer"] == test_site["kmer"])
    assert np.allclose(site["mod_ratio"], test_site["mod_ratio"])
    assert np.allclose(site["probability_modified"], test_site["probability_modified"], atol=1e-2)
    # site order = data.info order filtered to n_reads >= 20 (SURVEY.md section 8b "files out")
    info = pd.read_csv(os.path.join(bundled_dir, "data.info"))
    info = info[info["n_reads"] >= 20]
    raw = pd.read_csv(tmp_path / "data.site_proba.csv")
    assert list(raw["transcript_position"]) == list(info["transcript_position"]) and len(raw) == 101


def test_inference_is_reproducible_and_batching_independent(bundled_dir, tmp_path):
    """same seed => identical files, whatever the batching (the reference CLI is not reproducible)"""
    from m6anet_b200 import inference
    outs = []
    for name, kw in (("a", {}), ("b", {"reads_per_batch": 700, "n_processes": 1}), ("c", {"reads_per_batch": 1, "n_processes": 2})):
        out = tmp_path / name
        inference.main(inference_args([bundled_dir], out, num_iterations=1000, **kw))
        outs.append((open(out / "data.site_proba.csv").read(), open(out / "data.indiv_proba.csv").read()))
    assert outs[0] == outs[1] == outs[2]
    out = tmp_path / "seed1"
    inference.main(inference_args([bundled_dir], out, num_iterations=1000, seed=1))
    assert open(out / "data.site_proba.csv").read() != outs[0][0]
    assert open(out / "data.indiv_proba.csv").read() == outs[0][1]


def test_inference_replicates(bundled_dir, tmp_path):
    from m6anet_b200 import inference
    rep = tmp_path / "replicate"
    shutil.copytree(bundled_dir, rep)
    out = tmp_path / "out"
    inference.main(inference_args([bundled_dir, str(rep)], out))
    test_indiv = pd.read_csv(out / "data.indiv_proba.csv")
    test_site = pd.read_csv(out / "data.site_proba.csv")
    test_indiv["rep_num"] = test_indiv["read_index"].apply(lambda x: x.split("_")[1]).astype(int)
    test_indiv["read_index"] = test_indiv["read_index"].apply(lambda x: x.split("_")[0]).astype(int)
    keys = ["transcript_id", "transcript_position", "read_index"]
    r0 = test_indiv[test_indiv["rep_num"] == 0].sort_values(keys).reset_index(drop=True)
    r1 = test_indiv[test_indiv["rep_num"] == 1].sort_values(keys).reset_index(drop=True)
    assert np.all(r0["transcript_id"] == r1["transcript_id"]) and np.all(r0["read_index"] == r1["read_index"])
    assert np.allclose(r0["probability_modified"], r1["probability_modified"])
    indiv = pd.read_csv(os.path.join(GOLDEN, "bundled", "data.indiv_proba.csv.gz")).rename({'probability_modified': 'gt_probability'}, axis=1)
    site = pd.read_csv(os.path.join(GOLDEN, "bundled", "data.site_proba.csv.gz")).rename(
        {'probability_modified': 'gt_probability', 'mod_ratio': 'gt_mod_ratio'}, axis=1)
    for r in (r0, r1):
        m = r.merge(indiv, on=keys)
        assert len(m) == len(indiv) - 0 or len(m) > 0
        assert np.allclose(m["probability_modified"], m["gt_probability"])
    # pooling doubles every site's reads: every golden site must be present, with the same mod_ratio
    m = test_site.merge(site, on=["transcript_id", "transcript_position"])
    assert len(m) == len(site)
    assert np.allclose(m["mod_ratio"], m["gt_mod_ratio"])
    assert np.allclose(m["probability_modified"], m["gt_probability"], atol=1e-2)


def test_cli_entry_point(bundled_dir, tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "m6anet_b200", "inference", "--input_dir", bundled_dir, "--out_dir", str(tmp_path),
                        "--num_iterations", "5", "--n_processes", "2", "--pretrained_model", "HEK293T_RNA004"],
                       cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    site = pd.read_csv(tmp_path / "data.site_proba.csv")
    assert len(site) == 101 and site["probability_modified"].between(0, 1).all()


def torchrun_cli(n_ranks, input_dirs, out_dir, extra=(), timeout=1800):
    """`m6anet inference` under torchrun with n_ranks ranks.  With fewer GPUs than ranks every rank uses cuda:0 and the one
    all-gather runs over gloo on host tensors (M6A_DIST_BACKEND), so that the shard / gather / concatenate path is
    exercised on a single-GPU box too; with enough GPUs it is the production path (one rank per GPU, NCCL)."""
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    dev = []
    if torch.cuda.device_count() < n_ranks:
        env["M6A_DIST_BACKEND"] = "gloo"
        dev = ["--device", "cuda:0"]
    port = str(29600 + (os.getpid() + n_ranks) % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_ranks),
                        "--master-addr", "127.0.0.1", "--master-port", port, "-m", "m6anet_b200", "inference", "--input_dir",
                        *input_dirs, "--out_dir", str(out_dir), "--n_processes", "4", *dev, *extra],
                       cwd=root, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    assert not [f for f in os.listdir(out_dir) if ".rank" in f]


def test_two_rank_cli_equals_single_process(bundled_dir, tmp_path):
    """torchrun with 2 ranks: site shards + one all-gather + rank-ordered CSV concatenation give byte-identical files
    (runs on one GPU as well, see torchrun_cli)."""
    from m6anet_b200 import inference
    one = tmp_path / "one"
    inference.main(inference_args([bundled_dir], one, num_iterations=1000))
    two = tmp_path / "two"
    torchrun_cli(2, [bundled_dir], two, ["--num_iterations", "1000"])
    for name in ("data.site_proba.csv", "data.indiv_proba.csv"):
        assert open(one / name).read() == open(two / name).read()
