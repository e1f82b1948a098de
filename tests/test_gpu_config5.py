"""BASELINE config 5 as written: several synthetic input directories pooled through
NanopolishReplicateDS (outer join over data.info, reads of a site concatenated over the directories that hold it;
reference utils/data_utils.py:341-375,395-427), arabidopsis_RNA002 weights, 1000 iterations -- through the drop-in CLI,
every site against the C oracle, and 4 ranks against 1.  Default size 8 000 sites per directory (seconds);
M6A_CFG5_SITES=250000 runs the BASELINE size (tools/gpu/r02_config5_full.sh keeps that run's record under profiles/)."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu
SITES = int(os.environ.get("M6A_CFG5_SITES", "8000"))
THR = 0.0032978046219796


@pytest.fixture(scope="module", params=[0, 10], ids=["same-keys", "10pct-disjoint"])
def replicate_dirs(request, tmp_path_factory):
    from make_synthetic_dataset import write_replicates
    root = tmp_path_factory.mktemp(f"cfg5_{request.param}")
    dirs, _ = write_replicates(str(root), 4, SITES, 20, seed=5, disjoint=request.param)
    return dirs, request.param


def run_cli(dirs, out_dir, **kw):
    from m6anet_b200 import inference
    argv = ["--input_dir", *dirs, "--out_dir", str(out_dir), "--n_processes", "8", "--num_iterations", "1000", "--device", "cuda",
            "--seed", "0", "--pretrained_model", "arabidopsis_RNA002"]
    args = inference.argparser().parse_args(argv)
    for k, v in kw.items():
        setattr(args, k, v)
    inference.main(args)


def test_config5_replicate_dirs_every_site_and_four_ranks(replicate_dirs, tmp_path):
    from conftest import oracle_params
    from m6anet_b200.constants import PRETRAINED_CONFIGS
    from m6anet_b200.data import NanopolishReplicateDS
    from oracle import c_oracle, read_probabilities_float64
    from test_gpu_cli import torchrun_cli
    dirs, disjoint = replicate_dirs
    one = tmp_path / "one"
    run_cli(dirs, one)
    site = pd.read_csv(one / "data.site_proba.csv")
    indiv = pd.read_csv(one / "data.indiv_proba.csv")

    # the pooled sites as the dataset hands them to the kernel, scored by the C oracle on the shared index stream
    ds = NanopolishReplicateDS(dirs, 20, PRETRAINED_CONFIGS["arabidopsis_RNA002"][2], mode='Inference')
    batch = ds.load_sites(0, len(ds))
    n_reads = np.diff(batch.read_off)
    assert len(site) == len(ds) == SITES
    if disjoint:
        assert set(np.unique(n_reads)) == {60, 80} and abs((n_reads == 60).mean() - disjoint / 100) < 0.01
    else:
        assert np.all(n_reads == 80)
    P = oracle_params("arabidopsis_RNA002")
    orp, osp, omc = c_oracle.mil_inference(P, batch.feats, batch.read_off, batch.kmer_idx, 1000, seed=0, site_id_base=0,
                                           n_samples=20, read_threshold=THR)
    p64 = read_probabilities_float64(P, batch.feats, np.repeat(batch.kmer_idx, n_reads, axis=0))
    assert np.array_equal(site["n_reads"].to_numpy(), n_reads)
    assert np.max(np.abs(site["probability_modified"].to_numpy() - osp.astype(np.float64))) <= 1e-4          # every site
    rp = indiv["probability_modified"].to_numpy()
    assert len(rp) == len(orp) and np.max(np.abs(rp - p64)) <= 3e-6 and np.max(np.abs(rp - orp)) <= 5e-6       # every read
    near = np.abs(orp.astype(np.float64) - float(np.float32(THR))) < 1e-6
    slack = np.add.reduceat(near.astype(np.int64), batch.read_off[:-1]) / n_reads
    assert np.all(np.abs(site["mod_ratio"].to_numpy() - omc / n_reads) <= slack + 1e-12)
    # read ids carry the replicate suffix "{id}_{rep}" (reference utils/data_utils.py:421-423)
    assert indiv["read_index"].astype(str).str.match(r"^\d+_[0-3]$").all()

    # 4 ranks (BASELINE: 4 x B200): byte-identical files
    four = tmp_path / "four"
    torchrun_cli(4, dirs, four, ["--num_iterations", "1000", "--pretrained_model", "arabidopsis_RNA002", "--n_processes", "8"],
                 timeout=3600)
    for name in ("data.site_proba.csv", "data.indiv_proba.csv"):
        with open(one / name, "rb") as f, open(four / name, "rb") as g:
            assert f.read() == g.read(), name
