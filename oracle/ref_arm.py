"""The reference arm: the UNMODIFIED reference (oracle/_ref, installed by oracle/make_ref.sh) timed on host cores.
TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py) -- only bench.py's `--impl reference` / `cpu_baseline` legs and
tests/ import this module.

Two timings of the reference's own code, inputs already in host memory (the hot path, no JSON ingest):

  time_stock_functions   exactly the calls of the loop body of `run_inference` (utils/inference_utils.py:35-37,53-54):
                         `model.get_read_representation({'X', 'kmer'})` + `model.pooling_filter.probability_layer` in the
                         reference's 16-site batches (scripts/inference.py:104 batch_size default), `group_results`,
                         the mod_ratio line, and ONE `calculate_site_proba(read_probs, n_iters, 20, n_processes)` for the
                         whole sample (one Pool; the CLI re-forks a Pool every 32 sites, so this is kinder to the CPU).
  time_stock_run_inference  the literal `run_inference(model, dl, args)` on a DataLoader over an in-memory dataset that
                         yields what `NanopolishDS.__getitem__` yields, collated by the stock `inference_collate`,
                         CSV rows written to a scratch directory (flush cadence, Pool per flush and row formatting
                         included) -- context for the first number.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
# reference registry name -> bench.py model name is the identity (utils/constants.py:24-27)


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "m6anet"))


_REF = None


def _ref():
    """Import the installed reference once (oracle/_ref first on sys.path; ujson shim lives there)."""
    global _REF
    if _REF is None:
        if not available():
            raise RuntimeError("oracle/_ref is missing: run oracle/make_ref.sh in the build container")
        warnings.filterwarnings("ignore", message="pkg_resources is deprecated")
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import toml
        import torch
        from m6anet.model.model import MILModel
        from m6anet.utils import constants, data_utils, inference_utils
        _REF = dict(toml=toml, torch=torch, MILModel=MILModel, constants=constants, data_utils=data_utils,
                    inference_utils=inference_utils)
    return _REF


def stock_model(pretrained: str):
    """scripts/inference.py:88-90 on the CPU: MILModel(toml) + load_state_dict(map_location)."""
    R = _ref()
    torch, C = R["torch"], R["constants"]
    model = R["MILModel"](R["toml"].load(C.DEFAULT_MODEL_CONFIG)).to("cpu")
    model.load_state_dict(torch.load(C.PRETRAINED_CONFIGS[pretrained][0], map_location=torch.device("cpu")))
    model.eval()
    return model, float(C.PRETRAINED_CONFIGS[pretrained][1])


def _batches(feats, read_off, kmer_idx, batch_sites):
    """What inference_collate hands to the loop (utils/data_utils.py:498-506): X [N,9] f32, kmer [N,3] i64, n_reads."""
    torch = _ref()["torch"]
    n_sites = len(read_off) - 1
    n_reads = np.diff(read_off)
    for a in range(0, n_sites, batch_sites):
        b = min(a + batch_sites, n_sites)
        sl = slice(int(read_off[a]), int(read_off[b]))
        x = torch.from_numpy(np.ascontiguousarray(feats[sl], dtype=np.float32))
        k = torch.from_numpy(np.repeat(np.asarray(kmer_idx[a:b], dtype=np.int64), n_reads[a:b], axis=0))
        yield x, k, n_reads[a:b]


def stock_read_probs(model, feats, read_off, kmer_idx, batch_sites=16):
    torch = _ref()["torch"]
    out = []
    with torch.no_grad():
        for x, k, _ in _batches(feats, read_off, kmer_idx, batch_sites):
            f = model.get_read_representation({"X": x, "kmer": k})                 # utils/inference_utils.py:35-36
            out.append(model.pooling_filter.probability_layer(f).flatten().detach().cpu().numpy())   # :37,41
    return np.concatenate(out) if out else np.zeros(0, np.float32)


def time_stock_functions(pretrained, feats, read_off, kmer_idx, n_iters, n_procs, batch_sites=16):
    R = _ref()
    IU = R["inference_utils"]
    R["torch"].set_num_threads(n_procs)
    model, thr = stock_model(pretrained)
    read_off = np.asarray(read_off, dtype=np.int64)
    n_sites = len(read_off) - 1
    warm = min(n_sites, 2 * batch_sites)          # first-call costs (thread pool start, allocator) stay out of the timing
    stock_read_probs(model, feats[: int(read_off[warm])], read_off[: warm + 1], kmer_idx[:warm], batch_sites)
    t0 = time.perf_counter()
    p = stock_read_probs(model, feats, read_off, kmer_idx, batch_sites)
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    n_reads = np.diff(read_off)
    dummy = np.zeros(len(p), dtype=np.int64)
    read_probs, *_ = IU.group_results(p, dummy, dummy, dummy, dummy, n_reads)                       # :48-51
    mod_ratios = np.array([np.mean(x >= thr) for x in read_probs])                                 # :53
    site = IU.calculate_site_proba(read_probs, n_iters, 20, n_procs)                               # :54
    t_mc = time.perf_counter() - t0
    return dict(sites_per_s=n_sites / (t_enc + t_mc), t_encoder_s=t_enc, t_mc_s=t_mc, cores=n_procs, n_sites=n_sites,
                read_prob=p, site_prob=np.asarray(site, dtype=np.float32), mod_ratio=mod_ratios)


class _ResidentSites:
    """In-memory stand-in for NanopolishDS: same __getitem__ tuple (utils/data_utils.py:225-231), same int_to_kmer."""

    def __init__(self, feats, read_off, kmer_idx):
        torch = _ref()["torch"]
        self.torch = torch
        self.feats, self.off, self.kmer = feats, np.asarray(read_off, np.int64), np.asarray(kmer_idx, np.int64)
        C = _ref()["constants"]
        self.int_to_kmer = {v: k for k, v in C.KMER_TO_INT.items()} if hasattr(C, "KMER_TO_INT") else {}

    def __len__(self):
        return len(self.off) - 1

    def __getitem__(self, i):
        a, b = int(self.off[i]), int(self.off[i + 1])
        n = b - a
        x = self.torch.from_numpy(np.ascontiguousarray(self.feats[a:b], dtype=np.float32))
        k = self.torch.from_numpy(np.repeat(self.kmer[i][None, :], n, axis=0))
        return x, k, np.array(["tx"] * n), np.array([i] * n), np.arange(n)


def time_stock_run_inference(pretrained, feats, read_off, kmer_idx, n_iters, n_procs, batch_size=16, save_per_batch=2):
    """The reference's run_inference itself (utils/inference_utils.py:14-71) on resident inputs."""
    R = _ref()
    from torch.utils.data import DataLoader
    model, thr = stock_model(pretrained)
    ds = _ResidentSites(feats, read_off, kmer_idx)
    dl = DataLoader(ds, num_workers=0, collate_fn=R["data_utils"].inference_collate, batch_size=batch_size, shuffle=False)

    class Args:
        pass
    args = Args()
    with tempfile.TemporaryDirectory() as out_dir:
        args.out_dir, args.device, args.save_per_batch = out_dir, "cpu", save_per_batch
        args.read_proba_threshold, args.num_iterations, args.n_processes = thr, n_iters, n_procs
        t0 = time.perf_counter()
        R["inference_utils"].run_inference(model, dl, args)
        dt = time.perf_counter() - t0
        with open(os.path.join(out_dir, "data.site_proba.csv")) as fh:
            rows = sum(1 for _ in fh) - 1          # header line
    return dict(sites_per_s=rows / dt if dt > 0 else 0.0, seconds=dt, sites_written=rows, sites_given=len(ds), cores=n_procs)


class _ResidentBags:
    """In-memory stand-in for the reference's evaluation datasets: __getitem__ draws `min_reads` reads of the site without
    replacement exactly like NanopolishDS in 'Val' / 'Test' mode (utils/data_utils.py:213-214) and returns
    (features, kmer, label); batches are formed by the stock train_collate."""

    def __init__(self, feats, read_off, kmer_idx, labels, min_reads=20):
        self.torch = _ref()["torch"]
        self.feats, self.off, self.kmer = feats, np.asarray(read_off, np.int64), np.asarray(kmer_idx, np.int64)
        self.labels, self.min_reads = labels, min_reads

    def __len__(self):
        return len(self.off) - 1

    def __getitem__(self, i):
        a, b = int(self.off[i]), int(self.off[i + 1])
        f = self.feats[a:b]
        f = f[np.random.choice(len(f), self.min_reads, replace=False), :]             # utils/data_utils.py:214
        x = self.torch.Tensor(np.ascontiguousarray(f))
        k = self.torch.LongTensor(np.repeat(self.kmer[i][None, :], len(f), axis=0))
        return x, k, self.labels[i]


def time_stock_validate(pretrained, feats, read_off, kmer_idx, n_iterations, n_procs, batch_size=512):
    """The reference's validate() itself (utils/training_utils.py:213-268): n_iterations passes over the sites, each pass a
    fresh bag of 20 reads per site through MILModel.forward; DataLoader with the stock train_collate, resident inputs."""
    R = _ref()
    from torch.utils.data import DataLoader
    from m6anet.utils.training_utils import validate
    torch = R["torch"]
    torch.set_num_threads(n_procs)
    model, _ = stock_model(pretrained)
    n_sites = len(read_off) - 1
    labels = (np.arange(n_sites) % 2).astype(np.int64)
    ds = _ResidentBags(feats, read_off, kmer_idx, labels)
    dl = DataLoader(ds, num_workers=0, collate_fn=R["data_utils"].train_collate, batch_size=batch_size, shuffle=False)
    t0 = time.perf_counter()
    res = validate(model, dl, "cpu", torch.nn.BCELoss(), n_iterations)
    dt = time.perf_counter() - t0
    return dict(sites_per_s=n_sites / dt, seconds=dt, n_sites=n_sites, n_iterations=n_iterations, cores=n_procs,
                roc_auc=float(res["roc_auc"]))


if __name__ == "__main__":   # python -m oracle.ref_arm : quick self-check on synthetic sites
    rng = np.random.default_rng(0)
    S, n = int(os.environ.get('REF_ARM_SITES', '512')), 50
    f = rng.standard_normal((S * n, 9), dtype=np.float32)
    o = np.arange(S + 1, dtype=np.int64) * n
    k = rng.integers(0, 66, size=(S, 3))
    cores = len(os.sched_getaffinity(0))
    r = time_stock_functions("HCT116_RNA002", f, o, k, 1000, cores)
    r2 = time_stock_run_inference("HCT116_RNA002", f, o, k, 1000, cores)
    print(json.dumps({"functions": {x: r[x] for x in ("sites_per_s", "t_encoder_s", "t_mc_s", "cores")}, "run_inference": r2}))
