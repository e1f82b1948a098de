#!/usr/bin/env python
"""The drop-in CLI on a synthetic data.json (anchor: reference README.md:245-249 -- 95 030 sites in
408 s = 233 sites/s for the whole CLI): wall time of `python -m m6anet_b200 inference` and the rates of its stages measured
in-process (index, native ingest, H2D + kernel + D2H, native CSV emit).  Prints one JSON line.
    python tools/gpu_cli_timing.py [sites] [reads] [work_dir]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_synthetic_dataset import write_dataset       # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
work = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp(prefix="m6a_cli_")
data_dir, out_dir = os.path.join(work, "data"), os.path.join(work, "out")
t0 = time.perf_counter()
# the generator is single-threaded Python: write 8 slices in parallel and concatenate them into one directory
from concurrent.futures import ProcessPoolExecutor       # noqa: E402


def _slice(args):
    i, lo, hi = args
    d = os.path.join(work, f"part{i}")
    write_dataset(d, hi, n, seed=100 + i, kmer_seed=1, skip=lambda s: s < lo)
    return d


parts = 8
edges = [S * i // parts for i in range(parts + 1)]
with ProcessPoolExecutor(parts) as ex:
    dirs = list(ex.map(_slice, [(i, edges[i], edges[i + 1]) for i in range(parts)]))
os.makedirs(data_dir, exist_ok=True)
off = 0
with open(os.path.join(data_dir, "data.json"), "wb") as fj, open(os.path.join(data_dir, "data.info"), "w") as fi:
    fi.write("transcript_id,transcript_position,start,end,n_reads\n")
    for d in dirs:
        with open(os.path.join(d, "data.json"), "rb") as f:
            blob = f.read()
        fj.write(blob)
        with open(os.path.join(d, "data.info")) as f:
            next(f)
            for line in f:
                tx, pos, a, b, nr = line.rstrip("\n").split(",")
                fi.write(f"{tx},{pos},{int(a) + off},{int(b) + off},{nr}\n")
        off += len(blob)
t_gen = time.perf_counter() - t0
json_mb = off / 1e6

env = dict(os.environ, PYTHONPATH=ROOT)
t0 = time.perf_counter()
r = subprocess.run([sys.executable, "-m", "m6anet_b200", "inference", "--input_dir", data_dir, "--out_dir", out_dir,
                    "--num_iterations", "1000", "--n_processes", "16"], cwd=ROOT, env=env, capture_output=True, text=True)
wall = time.perf_counter() - t0
assert r.returncode == 0, r.stderr[-2000:]
rows = sum(1 for _ in open(os.path.join(out_dir, "data.indiv_proba.csv"))) - 1
csv_mb = (os.path.getsize(os.path.join(out_dir, "data.indiv_proba.csv")) + os.path.getsize(os.path.join(out_dir, "data.site_proba.csv"))) / 1e6

# stages in-process
from m6anet_b200 import _cabi, inference, weights as W     # noqa: E402
from m6anet_b200.constants import PRETRAINED_CONFIGS       # noqa: E402
from m6anet_b200.data import NanopolishDS                  # noqa: E402
from m6anet_b200.engine import MilEngine                   # noqa: E402
t0 = time.perf_counter()
ds = NanopolishDS(data_dir, 20, PRETRAINED_CONFIGS["HCT116_RNA002"][2], mode="Inference")
t_index = time.perf_counter() - t0
pool = _cabi.PinnedPool()
t0 = time.perf_counter()
pool.give_back(pool.empty(1 << 20, np.uint8))            # creates the CUDA context (part of the CLI wall time, not of a stage)
t_context = time.perf_counter() - t0
t0 = time.perf_counter()
batch = ds.load_sites(0, len(ds), n_threads=16, alloc=pool.empty)
t_ingest = time.perf_counter() - t0
eng = MilEngine(W.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna002_hct116.npz")), 0)
eng.infer_host(batch.feats[:1000 * n], batch.read_off[:1001], batch.kmer_idx[:1000], 1000)
t0 = time.perf_counter()
rp, sp, mc = eng.infer_host(batch.feats, batch.read_off, batch.kmer_idx, 1000, alloc=pool.empty)
t_gpu = time.perf_counter() - t0
t0 = time.perf_counter()
with open(os.path.join(work, "site.csv"), "wb") as f, open(os.path.join(work, "indiv.csv"), "wb") as g:
    inference.write_site_rows(f, batch, sp, mc, 16)
    inference.write_indiv_rows(g, batch, rp, 16)
t_emit = time.perf_counter() - t0
line = {"what": "drop-in CLI on a synthetic data.json, num_iterations=1000, HCT116_RNA002", "sites": S, "reads_per_site": n,
        "data_json_mb": json_mb, "csv_mb": csv_mb, "indiv_rows": rows, "cli_wall_s": wall, "cli_sites_per_s": S / wall,
        "reference_published_cli_sites_per_s": 232.8,
        "stages": {"cuda_context_s": t_context, "index_s": t_index, "ingest_s": t_ingest, "ingest_sites_per_s": S / t_ingest, "ingest_json_gbs": json_mb / 1e3 / t_ingest,
                   "h2d_kernel_d2h_s": t_gpu, "emit_s": t_emit, "emit_rows_per_s": (rows + S) / t_emit},
        "host_threads": 16, "dataset_generation_s": t_gen}
print(json.dumps(line))
