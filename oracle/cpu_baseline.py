"""Timed CPU port of the reference hot path (TEST/BENCH INFRASTRUCTURE; see oracle/__init__.py).

This is the `cpu_baseline` / `--impl reference` leg of bench.py: the same library calls the
reference makes on its CPU path, on all host cores:

  read encoder   reference: torch CPU ops in 16-site batches (utils/inference_utils.py:35-37,
                 scripts/inference.py:104 batch_size=16).  Port: the same torch.nn.functional calls the
                 reference's modules make (embedding, cat, linear, batch_norm(eval), relu, sigmoid), all
                 host threads, timed both in 16-site batches and as one large batch; the faster is reported.
                 (tests/test_oracle.py checks this torch port against the NumPy restatement.)
  MC pooling     reference: calculate_site_proba -> multiprocessing.Pool(n_processes).imap of
                 _calculate_site_proba = np.random.choice(...).reshape(n_iters, 20); (1-prod(1-p)).mean()
                 (utils/inference_utils.py:74-104).  Port: identical calls in a Pool over all cores.

The reference itself cannot travel to the GPU box (/root/reference does not exist there), hence a port.
"""
from __future__ import annotations

import os
import time
from multiprocessing import Pool

import numpy as np

from .mil_oracle import ReadEncoderParams


def read_probabilities_torch(params: ReadEncoderParams, feats: np.ndarray, kmer_rows):
    """Same ops, same order as the reference modules (model_blocks/blocks.py:126,204-205,65,249-255;
    model_blocks/pooling_blocks.py:52) through torch.nn.functional on CPU."""
    import torch
    import torch.nn.functional as F
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    with torch.no_grad():
        x = t(feats).view(-1, 9)
        if params.emb is not None:
            k = torch.from_numpy(np.ascontiguousarray(kmer_rows, dtype=np.int64)).view(-1, 1)
            e = F.embedding(k, t(params.emb)).reshape(-1, 3 * params.emb.shape[1])
            x = torch.cat([x, e], dim=1)
        h = F.linear(x, t(params.w1), t(params.b1))
        h = F.batch_norm(h, t(params.bn_mean), t(params.bn_var), t(params.bn_gamma), t(params.bn_beta),
                         training=False, eps=params.bn_eps)
        h = F.relu(h)
        h = F.relu(F.linear(h, t(params.w2), t(params.b2)))
        return torch.sigmoid(F.linear(h, t(params.w3).view(1, -1), t(params.b3))).flatten().numpy()


def _site_proba_task(task):
    # body of reference _calculate_site_proba (utils/inference_utils.py:84-87)
    proba, n_iters, n_samples = task
    proba = np.random.choice(proba, n_iters * n_samples, replace=True).reshape(n_iters, n_samples)
    return (1 - np.prod(1 - proba, axis=1)).mean()


def _site_proba_chunk(args):
    chunk, n_iters, n_samples = args
    return [_site_proba_task((p, n_iters, n_samples)) for p in chunk]


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def time_reference_port(params: ReadEncoderParams, feats: np.ndarray, read_off: np.ndarray, kmer_idx: np.ndarray,
                        n_iters: int, n_samples: int = 20, n_procs: int | None = None, read_threshold: float = 0.033379376):
    """Run the CPU port once over the given sites; returns dict(sites_per_s, t_encoder, t_mc, cores, ...)."""
    n_procs = n_procs or host_cores()
    read_off = np.asarray(read_off, dtype=np.int64)
    n_sites = len(read_off) - 1
    n_reads = np.diff(read_off)
    kmer_rows = np.repeat(np.asarray(kmer_idx).reshape(n_sites, 3), n_reads, axis=0) if params.emb is not None else None

    import torch
    torch.set_num_threads(n_procs)
    read_probabilities_torch(params, feats[:1024], None if kmer_rows is None else kmer_rows[:1024])  # warm-up
    # encoder, one large batch (all host threads)
    t0 = time.perf_counter()
    p = read_probabilities_torch(params, feats, kmer_rows)
    t_big = time.perf_counter() - t0
    # encoder, the reference's 16-site batches, on a slice then scaled (bounded time)
    n_b = min(n_sites, 16 * 256)
    t0 = time.perf_counter()
    for a in range(0, n_b, 16):
        b = min(a + 16, n_b)
        sl = slice(read_off[a], read_off[b])
        read_probabilities_torch(params, feats[sl], None if kmer_rows is None else kmer_rows[sl])
    t_small = (time.perf_counter() - t0) * (n_sites / max(1, n_b))
    t_enc = min(t_big, t_small)

    # mod_ratio (utils/inference_utils.py:53) + MC pooling in a Pool (utils/inference_utils.py:102-104)
    t0 = time.perf_counter()
    per_site = [p[read_off[s]:read_off[s + 1]] for s in range(n_sites)]
    _ = np.array([np.mean(x >= read_threshold) for x in per_site])
    chunks = [(per_site[i:i + 64], n_iters, n_samples) for i in range(0, n_sites, 64)]
    if n_procs > 1:
        with Pool(n_procs) as pool:
            out = pool.map(_site_proba_chunk, chunks)
    else:
        out = [_site_proba_chunk(c) for c in chunks]
    site = np.array([v for c in out for v in c], dtype=np.float32)
    t_mc = time.perf_counter() - t0
    return dict(sites_per_s=n_sites / (t_enc + t_mc), t_encoder_s=t_enc, t_encoder_big_batch_s=t_big,
                t_encoder_16site_batches_s=t_small, t_mc_s=t_mc, cores=n_procs, n_sites=n_sites,
                site_prob_mean=float(site.mean()))
