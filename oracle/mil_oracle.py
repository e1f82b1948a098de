"""NumPy restatement of the reference MIL-inference hot path (TEST INFRASTRUCTURE).

Every function cites the reference lines it follows (paths relative to
``/root/reference/m6anet``).  Arithmetic is float32 wherever the reference's is.
Not imported by the product package; see ``oracle/__init__.py``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .philox import bag_indices_many, sample_indices_many


@dataclass
class ReadEncoderParams:
    """Unfolded parameters exactly as stored in a reference ``state_dict``
    (keys listed in SURVEY.md section 8b).  ``emb`` is None for the signal-only topology
    (reference model/configs/model_configs/prod_pooling_signal.toml)."""

    emb: Optional[np.ndarray]      # [66, 2]     read_level_encoder.1.embedding_layer.weight
    w1: np.ndarray                 # [150, 15|9] read_level_encoder.3.layers.0.weight
    b1: np.ndarray                 # [150]
    bn_gamma: np.ndarray           # [150]       read_level_encoder.3.layers.1.weight
    bn_beta: np.ndarray            # [150]
    bn_mean: np.ndarray            # [150]       running_mean
    bn_var: np.ndarray             # [150]       running_var
    w2: np.ndarray                 # [32, 150]   read_level_encoder.4.layers.0.weight
    b2: np.ndarray                 # [32]
    w3: np.ndarray                 # [1, 32]     pooling_filter.probability_layer.0.weight
    b3: np.ndarray                 # [1]
    bn_eps: float = 1e-5           # torch.nn.BatchNorm1d default (model_blocks/blocks.py:251)

    @classmethod
    def from_npz(cls, path) -> "ReadEncoderParams":
        z = np.load(path)
        f = lambda k: np.ascontiguousarray(z[k], dtype=np.float32)
        emb = f("emb") if "emb" in z.files and z["emb"].size else None
        return cls(emb, f("w1"), f("b1"), f("bn_gamma"), f("bn_beta"), f("bn_mean"), f("bn_var"),
                   f("w2"), f("b2"), f("w3"), f("b3"), float(z["bn_eps"]) if "bn_eps" in z.files else 1e-5)


def read_probabilities(params: ReadEncoderParams, feats: np.ndarray, kmer: Optional[np.ndarray]) -> np.ndarray:
    """Per-read modification probability ``p_r``.

    Restates ``model.pooling_filter.probability_layer(model.get_read_representation(
    {'X','kmer'})).flatten()`` (utils/inference_utils.py:35-37) for the shipped topologies:

      DeaggregateNanopolish   X.view(-1, 9), kmer.view(-1, 1)          model_blocks/blocks.py:126
      KmerMultipleEmbedding   Embedding(66, 2)(kmer).reshape(-1, 6)    model_blocks/blocks.py:204-205
      ConcatenateFeatures     cat([X, emb], axis=1) (X first)          model_blocks/blocks.py:65
      Linear(15,150)+BatchNorm1d(eval)+ReLU(+Dropout p=0)              model_blocks/blocks.py:249-255
      Linear(150,32)+ReLU                                              model_blocks/blocks.py:249-255
      Linear(32,1)+Sigmoid                                             model_blocks/pooling_blocks.py:52

    feats: [N, 9] float32 normalised signal features; kmer: [N, 3] integer ids (or None when
    ``params.emb`` is None: ExtractSignal, model_blocks/blocks.py:86).  Returns float32 [N].
    """
    f32 = np.float32
    x = np.ascontiguousarray(feats, dtype=f32).reshape(-1, params.w1.shape[1] - (0 if params.emb is None else 6))
    if params.emb is not None:
        k = np.asarray(kmer).reshape(-1, 3).astype(np.int64)
        e = params.emb.astype(f32)[k].reshape(-1, 6)
        x = np.concatenate([x, e], axis=1)
    h = x @ params.w1.astype(f32).T + params.b1.astype(f32)
    # eval-mode BatchNorm1d: (h - running_mean) / sqrt(running_var + eps) * weight + bias
    inv_std = (f32(1.0) / np.sqrt(params.bn_var.astype(f32) + f32(params.bn_eps))).astype(f32)
    h = (h - params.bn_mean.astype(f32)) * inv_std * params.bn_gamma.astype(f32) + params.bn_beta.astype(f32)
    h = np.maximum(h, f32(0))
    h = np.maximum(h @ params.w2.astype(f32).T + params.b2.astype(f32), f32(0))
    z = (h @ params.w3.astype(f32).reshape(-1, 1)).reshape(-1) + params.b3.astype(f32).reshape(())
    with np.errstate(over="ignore"):
        p = f32(1.0) / (f32(1.0) + np.exp(-z, dtype=f32))
    return p.astype(f32)



def read_probabilities_float64(params: ReadEncoderParams, feats: np.ndarray, kmer: Optional[np.ndarray]) -> np.ndarray:
    """The read encoder of `read_probabilities` evaluated in float64 (exact to ~1e-15): the yardstick for float32 error.
    Over millions of reads every float32 evaluation -- the reference's own torch calls included -- sits up to ~1.9e-6
    from this value, so full-size checks hold the kernel to it (3e-6) instead of to another float32 evaluation."""
    f64 = np.float64
    x = np.asarray(feats).astype(f64)
    if params.emb is not None:
        x = np.concatenate([x, params.emb.astype(f64)[np.asarray(kmer)].reshape(-1, 3 * params.emb.shape[1])], axis=1)
    h = x @ params.w1.astype(f64).T + params.b1
    h = (h - params.bn_mean) / np.sqrt(params.bn_var.astype(f64) + params.bn_eps) * params.bn_gamma + params.bn_beta
    h = np.maximum(h, 0)
    h = np.maximum(h @ params.w2.astype(f64).T + params.b2, 0)
    z = h @ params.w3.astype(f64).reshape(-1) + float(params.b3.reshape(-1)[0])
    return 1.0 / (1.0 + np.exp(-z))

def noisy_or_site_probability(read_prob: np.ndarray, idx: np.ndarray) -> np.float32:
    """Monte-Carlo noisy-OR for ONE site on a given index stream.

    ``_calculate_site_proba`` (utils/inference_utils.py:85-86) with the random draw replaced
    by the supplied indices ``idx`` [n_iters, n_samples]:
        proba = proba[idx]                       # np.random.choice(..., replace=True).reshape
        (1 - np.prod(1 - proba, axis=1)).mean()  # all float32
    Identical to ``SigmoidProdPooling.forward`` on the gathered bags followed by the mean over
    iterations (model_blocks/pooling_blocks.py:127-129; utils/training_utils.py:236-256).
    """
    proba = np.asarray(read_prob, dtype=np.float32)[np.asarray(idx)]
    return (1 - np.prod(1 - proba, axis=1)).mean()


def mod_ratio(read_prob: np.ndarray, threshold: float) -> float:
    """``np.mean(x >= args.read_proba_threshold)`` (utils/inference_utils.py:53).  The float32
    array against a Python float compares in float32 (threshold rounded to float32)."""
    x = np.asarray(read_prob, dtype=np.float32)
    return float(np.mean(x >= np.float32(threshold)))


def closed_form_site_probability(read_prob: np.ndarray, n_samples: int = 20) -> float:
    """Expectation of the with-replacement estimator, 1 - (1 - mean p)^n.
    Statistical cross-check of the device RNG only; never the parity oracle."""
    p = np.asarray(read_prob, dtype=np.float64)
    return float(1.0 - (1.0 - p.mean()) ** n_samples)


def mil_inference(params: ReadEncoderParams, feats: np.ndarray, read_off: np.ndarray, kmer_idx: Optional[np.ndarray],
                  n_iters: int, seed: int = 0, site_id_base: int = 0, n_samples: int = 20,
                  read_threshold: float = 0.033379376, sample_idx: Optional[Sequence[np.ndarray]] = None):
    """Whole hot path on flat site-contiguous buffers, the oracle twin of the C-ABI call
    ``m6a_mil_infer_f32`` (include/m6anet_b200.h).

    feats [total_reads, 9] f32; read_off [n_sites+1] CSR offsets; kmer_idx [n_sites, 3].
    Follows ``run_inference`` (utils/inference_utils.py:35-54): read encoder over every read,
    ``group_results`` split by n_reads (:107-140), ``mod_ratio`` (:53), MC noisy-OR (:54,74-87).
    Index stream: ``sample_idx[s]`` if given, else the Philox stream of site
    ``site_id_base + s`` (oracle/philox.py: Philox-seeded MWC64X lane streams).

    Returns (read_prob f32 [total_reads], site_prob f32 [n_sites], mod_count i32 [n_sites]).
    """
    read_off = np.asarray(read_off, dtype=np.int64)
    n_sites = len(read_off) - 1
    n_reads = np.diff(read_off)
    kmer_rows = None
    if params.emb is not None:
        kmer_rows = np.repeat(np.asarray(kmer_idx).reshape(n_sites, 3), n_reads, axis=0)  # utils/data_utils.py:223-224
    read_prob = read_probabilities(params, feats, kmer_rows)
    site_prob = np.zeros(n_sites, dtype=np.float32)
    mod_count = np.zeros(n_sites, dtype=np.int32)
    thr = np.float32(read_threshold)
    CH = 64  # sites per vectorised index-stream batch
    for s0 in range(0, n_sites, CH):
        s1 = min(n_sites, s0 + CH)
        idx_b = None
        if sample_idx is None:
            idx_b = sample_indices_many(seed, site_id_base + np.arange(s0, s1), np.maximum(n_reads[s0:s1], 1), n_iters, n_samples)
        for s in range(s0, s1):
            p = read_prob[read_off[s]:read_off[s + 1]]
            mod_count[s] = int(np.count_nonzero(p >= thr))
            if len(p) == 0:
                site_prob[s] = np.float32("nan")
                continue
            idx = sample_idx[s] if sample_idx is not None else idx_b[s - s0]
            site_prob[s] = noisy_or_site_probability(p, idx)
    return read_prob, site_prob, mod_count


def pool_bags(read_prob: np.ndarray, idx: np.ndarray, pool: str = "prod") -> np.ndarray:
    """Site-level output of the pooling filter on gathered bags, float32 [n_bags].

    ``idx`` [n_bags, n_reads_per_site] selects the reads of each bag; the three instance-based pooling blocks are
        prod  1 - prod(1 - p, axis=1)   SigmoidProdPooling.forward  model_blocks/pooling_blocks.py:127-129
        mean  mean(p, axis=1)           SigmoidMeanPooling.forward  model_blocks/pooling_blocks.py:96-98
        max   max(p, axis=1)            SigmoidMaxPooling.forward   model_blocks/pooling_blocks.py:158-160
    (``MILModel.forward`` -> ``get_site_probability`` = pooling_filter(read_representation) with an empty decoder,
    model/model.py:140-164).  All float32 like the torch ops.
    """
    p = np.asarray(read_prob, dtype=np.float32)[np.asarray(idx)]
    if pool == "prod":
        return (np.float32(1) - np.prod(np.float32(1) - p, axis=1)).astype(np.float32)
    if pool == "mean":
        return np.mean(p, axis=1, dtype=np.float32)
    if pool == "max":
        return np.max(p, axis=1)
    raise ValueError(f"unknown pooling {pool!r}")


def mil_validate(params: ReadEncoderParams, feats: np.ndarray, read_off: np.ndarray, kmer_idx: Optional[np.ndarray],
                 n_iters: int, seed: int = 0, site_id_base: int = 0, n_samples: int = 20, pool: str = "prod",
                 read_threshold: float = 0.033379376, sample_idx: Optional[Sequence[np.ndarray]] = None):
    """validate()-style literal MIL forward on flat buffers, the oracle twin of ``m6a_mil_validate_f32``.

    Restates the evaluation loop of the reference (utils/training_utils.py:236-256): every pass draws, per site, a bag
    of ``min_reads`` reads WITHOUT replacement (utils/data_utils.py:213-214), runs ``MILModel.forward`` on it
    (model/model.py:155-164: read encoder -> pooling filter) and the passes are averaged
    (``np.mean(all_y_pred, axis=0)``, float32 rows added in pass order).  The read encoder is evaluated once per read
    (eval mode: BatchNorm running statistics, no dropout => a read's probability does not depend on its bag).
    Bags: ``sample_idx[s]`` [n_iters, n_samples] if given, else the device Floyd stream of site ``site_id_base + s``
    (oracle/philox.py: bag_indices).

    Returns (read_prob f32 [R], bag_prob f32 [n_sites, n_iters], site_mean f32 [n_sites], mod_count i32 [n_sites]).
    """
    read_off = np.asarray(read_off, dtype=np.int64)
    n_sites = len(read_off) - 1
    n_reads = np.diff(read_off)
    kmer_rows = None
    if params.emb is not None:
        kmer_rows = np.repeat(np.asarray(kmer_idx).reshape(n_sites, 3), n_reads, axis=0)
    read_prob = read_probabilities(params, feats, kmer_rows)
    bag_prob = np.full((n_sites, n_iters), np.nan, dtype=np.float32)
    mod_count = np.zeros(n_sites, dtype=np.int32)
    thr = np.float32(read_threshold)
    ok = n_reads >= n_samples
    idx_all = None
    if sample_idx is None and ok.any():
        sel = np.nonzero(ok)[0]
        idx_all = dict(zip(sel.tolist(), bag_indices_many(seed, site_id_base + sel, n_reads[sel], n_iters, n_samples)))
    for s in range(n_sites):
        p = read_prob[read_off[s]:read_off[s + 1]]
        mod_count[s] = int(np.count_nonzero(p >= thr))
        if sample_idx is not None:
            if len(p):
                bag_prob[s] = pool_bags(p, sample_idx[s], pool)
        elif ok[s]:
            bag_prob[s] = pool_bags(p, idx_all[s], pool)
    # np.mean(list of per-pass float32 lists, axis=0): rows are added in pass order, then divided (float32)
    acc = np.zeros(n_sites, dtype=np.float32)
    for it in range(n_iters):
        acc = acc + bag_prob[:, it]
    return read_prob, bag_prob, (acc / np.float32(n_iters)).astype(np.float32), mod_count
