"""validate(): the reference's evaluation loop on the B200 kernel.

Drop-in mirror of reference m6anet/utils/training_utils.py
    validate(model, val_dl, device, criterion, n_iterations)   :213-268
    get_roc_auc / get_pr_auc / get_accuracy                    :15-61   (sklearn roc_curve / precision_recall_curve / auc)
and of the loss wrappers in m6anet/utils/loss_functions/loss_functions.py:8-40 (torch BCELoss).

What the reference does per validation pass: every site's dataset item is a bag of `min_reads` reads drawn WITHOUT
replacement (utils/data_utils.py:213-214), `MILModel.forward` runs the read encoder on the bag and pools it with the
model's pooling block (model/model.py:155-164); the passes are averaged and scored against the labels.
Here the sites are ingested once (native parser), the read encoder runs once per read and the bags of all passes are
drawn and pooled on the device by `m6a_mil_validate_host_f32` (include/m6anet_b200.h): same numbers on the same bags
(tests/test_validate.py holds the kernel to the outputs of the reference's own validate() on a replayed MT19937
stream), without n_iterations re-reads of data.json.  The bags come from the counter-based device stream keyed by
(seed, site index), so the result does not depend on batching or GPU count.  There is no CPU path.
Under torchrun (WORLD_SIZE > 1) the sites are sharded by read count over the ranks and ONE all-gather of the per-pass
predictions (4 * n_iterations bytes per site) gives every rank the full matrix; metrics are computed on every rank.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Optional

import numpy as np

from .inference import DEFAULT_READS_PER_BATCH, _resolve_device, plan_batches

_POOL_OF_BLOCK = {"SigmoidProdPooling": "prod", "SigmoidMeanPooling": "mean", "SigmoidMaxPooling": "max"}


# ---- metrics (reference utils/training_utils.py:15-61; sklearn.metrics semantics restated in NumPy) ---------------------
def _binary_clf_curve(y_true: np.ndarray, y_score: np.ndarray):
    """Cumulative false / true positives at every distinct score, scores descending (sklearn's _binary_clf_curve with
    pos_label = 1: stable descending sort, one point per run of tied scores)."""
    y_true = np.asarray(y_true).reshape(-1)
    y_score = np.asarray(y_score, dtype=np.float64).reshape(-1)
    if y_true.shape != y_score.shape:
        raise ValueError("y_true and y_pred must have the same length")
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_pos = y_score[order], (y_true[order] == 1)
    last_of_run = np.r_[np.nonzero(np.diff(y_score))[0], y_pos.size - 1]
    tps = np.cumsum(y_pos, dtype=np.float64)[last_of_run]
    fps = 1 + last_of_run - tps
    return fps, tps, y_score[last_of_run]


def _trapezoid_auc(x: np.ndarray, y: np.ndarray) -> float:
    """sklearn.metrics.auc: trapezoidal rule on a monotonic x (either direction)."""
    if len(x) < 2:
        raise ValueError("At least 2 points are needed to compute area under curve")
    dx = np.diff(x)
    direction = 1.0
    if np.any(dx < 0):
        if np.all(dx <= 0):
            direction = -1.0
        else:
            raise ValueError("x is neither increasing nor decreasing")
    return float(direction * np.sum(dx * (y[1:] + y[:-1]) / 2.0))


def get_roc_auc(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """auc(*roc_curve(y_true, y_pred)[:2]) (reference utils/training_utils.py:26-28).  NaN when only one class is
    present (sklearn warns and returns NaN rates there)."""
    fps, tps, _ = _binary_clf_curve(y_true, y_pred)
    # roc_curve drops collinear interior points (drop_intermediate=True); they do not change the trapezoid area, so
    # they are kept here.  The curve starts at (0, 0).
    fps, tps = np.r_[0.0, fps], np.r_[0.0, tps]
    if fps[-1] <= 0 or tps[-1] <= 0:
        return float("nan")
    return _trapezoid_auc(fps / fps[-1], tps / tps[-1])


def get_pr_auc(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """auc(recall, precision) of precision_recall_curve(y_true, y_pred, pos_label=1) (reference
    utils/training_utils.py:42-44): points in increasing-threshold order followed by (recall 0, precision 1)."""
    fps, tps, _ = _binary_clf_curve(y_true, y_pred)
    ps = tps + fps
    precision = np.divide(tps, ps, out=np.zeros_like(tps), where=ps != 0)
    recall = np.ones_like(tps) if tps[-1] == 0 else tps / tps[-1]
    return _trapezoid_auc(np.r_[recall[::-1], 0.0], np.r_[precision[::-1], 1.0])


def get_accuracy(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """accuracy_score(y_true, y_pred) on predicted LABELS (reference utils/training_utils.py:47-59)."""
    return float(np.mean(np.asarray(y_true).reshape(-1) == np.asarray(y_pred).reshape(-1)))


# ---- loss wrappers (reference utils/loss_functions/loss_functions.py; torch.nn.BCELoss clamps log at -100) -------------
def _bce_terms(y_pred: np.ndarray, y_true: np.ndarray) -> np.ndarray:
    p = np.asarray(y_pred, dtype=np.float32).reshape(-1).astype(np.float64)
    y = np.asarray(y_true, dtype=np.float64).reshape(-1)
    with np.errstate(divide="ignore"):
        log_p = np.maximum(np.log(p), -100.0)
        log_q = np.maximum(np.log1p(-p), -100.0)
    return -(y * log_p + (1.0 - y) * log_q)


def binary_cross_entropy_loss(y_pred, y_true) -> float:
    """BCELoss()(y_pred.flatten(), y_true.float()) (loss_functions.py:8-20), evaluated in float64 on the float32 inputs."""
    return float(np.mean(_bce_terms(y_pred, y_true)))


def weighted_binary_cross_entropy_loss(y_pred, y_true) -> float:
    """loss_functions.py:23-40: per-sample BCE weighted by the count of the OTHER class, then the plain mean
    (`pos_weight, neg_weight = counts` of torch.unique(y_true) = (#negatives, #positives); negatives are weighted by
    `neg_weight` = #positives and positives by `pos_weight` = #negatives)."""
    y = np.asarray(y_true).reshape(-1)
    counts = np.unique(y, return_counts=True)[1]
    if len(counts) != 2:
        raise ValueError("weighted_binary_cross_entropy_loss needs both classes in y_true")
    pos_weight, neg_weight = counts
    weights = np.where(y == 0, neg_weight, pos_weight).astype(np.float64)
    return float(np.mean(_bce_terms(y_pred, y) * weights))


binary_cross_entropy_loss.accepts_numpy = True
weighted_binary_cross_entropy_loss.accepts_numpy = True
LOSS_FUNCTIONS = {"binary_cross_entropy_loss": binary_cross_entropy_loss,
                  "weighted_binary_cross_entropy_loss": weighted_binary_cross_entropy_loss}


def build_loss_function(config: Dict) -> Callable:
    """reference utils/builder.py:93-107: {'loss_function_type': name, **kwargs} -> callable."""
    from functools import partial
    if "loss_function_type" not in config:
        raise ValueError("Config must specify loss_function_type")
    config = dict(config)
    fn = LOSS_FUNCTIONS[config.pop("loss_function_type")]
    if not config:
        return fn
    out = partial(fn, **config)
    out.accepts_numpy = True
    return out


def _apply_criterion(criterion: Callable, y_pred_avg: np.ndarray, y_true: np.ndarray) -> float:
    if getattr(criterion, "accepts_numpy", False):
        return float(criterion(y_pred_avg, y_true))
    import torch   # a torch criterion, called exactly like the reference does (training_utils.py:266)
    return float(criterion(torch.Tensor(np.asarray(y_pred_avg)), torch.Tensor(np.asarray(y_true))).item())


# ---- the evaluation loop ------------------------------------------------------------------------------------------
def predict_bags(model, ds, device="cuda", n_iterations: int = 1, seed: int = 0, lo: int = 0, hi: Optional[int] = None,
                 reads_per_batch: int = DEFAULT_READS_PER_BATCH, n_threads: int = 0) -> np.ndarray:
    """Pooled probability of every (pass, site) for sites [lo, hi) of a labelled dataset: float32 [n_iterations, n_sites].
    One bag of ds.min_reads reads per site and pass, drawn without replacement on the device."""
    hi = len(ds) if hi is None else hi
    pool = _POOL_OF_BLOCK.get(model.pooling_filter.block_type)
    if pool is None:
        raise NotImplementedError(f"pooling block {model.pooling_filter.block_type!r} is not supported")
    if model.n_reads_per_site != ds.min_reads:
        raise ValueError(f"the pooling block expects {model.n_reads_per_site} reads per site but the dataset draws {ds.min_reads}")
    eng = model.engine(device)
    out = np.empty((n_iterations, hi - lo), dtype=np.float32)
    for a, b in plan_batches(ds.n_reads, lo, hi, reads_per_batch):
        batch = ds.load_sites(a, b, n_threads=n_threads)
        _, bag_prob, _, _ = eng.validate_host(batch.feats, batch.read_off, batch.kmer_idx, n_iterations, seed=seed,
                                              site_id_base=a, n_samples=ds.min_reads, pooling=pool, replace=False)
        out[:, a - lo:b - lo] = bag_prob.T
    return out


def validate(model, val_dl, device: str, criterion: Callable, n_iterations: Optional[int] = 1, seed: int = 0) -> Dict:
    """Validate the model on a labelled dataset (reference utils/training_utils.py:213-268).

    `val_dl` is a labelled NanopolishDS / NanopolishReplicateDS (mode 'Val' / 'Test' / 'Train') or anything with a
    `.dataset` attribute holding one (a DataLoader in the reference).  Returns the reference's dictionary:
    y_pred (list of n_iterations per-pass prediction lists), y_true, compute_time, roc_auc, pr_auc, avg_loss.
    `seed` keys the device bag stream (the reference draws from the process-global NumPy stream instead).
    """
    from .dist import env_world
    ds = getattr(val_dl, "dataset", val_dl)
    if getattr(ds, "labels", None) is None:
        raise ValueError("validate() needs a labelled dataset (mode 'Train', 'Val' or 'Test')")
    rank, world, local_rank = env_world()
    dev = _resolve_device(device, local_rank, world)
    model.eval()
    start = time.time()
    if world > 1:     # site shards + one all-gather (torch.distributed over NCCL), like run_inference
        import os
        import torch
        import torch.distributed as dist
        from .dist import all_gather_rows, shard_bounds
        torch.cuda.set_device(dev)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        bounds = shard_bounds(ds.n_reads, world)
        local = predict_bags(model, ds, dev, int(n_iterations), seed, bounds[rank], bounds[rank + 1])
        rows = torch.from_numpy(np.ascontiguousarray(local.T)).to(torch.device("cuda", dev))
        y_pred = np.ascontiguousarray(all_gather_rows(rows, bounds).cpu().numpy().T)
    else:
        y_pred = predict_bags(model, ds, dev, int(n_iterations), seed)
    compute_time = time.time() - start
    # np.mean(all_y_pred, axis=0) on the list of per-pass float32 lists: rows added in pass order, float32
    acc = np.zeros(y_pred.shape[1], dtype=np.float32)
    for row in y_pred:
        acc = acc + row
    y_pred_avg = acc / np.float32(len(y_pred))
    y_true = np.asarray(ds.labels).flatten()
    return {"y_pred": [list(row) for row in y_pred],
            "y_true": y_true,
            "compute_time": compute_time,
            "roc_auc": get_roc_auc(y_true, y_pred_avg),
            "pr_auc": get_pr_auc(y_true, y_pred_avg),
            "avg_loss": _apply_criterion(criterion, y_pred_avg, y_true)}
