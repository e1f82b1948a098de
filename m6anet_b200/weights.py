"""Read-encoder parameters: extraction from a reference-format state_dict and BatchNorm folding.

The kernel evaluates Linear-1 with the eval-mode BatchNorm folded in (reference
model_blocks/blocks.py:249-251: Linear -> BatchNorm1d -> ReLU):
    y = (W1 x + b1 - mean) / sqrt(var + eps) * gamma + beta  =  (s * W1) x + (s * (b1 - mean) + beta)
with s = gamma / sqrt(var + eps).  The fold is done in float64 and rounded once to float32.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Mapping, Optional

import numpy as np

# state_dict keys of the two shipped topologies
_KEYS_EMB = {
    "emb": "read_level_encoder.1.embedding_layer.weight",
    "w1": "read_level_encoder.3.layers.0.weight", "b1": "read_level_encoder.3.layers.0.bias",
    "bn_gamma": "read_level_encoder.3.layers.1.weight", "bn_beta": "read_level_encoder.3.layers.1.bias",
    "bn_mean": "read_level_encoder.3.layers.1.running_mean", "bn_var": "read_level_encoder.3.layers.1.running_var",
    "w2": "read_level_encoder.4.layers.0.weight", "b2": "read_level_encoder.4.layers.0.bias",
    "w3": "pooling_filter.probability_layer.0.weight", "b3": "pooling_filter.probability_layer.0.bias",
}
_KEYS_SIGNAL = {
    "w1": "read_level_encoder.2.layers.0.weight", "b1": "read_level_encoder.2.layers.0.bias",
    "bn_gamma": "read_level_encoder.2.layers.1.weight", "bn_beta": "read_level_encoder.2.layers.1.bias",
    "bn_mean": "read_level_encoder.2.layers.1.running_mean", "bn_var": "read_level_encoder.2.layers.1.running_var",
    "w2": "read_level_encoder.3.layers.0.weight", "b2": "read_level_encoder.3.layers.0.bias",
    "w3": "pooling_filter.probability_layer.0.weight", "b3": "pooling_filter.probability_layer.0.bias",
}


@dataclass
class EncoderWeights:
    """Folded float32 parameters in the layout of ``m6a_weights_t`` (include/m6anet_b200.h)."""
    emb: Optional[np.ndarray]   # [n_kmer, emb_dim] or None
    w1: np.ndarray              # [h1, 9 + 3*emb_dim]  (BatchNorm folded)
    b1: np.ndarray              # [h1]
    w2: np.ndarray              # [h2, h1]
    b2: np.ndarray              # [h2]
    w3: np.ndarray              # [h2]
    b3: np.ndarray              # [1]

    @property
    def emb_dim(self) -> int:
        return 0 if self.emb is None else int(self.emb.shape[1])

    @property
    def n_kmer(self) -> int:
        return 0 if self.emb is None else int(self.emb.shape[0])

    @property
    def h1(self) -> int:
        return int(self.w1.shape[0])

    @property
    def h2(self) -> int:
        return int(self.w2.shape[0])

    @property
    def n_sig(self) -> int:
        return int(self.w1.shape[1]) - 3 * self.emb_dim


def fold_batchnorm(w1, b1, gamma=None, beta=None, mean=None, var=None, eps: float = 1e-5):
    w1 = np.asarray(w1, dtype=np.float64)
    b1 = np.asarray(b1, dtype=np.float64)
    if gamma is None:
        return w1.astype(np.float32), b1.astype(np.float32)
    s = np.asarray(gamma, np.float64) / np.sqrt(np.asarray(var, np.float64) + eps)
    w = w1 * s[:, None]
    b = (b1 - np.asarray(mean, np.float64)) * s + np.asarray(beta, np.float64)
    return w.astype(np.float32), b.astype(np.float32)


def from_raw(raw: Mapping[str, np.ndarray], bn_eps: float = 1e-5) -> EncoderWeights:
    """raw: unfolded arrays keyed emb,w1,b1,bn_gamma,bn_beta,bn_mean,bn_var,w2,b2,w3,b3 (emb/bn_* optional)."""
    f = lambda k: np.ascontiguousarray(np.asarray(raw[k]), dtype=np.float32)
    has_bn = "bn_gamma" in raw and raw["bn_gamma"] is not None
    if has_bn:
        w1, b1 = fold_batchnorm(raw["w1"], raw["b1"], raw["bn_gamma"], raw["bn_beta"], raw["bn_mean"], raw["bn_var"],
                                bn_eps)
    else:
        w1, b1 = fold_batchnorm(raw["w1"], raw["b1"])
    emb = f("emb") if ("emb" in raw and raw["emb"] is not None and np.size(raw["emb"])) else None
    return EncoderWeights(emb, np.ascontiguousarray(w1), np.ascontiguousarray(b1), f("w2"), f("b2"),
                          f("w3").reshape(-1), f("b3").reshape(-1))


def from_npz(path: str) -> EncoderWeights:
    z = np.load(path)
    raw = {k: z[k] for k in z.files if k != "bn_eps"}
    return from_raw(raw, float(z["bn_eps"]) if "bn_eps" in z.files else 1e-5)


def from_state_dict(sd: Mapping[str, "np.ndarray"], bn_eps: float = 1e-5) -> EncoderWeights:
    """Reference-format state_dict (torch tensors or arrays) of either shipped topology."""
    def arr(v):
        return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    keys = _KEYS_EMB if _KEYS_EMB["emb"] in sd else _KEYS_SIGNAL
    missing = [k for k in keys.values() if k not in sd]
    if missing:
        raise KeyError(f"state_dict is not one of the supported m6anet read-encoder topologies; missing {missing}")
    return from_raw({k: arr(sd[v]) for k, v in keys.items()}, bn_eps)
