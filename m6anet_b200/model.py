"""Model-config plugin surface: TOML ``[[block]]`` list -> read-encoder parameters for the kernel.

Mirrors reference m6anet/model/model.py (MILModel: build_model :40-69, _build_block :71-83) for the
block classes that the inference path can reach (reference m6anet/model/model_blocks/):

    DeaggregateNanopolish      blocks.py:89-126
    KmerMultipleEmbedding      blocks.py:165-205     } or  ExtractSignal  blocks.py:69-86
    ConcatenateFeatures        blocks.py:48-66       }
    Linear (x2)                blocks.py:208-266     (Linear [+ BatchNorm1d] + ReLU [+ Dropout: identity in eval])
    SigmoidProdPooling / SigmoidMeanPooling / SigmoidMaxPooling   pooling_blocks.py:70-160
        (run_inference only uses their shared `probability_layer` = Linear(h2, 1) + Sigmoid,
         pooling_blocks.py:52; the site pooling is always noisy-OR, utils/inference_utils.py:54,85-86)

The model holds plain NumPy parameters under the reference's state_dict key names, so a reference
``.pt`` state_dict (or this package's ``.npz`` conversion) loads unchanged.  There is no torch module
and no CPU forward: evaluation happens in the CUDA kernel through ``MILModel.engine()``.
Attention / KDE / summary-statistics pooling blocks are not reachable from the reference's
``run_inference`` (they define no ``probability_layer``) and raise NotImplementedError here.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Mapping, Optional

import numpy as np

from .weights import EncoderWeights, fold_batchnorm

_POOLING = ("SigmoidProdPooling", "SigmoidMeanPooling", "SigmoidMaxPooling")
_KNOWN_UNSUPPORTED = ("Flatten", "SummaryStatsAggregator", "ProbabilityAttention", "SummaryStatsProbability",
                      "MeanAggregator", "Attention", "GatedAttention", "KDELayer", "KDEAttentionLayer",
                      "KDEGatedAttentionLayer")
_BN_EPS = 1e-5   # torch.nn.BatchNorm1d default, used by reference blocks.py:251


def load_model_config(path: str) -> Dict:
    """toml.load equivalent (reference scripts/inference.py:88)."""
    try:
        import tomllib
        with open(path, "rb") as f:
            return tomllib.load(f)
    except ModuleNotFoundError:  # pragma: no cover  (python < 3.11)
        import toml
        return toml.load(path)


class _Block:
    def __init__(self, block_type: str, kwargs: Dict):
        self.block_type = block_type
        self.kwargs = kwargs

    def __repr__(self):
        return f"{self.block_type}({', '.join(f'{k}={v!r}' for k, v in self.kwargs.items())})"


class MILModel:
    """Configuration-built m6Anet model (reference model/model.py:8-164), inference path only."""

    def __init__(self, model_config: Mapping):
        self.model_config = copy.deepcopy(dict(model_config))
        self.read_level_encoder: List[_Block] = []
        self.pooling_filter: Optional[_Block] = None
        self.params: Dict[str, np.ndarray] = {}
        self._engines = {}
        self.build_model()

    # ---- construction -----------------------------------------------------------------------------
    def build_model(self):
        blocks = self.model_config.get("block")
        if not blocks:
            raise ValueError("model config has no [[block]] entries")
        enc: List[_Block] = []
        for raw in blocks:
            kwargs = dict(raw)
            block_type = kwargs.pop("block_type")       # model.py:46
            if block_type in _KNOWN_UNSUPPORTED:
                raise NotImplementedError(
                    f"block_type {block_type!r} is not supported by the B200 inference kernel: the reference's "
                    f"run_inference needs pooling_filter.probability_layer (utils/inference_utils.py:37), which only the "
                    f"Sigmoid*Pooling blocks define")
            if block_type in _POOLING:
                if self.pooling_filter is not None:
                    raise ValueError("more than one pooling block")
                self.pooling_filter = _Block(block_type, kwargs)
                continue
            if self.pooling_filter is not None:
                raise NotImplementedError("decoder blocks after the pooling filter are not supported on the inference path")
            if block_type not in ("DeaggregateNanopolish", "KmerMultipleEmbedding", "ConcatenateFeatures", "ExtractSignal",
                                  "Linear"):
                raise AttributeError(f"module 'model_blocks' has no attribute {block_type!r}")   # getattr failure, model.py:82
            enc.append(_Block(block_type, kwargs))
        if self.pooling_filter is None:
            raise NotImplementedError("a Sigmoid*Pooling block is required (run_inference uses its probability_layer)")
        self.read_level_encoder = enc
        self._analyse()

    def _analyse(self):
        enc = self.read_level_encoder
        types = [b.block_type for b in enc]
        if not types or types[0] != "DeaggregateNanopolish":
            raise NotImplementedError(f"encoder must start with DeaggregateNanopolish, got {types}")
        d = enc[0].kwargs
        self.num_neighboring_features = int(d["num_neighboring_features"])
        self.n_sig = int(d.get("n_features", 3)) * (2 * self.num_neighboring_features + 1)      # blocks.py:111
        rest = types[1:]
        if rest[:2] == ["KmerMultipleEmbedding", "ConcatenateFeatures"]:
            e = enc[1].kwargs
            self.n_kmer, self.emb_dim = int(e["input_channel"]), int(e["output_channel"])
            n_pos = 2 * int(e.get("num_neighboring_features", 1)) + 1
            if n_pos != 3:
                raise NotImplementedError("only 3 five-mers per site (num_neighboring_features=1) are supported")
            self._idx_emb, lin = 1, 3
        elif rest[:1] == ["ExtractSignal"]:
            self.n_kmer, self.emb_dim, self._idx_emb, lin = 0, 0, None, 2
        else:
            raise NotImplementedError(f"unsupported encoder layout {types}; supported: the shipped m6anet.toml and "
                                      f"prod_pooling_signal.toml topologies")
        if types[lin:] != ["Linear", "Linear"]:
            raise NotImplementedError(f"encoder must end with exactly two Linear blocks, got {types[lin:]}")
        self._idx_lin = (lin, lin + 1)
        l1, l2 = enc[lin].kwargs, enc[lin + 1].kwargs
        for l in (l1, l2):
            if l.get("activation", "relu") != "relu":
                raise NotImplementedError("only relu activations are supported by the kernel")
        self.in1, self.h1 = int(l1["input_channel"]), int(l1["output_channel"])
        self.h2 = int(l2["output_channel"])
        self.bn1, self.bn2 = bool(l1.get("batch_norm", True)), bool(l2.get("batch_norm", True))
        if int(l2["input_channel"]) != self.h1:
            raise ValueError("Linear blocks do not chain")
        if self.in1 != self.n_sig + 3 * self.emb_dim:
            raise ValueError(f"first Linear expects {self.in1} inputs but the encoder produces {self.n_sig + 3 * self.emb_dim}")
        p = self.pooling_filter.kwargs
        if int(p["input_channel"]) != self.h2:
            raise ValueError("pooling input_channel does not match the last Linear block")
        self.n_reads_per_site = int(p.get("n_reads_per_site", 20))

    # ---- state_dict (reference key names) ----------------------------------------------------------------
    def expected_keys(self) -> Dict[str, tuple]:
        i1, i2 = self._idx_lin
        keys = {}
        if self._idx_emb is not None:
            keys[f"read_level_encoder.{self._idx_emb}.embedding_layer.weight"] = (self.n_kmer, self.emb_dim)
        for i, bn, (o, n_in) in ((i1, self.bn1, (self.h1, self.in1)), (i2, self.bn2, (self.h2, self.h1))):
            keys[f"read_level_encoder.{i}.layers.0.weight"] = (o, n_in)
            keys[f"read_level_encoder.{i}.layers.0.bias"] = (o,)
            if bn:
                for nm in ("weight", "bias", "running_mean", "running_var"):
                    keys[f"read_level_encoder.{i}.layers.1.{nm}"] = (o,)
        keys["pooling_filter.probability_layer.0.weight"] = (1, self.h2)
        keys["pooling_filter.probability_layer.0.bias"] = (1,)
        return keys

    def load_state_dict(self, state_dict: Mapping, strict: bool = True):
        def arr(v):
            return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        want = self.expected_keys()
        missing = [k for k in want if k not in state_dict]
        extra = [k for k in state_dict if k not in want and not k.endswith("num_batches_tracked")]
        if missing or (strict and extra):
            raise RuntimeError(f"Error(s) in loading state_dict for MILModel: missing keys {missing}, unexpected keys {extra}")
        params = {}
        for k, shape in want.items():
            a = np.ascontiguousarray(arr(state_dict[k]), dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {k}: expected {shape}, got {tuple(a.shape)}")
            params[k] = a
        self.params = params
        self._engines = {}
        return self

    def state_dict(self) -> Dict[str, np.ndarray]:
        return dict(self.params)

    def load_weights(self, path: str):
        """Load a weights file: this package's .npz (tools/convert_assets.py) or a reference-format torch .pt."""
        if str(path).endswith(".npz"):
            z = np.load(path)
            i1, i2 = self._idx_lin
            sd = {}
            if self._idx_emb is not None:
                sd[f"read_level_encoder.{self._idx_emb}.embedding_layer.weight"] = z["emb"]
            sd.update({
                f"read_level_encoder.{i1}.layers.0.weight": z["w1"], f"read_level_encoder.{i1}.layers.0.bias": z["b1"],
                f"read_level_encoder.{i1}.layers.1.weight": z["bn_gamma"], f"read_level_encoder.{i1}.layers.1.bias": z["bn_beta"],
                f"read_level_encoder.{i1}.layers.1.running_mean": z["bn_mean"],
                f"read_level_encoder.{i1}.layers.1.running_var": z["bn_var"],
                f"read_level_encoder.{i2}.layers.0.weight": z["w2"], f"read_level_encoder.{i2}.layers.0.bias": z["b2"],
                "pooling_filter.probability_layer.0.weight": z["w3"], "pooling_filter.probability_layer.0.bias": z["b3"]})
            return self.load_state_dict(sd)
        import torch
        return self.load_state_dict(torch.load(path, map_location="cpu"))   # scripts/inference.py:89-90

    # ---- torch.nn.Module-like no-ops kept for drop-in call sites ------------------------------------------
    def to(self, device):
        return self

    def eval(self):
        return self

    # ---- kernel-side view ------------------------------------------------------------------------------------
    def encoder_weights(self) -> EncoderWeights:
        if not self.params:
            raise RuntimeError("MILModel has no weights: call load_state_dict/load_weights first")
        P = self.params
        i1, i2 = self._idx_lin

        def lin(i, bn):
            w, b = P[f"read_level_encoder.{i}.layers.0.weight"], P[f"read_level_encoder.{i}.layers.0.bias"]
            if not bn:
                return fold_batchnorm(w, b)
            pre = f"read_level_encoder.{i}.layers.1."
            return fold_batchnorm(w, b, P[pre + "weight"], P[pre + "bias"], P[pre + "running_mean"], P[pre + "running_var"], _BN_EPS)

        w1, b1 = lin(i1, self.bn1)
        w2, b2 = lin(i2, self.bn2)
        emb = P[f"read_level_encoder.{self._idx_emb}.embedding_layer.weight"] if self._idx_emb is not None else None
        return EncoderWeights(emb, np.ascontiguousarray(w1), np.ascontiguousarray(b1), np.ascontiguousarray(w2),
                              np.ascontiguousarray(b2), P["pooling_filter.probability_layer.0.weight"].reshape(-1).copy(),
                              P["pooling_filter.probability_layer.0.bias"].reshape(-1).copy())

    def engine(self, device="cuda"):
        """MilEngine (packed weights resident on `device`), cached per device."""
        from .engine import MilEngine
        key = str(device)
        if key not in self._engines:
            self._engines[key] = MilEngine(self.encoder_weights(), device)
        return self._engines[key]
