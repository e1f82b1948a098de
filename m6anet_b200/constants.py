"""Defaults, the pretrained-model registry and the k-mer tables of the inference path.

Mirrors reference m6anet/utils/constants.py:8-37 (names and meaning kept so that code written
against the reference keeps working); the asset files are this package's .npz conversions of the
reference's .pt/.joblib files (tools/convert_assets.py), bit-identical in value.
"""
from __future__ import annotations

import os
from itertools import product

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def _asset(*parts: str) -> str:
    return os.path.join(_ASSETS, *parts)


DEFAULT_PRETRAINED_MODELS = ['HCT116_RNA002', 'arabidopsis_RNA002', 'HEK293T_RNA004']   # constants.py:8
DEFAULT_PRETRAINED_MODEL = 'HCT116_RNA002'
DEFAULT_MODEL_CONFIG = _asset("model_configs", "m6anet.toml")
DEFAULT_MODEL_WEIGHTS = _asset("model_states", "rna002_hct116.npz")
DEFAULT_NORM_PATH = _asset("norm_factors", "rna002_hct116.npz")
DEFAULT_MIN_READS = 20
DEFAULT_READ_THRESHOLD = 0.033379376

ARABIDOPSIS_MODEL_WEIGHTS = _asset("model_states", "rna002_arabidopsis_virc.npz")
ARABIDOPSIS_NORM_PATH = _asset("norm_factors", "rna002_arabidopsis_virc.npz")
ARABIDOPSIS_READ_THRESHOLD = 0.0032978046219796

HEK293TRNA004_GLORI_MODEL_WEIGHTS = _asset("model_states", "rna004_hek293t_glori.npz")
HEK293TRNA004_M6ACE_MODEL_WEIGHTS = _asset("model_states", "rna004_hek293t_m6ace.npz")

# name -> (weights, read probability threshold, norm factors)          constants.py:24-27
PRETRAINED_CONFIGS = {
    'HCT116_RNA002': (DEFAULT_MODEL_WEIGHTS, DEFAULT_READ_THRESHOLD, DEFAULT_NORM_PATH),
    'arabidopsis_RNA002': (ARABIDOPSIS_MODEL_WEIGHTS, ARABIDOPSIS_READ_THRESHOLD, ARABIDOPSIS_NORM_PATH),
    'HEK293T_RNA004': (HEK293TRNA004_GLORI_MODEL_WEIGHTS, DEFAULT_READ_THRESHOLD, DEFAULT_NORM_PATH),
    'HEK293T_RNA004_M6ACE': (HEK293TRNA004_M6ACE_MODEL_WEIGHTS, DEFAULT_READ_THRESHOLD, DEFAULT_NORM_PATH),
}

NUM_NEIGHBORING_FEATURES = 1
CENTER_MOTIFS = [['A', 'G', 'T'], ['G', 'A'], ['A'], ['C'], ['A', 'C', 'T']]      # DRACH


def kmer_table(num_neighboring_features: int = NUM_NEIGHBORING_FEATURES):
    """Sorted unique 5-mers of every DRACH-centred (5 + 2n)-mer (constants.py:29-36: np.unique => lexicographic)."""
    flank = [['G', 'A', 'C', 'T']] * num_neighboring_features
    full = ["".join(x) for x in product(*(flank + CENTER_MOTIFS + flank))]
    five = sorted({s[i:i + 5] for s in full for i in range(len(s) - 4)})
    return np.array(five)


ALL_KMERS = kmer_table()
KMER_TO_INT = {str(k): i for i, k in enumerate(ALL_KMERS)}
INT_TO_KMER = {i: str(k) for i, k in enumerate(ALL_KMERS)}
M6A_KMERS = ["".join(x) for x in product(*CENTER_MOTIFS)]
