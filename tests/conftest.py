import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(ROOT, "m6anet_b200", "assets")

MODEL_FILES = {
    "HCT116_RNA002": "rna002_hct116.npz",
    "arabidopsis_RNA002": "rna002_arabidopsis_virc.npz",
    "HEK293T_RNA004": "rna004_hek293t_glori.npz",
    "HEK293T_RNA004_M6ACE": "rna004_hek293t_m6ace.npz",
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass without a device.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The suites need the in-tree C-ABI library (nvcc cross-compiles without a GPU); build it once if absent."""
    from m6anet_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    return _cabi.LIB_PATH


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def synthetic_inputs():
    z = np.load(os.path.join(GOLDEN, "synthetic_inputs.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def bundled_flat():
    z = np.load(os.path.join(GOLDEN, "bundled_flat.npz"))
    return {k: z[k] for k in z.files}


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, f"synthetic_{tag}.npz"))
    return {k: z[k] for k in z.files}


def oracle_params(tag):
    """ReadEncoderParams for a registry model name or 'signal_only' (weights stored in its golden file)."""
    from oracle import ReadEncoderParams
    if tag == "signal_only":
        g = load_golden(tag)
        return ReadEncoderParams(None, g["w_w1"], g["w_b1"], g["w_bn_gamma"], g["w_bn_beta"], g["w_bn_mean"],
                                 g["w_bn_var"], g["w_w2"], g["w_b2"], g["w_w3"], g["w_b3"])
    return ReadEncoderParams.from_npz(os.path.join(ASSETS, "model_states", MODEL_FILES[tag]))


ALL_TAGS = list(MODEL_FILES) + ["signal_only"]
