"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/m6anet_b200.h declares (no compute calls without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "m6anet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(m6a_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from m6anet_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    return _cabi.lib()


def test_header_symbols_are_exported(lib):
    from m6anet_b200 import _cabi
    syms = declared_symbols()
    assert syms and set(syms) == set(_cabi.EXPORTS)
    for s in syms:
        assert getattr(lib, s) is not None


def test_version_and_strerror(lib):
    assert lib.m6a_version() == 100
    assert lib.m6a_strerror(0) == b"ok"
    assert b"invalid" in lib.m6a_strerror(-1)
    assert b"supported" in lib.m6a_strerror(-2)


def test_library_targets_sm100a():
    import subprocess
    from m6anet_b200 import _cabi
    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from m6anet_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.lib()


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from m6anet_b200 import weights as W
    from m6anet_b200.engine import MilEngine
    w = W.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna002_hct116.npz"))
    with pytest.raises(RuntimeError, match="CUDA"):
        MilEngine(w)


def test_batchnorm_fold_matches_unfolded_oracle():
    """host logic: the float64 BatchNorm fold reproduces the unfolded restatement to float32 rounding."""
    import numpy as np
    from m6anet_b200 import weights as W
    from oracle import ReadEncoderParams, read_probabilities
    npz = os.path.join(ROOT, "m6anet_b200", "assets", "model_states", "rna004_hek293t_glori.npz")
    w = W.from_npz(npz)
    assert (w.h1, w.h2, w.n_sig, w.emb_dim, w.n_kmer) == (150, 32, 9, 2, 66)
    raw = ReadEncoderParams.from_npz(npz)
    folded = ReadEncoderParams(w.emb, w.w1, w.b1, np.ones(150, np.float32), np.zeros(150, np.float32),
                               np.zeros(150, np.float32), np.ones(150, np.float32), w.w2, w.b2, w.w3.reshape(1, -1), w.b3,
                               bn_eps=0.0)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4096, 9), dtype=np.float32)
    k = rng.integers(0, 66, size=(4096, 3))
    assert np.max(np.abs(read_probabilities(raw, x, k) - read_probabilities(folded, x, k))) <= 2e-6


def test_sass_shows_the_blackwell_native_paths():
    """cuobjdump evidence (no GPU needed).  CUDA-core kernel: TMA bulk copy + mbarrier for the feature tiles, packed FFMA2 fed
    from uniform registers (weights as a __grid_constant__ parameter), no tensor-core instruction.  Tensor-core kernel:
    tcgen05.mma (UTCHMMA) with TMEM loads / stores (LDTM / STTM) and MMA-completion barriers (UTCBAR); never the legacy
    mma.sync path (HMMA)."""
    import re
    import subprocess
    from m6anet_b200 import _cabi
    sass = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass                       # cp.async.bulk + mbarrier transaction count
    assert re.search(r"FFMA2 .*UR\d+\.F32x2", sass)                    # uniform-register pair operand
    assert re.search(r"LDCU\.64 UR\d+, c\[0x0\]\[UR\d+", sass)          # weights streamed from the parameter bank
    ffma = sass.split("mil_infer_kernelILi20")[1].split("Function :")[0]
    assert "UTCHMMA" not in ffma and "LDL" not in ffma                 # CUDA-core kernel: no tensor cores, no spills
    tc = sass.split("mil_infer_tc_kernelILi20")[1].split("Function :")[0]
    assert tc.count("UTCHMMA") >= 14 and "LDTM" in tc and "STTM" in tc and "UTCBAR" in tc and "LDGSTS" in tc
    # the MMAs of a Linear-2 chunk are issued back to back by one elected lane of a converged warp (not one lane-loop per MMA)
    ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", tc)
    runs = max(len(m.group(0).split()) for m in re.finditer(r"(?:UTCHMMA )+", " ".join(ops) + " "))
    assert runs >= 8 and "ELECT" in tc
    assert not re.search(r"\bHMMA\b", sass)                            # no legacy mma.sync anywhere


def test_auto_tile_reads_policy(lib):
    """Rows per tile: a multiple of the site depth near 1000 (constant depth), 1000 for uneven sites, ~500 for small jobs."""
    t = lambda s, r: lib.m6a_auto_tile_reads(s, r, 148)
    assert t(1_000_000, 50_000_000) == 1000            # headline job: 20 sites of 50 reads
    assert t(1_000_000, 20_000_000) == 1000            # 50 sites of 20 reads (<= 64 sites per slice)
    assert t(500_000, 15_000_000) == 990               # 33 sites of 30 reads
    assert t(250_000, 20_000_000) == 960               # 12 pooled sites of 80 reads
    assert t(125_000, 6_250_000) == 500                # one N=8 shard: smaller tiles, >= 48 tiles per CTA
    assert t(1_000_000, 47_900_123) == 1000            # uneven sites
    assert t(10, 7) == 500 and t(0, 0) >= 64
    assert lib.m6a_mil_workspace_bytes(50_000_000) >= (50_000_000 // 500 + 2) * 8


def test_auto_tile_reads_policy_tensor_core_kernel(lib):
    """Rows per tile of the tensor-core kernel: 64 sites' worth (no padded MMA rows for even depths), and for jobs of fewer
    than 64 rounds of tiles a tile shrunk so that every SM gets the same number of whole tiles (multiples of 128 rows)."""
    t = lambda s, r: lib.m6a_auto_tile_reads_tc(s, r, 148)
    assert t(1_000_000, 50_000_000) == 3200            # headline job: 64 sites of 50 reads, 105.6 -> 106 rounds: unchanged
    assert t(1_000_000, 20_000_000) == 1280            # config 2 shape: 64 sites of 20 reads
    assert t(125_000, 6_250_000) == 3072               # one N=8 shard: 13.2 rounds of 3200 -> 14 rounds of 3072
    assert t(1000, 50_000) == 384                      # one tile of 3 MMA tiles per SM (the shape that once lost a hand-over)
    assert t(288, 14_108) == 2048                      # ragged and tiny: below 256 rows per SM the default stays
    assert t(1_000_000, 47_900_123) == 2048            # uneven sites, long job
    for s, r in [(1, 20), (10, 7), (0, 0), (5_000, 250_000), (40, 16_000), (3, 30_000)]:
        v = t(s, r)
        assert 64 <= v <= 6144
        assert lib.m6a_mil_workspace_bytes(r) >= (r // v + 3) * 8          # bounds + counter fit the documented workspace
