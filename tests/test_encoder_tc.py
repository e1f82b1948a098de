"""EXPERIMENTAL tensor-core read encoder (m6anet_b200/csrc/experimental/m6a_encoder_tc.cu, its own library
libm6a_encoder_tc.so; nothing in the product loads it).

The kernel was written after round 1's GPU budget was spent and has not run on hardware yet, so its GPU test is opt-in:
    M6A_TEST_TC=1 python -m pytest tests/test_encoder_tc.py -m gpu
It runs in a subprocess (a trapping kernel would poison the CUDA context of the whole pytest process) and holds the
tcgen05 3xTF32 encoder to the oracle within 5e-6 per read.
The CPU part checks what can be checked without a GPU: the library cross-compiles and exports its entry points, the packed
weight image read back through the UMMA K-major no-swizzle operand geometry (the very constants the shared-memory
descriptors are built from) is the folded weight matrix and its TF32 residual, the instruction descriptors decode to
M128 x N160 / N32 TF32 -> F32, and the 3-term TF32 arithmetic evaluated from that image keeps float32-level accuracy."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT, oracle_params

EXP_DIR = os.path.join(ROOT, "m6anet_b200", "csrc", "experimental")
EXP_LIB = os.path.join(EXP_DIR, "libm6a_encoder_tc.so")


@pytest.fixture(scope="module")
def tclib():
    res = subprocess.run(["make", "-C", EXP_DIR], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    L = C.CDLL(EXP_LIB)
    L.m6a_tc_image_bytes.restype = C.c_int64
    return L


def folded(tag):
    """product-side folded weights (BatchNorm folded in float64) as the C struct m6a_weights_t + the arrays behind it"""
    from m6anet_b200 import _cabi
    from test_gpu_parity import product_weights
    w = product_weights(tag)
    keep = {k: np.ascontiguousarray(getattr(w, k), dtype=np.float32) for k in ("w1", "b1", "w2", "b2", "w3", "b3")}
    emb = None if w.emb is None else np.ascontiguousarray(w.emb, dtype=np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    st = _cabi.M6AWeights(emb=ptr(emb), w1=ptr(keep["w1"]), b1=ptr(keep["b1"]), w2=ptr(keep["w2"]), b2=ptr(keep["b2"]),
                          w3=ptr(keep["w3"]), b3=ptr(keep["b3"]), n_kmer=w.n_kmer, emb_dim=w.emb_dim, n_sig=w.n_sig,
                          h1=w.h1, h2=w.h2)
    return st, keep, emb, w


def trunc_tf32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).reshape(x.shape)


def read_operand(img_f32, base_bytes, rows, k_total, lbo, sbo, kstep):
    """Element (row, k) exactly as a K-major SWIZZLE_NONE descriptor {start = base + (k // 8) * kstep, LBO, SBO} addresses
    it (cute::UMMA canonical layout ((8,m),(4,2)) : ((16 B, SBO), (4 B, LBO)))."""
    r, k = np.meshgrid(np.arange(rows), np.arange(k_total), indexing="ij")
    off = base_bytes + (k // 8) * kstep + ((k // 4) % 2) * lbo + (r // 8) * sbo + (r % 8) * 16 + (k % 4) * 4
    return img_f32[off // 4]


def test_library_exports(tclib):
    for sym in ("m6a_tc_create", "m6a_tc_destroy", "m6a_tc_read_probs_f32", "m6a_tc_mil_infer_f32", "m6a_tc_debug_image",
                "m6a_tc_geometry"):
        getattr(tclib, sym)
    assert tclib.m6a_tc_read_probs_f32(None, None, None, None, 0, 0, None, None) == -1        # M6A_EINVAL
    sass = subprocess.run(["cuobjdump", "-sass", EXP_LIB], capture_output=True, text=True).stdout
    # 6 + 12 tcgen05.mma per kernel: the stand-alone encoder and the two instantiations of the fused kernel
    assert sass.count("UTCHMMA") == 3 * 18 and "LDTM" in sass and "STTM" in sass and "UTCBAR" in sass
    assert tclib.m6a_tc_mil_infer_f32(None, None, None, None, 0, 0, 0, 20, 10, 0, C.c_float(0.5), None, None, None, None, 0,
                                      None) == -1


@pytest.mark.parametrize("tag", ["HCT116_RNA002", "HEK293T_RNA004", "signal_only"])
def test_packed_image_matches_descriptor_geometry_and_keeps_accuracy(tclib, tag, synthetic_inputs):
    from oracle import read_probabilities
    st, keep, emb, w = folded(tag)
    geo = np.zeros(16, dtype=np.int64)
    assert tclib.m6a_tc_geometry(geo.ctypes.data_as(C.c_void_p)) == 0
    o_w1, o_w1lo, o_w2, o_w2lo, o_b2, o_w3, o_b3, o_emb, sbo, lbo1, step1, lbo2, step2, idesc1, idesc2, tmem_cols = geo.tolist()
    img = np.zeros(int(tclib.m6a_tc_image_bytes()) // 4, dtype=np.float32)
    assert tclib.m6a_tc_debug_image(C.byref(st), img.ctypes.data_as(C.c_void_p), img.nbytes) == 0
    # instruction descriptors: c_format F32, a/b format TF32, K-major both, N >> 3, M >> 4
    for idesc, n in ((idesc1, 160), (idesc2, 32)):
        assert (idesc >> 4) & 3 == 1 and (idesc >> 7) & 7 == 2 and (idesc >> 10) & 7 == 2
        assert (idesc >> 15) & 3 == 0 and (idesc >> 17) & 63 == n >> 3 and (idesc >> 24) & 31 == 128 >> 4
    assert tmem_cols == 256 and sbo == 128
    # Linear-1 operand B: rows = hidden units (160), K = 16 = [9 signal | 3 x emb | 0.. | bias]
    E, h1 = w.emb_dim, w.h1
    in1 = 9 + 3 * E
    W1 = np.zeros((160, 16), dtype=np.float32)
    W1[:h1, :in1] = keep["w1"]
    W1[:h1, 15] = keep["b1"]
    # operands are stored TF32-exact (hi = trunc, lo = trunc(w - hi)): the split does not depend on operand conversion
    assert np.array_equal(read_operand(img, o_w1, 160, 16, lbo1, sbo, step1), trunc_tf32(W1))
    assert np.array_equal(read_operand(img, o_w1lo, 160, 16, lbo1, sbo, step1), trunc_tf32(W1 - trunc_tf32(W1)))
    W2 = np.zeros((32, 160), dtype=np.float32)
    W2[:, :h1] = keep["w2"]
    assert np.array_equal(read_operand(img, o_w2, 32, 160, lbo2, sbo, step2), trunc_tf32(W2))
    assert np.array_equal(read_operand(img, o_w2lo, 32, 160, lbo2, sbo, step2), trunc_tf32(W2 - trunc_tf32(W2)))
    assert np.array_equal(img[o_b2 // 4:o_b2 // 4 + 32], keep["b2"]) and np.array_equal(img[o_w3 // 4:o_w3 // 4 + 32], keep["w3"])
    assert img[o_b3 // 4] == keep["b3"][0]
    if emb is not None:
        assert np.array_equal(img[o_emb // 4:o_emb // 4 + emb.size], emb.reshape(-1))
    # the kernel's arithmetic from that image: A.B = A.B_hi + A_lo.B_hi + A.B_lo with TF32-truncated operands
    si = synthetic_inputs
    feats, off, kmer = si["feats"], si["read_off"], si["kmer_idx"]
    rows = np.repeat(kmer, np.diff(off), axis=0)
    X = np.zeros((len(feats), 16), dtype=np.float32)
    X[:, :9] = feats
    if emb is not None:
        X[:, 9:9 + 3 * E] = emb[rows].reshape(len(feats), 3 * E)
    X[:, 15] = 1.0

    def mm3(a, b):       # products exact, one float32 rounding per term (optimistic model of the TMEM accumulation)
        t, a64, b64 = trunc_tf32, a.astype(np.float64), None
        ah, bh = t(a).astype(np.float64), t(b).astype(np.float64)
        al, bl = t(a - t(a)).astype(np.float64), t(b - t(b)).astype(np.float64)
        return ((ah @ bh.T).astype(np.float32) + (al @ bh.T).astype(np.float32)) + (ah @ bl.T).astype(np.float32)

    H = np.maximum(mm3(X, W1), 0)
    D2 = mm3(H, W2)
    z = np.maximum(D2 + keep["b2"], 0) @ keep["w3"] + keep["b3"][0]
    p = 1 / (1 + np.exp(-z.astype(np.float32)))
    P = oracle_params(tag)
    want = read_probabilities(P, feats, None if P.emb is None else rows)
    assert np.abs(p - want).max() < 2e-6


def _cutlass_include():
    import importlib.util
    for pkg, sub in (("flashinfer", "data/cutlass/include"), ("tilelang", "3rdparty/cutlass/include")):
        spec = importlib.util.find_spec(pkg)
        if spec and spec.submodule_search_locations:
            path = os.path.join(list(spec.submodule_search_locations)[0], sub)
            if os.path.exists(os.path.join(path, "cute", "arch", "mma_sm100_desc.hpp")):
                return path
    return None


def test_descriptors_equal_cute_encoders(tclib, tmp_path):
    """The hand-packed instruction descriptors and the LBO / SBO / layout fields of the shared-memory descriptors equal
    what CuTe's own encoders (vendored CUTLASS headers) produce for the same shapes and operand layout; CuTe also
    statically asserts that the layout is a canonical UMMA K-major one (tools/cute_desc_check.cu, host run)."""
    inc = _cutlass_include()
    if inc is None:
        pytest.skip("no vendored CUTLASS headers in this environment")
    exe = tmp_path / "cute_desc_check"
    res = subprocess.run(["nvcc", "-std=c++17", f"-I{inc}", "-arch=sm_100a", "-o", str(exe),
                          os.path.join(ROOT, "tools", "cute_desc_check.cu")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout
    geo = np.zeros(16, dtype=np.int64)
    assert tclib.m6a_tc_geometry(geo.ctypes.data_as(C.c_void_p)) == 0
    sbo, lbo1, lbo2, idesc1, idesc2 = int(geo[8]), int(geo[9]), int(geo[11]), int(geo[13]), int(geo[14])
    idesc = {(int(a), int(b)): int(c) for _, a, b, c in (l.split() for l in out.splitlines() if l.startswith("idesc"))}
    ops = {int(l.split()[1]): tuple(int(x) for x in l.split()[2:]) for l in out.splitlines() if l.startswith("operand")}
    assert idesc[(128, 160)] == idesc1 and idesc[(128, 32)] == idesc2
    assert ops[160] == (lbo1 >> 4, sbo >> 4, 0, 1)          # W1: LBO, SBO in 16-byte units, SWIZZLE_NONE, version 1
    assert ops[32] == (lbo2 >> 4, sbo >> 4, 0, 1)           # W2
    assert ops[128] == (128 * 16 >> 4, sbo >> 4, 0, 1)      # X tile


WORKER = textwrap.dedent("""
    import ctypes as C, os, sys
    import numpy as np, torch
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from conftest import ALL_TAGS, GOLDEN, oracle_params
    from test_encoder_tc import EXP_LIB, folded
    from oracle import read_probabilities
    L = C.CDLL(EXP_LIB)
    z = np.load(os.path.join(GOLDEN, "synthetic_inputs.npz"))
    feats, off, kmer = z["feats"], z["read_off"], z["kmer_idx"]
    dev = torch.device("cuda:0")
    torch.zeros(1, device=dev)
    worst = 0.0
    for tag in ALL_TAGS:
        st, keep, emb, w = folded(tag)
        h = C.c_void_p()
        rc = L.m6a_tc_create(C.byref(st), C.byref(h)); assert rc == 0, rc
        for n_sites in (1, 5, len(off) - 1):                       # a partial tile, one tile, many tiles
            R = int(off[n_sites])
            f = torch.from_numpy(feats[:R].copy()).to(dev); o = torch.from_numpy(off[:n_sites + 1].copy()).to(dev)
            k = torch.from_numpy(kmer[:n_sites].copy()).to(dev)
            out = torch.full((R,), -1.0, dtype=torch.float32, device=dev)
            rc = L.m6a_tc_read_probs_f32(h, C.c_void_p(f.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(k.data_ptr()),
                                         C.c_int64(n_sites), C.c_int64(R), C.c_void_p(out.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, rc
            torch.cuda.synchronize()
            P = oracle_params(tag)
            rows = None if P.emb is None else np.repeat(kmer[:n_sites], np.diff(off[:n_sites + 1]), axis=0)
            want = read_probabilities(P, feats[:R], rows)
            d = float(np.abs(out.cpu().numpy() - want).max())
            print(tag, n_sites, R, "max|dp| = %.3e" % d, flush=True)
            worst = max(worst, d)
        L.m6a_tc_destroy(h)
    print("WORST %.3e" % worst)
    # the fused kernel: tensor-core encoder + the product's Monte-Carlo phase, against the oracle on the device stream
    from oracle import mil_inference
    st, keep, emb, w = folded("HCT116_RNA002")
    h = C.c_void_p(); assert L.m6a_tc_create(C.byref(st), C.byref(h)) == 0
    S, R = len(off) - 1, len(feats)
    f = torch.from_numpy(feats).to(dev); o = torch.from_numpy(off).to(dev); k = torch.from_numpy(kmer).to(dev)
    rp = torch.empty(R, device=dev); sp = torch.empty(S, device=dev); mc = torch.empty(S, dtype=torch.int32, device=dev)
    ws = torch.empty(R // 64 + 2, dtype=torch.int64, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())
    rc = L.m6a_tc_mil_infer_f32(h, vp(f), vp(o), vp(k), C.c_int64(S), C.c_int64(R), C.c_int64(7_000_000_000), 20, 200,
                                C.c_uint64(1234), C.c_float(0.033379376), vp(rp), vp(sp), vp(mc), vp(ws),
                                C.c_int64(ws.numel() * 8), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    torch.cuda.synchronize()
    orp, osp, omc = mil_inference(oracle_params("HCT116_RNA002"), feats, off, kmer, n_iters=200, seed=1234,
                                  site_id_base=7_000_000_000)
    d_read = float(np.abs(rp.cpu().numpy() - orp).max()); d_site = float(np.abs(sp.cpu().numpy() - osp).max())
    print("fused: max|dp_read| = %.3e  max|dp_site| = %.3e" % (d_read, d_site))
    sys.exit(0 if worst <= 5e-6 and d_read <= 5e-6 and d_site <= 1e-4 else 1)
""")


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("M6A_TEST_TC") != "1", reason="experimental kernel: opt in with M6A_TEST_TC=1")
def test_tc_encoder_matches_oracle(tmp_path, tclib):
    script = tmp_path / "tc_worker.py"
    script.write_text(WORKER.format(root=ROOT))
    res = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
