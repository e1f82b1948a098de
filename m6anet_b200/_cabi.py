"""ctypes binding of the C ABI declared in include/m6anet_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C m6anet_b200/csrc``.
There is deliberately no fallback: if the library is missing, using the symbols raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# M6A_LIB selects an alternative build of the same library (kernel A/B experiments); never a fallback.
LIB_PATH = os.environ.get("M6A_LIB") or os.path.join(_HERE, "libm6anet_b200.so")
CSRC = os.path.join(_HERE, "csrc")

EXPORTS = (
    "m6a_version", "m6a_strerror", "m6a_device_count", "m6a_set_device", "m6a_model_create", "m6a_model_destroy", "m6a_model_set_tile_reads", "m6a_auto_tile_reads", "m6a_auto_tile_reads_tc", "m6a_mil_workspace_bytes",
    "m6a_mil_infer_f32", "m6a_mil_infer_packed_f32", "m6a_model_set_encoder", "m6a_model_get_encoder", "m6a_debug_trap_record",
    "m6a_pinned_alloc", "m6a_pinned_free", "m6a_build_info",
    "m6a_mil_infer_host_f32", "m6a_sample_indices", "m6a_mil_validate_f32", "m6a_mil_validate_host_f32", "m6a_sample_bags",
    "m6a_last_launch", "m6a_ingest_parts", "m6a_ingest_parts_keyed", "m6a_info_count", "m6a_info_read",
    "m6a_write_site_csv",
    "m6a_write_indiv_csv",
)

# struct m6a_part_t as a NumPy record (56 bytes, natural alignment)
import numpy as _np
PART_DTYPE = _np.dtype([("file", "<i4"), ("rep", "<i4"), ("start", "<i8"), ("end", "<i8"), ("row_off", "<i8"),
                        ("n_rows", "<i8"), ("site", "<i8"), ("first_of_site", "<i4"), ("reserved", "<i4")], align=True)
assert PART_DTYPE.itemsize == 56


class M6AWeights(C.Structure):
    """struct m6a_weights_t"""
    _fields_ = [
        ("emb", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("n_kmer", C.c_int32), ("emb_dim", C.c_int32), ("n_sig", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32),
    ]


class M6AError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib().m6a_strerror(status).decode() if _lib is not None else "?"
        super().__init__(f"{where}: m6anet_b200 status {status} ({msg})")


_lib = None


def build(verbose: bool = False) -> str:
    """Compile libm6anet_b200.so for sm_100a with nvcc (in-tree)."""
    res = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libm6anet_b200.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension was not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C m6anet_b200/csrc`. m6anet_b200 has no CPU fallback for the hot path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    L.m6a_version.restype = C.c_int
    L.m6a_version.argtypes = []
    L.m6a_strerror.restype = C.c_char_p
    L.m6a_strerror.argtypes = [C.c_int]
    L.m6a_device_count.restype = C.c_int
    L.m6a_device_count.argtypes = [C.POINTER(i32)]
    L.m6a_set_device.restype = C.c_int
    L.m6a_set_device.argtypes = [i32]
    L.m6a_model_create.restype = C.c_int
    L.m6a_model_create.argtypes = [C.POINTER(M6AWeights), C.POINTER(vp)]
    L.m6a_model_destroy.restype = C.c_int
    L.m6a_model_destroy.argtypes = [vp]
    L.m6a_model_set_tile_reads.restype = C.c_int
    L.m6a_model_set_tile_reads.argtypes = [vp, i32]
    L.m6a_mil_infer_f32.restype = C.c_int
    L.m6a_mil_infer_f32.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, u64, vp, f32, vp, vp, vp, vp, i64, vp]
    L.m6a_mil_infer_packed_f32.restype = C.c_int
    L.m6a_mil_infer_packed_f32.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, u64, f32, vp, vp, vp, i64, vp]
    L.m6a_model_set_encoder.restype = C.c_int
    L.m6a_model_set_encoder.argtypes = [vp, i32]
    L.m6a_model_get_encoder.restype = C.c_int
    L.m6a_model_get_encoder.argtypes = [vp]
    L.m6a_debug_trap_record.restype = C.c_int
    L.m6a_debug_trap_record.argtypes = [vp, vp]
    L.m6a_pinned_alloc.restype = C.c_int
    L.m6a_pinned_alloc.argtypes = [C.POINTER(vp), i64]
    L.m6a_pinned_free.restype = C.c_int
    L.m6a_pinned_free.argtypes = [vp]
    L.m6a_build_info.restype = C.c_char_p
    L.m6a_build_info.argtypes = []
    L.m6a_auto_tile_reads.restype = i32
    L.m6a_auto_tile_reads.argtypes = [i64, i64, i32]
    L.m6a_auto_tile_reads_tc.restype = i32
    L.m6a_auto_tile_reads_tc.argtypes = [i64, i64, i32]
    L.m6a_mil_workspace_bytes.restype = i64
    L.m6a_mil_workspace_bytes.argtypes = [i64]
    L.m6a_mil_infer_host_f32.restype = C.c_int
    L.m6a_mil_infer_host_f32.argtypes = [vp, vp, vp, vp, i64, i64, i32, i32, u64, f32, vp, vp, vp, i32]
    L.m6a_sample_indices.restype = C.c_int
    L.m6a_sample_indices.argtypes = [u64, i64, i32, i32, i32, vp, vp]
    L.m6a_mil_validate_f32.restype = C.c_int
    L.m6a_mil_validate_f32.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, u64, vp, i32, i32, f32, vp, vp, vp, vp, vp,
                                       i64, vp]
    L.m6a_mil_validate_host_f32.restype = C.c_int
    L.m6a_mil_validate_host_f32.argtypes = [vp, vp, vp, vp, i64, i64, i32, i32, u64, i32, i32, f32, vp, vp, vp, vp, i32]
    L.m6a_sample_bags.restype = C.c_int
    L.m6a_sample_bags.argtypes = [u64, i64, i32, i32, i32, vp, vp]
    L.m6a_last_launch.restype = C.c_int
    L.m6a_last_launch.argtypes = [C.POINTER(i32)] * 5
    L.m6a_ingest_parts.restype = C.c_int
    L.m6a_ingest_parts.argtypes = [C.POINTER(C.c_char_p), i32, vp, i64, i32, vp, vp, vp, vp, vp, vp, i32, C.POINTER(i64)]
    L.m6a_ingest_parts_keyed.restype = C.c_int
    L.m6a_ingest_parts_keyed.argtypes = [C.POINTER(C.c_char_p), i32, vp, i64, i32, vp, vp, vp, C.c_char_p, vp, vp, vp, vp, vp, i32,
                                         C.POINTER(i64)]
    L.m6a_info_count.restype = C.c_int
    L.m6a_info_count.argtypes = [C.c_char_p, C.POINTER(i64), C.POINTER(i64)]
    L.m6a_info_read.restype = C.c_int
    L.m6a_info_read.argtypes = [C.c_char_p, i64, i64, vp, vp, vp, vp, vp, vp]
    L.m6a_write_site_csv.restype = C.c_int
    L.m6a_write_site_csv.argtypes = [i32, i64, vp, vp, vp, vp, vp, vp, vp, i32]
    L.m6a_write_indiv_csv.restype = C.c_int
    L.m6a_write_indiv_csv.argtypes = [i32, i64, vp, vp, vp, vp, vp, vp, vp, i32]
    _lib = L
    return L


def device_count() -> int:
    """Number of visible CUDA devices (0 when there is no driver / GPU); does not import torch."""
    n = C.c_int32(0)
    lib().m6a_device_count(C.byref(n))
    return int(n.value)


def parse_device(device) -> int:
    """'cuda' / 'cuda:N' / N / torch.device -> device index (None = current default 0)."""
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    idx = getattr(device, "index", None)
    if hasattr(device, "type"):
        if device.type != "cuda":
            raise ValueError(f"m6anet_b200 needs a cuda device, got {device}")
        return 0 if idx is None else int(idx)
    text = str(device)
    if not text.startswith("cuda"):
        raise ValueError(f"m6anet_b200 needs a cuda device, got {device!r}")
    return int(text.split(":", 1)[1]) if ":" in text else 0


def check(status: int, where: str) -> None:
    if status != 0:
        raise M6AError(status, where)


ENCODERS = {"ffma": 0, "tc": 1}     # M6A_ENCODER_FFMA / M6A_ENCODER_TC


class _PinnedOwner:
    """Keeps one cudaHostAlloc block alive for the NumPy array that views it."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr and _lib is not None:
                _lib.m6a_pinned_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype) -> "_np.ndarray":
    """np.empty in page-locked host memory (m6a_pinned_alloc): the H2D / D2H copies of the host-buffer entry points only
    overlap with the kernel when their buffers are pinned.  The array (and every view of it) keeps the block alive."""
    dtype = _np.dtype(dtype)
    shape = (shape,) if isinstance(shape, (int, _np.integer)) else tuple(int(x) for x in shape)
    n_bytes = int(_np.prod(shape, dtype=_np.int64)) * dtype.itemsize
    if n_bytes == 0:
        return _np.empty(shape, dtype=dtype)
    p = C.c_void_p()
    check(lib().m6a_pinned_alloc(C.byref(p), n_bytes), "m6a_pinned_alloc")
    owner = _PinnedOwner(p)
    buf = (C.c_char * n_bytes).from_address(p.value)
    buf._owner = owner                                    # the ctypes buffer is the base object of the array
    return _np.frombuffer(buf, dtype=dtype).reshape(shape)


class PinnedPool:
    """Recycles page-locked blocks (cudaHostAlloc costs milliseconds per call): `empty` hands out an array backed by a free
    block of at least the requested size, `give_back` returns the block of an array obtained from this pool."""

    def __init__(self):
        import threading
        self._free = []          # (n_bytes, uint8 array)
        self._lent = {}          # data pointer -> uint8 block
        self._lock = threading.Lock()     # ingest, GPU and emit threads share one pool

    def empty(self, shape, dtype):
        dtype = _np.dtype(dtype)
        shape = (shape,) if isinstance(shape, (int, _np.integer)) else tuple(int(x) for x in shape)
        n_bytes = int(_np.prod(shape, dtype=_np.int64)) * dtype.itemsize
        if n_bytes == 0:
            return _np.empty(shape, dtype=dtype)
        with self._lock:
            best = None
            for i, (cap, _) in enumerate(self._free):
                if cap >= n_bytes and (best is None or cap < self._free[best][0]):
                    best = i
            block = None if best is None else self._free.pop(best)[1]
        if block is None:
            block = pinned_empty(n_bytes + n_bytes // 4 + 4096, _np.uint8)
        out = block[:n_bytes].view(dtype).reshape(shape)
        with self._lock:
            self._lent[out.ctypes.data] = block
        return out

    def give_back(self, *arrays):
        for a in arrays:
            if a is None or not hasattr(a, "ctypes") or a.size == 0:
                continue
            with self._lock:
                block = self._lent.pop(a.ctypes.data, None)
                if block is not None:
                    self._free.append((block.nbytes, block))
