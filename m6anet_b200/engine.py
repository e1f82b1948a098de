"""MilEngine: Python handle on the C ABI (include/m6anet_b200.h).  PyTorch is used only for device
memory and streams; every number comes out of the CUDA kernel."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _cabi
from .weights import EncoderWeights

DEFAULT_N_SAMPLES = 20   # the literal 20 of reference utils/inference_utils.py:54
POOLING = {"prod": 0, "mean": 1, "max": 2}   # M6A_POOL_* (reference SigmoidProdPooling / SigmoidMeanPooling / SigmoidMaxPooling)


def _pool_code(pooling) -> int:
    if isinstance(pooling, str):
        if pooling not in POOLING:
            raise ValueError(f"unknown pooling {pooling!r}; one of {sorted(POOLING)}")
        return POOLING[pooling]
    return int(pooling)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class MilEngine:
    """One packed model resident on one CUDA device."""

    def __init__(self, weights: EncoderWeights, device: "int | str | None" = None):
        self._lib = _cabi.lib()
        if _cabi.device_count() < 1:
            raise RuntimeError("m6anet_b200 needs a CUDA device (the hot path has no CPU fallback)")
        self.device_index = _cabi.parse_device(device)
        if self.device_index >= _cabi.device_count():
            raise ValueError(f"cuda:{self.device_index} is not visible ({_cabi.device_count()} device(s))")
        self.weights = weights
        keep = {k: np.ascontiguousarray(getattr(weights, k), dtype=np.float32)
                for k in ("w1", "b1", "w2", "b2", "w3", "b3")}
        emb = None if weights.emb is None else np.ascontiguousarray(weights.emb, dtype=np.float32)
        w = _cabi.M6AWeights(emb=_ptr(emb), w1=_ptr(keep["w1"]), b1=_ptr(keep["b1"]), w2=_ptr(keep["w2"]),
                             b2=_ptr(keep["b2"]), w3=_ptr(keep["w3"]), b3=_ptr(keep["b3"]),
                             n_kmer=weights.n_kmer, emb_dim=weights.emb_dim, n_sig=weights.n_sig,
                             h1=weights.h1, h2=weights.h2)
        handle = C.c_void_p()
        self._select()
        _cabi.check(self._lib.m6a_model_create(C.byref(w), C.byref(handle)), "m6a_model_create")
        self._handle = handle

    def _select(self):
        """Make this engine's device current for the calling thread (cudaSetDevice)."""
        _cabi.check(self._lib.m6a_set_device(self.device_index), "m6a_set_device")

    @property
    def _torch(self):
        import torch       # only the tensor-facing calls need torch; the host-buffer path does not
        return torch

    @property
    def device(self):
        return self._torch.device("cuda", self.device_index)

    def set_encoder(self, encoder: str):
        """'tc' (tcgen05 tensor-core read encoder, the default) or 'ffma' (CUDA-core FFMA2 encoder)."""
        if encoder not in _cabi.ENCODERS:
            raise ValueError(f"unknown encoder {encoder!r}; one of {sorted(_cabi.ENCODERS)}")
        _cabi.check(self._lib.m6a_model_set_encoder(self._handle, _cabi.ENCODERS[encoder]), "m6a_model_set_encoder")

    @property
    def encoder(self) -> str:
        return "tc" if self._lib.m6a_model_get_encoder(self._handle) == 1 else "ffma"

    def trap_record(self):
        """Arms (first call) / reads the debug record of the tensor-core kernel's bounded waits: [site, block, thread, parity]."""
        out = (C.c_int32 * 64)()
        _cabi.check(self._lib.m6a_debug_trap_record(self._handle, out), "m6a_debug_trap_record")
        v = list(out)
        return [v[4 * i:4 * i + 4] for i in range(16) if v[4 * i] != 0]

    def set_tile_reads(self, tile_reads: int = 0):
        """Feature rows per tile (64..4096); 0 = automatic (a multiple of the site depth near 1000 rows)."""
        _cabi.check(self._lib.m6a_model_set_tile_reads(self._handle, int(tile_reads)), "m6a_model_set_tile_reads")

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.m6a_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device-resident call (m6a_mil_infer_f32) -------------------------------------------------
    def infer_device(self, feats, read_off, kmer_idx, n_iters: int, seed: int = 0, site_id_base: int = 0,
                     n_samples: int = DEFAULT_N_SAMPLES, read_threshold: float = 0.033379376,
                     sample_idx=None, out=None, stream=None):
        """feats [R,9] f32, read_off [S+1] i64, kmer_idx [S,3] i32 -- CUDA tensors on self.device.
        Enqueues on `stream` (default: torch's current stream); returns (read_prob, site_prob, mod_count)
        CUDA tensors without synchronising."""
        torch = self._torch

        def need(cond, msg):
            if not cond:
                raise ValueError("MilEngine.infer_device: " + msg)

        need(feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous() and feats.dim() == 2 and feats.shape[1] == 9,
             "feats must be a contiguous float32 CUDA tensor [reads, 9]")
        need(read_off.is_cuda and read_off.dtype == torch.int64 and read_off.is_contiguous() and read_off.numel() >= 1,
             "read_off must be a contiguous int64 CUDA tensor [sites + 1]")
        need(feats.device == self.device and read_off.device == self.device, f"tensors must live on {self.device}")
        n_sites = read_off.numel() - 1
        total_reads = feats.shape[0]
        if kmer_idx is not None:
            need(kmer_idx.is_cuda and kmer_idx.dtype == torch.int32 and kmer_idx.is_contiguous() and kmer_idx.numel() == 3 * n_sites,
                 "kmer_idx must be a contiguous int32 CUDA tensor [sites, 3]")
        if sample_idx is not None:
            need(sample_idx.is_cuda and sample_idx.dtype == torch.uint16 and sample_idx.is_contiguous()
                 and sample_idx.numel() == n_sites * n_iters * n_samples,
                 "sample_idx must be a contiguous uint16 CUDA tensor [sites, n_iters, n_samples]")
        if out is None:
            read_prob = torch.empty(total_reads, dtype=torch.float32, device=self.device)
            site_prob = torch.empty(n_sites, dtype=torch.float32, device=self.device)
            mod_count = torch.empty(n_sites, dtype=torch.int32, device=self.device)
        else:
            read_prob, site_prob, mod_count = out
        # scratch for the tile-boundary prepass; torch's caching allocator keeps it stream-safe and cheap
        ws_bytes = int(self._lib.m6a_mil_workspace_bytes(total_reads))
        workspace = torch.empty(ws_bytes // 8, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device) if stream is None else stream
            if stream is not None:
                workspace.record_stream(st)
            rc = self._lib.m6a_mil_infer_f32(
                self._handle, feats.data_ptr(), read_off.data_ptr(), None if kmer_idx is None else kmer_idx.data_ptr(),
                n_sites, total_reads, site_id_base, n_samples, n_iters, seed & 0xFFFFFFFFFFFFFFFF,
                None if sample_idx is None else sample_idx.data_ptr(), read_threshold,
                read_prob.data_ptr(), site_prob.data_ptr(), mod_count.data_ptr(), workspace.data_ptr(), ws_bytes,
                st.cuda_stream)
        _cabi.check(rc, "m6a_mil_infer_f32")
        return read_prob, site_prob, mod_count

    def infer_device_packed(self, feats, read_off, kmer_idx, n_iters: int, seed: int = 0, site_id_base: int = 0,
                            n_samples: int = DEFAULT_N_SAMPLES, read_threshold: float = 0.033379376, read_prob=None,
                            site_out=None, stream=None):
        """m6a_mil_infer_packed_f32: like infer_device, but the kernel writes the two per-site outputs interleaved into
        `site_out` [>= sites, 2] float32 words (site_prob, mod_count bits) -- the send buffer of the all-gather of a
        multi-GPU run, so no pack kernels run.  Returns (read_prob, site_out)."""
        torch = self._torch
        n_sites = read_off.numel() - 1
        total_reads = feats.shape[0]
        if not (feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous() and feats.dim() == 2 and feats.shape[1] == 9):
            raise ValueError("MilEngine.infer_device_packed: feats must be a contiguous float32 CUDA tensor [reads, 9]")
        if not (read_off.is_cuda and read_off.dtype == torch.int64 and read_off.is_contiguous() and read_off.numel() >= 1):
            raise ValueError("MilEngine.infer_device_packed: read_off must be a contiguous int64 CUDA tensor [sites + 1]")
        if kmer_idx is not None and not (kmer_idx.is_cuda and kmer_idx.dtype == torch.int32 and kmer_idx.is_contiguous()
                                         and kmer_idx.numel() == 3 * n_sites):
            raise ValueError("MilEngine.infer_device_packed: kmer_idx must be a contiguous int32 CUDA tensor [sites, 3]")
        if read_prob is None:
            read_prob = torch.empty(total_reads, dtype=torch.float32, device=self.device)
        if site_out is None:
            site_out = torch.empty((n_sites, 2), dtype=torch.float32, device=self.device)
        if not (site_out.is_cuda and site_out.dtype == torch.float32 and site_out.is_contiguous() and site_out.dim() == 2
                and site_out.shape[1] == 2 and site_out.shape[0] >= n_sites and read_prob.numel() >= total_reads):
            raise ValueError("MilEngine.infer_device_packed: site_out must be a contiguous float32 CUDA tensor [>= sites, 2]")
        ws_bytes = int(self._lib.m6a_mil_workspace_bytes(total_reads))
        workspace = torch.empty(ws_bytes // 8, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device) if stream is None else stream
            if stream is not None:
                workspace.record_stream(st)
            rc = self._lib.m6a_mil_infer_packed_f32(
                self._handle, feats.data_ptr(), read_off.data_ptr(), None if kmer_idx is None else kmer_idx.data_ptr(),
                n_sites, total_reads, site_id_base, n_samples, n_iters, seed & 0xFFFFFFFFFFFFFFFF, read_threshold,
                read_prob.data_ptr(), site_out.data_ptr(), workspace.data_ptr(), ws_bytes, st.cuda_stream)
        _cabi.check(rc, "m6a_mil_infer_packed_f32")
        return read_prob, site_out

    # ---- host-buffer call (m6a_mil_infer_host_f32): chunked H2D / kernel / D2H pipeline -------------
    def infer_host(self, feats: np.ndarray, read_off: np.ndarray, kmer_idx: Optional[np.ndarray], n_iters: int,
                   seed: int = 0, site_id_base: int = 0, n_samples: int = DEFAULT_N_SAMPLES,
                   read_threshold: float = 0.033379376, n_chunks: int = 0, out=None, alloc=None
                   ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """NumPy buffers (page-locked ones overlap H2D / kernel / D2H).  `alloc(shape, dtype)` supplies the output arrays
        (e.g. _cabi.PinnedPool.empty); default np.empty."""
        feats, read_off, kmer_idx, n_sites = self._check_host(feats, read_off, kmer_idx, "infer_host")
        empty = alloc or np.empty
        if out is None:
            read_prob = empty(feats.shape[0], np.float32)
            site_prob = empty(n_sites, np.float32)
            mod_count = empty(n_sites, np.int32)
        else:
            read_prob, site_prob, mod_count = out
            if not (read_prob.size >= feats.shape[0] and site_prob.size >= n_sites and mod_count.size >= n_sites):
                raise ValueError("MilEngine.infer_host: output arrays are too small")
        self._select()
        rc = self._lib.m6a_mil_infer_host_f32(
            self._handle, _ptr(feats), _ptr(read_off), _ptr(kmer_idx), n_sites, site_id_base, n_samples, n_iters,
            seed & 0xFFFFFFFFFFFFFFFF, read_threshold, _ptr(read_prob), _ptr(site_prob), _ptr(mod_count), n_chunks)
        _cabi.check(rc, "m6a_mil_infer_host_f32")
        return read_prob, site_prob, mod_count

    def _check_host(self, feats, read_off, kmer_idx, where):
        """Shape / consistency checks of the host-buffer entry points (the C ABI trusts its arguments)."""
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        if feats.ndim != 2 or feats.shape[1] != 9:
            raise ValueError(f"MilEngine.{where}: feats must be [reads, 9] float32 (num_neighboring_features = 1), got {feats.shape}")
        if read_off.ndim != 1 or read_off.size < 1 or read_off[0] != 0 or read_off[-1] != feats.shape[0]:
            raise ValueError(f"MilEngine.{where}: read_off must be [sites + 1] with read_off[0] == 0 and read_off[-1] == len(feats)")
        if read_off.size > 1 and np.any(np.diff(read_off) < 0):
            raise ValueError(f"MilEngine.{where}: read_off must be non-decreasing")
        n_sites = read_off.size - 1
        if kmer_idx is not None:
            kmer_idx = np.ascontiguousarray(kmer_idx, dtype=np.int32)
            if kmer_idx.shape != (n_sites, 3):
                raise ValueError(f"MilEngine.{where}: kmer_idx must be [sites, 3] int32, got {kmer_idx.shape}")
        elif self.weights.emb is not None:
            raise ValueError(f"MilEngine.{where}: this model has a k-mer embedding, kmer_idx is required")
        return feats, read_off, kmer_idx, n_sites

    # ---- validate()-style literal MIL forward (m6a_mil_validate_f32 / m6a_mil_validate_host_f32) ----------------
    def validate_device(self, feats, read_off, kmer_idx, n_iters: int, seed: int = 0, site_id_base: int = 0,
                        n_samples: int = DEFAULT_N_SAMPLES, pooling="prod", replace: bool = False,
                        read_threshold: float = 0.033379376, sample_idx=None, stream=None):
        """CUDA tensors as infer_device.  Every pass pools one bag of n_samples reads per site (drawn without
        replacement unless replace=True or explicit `sample_idx` [sites, n_iters, n_samples] uint16 bags are given).
        Returns (read_prob [R], bag_prob [sites, n_iters], site_mean [sites], mod_count [sites]) without synchronising."""
        torch = self._torch

        def need(cond, msg):
            if not cond:
                raise ValueError("MilEngine.validate_device: " + msg)

        need(feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous() and feats.dim() == 2 and feats.shape[1] == 9,
             "feats must be a contiguous float32 CUDA tensor [reads, 9]")
        need(read_off.is_cuda and read_off.dtype == torch.int64 and read_off.is_contiguous() and read_off.numel() >= 1,
             "read_off must be a contiguous int64 CUDA tensor [sites + 1]")
        need(feats.device == self.device and read_off.device == self.device, f"tensors must live on {self.device}")
        n_sites = read_off.numel() - 1
        total_reads = feats.shape[0]
        if kmer_idx is not None:
            need(kmer_idx.is_cuda and kmer_idx.dtype == torch.int32 and kmer_idx.is_contiguous() and kmer_idx.numel() == 3 * n_sites,
                 "kmer_idx must be a contiguous int32 CUDA tensor [sites, 3]")
        if sample_idx is not None:
            need(sample_idx.is_cuda and sample_idx.dtype == torch.uint16 and sample_idx.is_contiguous()
                 and sample_idx.numel() == n_sites * n_iters * n_samples,
                 "sample_idx must be a contiguous uint16 CUDA tensor [sites, n_iters, n_samples]")
        read_prob = torch.empty(total_reads, dtype=torch.float32, device=self.device)
        bag_prob = torch.empty((n_sites, n_iters), dtype=torch.float32, device=self.device)
        site_prob = torch.empty(n_sites, dtype=torch.float32, device=self.device)
        mod_count = torch.empty(n_sites, dtype=torch.int32, device=self.device)
        ws_bytes = int(self._lib.m6a_mil_workspace_bytes(total_reads))
        workspace = torch.empty(ws_bytes // 8, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device) if stream is None else stream
            if stream is not None:
                for t in (workspace, read_prob, bag_prob, site_prob, mod_count):
                    t.record_stream(st)
            rc = self._lib.m6a_mil_validate_f32(
                self._handle, feats.data_ptr(), read_off.data_ptr(), None if kmer_idx is None else kmer_idx.data_ptr(),
                n_sites, total_reads, site_id_base, n_samples, n_iters, seed & 0xFFFFFFFFFFFFFFFF,
                None if sample_idx is None else sample_idx.data_ptr(), _pool_code(pooling), 1 if replace else 0,
                read_threshold, read_prob.data_ptr(), bag_prob.data_ptr(), site_prob.data_ptr(), mod_count.data_ptr(),
                workspace.data_ptr(), ws_bytes, st.cuda_stream)
        _cabi.check(rc, "m6a_mil_validate_f32")
        return read_prob, bag_prob, site_prob, mod_count

    def validate_host(self, feats: np.ndarray, read_off: np.ndarray, kmer_idx: Optional[np.ndarray], n_iters: int,
                      seed: int = 0, site_id_base: int = 0, n_samples: int = DEFAULT_N_SAMPLES, pooling="prod",
                      replace: bool = False, read_threshold: float = 0.033379376, n_chunks: int = 0):
        """NumPy buffers; returns (read_prob [R], bag_prob [sites, n_iters], site_mean [sites], mod_count [sites])."""
        feats, read_off, kmer_idx, n_sites = self._check_host(feats, read_off, kmer_idx, "validate_host")
        read_prob = np.empty(feats.shape[0], dtype=np.float32)
        bag_prob = np.empty((n_sites, n_iters), dtype=np.float32)
        site_prob = np.empty(n_sites, dtype=np.float32)
        mod_count = np.empty(n_sites, dtype=np.int32)
        self._select()
        rc = self._lib.m6a_mil_validate_host_f32(
            self._handle, _ptr(feats), _ptr(read_off), _ptr(kmer_idx), n_sites, site_id_base, n_samples, n_iters,
            seed & 0xFFFFFFFFFFFFFFFF, _pool_code(pooling), 1 if replace else 0, read_threshold, _ptr(read_prob),
            _ptr(bag_prob), _ptr(site_prob), _ptr(mod_count), n_chunks)
        _cabi.check(rc, "m6a_mil_validate_host_f32")
        return read_prob, bag_prob, site_prob, mod_count

    def sample_bags(self, seed: int, site_id: int, n_reads: int, n_iters: int, n_samples: int = DEFAULT_N_SAMPLES):
        """The device without-replacement bags of one site, int32 CUDA tensor [n_iters, n_samples]."""
        torch = self._torch
        out = torch.empty((n_iters, n_samples), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._lib.m6a_sample_bags(seed & 0xFFFFFFFFFFFFFFFF, site_id, n_reads, n_iters, n_samples,
                                           out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(rc, "m6a_sample_bags")
        return out

    def sample_indices(self, seed: int, site_id: int, n_reads: int, n_iters: int, n_samples: int = DEFAULT_N_SAMPLES):
        torch = self._torch
        out = torch.empty((n_iters, n_samples), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self._lib.m6a_sample_indices(seed & 0xFFFFFFFFFFFFFFFF, site_id, n_reads, n_iters, n_samples,
                                              out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(rc, "m6a_sample_indices")
        return out

    def last_launch(self) -> dict:
        v = [C.c_int32() for _ in range(5)]
        self._lib.m6a_last_launch(*[C.byref(x) for x in v])
        return dict(zip(("grid", "block", "smem_bytes", "tile_reads", "n_launches"), (x.value for x in v)))
