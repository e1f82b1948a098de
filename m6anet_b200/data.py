"""Ingest of `m6anet dataprep` outputs (data.info + data.json) into flat site-contiguous buffers.

Mirrors reference m6anet/utils/data_utils.py for the inference path:
    NanopolishDS            :20-290   (data.info index :118-129, byte-range JSON read :169-190,
                                       9-column selection :105-116,166, float64 normalisation :210-218,233-248,
                                       k-mer ids :223-224)
    NanopolishReplicateDS   :293-427  (outer join of N data.info files :341-375, per-site concatenation and
                                       "{read_id}_{replicate}" ids :395-427)
    inference_collate       :498-506
Input format (reference utils/dataprep_utils.py:473-485): data.info is a csv
`transcript_id,transcript_position,start,end,n_reads`; data.json holds one line per site
`{"<tx>":{"<pos>":{"<7-mer>":[[f0..f8, read_id], ...]}}}` addressed by the `[start, end)` byte range.

Besides the reference's per-site `__getitem__` (Python json, reference-shaped), the datasets expose
`load_sites(lo, hi)`, which returns the flat buffers the C ABI consumes (feats [R,9] f32, read_off [S+1] i64,
kmer_idx [S,3] i32).  `load_sites` runs the multi-threaded native parser `m6a_ingest_parts`
(m6anet_b200/csrc/m6a_io.cpp); tests/test_host.py checks it bit for bit against the per-site Python path and
against the reference's own NanopolishDS output.
The evaluation modes of the reference ('Train' / 'Val' / 'Test': data.info.labelled filtered by set_type, labels from
modification_status, data_utils.py:124-126,102-103) are provided for `m6anet_b200.validation.validate`; training itself
and on-the-fly norm-factor computation are outside the path and not provided.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .constants import kmer_table


@dataclass
class SiteBatch:
    """Flat, site-contiguous view of sites [lo, hi) -- the layout of m6a_mil_infer_f32."""
    feats: np.ndarray        # [R, 9] float32, normalised
    read_off: np.ndarray     # [S+1] int64
    kmer_idx: np.ndarray     # [S, 3] int32
    read_ids: np.ndarray     # [R] int64
    tx_ids: np.ndarray       # [S] str
    tx_pos: np.ndarray       # [S] int64
    kmers: np.ndarray        # [S] centre 5-mer (str)
    read_rep: Optional[np.ndarray] = None   # [R] int32 replicate number (multi-directory input) -> "{id}_{rep}"
    tx_bytes: Optional[np.ndarray] = None   # [S] fixed-width bytes view of tx_ids (what the CSV writers consume)

    @property
    def n_sites(self) -> int:
        return len(self.read_off) - 1


def load_norm_factors(norm_path: str) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """5-mer -> (mean[3], std[3]) float64, order (dwell, sd, mean).  Accepts this package's .npz or a
    reference .joblib (reference utils/data_utils.py:79-80)."""
    if str(norm_path).endswith(".npz"):
        with np.load(norm_path) as z:
            kmers, mean, std = z["kmers"], z["mean"], z["std"]      # read each array once (NpzFile re-reads per access)
        return {str(k): (mean[i], std[i]) for i, k in enumerate(kmers)}
    import joblib
    nd = joblib.load(norm_path)
    return {str(k): (np.asarray(v[0], dtype=np.float64), np.asarray(v[1], dtype=np.float64)) for k, v in nd.items()}


def read_info(root_dir: str, name: str = "data.info"):
    """data.info -> (transcript_id bytes array 'S<w>', transcript_position, start, end, n_reads) int64 arrays.
    Native reader (m6a_info_count / m6a_info_read); the reference uses pd.read_csv (utils/data_utils.py:118-129)."""
    import ctypes as C
    from . import _cabi
    L = _cabi.lib()
    path = os.fsencode(os.path.join(root_dir, name))
    n, nb = C.c_int64(0), C.c_int64(0)
    _cabi.check(L.m6a_info_count(path, C.byref(n), C.byref(nb)), f"m6a_info_count({os.fsdecode(path)})")
    n, nb = int(n.value), int(nb.value)
    buf = np.zeros(max(nb, 1), dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.int64)
    cols = [np.zeros(max(n, 1), dtype=np.int64) for _ in range(4)]
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _cabi.check(L.m6a_info_read(path, n, nb, vp(buf), vp(off), *[vp(c) for c in cols]), f"m6a_info_read({os.fsdecode(path)})")
    # variable-length ids -> fixed-width bytes array (vectorised scatter into a zero-padded [n, width] matrix)
    lengths = np.diff(off)
    width = int(lengths.max()) if n else 1
    mat = np.zeros((n, max(width, 1)), dtype=np.uint8)
    if nb:
        rows = np.repeat(np.arange(n), lengths)
        mat[rows, np.arange(nb) - np.repeat(off[:-1], lengths)] = buf[:nb]
    tx = mat.view(f"S{max(width, 1)}").reshape(n)
    return tx, cols[0][:n], cols[1][:n], cols[2][:n], cols[3][:n]


def read_labels(root_dir: str, name: str = "data.info.labelled"):
    """(modification_status int64 [n], set_type str [n]) of data.info.labelled, rows aligned with read_info(root_dir, name)
    (blank lines skipped like the native reader does).  Reference: pd.read_csv at utils/data_utils.py:125."""
    import csv
    status, set_type = [], []
    with open(os.path.join(root_dir, name), newline="") as f:
        rows = csv.reader(f)
        header = [h.strip() for h in next(rows)]
        try:
            i_status, i_set = header.index("modification_status"), header.index("set_type")
        except ValueError:
            raise ValueError(f"{os.path.join(root_dir, name)} needs the columns modification_status and set_type")
        for row in rows:
            if not row or (len(row) == 1 and not row[0].strip()):
                continue
            status.append(int(float(row[i_status])))
            set_type.append(row[i_set].strip())
    return np.array(status, dtype=np.int64), np.array(set_type, dtype=str)


class NanopolishDS:
    allowed_mode = ('Train', 'Test', 'Val', 'Inference')     # reference data_utils.py:42

    def __init__(self, root_dir, min_reads: int = 20, norm_path: Optional[str] = None, num_neighboring_features: int = 1,
                 mode: str = 'Inference', n_processes: int = 1):
        if mode not in self.allowed_mode:
            raise ValueError(f"Invalid mode passed to dataset, must be one of {self.allowed_mode}")
        if root_dir is None:
            raise ValueError("Either root directory or data info must be given")
        if norm_path is None:
            raise NotImplementedError("computing norm factors from the data is not part of the inference path; pass norm_path")
        if num_neighboring_features > 5:
            raise ValueError(f"Invalid neighboring features number {num_neighboring_features}")
        self.root_dir = root_dir
        self.min_reads = min_reads
        self.mode = mode
        self.num_neighboring_features = num_neighboring_features
        self.norm_dict = load_norm_factors(norm_path)
        self.all_kmers = kmer_table(num_neighboring_features)
        self.kmer_to_int = {str(k): i for i, k in enumerate(self.all_kmers)}
        self.int_to_kmer = {i: str(k) for i, k in enumerate(self.all_kmers)}
        self._files: Dict[str, object] = {}
        self.initialize_data_info()
        self.set_feature_indices()

    # ---- index ------------------------------------------------------------------------------------
    def initialize_data_info(self):
        if self.mode == 'Inference':
            tx, pos, start, end, n_reads = read_info(self.root_dir)
            keep = n_reads >= self.min_reads                                               # data_utils.py:129
            self.labels = None
        else:    # labelled index filtered by set_type (data_utils.py:124-126); labels = modification_status (:102-103)
            tx, pos, start, end, n_reads = read_info(self.root_dir, "data.info.labelled")
            status, set_type = read_labels(self.root_dir)
            if len(status) != len(pos):
                raise ValueError("data.info.labelled: label columns and index columns disagree on the number of rows")
            keep = (set_type == self.mode) & (n_reads >= self.min_reads)
            self.labels = status[keep]
        self.data_fpath = os.path.join(self.root_dir, "data.json")
        self._tx_bytes = tx[keep]
        self._pos = pos[keep]
        self._n_reads = n_reads[keep]
        # part tables (one part per site here): CSR pointer over sites + per-part file / byte range / replicate / rows
        S = len(self._pos)
        self._paths = [self.data_fpath]
        self._part_ptr = np.arange(S + 1, dtype=np.int64)
        self._part_file = np.zeros(S, dtype=np.int32)
        self._part_rep = np.zeros(S, dtype=np.int32)
        self._part_start = start[keep]
        self._part_end = end[keep]
        self._part_rows = self._n_reads.copy()
        self._multi = False

    @property
    def _tx(self) -> np.ndarray:
        """transcript ids as a str array (decoded on demand; the index keeps them as fixed-width bytes)"""
        return self._tx_bytes.astype(str)

    @property
    def data_info(self):
        """pandas view of the site index (reference attribute name); built on demand"""
        import pandas as pd
        df = pd.DataFrame({"transcript_id": self._tx, "transcript_position": self._pos, "n_reads": self._n_reads})
        if self.labels is not None:
            df["modification_status"] = self.labels
            df["set_type"] = self.mode
        return df

    def __len__(self) -> int:
        return len(self._pos)

    @property
    def n_reads(self) -> np.ndarray:
        return self._n_reads

    def set_feature_indices(self):
        """Column indices of the 3 x (2n+1) signal features inside a data.json row (data_utils.py:105-116)."""
        self.total_neighboring_features = self.get_total_neighboring_features()
        T, n = self.total_neighboring_features, self.num_neighboring_features
        left = [(T - n + j) * 3 + i for j in range(n) for i in range(3)]
        centre = [T * 3 + i for i in range(3)]
        right = [(T + j) * 3 + i for j in range(1, n + 1) for i in range(3)]
        self.indices = np.array(left + centre + right, dtype=int)

    def get_total_neighboring_features(self) -> int:
        if len(self) == 0:
            return self.num_neighboring_features
        path, s, e, _ = self._site_parts(0)[0]
        kmer, _ = self._load_data(path, self._tx_bytes[0].decode(), int(self._pos[0]), s, e)
        return (len(kmer) - 5) // 2

    def _site_parts(self, idx: int):
        a, b = int(self._part_ptr[idx]), int(self._part_ptr[idx + 1])
        return [(self._paths[self._part_file[i]], int(self._part_start[i]), int(self._part_end[i]), int(self._part_rep[i]))
                for i in range(a, b)]

    # ---- raw access ------------------------------------------------------------------------------------
    def _handle(self, path: str) -> int:
        fd = self._files.get(path)
        if fd is None:
            fd = os.open(path, os.O_RDONLY)
            self._files[path] = fd
        return fd

    def _load_data(self, data_fpath: str, tx_id: str, tx_pos: int, start_pos: int, end_pos: int):
        """One site's JSON line -> (sequence, float64 array [n, 3*(2T+1)+1]) (data_utils.py:169-190)."""
        # positioned read: forked ingest workers share the descriptor, so no seek (shared file offset)
        raw = os.pread(self._handle(data_fpath), end_pos - start_pos, start_pos)
        pos_info = json.loads(raw)[tx_id][str(tx_pos)]
        assert len(pos_info.keys()) == 1
        kmer, features = next(iter(pos_info.items()))
        return kmer, np.array(features, dtype=np.float64)

    def _read_id_array(self, raw: np.ndarray, rep: int):
        return raw.astype(np.int64)

    def load_data(self, idx: int):
        """(tx_id, tx_pos, read_ids, raw selected features float64 [n, 9], sequence)."""
        feats, ids, seq = [], [], None
        tx_id = self._tx_bytes[idx].decode()
        for path, s, e, rep in self._site_parts(idx):
            kmer, arr = self._load_data(path, tx_id, int(self._pos[idx]), s, e)
            if seq is None:
                seq = kmer
            else:
                assert seq == kmer
            feats.append(arr[:, self.indices])
            ids.append(self._read_id_array(arr[:, -1], rep))
        return tx_id, int(self._pos[idx]), np.concatenate(ids), np.concatenate(feats), seq

    def _retrieve_full_sequence(self, kmer: str, n_neighboring_features: int = 1) -> str:
        # centre 5-mer plus n flanks out of a (5 + 2T)-mer.  (The reference's slice at data_utils.py:262 returns a
        # 4-mer for T > n and then fails on the k-mer lookup, so only T == n is exercised there.)
        T = self.total_neighboring_features
        if n_neighboring_features < T:
            return kmer[T - n_neighboring_features: T + 5 + n_neighboring_features]
        return kmer

    def get_norm_factor(self, list_of_kmers: Sequence[str]):
        mean = np.concatenate([self.norm_dict[k][0] for k in list_of_kmers])
        std = np.concatenate([self.norm_dict[k][1] for k in list_of_kmers])
        return mean, std

    def _site(self, idx: int):
        tx_id, tx_pos, read_ids, raw, seq = self.load_data(idx)
        seq = self._retrieve_full_sequence(seq, self.num_neighboring_features)
        five = [seq[i:i + 5] for i in range(2 * self.num_neighboring_features + 1)]
        mean, std = self.get_norm_factor(five)
        feats = ((raw - mean) / std).astype(np.float32)            # float64 arithmetic, one rounding (data_utils.py:216-218)
        kid = np.array([self.kmer_to_int[k] for k in five], dtype=np.int64)
        return tx_id, tx_pos, read_ids, feats, kid

    def __getitem__(self, idx: int):
        """Reference-shaped item (data_utils.py:192-231).  Inference: features [n,9] f32, kmer [n,3] i64 (row repeated),
        tx_id [n], tx_pos [n], read_ids [n].  Labelled modes: a bag of min_reads reads drawn without replacement from
        the process-global NumPy stream (:213-214), its k-mer rows and the site's label -- the reference-shaped view
        only; `validation.validate` draws its bags on the device instead."""
        tx_id, tx_pos, read_ids, feats, kid = self._site(idx)
        if self.mode != 'Inference':
            feats = feats[np.random.choice(len(feats), self.min_reads, replace=False), :]
            return feats, np.repeat(kid[None, :], len(feats), axis=0), int(self.labels[idx])
        n = len(feats)
        return feats, np.repeat(kid[None, :], n, axis=0), np.repeat(tx_id, n), np.repeat(tx_pos, n), read_ids

    # ---- flat access (native parser) -----------------------------------------------------------------------
    def _native_tables(self):
        t = getattr(self, "_tables", None)
        if t is None:
            code = {"A": 0, "C": 1, "G": 2, "T": 3}
            enc = lambda k: sum(code[ch] * 4 ** (4 - i) for i, ch in enumerate(k))
            mean = np.full((1024, 3), np.nan)
            std = np.full((1024, 3), np.nan)
            for k, (m, sd) in self.norm_dict.items():
                if len(k) == 5 and set(k) <= set("ACGT"):
                    mean[enc(k)], std[enc(k)] = m, sd
            kid = np.full(1024, -1, dtype=np.int32)
            for k, i in self.kmer_to_int.items():
                kid[enc(k)] = i
            t = self._tables = (np.ascontiguousarray(mean), np.ascontiguousarray(std), kid)
        return t

    def load_sites(self, lo: int, hi: int, n_threads: int = 0, alloc=None) -> SiteBatch:
        import ctypes as C
        from . import _cabi
        hi = min(hi, len(self))
        S = max(hi - lo, 0)
        n_pos = 2 * self.num_neighboring_features + 1
        a, b = int(self._part_ptr[lo]), int(self._part_ptr[lo + S])
        rows = self._part_rows[a:b]
        parts = np.zeros(b - a, dtype=_cabi.PART_DTYPE)
        parts["file"] = self._part_file[a:b]
        parts["rep"] = self._part_rep[a:b]
        parts["start"] = self._part_start[a:b]
        parts["end"] = self._part_end[a:b]
        parts["n_rows"] = rows
        row_off = np.zeros(b - a + 1, dtype=np.int64)
        np.cumsum(rows, out=row_off[1:])
        parts["row_off"] = row_off[:-1]
        counts = np.diff(self._part_ptr[lo:lo + S + 1])
        parts["site"] = np.repeat(np.arange(S, dtype=np.int64), counts)
        first = np.zeros(b - a, dtype=np.int32)
        first[(self._part_ptr[lo:lo + S] - a)[counts > 0]] = 1
        parts["first_of_site"] = first
        R = int(row_off[-1])
        # `alloc(shape, dtype)`: where the feature rows live (run_inference passes a page-locked pool so that the H2D copies
        # of the host-buffer call overlap with the kernel); default: ordinary memory
        feats = (alloc or np.empty)((R, 3 * n_pos), np.float32)
        read_ids = np.empty(R, dtype=np.int64)
        kmer_idx = np.zeros((S, n_pos), dtype=np.int32)
        mean, std, kid = self._native_tables()
        paths = (C.c_char_p * len(self._paths))(*[os.fsencode(p) for p in self._paths])
        bad = C.c_int64(-1)
        vp = lambda arr: arr.ctypes.data_as(C.c_void_p)
        # keys of the sites: the line of every part must name its site's transcript id and position (a stale data.info
        # raises KeyError in the reference, utils/data_utils.py:185; here M6A_EPARSE)
        txb = self._tx_bytes[lo:lo + S]
        lens = np.char.str_len(txb).astype(np.int64) if S else np.zeros(0, np.int64)
        tx_off = np.zeros(S + 1, dtype=np.int64)
        np.cumsum(lens, out=tx_off[1:])
        tx_buf = b"".join(txb.tolist())
        tx_pos = np.ascontiguousarray(self._pos[lo:lo + S], dtype=np.int64)
        rc = _cabi.lib().m6a_ingest_parts_keyed(paths, len(self._paths), vp(parts), len(parts), self.num_neighboring_features,
                                                vp(mean), vp(std), vp(kid), tx_buf, vp(tx_off), vp(tx_pos), vp(feats),
                                                vp(read_ids), vp(kmer_idx), n_threads, C.byref(bad))
        if rc != 0:
            where = ""
            if bad.value >= 0:
                i = a + bad.value
                site = lo + int(parts["site"][bad.value])
                where = (f" (site {self._tx_bytes[site].decode()}:{self._pos[site]}, file {self._paths[self._part_file[i]]}, bytes "
                         f"[{self._part_start[i]}, {self._part_end[i]}))")
            raise _cabi.M6AError(rc, "m6a_ingest_parts" + where)
        read_off = np.zeros(S + 1, dtype=np.int64)
        read_off[1:] = row_off[1:][np.cumsum(counts) - 1] if S and (b - a) else 0
        centre = np.asarray(self.all_kmers)[kmer_idx[:, n_pos // 2]] if S else np.array([], dtype=str)    # == int_to_kmer
        rep = np.repeat(self._part_rep[a:b], rows).astype(np.int32) if self._multi else None
        return SiteBatch(feats=feats, read_off=read_off, kmer_idx=kmer_idx, read_ids=read_ids,
                         tx_ids=self._tx_bytes[lo:hi].astype(str),
                         tx_pos=self._pos[lo:hi], kmers=centre, read_rep=rep, tx_bytes=self._tx_bytes[lo:hi])

    def close(self):
        for fd in self._files.values():
            os.close(fd)
        self._files = {}

    def __getstate__(self):   # file handles do not cross process boundaries (worker pools)
        d = dict(self.__dict__)
        d["_files"] = {}
        return d


class NanopolishReplicateDS(NanopolishDS):
    """Several input directories pooled per (transcript_id, transcript_position) -- reference
    data_utils.py:293-427.  Site order: first appearance over the directories in the order given (the
    reference's order comes from a pandas outer join and is version dependent; its tests sort before comparing)."""

    def __init__(self, root_dir: List[str], min_reads: int = 20, norm_path: Optional[str] = None,
                 num_neighboring_features: int = 1, mode: str = 'Inference', n_processes: int = 1):
        super().__init__(list(root_dir), min_reads, norm_path, num_neighboring_features, mode, n_processes)

    def initialize_data_info(self):
        # outer join of the directories on (transcript_id, transcript_position), sites in first-appearance order
        labelled = self.mode != 'Inference'
        cols = [read_info(d, "data.info.labelled" if labelled else "data.info") for d in self.root_dir]
        width = max(c[0].dtype.itemsize for c in cols)
        tx = np.concatenate([c[0].astype(f"S{width}") for c in cols])
        pos, start, end, n_reads = (np.concatenate([c[k] for c in cols]) for k in (1, 2, 3, 4))
        rep = np.concatenate([np.full(len(c[1]), r, dtype=np.int32) for r, c in enumerate(cols)])
        if labelled:   # the join key also carries modification_status and set_type (data_utils.py:349-350)
            lab = [read_labels(d) for d in self.root_dir]
            status = np.concatenate([l[0] for l in lab])
            set_type = np.concatenate([l[1] for l in lab])
            if len(status) != len(pos):
                raise ValueError("data.info.labelled: label columns and index columns disagree on the number of rows")
            key = np.rec.fromarrays([tx, pos, status, set_type], names="tx,pos,status,set_type")
        else:
            key = np.rec.fromarrays([tx, pos], names="tx,pos")
        uniq, first, inverse = np.unique(key, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")                 # unique keys by first appearance
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[order] = np.arange(len(uniq))
        site = rank[inverse.reshape(-1)]
        total = np.bincount(site, weights=n_reads, minlength=len(uniq)).astype(np.int64)
        keep = total >= self.min_reads                                                   # data_utils.py:373
        if labelled:
            keep &= (uniq["set_type"][order] == self.mode)                                  # data_utils.py:370-371
            self.labels = uniq["status"][order][keep].astype(np.int64)
        else:
            self.labels = None
        new_id = np.cumsum(keep) - 1
        self._tx_bytes = uniq["tx"][order][keep]
        self._pos = uniq["pos"][order][keep].astype(np.int64)
        self._n_reads = total[keep]
        rows = np.nonzero(keep[site])[0]
        rows = rows[np.lexsort((rep[rows], site[rows]))]         # by site, then directory
        site_new = new_id[site[rows]]
        self._paths = [os.path.join(d, "data.json") for d in self.root_dir]
        self._part_file = rep[rows].astype(np.int32)
        self._part_rep = rep[rows].astype(np.int32)
        self._part_start = start[rows]
        self._part_end = end[rows]
        self._part_rows = n_reads[rows]
        self._part_ptr = np.zeros(int(keep.sum()) + 1, dtype=np.int64)
        np.cumsum(np.bincount(site_new, minlength=int(keep.sum())), out=self._part_ptr[1:])
        self._multi = True
        self.fpath_mapping = {d: rep_ for rep_, d in enumerate(self.root_dir)}
        self.data_fpath = None

    def _read_id_array(self, raw: np.ndarray, rep: int):
        """reference-shaped ids of __getitem__: "{read_id}_{replicate}" (data_utils.py:421-423)"""
        return np.array([f"{int(r)}_{rep}" for r in raw], dtype=object)


def inference_collate(batch):
    """Reference-shaped collate (data_utils.py:498-506) over __getitem__ items, NumPy outputs."""
    n_reads = np.array([len(item[0]) for item in batch], dtype=np.int64)
    return (np.concatenate([item[0] for item in batch]), np.concatenate([item[1] for item in batch]), n_reads,
            np.concatenate([item[2] for item in batch]), np.concatenate([item[3] for item in batch]),
            np.concatenate([item[4] for item in batch]))
