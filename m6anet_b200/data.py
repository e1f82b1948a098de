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

Besides the reference's per-site `__getitem__`, the datasets expose `load_sites(lo, hi)`, which returns
the flat buffers the C ABI consumes (feats [R,9] f32, read_off [S+1] i64, kmer_idx [S,3] i32).
Training modes and on-the-fly norm-factor computation are outside the inference path and not provided.
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
    read_ids: np.ndarray     # [R] int64 (single dir) or object/str "{id}_{rep}" (replicates)
    tx_ids: np.ndarray       # [S] str
    tx_pos: np.ndarray       # [S] int64
    kmers: np.ndarray        # [S] centre 5-mer (str)

    @property
    def n_sites(self) -> int:
        return len(self.read_off) - 1


def load_norm_factors(norm_path: str) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """5-mer -> (mean[3], std[3]) float64, order (dwell, sd, mean).  Accepts this package's .npz or a
    reference .joblib (reference utils/data_utils.py:79-80)."""
    if str(norm_path).endswith(".npz"):
        z = np.load(norm_path)
        return {str(k): (z["mean"][i], z["std"][i]) for i, k in enumerate(z["kmers"])}
    import joblib
    nd = joblib.load(norm_path)
    return {str(k): (np.asarray(v[0], dtype=np.float64), np.asarray(v[1], dtype=np.float64)) for k, v in nd.items()}


def _read_info(root_dir: str):
    import pandas as pd
    return pd.read_csv(os.path.join(root_dir, "data.info"))


class NanopolishDS:
    allowed_mode = ('Inference',)

    def __init__(self, root_dir, min_reads: int = 20, norm_path: Optional[str] = None, num_neighboring_features: int = 1,
                 mode: str = 'Inference', n_processes: int = 1):
        if mode not in self.allowed_mode:
            raise ValueError(f"Invalid mode passed to dataset, must be one of {self.allowed_mode} "
                             f"(training modes are outside the m6anet_b200 inference path)")
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
        info = _read_info(self.root_dir)
        self.data_fpath = os.path.join(self.root_dir, "data.json")
        self.data_info = info[info["n_reads"] >= self.min_reads].reset_index(drop=True)     # data_utils.py:129
        self._tx = self.data_info["transcript_id"].to_numpy().astype(str)
        self._pos = self.data_info["transcript_position"].to_numpy(dtype=np.int64)
        self._n_reads = self.data_info["n_reads"].to_numpy(dtype=np.int64)
        # per site: list of (file path, start, end, replicate number)
        start = self.data_info["start"].to_numpy(dtype=np.int64)
        end = self.data_info["end"].to_numpy(dtype=np.int64)
        self._parts = [[(self.data_fpath, int(s), int(e), 0)] for s, e in zip(start, end)]

    def __len__(self) -> int:
        return len(self._tx)

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
        path, s, e, _ = self._parts[0][0]
        kmer, _ = self._load_data(path, self._tx[0], int(self._pos[0]), s, e)
        return (len(kmer) - 5) // 2

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
        for path, s, e, rep in self._parts[idx]:
            kmer, arr = self._load_data(path, self._tx[idx], int(self._pos[idx]), s, e)
            if seq is None:
                seq = kmer
            else:
                assert seq == kmer
            feats.append(arr[:, self.indices])
            ids.append(self._read_id_array(arr[:, -1], rep))
        return self._tx[idx], int(self._pos[idx]), np.concatenate(ids), np.concatenate(feats), seq

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
        """Reference-shaped item (data_utils.py:192-231): features [n,9] f32, kmer [n,3] i64 (row repeated),
        tx_id [n], tx_pos [n], read_ids [n]."""
        tx_id, tx_pos, read_ids, feats, kid = self._site(idx)
        n = len(feats)
        return feats, np.repeat(kid[None, :], n, axis=0), np.repeat(tx_id, n), np.repeat(tx_pos, n), read_ids

    # ---- flat access ------------------------------------------------------------------------------------
    def load_sites(self, lo: int, hi: int) -> SiteBatch:
        hi = min(hi, len(self))
        feats, ids, kids, n_reads = [], [], [], []
        for idx in range(lo, hi):
            _, _, read_ids, f, kid = self._site(idx)
            feats.append(f)
            ids.append(read_ids)
            kids.append(kid)
            n_reads.append(len(f))
        S = hi - lo
        read_off = np.zeros(S + 1, dtype=np.int64)
        np.cumsum(n_reads, out=read_off[1:])
        kid = np.array(kids, dtype=np.int32).reshape(S, 3) if S else np.zeros((0, 3), np.int32)
        centre = np.array([self.int_to_kmer[int(k)] for k in kid[:, kid.shape[1] // 2]]) if S else np.array([], dtype=str)
        return SiteBatch(
            feats=np.concatenate(feats) if feats else np.zeros((0, 9), np.float32), read_off=read_off, kmer_idx=kid,
            read_ids=np.concatenate(ids) if ids else np.zeros(0, np.int64), tx_ids=self._tx[lo:hi], tx_pos=self._pos[lo:hi],
            kmers=centre)

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
        import pandas as pd
        keys = ["transcript_id", "transcript_position"]
        frames = []
        for rep, d in enumerate(self.root_dir):
            df = _read_info(d)
            df["rep"] = rep
            frames.append(df)
        allrows = pd.concat(frames, ignore_index=True)
        site_key = allrows[keys].drop_duplicates(keep="first").reset_index(drop=True)      # outer join, first-appearance order
        site_key["site"] = np.arange(len(site_key))
        allrows = allrows.merge(site_key, on=keys, how="left", sort=False)
        total = allrows.groupby("site", sort=True)["n_reads"].sum().to_numpy()
        keep = total >= self.min_reads                                                   # data_utils.py:373
        new_id = np.cumsum(keep) - 1
        self._tx = site_key["transcript_id"].to_numpy().astype(str)[keep]
        self._pos = site_key["transcript_position"].to_numpy(dtype=np.int64)[keep]
        self._n_reads = total[keep].astype(np.int64)
        parts: List[list] = [[] for _ in range(int(keep.sum()))]
        rows = allrows[keep[allrows["site"].to_numpy()]].sort_values(["site", "rep"], kind="stable")
        for site, rep, s, e in zip(rows["site"].to_numpy(), rows["rep"].to_numpy(), rows["start"].to_numpy(), rows["end"].to_numpy()):
            parts[new_id[site]].append((os.path.join(self.root_dir[rep], "data.json"), int(s), int(e), int(rep)))
        self._parts = parts
        self.data_info = pd.DataFrame({"transcript_id": self._tx, "transcript_position": self._pos, "n_reads": self._n_reads})
        self.fpath_mapping = {d: rep for rep, d in enumerate(self.root_dir)}
        self.data_fpath = None

    def _read_id_array(self, raw: np.ndarray, rep: int):
        return np.array([f"{int(r)}_{rep}" for r in raw], dtype=object)                  # data_utils.py:421-423


def inference_collate(batch):
    """Reference-shaped collate (data_utils.py:498-506) over __getitem__ items, NumPy outputs."""
    n_reads = np.array([len(item[0]) for item in batch], dtype=np.int64)
    return (np.concatenate([item[0] for item in batch]), np.concatenate([item[1] for item in batch]), n_reads,
            np.concatenate([item[2] for item in batch]), np.concatenate([item[3] for item in batch]),
            np.concatenate([item[4] for item in batch]))
