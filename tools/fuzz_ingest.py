#!/usr/bin/env python
"""Mutation fuzzer for the native data.json / data.info readers (m6a_ingest_parts, m6a_info_*).

    python tools/fuzz_ingest.py [iterations] [seed]

Every iteration corrupts the bundled data.json (byte flips, truncations, deletions, duplicated spans, wrong byte ranges
or row counts in the part table) and ingests it.  The parser must return a status (0 or M6A_EPARSE / M6A_EIO / ...) --
never crash, hang or write outside the output buffers (guard bands are checked here; run with an ASan build of the
library, `M6A_LIB=... LD_PRELOAD=libasan.so`, for the reads).  Successful parses must equal Python's json on the same bytes.
"""
import ctypes as C
import gzip
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m6anet_b200 import _cabi                      # noqa: E402
from m6anet_b200.data import NanopolishDS           # noqa: E402

GUARD = 64


def mutate(rng, raw: bytearray) -> bytearray:
    kind = rng.integers(0, 6)
    n = len(raw)
    if kind == 0:       # flip bytes
        for _ in range(int(rng.integers(1, 8))):
            raw[int(rng.integers(0, n))] = int(rng.integers(0, 256))
    elif kind == 1:     # structural characters
        for _ in range(int(rng.integers(1, 6))):
            raw[int(rng.integers(0, n))] = int(rng.choice(list(b'[]{},:"-+.eE0 \n')))
    elif kind == 2:     # delete a span
        a = int(rng.integers(0, n))
        del raw[a:a + int(rng.integers(1, 40))]
    elif kind == 3:     # duplicate a span
        a = int(rng.integers(0, n))
        b = a + int(rng.integers(1, 40))
        raw[a:a] = raw[a:b]
    elif kind == 4:     # truncate
        del raw[int(rng.integers(0, n)):]
    return raw          # kind 5: bytes untouched (the part table is mutated instead)


def run(iterations=300, seed=0, verbose=True):
    rng = np.random.default_rng(seed)
    src = os.path.join(ROOT, "tests", "golden", "bundled")
    norm = os.path.join(ROOT, "m6anet_b200", "assets", "norm_factors", "rna002_hct116.npz")
    L = _cabi.lib()
    stats = {"ok": 0, "rejected": 0}
    with tempfile.TemporaryDirectory() as d:
        with gzip.open(os.path.join(src, "data.json.gz"), "rb") as f:
            good = f.read()
        with open(os.path.join(d, "data.json"), "wb") as g:
            g.write(good)
        with open(os.path.join(src, "data.info"), "rb") as f, open(os.path.join(d, "data.info"), "wb") as g:
            g.write(f.read())
        ds = NanopolishDS(d, 20, norm)
        mean, std, kid = ds._native_tables()
        S = len(ds)
        fuzz_path = os.path.join(d, "fuzz.json")
        paths = (C.c_char_p * 1)(os.fsencode(fuzz_path))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        for it in range(iterations):
            sites = np.sort(rng.choice(S, size=int(rng.integers(1, 6)), replace=False))
            raw = bytearray()
            parts = np.zeros(len(sites), dtype=_cabi.PART_DTYPE)
            row = 0
            for i, s in enumerate(sites):
                a, b = int(ds._part_start[s]), int(ds._part_end[s])
                parts[i] = (0, 0, len(raw), len(raw) + (b - a), row, int(ds._n_reads[s]), i, 1, 0)
                raw += good[a:b]
                row += int(ds._n_reads[s])
            before = bytes(raw)
            raw = mutate(rng, raw)
            if bytes(raw) == before or rng.random() < 0.2:      # corrupt the part table
                j = int(rng.integers(0, len(parts)))
                field = rng.choice(["start", "end", "n_rows"])
                parts[field][j] += int(rng.integers(-30, 31))
                if field != "n_rows":
                    parts[field][j] = max(0, parts[field][j])
            with open(fuzz_path, "wb") as g:
                g.write(raw)
            rows = row
            feats = np.full((rows + 2 * GUARD, 9), 7777.0, dtype=np.float32)
            ids = np.full(rows + 2 * GUARD, -7777, dtype=np.int64)
            kmer = np.full((len(sites) + 2 * GUARD, 3), -7777, dtype=np.int32)
            bad = C.c_int64(-1)
            rc = L.m6a_ingest_parts(paths, 1, vp(parts), len(parts), 1, vp(mean), vp(std), vp(kid), vp(feats[GUARD:]),
                                    vp(ids[GUARD:]), vp(kmer[GUARD:]), int(rng.integers(1, 4)), C.byref(bad))
            assert rc <= 0, f"iteration {it}: unexpected status {rc}"
            for buf, fill in ((feats, 7777.0), (ids, -7777), (kmer, -7777)):      # nothing outside the output window
                assert (buf[:GUARD] == fill).all() and (buf[len(buf) - GUARD:] == fill).all(), f"iteration {it}: guard band"
            if rc == 0:
                stats["ok"] += 1
                # an accepted input must mean what Python's json says it means
                for i in range(len(parts)):
                    a, b = int(parts["start"][i]), int(parts["end"][i])
                    obj = json.loads(bytes(raw[a:b]))
                    (tx, inner), = obj.items()
                    (pos, inner2), = inner.items()
                    (seq, rows_), = inner2.items()
                    arr = np.array(rows_, dtype=np.float64)
                    assert arr.shape[0] == parts["n_rows"][i]
                    five = [seq[t:t + 5] for t in range(3)]
                    m, sd = ds.get_norm_factor(five)
                    want = ((arr[:, :9] - m) / sd).astype(np.float32)
                    r0 = int(parts["row_off"][i])
                    got = feats[GUARD + r0: GUARD + r0 + len(want)]
                    assert np.array_equal(got, want, equal_nan=True), f"iteration {it}: accepted input parsed differently"
            else:
                stats["rejected"] += 1
                assert 0 <= bad.value < len(parts) or rc != -6, f"iteration {it}: bad_part not set"
    if verbose:
        print(f"fuzz_ingest: {iterations} inputs, {stats['ok']} accepted (== json), {stats['rejected']} rejected, no crash")
    return stats


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
