#!/usr/bin/env python
"""Write synthetic `m6anet dataprep`-format directories (data.json + data.info) for ingest / CLI / replicate benchmarks.

    python tools/make_synthetic_dataset.py OUT_DIR --sites 20000 --reads 50 [--seed 0]
    python tools/make_synthetic_dataset.py OUT_ROOT --replicates 4 --sites 250000 --reads 20 [--disjoint 10]

Row format and index follow reference utils/dataprep_utils.py:473-485: one JSON line per site
{"<tx>":{"<pos>":{"<7-mer>":[[dwell,sd,mean x3, read_id], ...]}}} and a csv row
transcript_id,transcript_position,start,end,n_reads with the line's byte range.

--replicates R writes OUT_ROOT/rep0 .. rep{R-1} (BASELINE config 5: the input directories of one
NanopolishReplicateDS run): the same site keys and 7-mers in every directory, different reads.  --disjoint P makes P % of
the keys absent from one directory each (site s is missing from directory (s // P') % R when s % P' == 0, P' = 100 / P),
which exercises the outer join of reference utils/data_utils.py:341-375."""
import argparse
import itertools
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

CENTRE = ["".join(c) for c in itertools.product("AGT", "GA", "A", "C", "ACT")]
SEVEN = [x + c + y for x in "GACT" for c in CENTRE for y in "GACT"]


def site_key(s: int):
    return f"ENST{s // 50:011d}.1", 100 + 7 * (s % 50)


def write_dataset(out_dir: str, sites: int, reads: int, seed: int = 0, kmer_seed: int = 0, skip=None) -> float:
    """One directory.  `skip(s) -> bool` drops site s.  7-mers depend on (kmer_seed, s) only, reads on `seed`."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    kmers = np.random.default_rng([kmer_seed, 7]).integers(len(SEVEN), size=sites)
    row_fmt = "[" + ",".join(["%.5f"] * 9) + ",%d.0]"
    off = 0
    chunk = 512
    with open(os.path.join(out_dir, "data.json"), "w") as fj, open(os.path.join(out_dir, "data.info"), "w") as fi:
        fi.write("transcript_id,transcript_position,start,end,n_reads\n")
        for s0 in range(0, sites, chunk):
            n = min(chunk, sites - s0)
            vals = rng.normal([0.008, 3.5, 105.0] * 3, [0.004, 1.5, 8.0] * 3, size=(n, reads, 9))
            vals[..., 0::3] = np.abs(vals[..., 0::3]) + 1e-3
            vals[..., 1::3] = np.abs(vals[..., 1::3]) + 0.5
            ids = rng.integers(1, 10**6, size=(n, reads))
            for i in range(n):
                s = s0 + i
                if skip is not None and skip(s):
                    continue
                rows = ",".join(row_fmt % (*r, j) for r, j in zip(vals[i].tolist(), ids[i].tolist()))
                tx, pos = site_key(s)
                line = '{"%s":{"%d":{"%s":[%s]}}}\n' % (tx, pos, SEVEN[kmers[s]], rows)
                fj.write(line)
                fi.write(f"{tx},{pos},{off},{off + len(line)},{reads}\n")
                off += len(line)
    return off / 1e6


def _one_replicate(args):
    root, r, n_rep, sites, reads, seed, disjoint = args
    skip = None
    if disjoint > 0:
        every = max(1, 100 // disjoint)
        skip = lambda s: s % every == 0 and (s // every) % n_rep == r          # noqa: E731
    return write_dataset(os.path.join(root, f"rep{r}"), sites, reads, seed=seed + 1000 * (r + 1), kmer_seed=seed, skip=skip)


def write_replicates(root: str, n_rep: int, sites: int, reads: int, seed: int = 0, disjoint: int = 0, workers: int = 0):
    os.makedirs(root, exist_ok=True)
    jobs = [(root, r, n_rep, sites, reads, seed, disjoint) for r in range(n_rep)]
    with ProcessPoolExecutor(max_workers=workers or n_rep) as ex:
        mbs = list(ex.map(_one_replicate, jobs))
    return [os.path.join(root, f"rep{r}") for r in range(n_rep)], mbs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("--sites", type=int, default=20000)
    ap.add_argument("--reads", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--replicates", type=int, default=0)
    ap.add_argument("--disjoint", type=int, default=0, help="percent of site keys missing from one directory each")
    a = ap.parse_args()
    if a.replicates > 0:
        dirs, mbs = write_replicates(a.out_dir, a.replicates, a.sites, a.reads, a.seed, a.disjoint)
        print(f"wrote {a.replicates} directories x {a.sites} sites x {a.reads} reads ({sum(mbs):.1f} MB of data.json) -> {a.out_dir}")
    else:
        mb = write_dataset(a.out_dir, a.sites, a.reads, a.seed, kmer_seed=a.seed)
        print(f"wrote {a.sites} sites x {a.reads} reads, data.json {mb:.1f} MB -> {a.out_dir}")


if __name__ == "__main__":
    main()
