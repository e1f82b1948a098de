#!/usr/bin/env python
"""Write a synthetic `m6anet dataprep`-format directory (data.json + data.info) for ingest/CLI benchmarks.

    python tools/make_synthetic_dataset.py OUT_DIR --sites 20000 --reads 50 [--seed 0]

Row format and index follow reference utils/dataprep_utils.py:473-485: one JSON line per site
{"<tx>":{"<pos>":{"<7-mer>":[[dwell,sd,mean x3, read_id], ...]}}} and a csv row
transcript_id,transcript_position,start,end,n_reads with the line's byte range."""
import argparse
import itertools
import os

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("--sites", type=int, default=20000)
    ap.add_argument("--reads", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    os.makedirs(a.out_dir, exist_ok=True)
    rng = np.random.default_rng(a.seed)
    centre = ["".join(c) for c in itertools.product("AGT", "GA", "A", "C", "ACT")]
    seven = [x + c + y for x in "GACT" for c in centre for y in "GACT"]
    off = 0
    with open(os.path.join(a.out_dir, "data.json"), "w") as fj, open(os.path.join(a.out_dir, "data.info"), "w") as fi:
        fi.write("transcript_id,transcript_position,start,end,n_reads\n")
        for s in range(a.sites):
            k = seven[int(rng.integers(len(seven)))]
            vals = rng.normal([0.008, 3.5, 105.0] * 3, [0.004, 1.5, 8.0] * 3, size=(a.reads, 9))
            vals[:, 0::3] = np.abs(vals[:, 0::3]) + 1e-3
            vals[:, 1::3] = np.abs(vals[:, 1::3]) + 0.5
            ids = rng.integers(1, 10**6, a.reads)
            rows = ",".join("[" + ",".join(repr(round(float(v), 5)) for v in r) + f",{float(i)!r}]" for r, i in zip(vals, ids))
            tx, pos = f"ENST{s // 50:011d}.1", 100 + 7 * (s % 50)
            line = '{"%s":{"%d":{"%s":[%s]}}}\n' % (tx, pos, k, rows)
            fj.write(line)
            fi.write(f"{tx},{pos},{off},{off + len(line)},{a.reads}\n")
            off += len(line)
    print(f"wrote {a.sites} sites x {a.reads} reads, data.json {off / 1e6:.1f} MB -> {a.out_dir}")


if __name__ == "__main__":
    main()
