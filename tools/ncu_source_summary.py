#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: executed instructions by opcode and by address range,
top stall sites.  Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_source_summary.py src.csv"""
import csv
import sys
from collections import Counter


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    tot_inst = sum(int(r[col["Instructions Executed"]]) for r in body)
    tot_samp = sum(int(r[col["# Samples"]]) for r in body)
    by_op = Counter()
    samp_op = Counter()
    for r in body:
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[0] if not toks[0].startswith("@") else toks[1]
        op = op.split(".")[0] if not op.startswith(("LDS", "LDG", "STG", "STS", "LDL", "STL", "IMAD", "ATOMS")) else ".".join(op.split(".")[:2])
        by_op[op] += int(r[col["Instructions Executed"]])
        samp_op[op] += int(r[col["# Samples"]])
    print(f"total warp instructions {tot_inst:,}   total samples {tot_samp:,}   SASS lines {len(body)}")
    print(f"{'opcode':14s} {'inst':>15s} {'%inst':>7s} {'%samples':>9s}")
    for op, n in by_op.most_common(top):
        print(f"{op:14s} {n:15,d} {100*n/tot_inst:7.2f} {100*samp_op[op]/max(1,tot_samp):9.2f}")
    # stall reasons overall
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    st = Counter()
    for r in body:
        for h in stall_cols:
            st[h] += int(r[col[h]] or 0)
    tot = sum(st.values())
    print("\nstall reasons (all samples):")
    for h, n in st.most_common(10):
        print(f"  {h:28s} {100*n/max(1,tot):6.2f}%")
    # hottest lines
    print("\nhottest SASS lines by samples:")
    hot = sorted(body, key=lambda r: -int(r[col["# Samples"]]))[:top]
    for r in hot:
        print(f"  {int(r[col['# Samples']]):8d}  {int(r[col['Instructions Executed']]):12,d}  {r[col['Source']].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)


def regions(path):
    """Bucket samples into the read-encoder loop (FFMA2 range), the MC loop (MWC multiplier range) and the rest."""
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    src = [r[col["Source"]] for r in body]
    fa = [i for i, s in enumerate(src) if "FFMA2" in s]
    fb = [i for i, s in enumerate(src) if "147e5" in s]
    A = (min(fa) - 12, max(fa) + 12) if fa else (0, -1)
    B = (min(fb) - 30, max(fb) + 40) if fb else (0, -1)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[col["# Samples"]]) for r in body)
    print(f"\nregions (SASS index ranges): encoder loop {A}, MC loop {B}; total samples {tot}")
    for name, sel in (("encoder loop", lambda i: A[0] <= i <= A[1]), ("MC loop", lambda i: B[0] <= i <= B[1]),
                      ("other", lambda i: not (A[0] <= i <= A[1]) and not (B[0] <= i <= B[1]))):
        rs = [r for i, r in enumerate(body) if sel(i)]
        n = sum(int(r[col["# Samples"]]) for r in rs)
        inst = sum(int(r[col["Instructions Executed"]]) for r in rs)
        st = Counter()
        for r in rs:
            for h in stall_cols:
                st[h] += int(r[col[h]] or 0)
        top = ", ".join(f"{h[6:]} {100*v/max(1,n):.0f}%" for h, v in st.most_common(6))
        print(f"  {name:13s} samples {100*n/max(1,tot):5.1f}%  warp-instr {inst:15,d}   {top}")


if __name__ == "__main__" and len(sys.argv) > 1:
    regions(sys.argv[1])
