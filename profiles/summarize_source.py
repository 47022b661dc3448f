#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall reasons, opcode mix, hottest source lines."""
import csv, collections, sys

def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    idx = {}
    for i, h in enumerate(hdr):
        idx.setdefault(h, i)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter(); ops = collections.Counter(); opinst = collections.Counter(); thr = collections.Counter()
    samples = 0
    def num(x):
        try: return int(float(x))
        except Exception: return 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] == "Address": continue
        for s in stalls: tot[s] += num(r[idx[s]])
        toks = r[idx["Source"]].strip().split()
        if not toks: continue
        o = toks[0] if not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else toks[0])
        o = o.split(".")[0]
        ns = num(r[idx["# Samples"]]); samples += ns
        ops[o] += ns; opinst[o] += num(r[idx["Instructions Executed"]]); thr[o] += num(r[idx["Thread Instructions Executed"]])
    print("SASS lines", len(rows) - hi - 1, "total samples", samples)
    for s, v in tot.most_common(12): print(f"  {s:26s} {v:8d} {100*v/max(samples,1):5.1f}%")
    ti = sum(opinst.values())
    print("warp instructions", ti, "thread instructions", sum(thr.values()), "avg active", sum(thr.values())/max(ti,1))
    for o, v in opinst.most_common(top):
        print(f"  {o:10s} inst {v:12d} {100*v/ti:5.1f}%   samples {ops[o]:7d} {100*ops[o]/max(samples,1):5.1f}%")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
