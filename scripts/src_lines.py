#!/usr/bin/env python
"""Attribute an `ncu --page source --csv` export to source lines: nvdisasm -g of the same kernel gives the line of every SASS
instruction; the i-th instruction of the export is the i-th instruction of the disassembly.
usage: src_lines.py export.csv cubin mangled-kernel-substring [top]"""
import csv, collections, re, subprocess, sys

def disasm_lines(cubin, sub):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    res, on, cur = [], False, ("?", 0)
    for ln in out:
        if ln.startswith("//--------------------- .text."):
            on = sub in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            res.append(cur)
    return res

def main():
    path, cubin, sub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = disasm_lines(cubin, sub)
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    idx = {}
    for i, h in enumerate(hdr):
        idx.setdefault(h, i)
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0] != "Address"]
    n = len(lines)
    if len(data) % n != 0:
        print("warning: %d export rows vs %d disassembled instructions" % (len(data), n))
    def num(x):
        try: return int(float(x))
        except Exception: return 0
    samp, inst, fp64 = collections.Counter(), collections.Counter(), collections.Counter()
    for i, r in enumerate(data[:n]):
        k = lines[i] if i < n else ("?", 0)
        samp[k] += num(r[idx["# Samples"]])
        inst[k] += num(r[idx["Instructions Executed"]])
        op = r[idx["Source"]].split()
        o = (op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")).split(".")[0]
        if o in ("DFMA", "DMUL", "DADD", "DSETP"):
            fp64[k] += num(r[idx["Instructions Executed"]])
    ts, ti = sum(samp.values()), sum(inst.values())
    print("total samples %d, warp instructions %d" % (ts, ti))
    for k, v in samp.most_common(top):
        print("%-16s:%-5d samples %6.2f%%  inst %6.2f%%  fp64 share of line %4.0f%%" % (k[0], k[1], 100 * v / ts, 100 * inst[k] / ti, 100 * fp64[k] / max(inst[k], 1)))

if __name__ == "__main__":
    main()
