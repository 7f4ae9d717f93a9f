#!/usr/bin/env python
"""Attribute ncu's per-SASS-instruction counters to CUDA source lines.

ncu's CSV source page lists SASS only; nvdisasm --print-line-info gives the line of every SASS
instruction.  Both list the kernel's instructions in address order, so they are joined by index.
Usage: line_profile.py <ncu source csv> <nvdisasm listing> <mangled kernel name> [top_n]"""
import csv, re, sys, collections
src_csv, listing, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(listing).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(f".text.{kern}:"))
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith("//---") : break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = int(m.group(2)) if "inlined at" not in m.group(3) else cur
        if "inlined at" in m.group(3):
            cur = int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
recs = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(recs) == len(seq), (len(recs), len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
fp64 = ("DFMA", "DMUL", "DADD", "DSETP")
tot = 0
for (ln, sass), r in zip(seq, recs):
    inst = int(r[idx["Instructions Executed"]] or 0); thr = int(r[idx["Predicated-On Thread Instructions Executed"]] or 0)
    a = agg[ln]; a[0] += inst; a[1] += thr; a[2] += int(r[idx["# Samples"]] or 0)
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    if op.split(".")[0] in fp64: a[3] += inst
    tot += inst
src = open("/root/repo/raytracerfortran_b200/csrc/rt_kernels.cu").read().splitlines()
print(f"total warp-inst {tot:.3e}")
print(f"{'line':>5s} {'warp-inst':>10s} {'share':>6s} {'fp64':>6s} {'lane':>5s} {'samp':>8s}  source")
for ln, (i, t, s, f) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[ln - 1].strip()[:90] if ln and ln <= len(src) else "?"
    print(f"{ln or 0:5d} {i:10.3e} {i/tot:6.3f} {f/max(i,1):6.2f} {t/32/max(i,1):5.2f} {s:8d}  {text}")
