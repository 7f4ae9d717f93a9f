#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: instruction mix and lane efficiency per opcode,
and the hottest SASS ranges.  Usage: sass_summary.py src.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
tot_inst = tot_thr = tot_samp = 0
by_op = collections.defaultdict(lambda: [0, 0, 0])
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    sass = r[idx["Source"]].strip()
    inst = int(r[idx["Instructions Executed"]] or 0); thr = int(r[idx["Predicated-On Thread Instructions Executed"]] or 0)
    samp = int(r[idx["# Samples"]] or 0)
    toks = sass.split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0]
    by_op[op][0] += inst; by_op[op][1] += thr; by_op[op][2] += samp
    tot_inst += inst; tot_thr += thr; tot_samp += samp
    recs.append((inst, thr, samp, sass))
print(f"total warp-inst {tot_inst:.3e}  thread-inst {tot_thr:.3e}  lane-eff {tot_thr/32/tot_inst:.3f} samples {tot_samp}")
print(f"{'op':12s} {'warp-inst':>12s} {'share':>7s} {'lane-eff':>8s} {'samples%':>8s}")
for op, (i, t, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{op:12s} {i:12.3e} {i/tot_inst:7.3f} {t/32/max(i,1):8.3f} {100*s/max(tot_samp,1):8.2f}")
