#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full report of the config-2 launch: the DRAM traffic and
FP64-pipe figures bench.py quotes, stamped with the sha-256 of the kernel sources they were taken
on (bench.py refuses to quote them for other sources).

    python profiles/make_traffic.py report.ncu-rep <commit> [out.json ...]"""
import csv, datetime, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_sources_sha

rep, commit, outs = sys.argv[1], sys.argv[2], sys.argv[3:] or [os.path.join(ROOT, "profiles", "traffic.json")]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, val = rows[0], rows[1], rows[2]
def get(name, scale=True):
    i = hdr.index(name)
    x = float(val[i].replace(",", ""))
    u = units[i].lower()
    if scale:
        x *= {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    return x
dur = get("gpu__time_duration.sum", scale=False) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[units[hdr.index("gpu__time_duration.sum")]]
kern = val[hdr.index("Kernel Name")]
kern = kern[kern.index("rt_batch_kernel"):kern.index(">") + 1].replace("(int)", "")
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
d = {"source": "profiles/r02_c2_ncu_full.txt (ncu --set full --clock-control none of `python bench.py --steps 2 "
               f"--warmup 1 --no-cpu`, {kern}, third launch; recipe: profiles/capture_r02.sh)",
     "kernel": kern, "duration_ms": dur, "dram_bytes_read": rd, "dram_bytes_write": wr,
     "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": 188000000,
     "fp64_pipe_active_pct": round(get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False), 2),
     "issue_active_pct": round(get("smsp__issue_active.avg.pct_of_peak_sustained_active", False), 2),
     "warp_instructions": int(get("smsp__inst_executed.sum", False)),
     "threads_per_warp_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio", False),
     "threads_per_warp_instruction_note": "ncu's thread_inst/inst; idle lanes execute the convergent state update on dead "
                                          "values, so this is not the share of useful lanes (profiles/lane_hist.py measures that)",
     "kernel_sources_sha": kernel_sources_sha(), "commit": commit, "date": datetime.date.today().isoformat()}
for o in outs:
    json.dump(d, open(o, "w"), indent=1)
print(json.dumps(d))
