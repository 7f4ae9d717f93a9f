python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/bench_ref_final.json").read().strip().splitlines()[-1])
print("same_config:", d["config"]==r["config"])
print("value %.3e ms %.3f e2e %.3e pageable %.3e launches %d traffic %s pipe %s note %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["pageable"]["value"],d["gpu_launches"],d["roofline"]["traffic"],d["roofline"]["fp64_pipe_active_pct_ncu"],d["roofline"]["ncu_note"]))
print("frac %.4f cpu %.3e cores %d ref %.3e"%(d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], r["value"]))
print("kernel", d["kernel"])
print("clocks", d["clocks"])
for k,c in d["configs"].items():
    print(k, "%.3f ms"%c["ms_per_step"], "%.3e evals/s"%c["evals_per_s"], "frac %.4f"%c["roofline"]["frac"], "W %.1f"%c["roofline"]["flops_per_eval_min"])
PY
