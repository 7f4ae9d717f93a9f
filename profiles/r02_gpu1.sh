set +e
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -3 gpurun_out/bench_a.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1])
print("value %.3e ms %.3f e2e %.3e pageable %.3e (%.2f of pinned)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["pageable"]["value"],d["e2e"]["pageable"]["frac_of_pinned"]))
for k,c in d["configs"].items():
    print(k, "%.3f ms"%c["ms_per_step"], "%.3e evals/s"%c["evals_per_s"], "frac %.4f"%c["roofline"]["frac"], {kk:vv for kk,vv in c.items() if kk in ("swap_exposed_ms_per_round","ms_per_step_without_swap","mh_moves_per_s")})
PY
