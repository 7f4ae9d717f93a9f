N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/swap_check_multi.py 2>&1 | tail -2
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -c "NCCL INFO" gpurun_out/bench_n$N.err; grep -m2 "NCCL INFO.*ranks\|nranks" gpurun_out/bench_n$N.err | cut -c1-200
python - $N <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_n%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print("N",d["n_gpus"],"value %.3e ms %.3f e2e %.3e pageable %.3e"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["pageable"]["value"]))
for k,c in d["configs"].items():
    print(k, "%.3f ms"%c["ms_per_step"], "%.3e evals/s"%c["evals_per_s"], "frac %.4f"%c["roofline"]["frac"], {kk:vv for kk,vv in c.items() if kk in ("swap_exposed_ms_per_round","ms_per_step_without_swap","mh_moves_per_s","collective")})
PY
