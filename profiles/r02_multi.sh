N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 profiles/swap_check_multi.py 2>/dev/null | grep "^{" | tee gpurun_out/swap_check_n$N.json
run 29512 profiles/config5_full.py $((2000000*N)) 2>/dev/null | grep "^{" | tee gpurun_out/config5_full_n$N.json
run 29513 profiles/config4_chains.py --sweeps 30 --moves 20 2>/dev/null | grep "^{" | tee gpurun_out/config4_chains_n$N.json
NCCL_DEBUG=INFO run 29514 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.out 2> gpurun_out/bench_n$N.err
grep -h "NCCL INFO.*nranks" gpurun_out/bench_n$N.out gpurun_out/bench_n$N.err | grep "Init COMPLETE" | head -2 | cut -c1-160
grep "^{" gpurun_out/bench_n$N.out | tail -1 > gpurun_out/bench_n$N.json
python - $N <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_n%s.json"%sys.argv[1]).read())
print("N",d["n_gpus"],"value %.3e ms %.3f e2e %.3e pageable %.3e"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["pageable"]["value"]))
for k,c in d["configs"].items():
    print(k, "%.3f ms"%c["ms_per_step"], "%.3e evals/s"%c["evals_per_s"], "frac %.4f"%c["roofline"]["frac"], {kk:vv for kk,vv in c.items() if kk in ("swap_exposed_ms_per_round","ms_per_step_without_swap","mh_moves_per_s","collective")})
PY
