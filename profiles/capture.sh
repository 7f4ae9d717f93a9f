# Capture recipe for the profiles/r01_h_* files (run on a B200 through gpurun from the repo root).
# Every ncu pass repeats a command that has just exited 0 without ncu.
set -e
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_h.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01_h python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_h.log 2>&1
python profiles/uniform_workload.py > gpurun_out/uniform_h.txt 2>&1
python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_h.json 2>/dev/null
python profiles/latency_config1.py > gpurun_out/latency_h.json 2>&1 || true
python profiles/other_configs.py --steps 50 > gpurun_out/other_configs_h.jsonl 2>/dev/null || true
python profiles/config4_chains.py --sweeps 20 --moves 20 > gpurun_out/config4_chains_h.json 2>/dev/null || true
python profiles/e2e_with_times.py > gpurun_out/e2e_with_times_h.json 2>/dev/null || true
tail -1 gpurun_out/bench_h.json
cat gpurun_out/uniform_h.txt
