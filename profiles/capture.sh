set -e
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_h.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01_h python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_h.log 2>&1
python profiles/uniform_workload.py > gpurun_out/uniform_h.txt 2>&1
python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_h.json 2>/dev/null
python profiles/latency_config1.py > gpurun_out/latency_h.json 2>&1 || true
tail -1 gpurun_out/bench_h.json
cat gpurun_out/uniform_h.txt
