# config-2 capture and bench only (after a host-side change: the kernels are the same, the stamp is not)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 2 --warmup 1 --no-cpu"
$B > gpurun_out/plain_r02.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launch_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_c2 $B > gpurun_out/ncu_full_r02.log 2>&1
python profiles/make_traffic.py gpurun_out/prof_r02_c2.ncu-rep ${1:-5e1e4e1} profiles/traffic.json gpurun_out/traffic.json > /dev/null
python profiles/extract_metrics.py gpurun_out/prof_r02_c2.ncu-rep > gpurun_out/r02_c2_ncu_full.txt
bash profiles/r02_final.sh
