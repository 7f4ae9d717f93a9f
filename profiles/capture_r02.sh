# Capture recipe for the profiles/r02_* evidence (run on a B200 through gpurun from the repo root).
# Every ncu pass repeats a command that has just exited 0 without ncu.
set -e
B="python bench.py --steps 2 --warmup 1 --no-cpu"
$B > gpurun_out/plain_r02.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launch_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_c2 $B > gpurun_out/ncu_full_r02.log 2>&1
O="python profiles/other_configs.py --steps 5 --warmup 2 --only config3u"
$O > gpurun_out/plain_r02_c3u.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_c3u $O > gpurun_out/ncu_full_r02_c3u.log 2>&1
O="python profiles/other_configs.py --steps 5 --warmup 2 --only config4"
$O > gpurun_out/plain_r02_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_c4 $O > gpurun_out/ncu_full_r02_c4.log 2>&1
echo captured
