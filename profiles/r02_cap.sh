# ncu full capture of one batch kernel launch; usage: r02_cap.sh <tag> <bench args...>
tag=$1; shift
python bench.py --steps 2 --warmup 1 --no-cpu "$@" > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 2 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
