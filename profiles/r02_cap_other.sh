# ncu full capture of the batch kernel on a small trans-dimensional shape; usage: r02_cap_other.sh <tag> <shape substring>
tag=$1; shape="$2"
python profiles/other_configs.py --steps 5 --warmup 2 --only "$shape" > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rt_batch_kernel -s 3 -c 1 -f -o gpurun_out/prof_$tag python profiles/other_configs.py --steps 5 --warmup 2 --only "$shape" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log; cat gpurun_out/plain_$tag.log | cut -c1-300
