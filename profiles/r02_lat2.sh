set -e
gcc -O2 profiles/latency_c.c -o /tmp/latency_c -Lraytracerfortran_b200 -lraytrace_b200 -Loracle -loracle_raymod -Wl,-rpath,$PWD/raytracerfortran_b200 -Wl,-rpath,$PWD/oracle -lm
nvcc -O2 -gencode arch=compute_100a,code=sm_100a profiles/launch_floor.cu -o /tmp/launch_floor
/tmp/launch_floor | tee gpurun_out/launch_floor.json
/tmp/latency_c | tee gpurun_out/latency_c.json
python profiles/latency_config1.py | tee gpurun_out/latency_py.json
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
