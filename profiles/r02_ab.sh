# usage: r02_ab.sh [pytest-k-expr] -- "bench args" "bench args" ...
set +e
k="$1"; shift; shift
if [ -n "$k" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -m gpu -x -q -k "$k" 2>&1 | tail -5; fi
i=0
for args in "$@"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu $args 2>gpurun_out/ab_$i.err | tail -1 > gpurun_out/ab_$i.json
  python - "$args" gpurun_out/ab_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); c=d["config"]
    print(sys.argv[1],"| v",c["kernel_variant"],"%.3e"%d["value"],"%.3f ms"%d["ms_per_step"],"e2e %.3e"%d["e2e"]["value"],"ctas",c["ctas_per_sm"],"M",c["tile_models"],"smem",c["smem_bytes"],"thr",c["threads"])
except Exception as e:
    print(sys.argv[1],"FAILED",e)
PY
done
