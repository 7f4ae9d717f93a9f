export RTB200_LIB=$PWD/build/ab/lib_seg.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
for n in seg segv0 seg segv0; do
  if [ $n = default ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/build/ab/lib_$n.so; fi
  python bench.py --steps 8 --warmup 3 --no-cpu --models 1000001 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', '%.4f ms %.4e evals/s e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']))
"
  python profiles/other_configs.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   $n  %-40s %.4f ms %.3e evals/s'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s']))
"
done
