python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for v in 1 5 1 5; do
  python bench.py --steps 8 --warmup 3 --no-cpu --models 1000001 --variant $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $v', '%.4f ms %.4e evals/s e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']))
"
done
for B in 20000 60000 150000 400000; do for v in 1 5; do
  python bench.py --steps 20 --warmup 3 --no-cpu --models $B --variant $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B $B variant $v', '%.4f ms %.4e evals/s'%(d['ms_per_step'], d['value']))
"
done; done
python profiles/other_configs.py --steps 100 --warmup 10 --opt variant=5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   v5  %-40s %.4f ms %.3e evals/s'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s']))
"
