python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/r02_cap.sh v5 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', '%.4f ms %.4e evals/s e2e %.3e pageable %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['pageable']['value']), d['kernel'])
for k,c in d['configs'].items(): print(k, '%.3f ms'%c['ms_per_step'], '%.3e'%c['evals_per_s'])
"
