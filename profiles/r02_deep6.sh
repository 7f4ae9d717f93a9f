# variant 6 (the deep kernel with one list segment per warp) against variant 3 on the config-5 slice
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
for v in 3 6 3 6; do
python profiles/other_configs.py --steps 10 --warmup 3 --only config5 --opt variant=$v 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   v$v  %-40s %.4f ms %.3e evals/s frac %.4f'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s'], d['roofline_frac']))
"
done
