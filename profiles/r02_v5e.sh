# branch-free phase update (default) against the nested-conditional form ("old") and predicated
# final stores ("fin"): config 2, then the other shapes (variants 1, 3, 5 by the default rule)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
for n in default fin old default fin; do
  if [ $n = default ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/build/ab/lib_$n.so; fi
  python bench.py --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$n', '%.4f ms %.4e evals/s e2e %.3e v%d'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel']['kernel_variant']))
"
  python profiles/other_configs.py --steps 50 --warmup 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   $n  %-40s %.4f ms %.3e evals/s'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s']))
"
done
unset RTB200_LIB
