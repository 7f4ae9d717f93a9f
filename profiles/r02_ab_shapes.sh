# usage: r02_ab_shapes.sh name1 name2 ... : config 2, a 60 000-model batch (variant 1 by the default rule) and the
# other shapes with each library (build/ab/lib_<name>.so; "default" = the product library)
for n in "$@"; do
  if [ $n = default ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/build/ab/lib_$n.so; fi
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -1
  for B in 0 60000; do
  python bench.py --steps 8 --warmup 3 --no-cpu --models $B 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$n B=$B', '%.4f ms %.4e evals/s e2e %.3e v%d'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel']['kernel_variant']))
"
  done
  python profiles/other_configs.py --steps 50 --warmup 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   $n  %-40s %.4f ms %.3e evals/s'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s']))
"
done
unset RTB200_LIB
