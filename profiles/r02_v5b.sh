for n in default s2 s6 s8 w5 w9 default; do
  if [ $n = default ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/build/ab/lib_$n.so; fi
  python bench.py --steps 8 --warmup 3 --no-cpu --models 1000001 --variant 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', '%.4f ms %.4e evals/s e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']))
"
done
