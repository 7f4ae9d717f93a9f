# usage: r02_ab_generic.sh name1 name2 ...   (libraries build/ab/lib_<name>.so; "default" = the product library)
for n in "$@"; do
  if [ $n = default ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/build/ab/lib_$n.so; fi
  python bench.py --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$n', '%.4f ms %.4e evals/s e2e %.3e v%d'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel']['kernel_variant']))
"
done
unset RTB200_LIB
