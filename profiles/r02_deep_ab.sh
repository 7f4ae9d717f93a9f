for n in default u2c3 u3c3 u4c3 u2c2; do
  if [ $n = default ]; then unset RTB200_LIB; else export RTB200_LIB=$PWD/build/ab/lib_$n.so; fi
  for ctas in 0; do
  python profiles/other_configs.py --steps 6 --warmup 2 --only config5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$n', '%.3f ms %.3e evals/s frac %.4f M %d grid %d ctas %d smem %d'%(d['ms_per_step'], d['evals_per_s'], d['roofline_frac'], d['tile_models'], d['grid'], d['ctas_per_sm'], d['smem_bytes']))
"
  done
done
