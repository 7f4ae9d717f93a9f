for opts in "" "--opt threads=128 --opt tile_models=4" "--opt threads=128 --opt tile_models=6" "--opt threads=128 --opt tile_models=8" "--opt threads=192 --opt tile_models=6" "--opt variant=4" "--opt variant=4 --opt threads=128 --opt tile_models=4"; do
  echo "== $opts"
  python profiles/other_configs.py --steps 200 --warmup 10 $opts 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if 'config5' in d['shape']: continue
    print('  %-40s %.4f ms %.3e evals/s M %d grid %d ctas %d smem %d'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s'], d['tile_models'], d['grid'], d['ctas_per_sm'], d['smem_bytes']))
"
done
