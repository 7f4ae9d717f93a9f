for tm in 7 5 8; do
  echo "== tile_models=$tm"
  python profiles/other_configs.py --steps 200 --warmup 10 --opt tile_models=$tm 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('  %-40s %.4f ms %.3e evals/s frac %.4f M %d grid %d ctas %d'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s'], d['roofline_frac'], d['tile_models'], d['grid'], d['ctas_per_sm']))
"
done
