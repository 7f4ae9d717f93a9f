# after extending the variant-5 rule to single-wave launches: full GPU suite, small shapes, config 2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for o in "" "--opt variant=1" "--opt variant=5" ""; do
python profiles/other_configs.py --steps 100 --warmup 10 $o 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   [$o]  %-40s %.4f ms %.3e evals/s M=%d grid=%d'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s'], d['tile_models'], d['grid']))
"
done
python profiles/other_configs.py --steps 20 --warmup 5 --only config5 2>/dev/null | cut -c1-200
