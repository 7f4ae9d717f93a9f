# resident config-2 step time against the tile geometry (dynamic tile scheduling)
for o in "" "--opt tile_models=16" "--opt tile_models=24" "--opt tile_models=48" "--opt tile_models=64 --opt ctas_per_sm=2" "--opt ctas_per_sm=3 --opt tile_models=48" "--opt ctas_per_sm=3" "--opt ctas_per_sm=5" "--opt threads=128 --opt tile_models=16" "--opt threads=128 --opt tile_models=32"; do
  python bench.py --steps 10 --warmup 3 --no-cpu $o | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RES', '$o', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
