# small batches (configs 3 and 4): step time against models per tile
for tm in 0 2 4 6 8 12 16; do
  for c in config3 config4; do
    python profiles/other_configs.py --only $c --opt tile_models=$tm | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RES tile_models=$tm', d['shape'][:8], round(d['ms_per_step'],4), d['tile_models'], d['grid'], d['ctas_per_sm'])"
  done
done
