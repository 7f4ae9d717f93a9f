for o in "--opt static_tiles=1" "" "--opt static_tiles=1 --opt comp_streams=1 --opt chunk_models=65536"; do
  python bench.py --steps 10 --warmup 3 --no-cpu $o | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RES', '$o', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
