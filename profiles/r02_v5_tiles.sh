# variant 5 on config 2: tile geometry sweep (models per tile, CTAs per SM, threads)
run() { python bench.py --steps 8 --warmup 3 --no-cpu --variant 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['kernel']
print('$*', '%.4f ms %.4e evals/s'%(d['ms_per_step'], d['value']), k['tile_models'], k['tile_sources'], k['threads'], k['ctas_per_sm'], k['smem_bytes'])
"; }
run
run --opt tile_models=24
run --opt tile_models=28
run --opt tile_models=36
run --opt tile_models=40
run --opt tile_models=42
run --opt tile_models=48 --opt ctas_per_sm=2
run --opt tile_models=64 --opt ctas_per_sm=2
run --opt threads=192
run --opt threads=192 --opt tile_models=24
run --opt threads=128 --opt tile_models=16
run
