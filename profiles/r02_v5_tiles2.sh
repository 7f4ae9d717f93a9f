run() { python bench.py --steps 8 --warmup 3 --no-cpu --variant 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['kernel']
print('$*', '%.4f ms %.4e evals/s'%(d['ms_per_step'], d['value']), k['tile_models'], k['tile_sources'], k['threads'], k['ctas_per_sm'], k['smem_bytes'])
"; }
run --opt threads=224 --opt tile_models=32
run --opt threads=192 --opt tile_models=32
run --opt threads=160 --opt tile_models=32
run --opt threads=224 --opt tile_models=28
run --opt threads=128 --opt tile_models=32
run --opt threads=128 --opt tile_models=24
run --opt tile_models=34
run --opt tile_models=30
