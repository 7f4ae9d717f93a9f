# config 2: the next tile's rows prefetched into their own staging buffer (default) against rows fetched at the
# start of the tile into the travel-time tile's memory, which leaves room for more models per tile
run() { python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); k=d['kernel']
print('$*', '%.4f ms %.4e evals/s e2e %.3e'%(d['ms_per_step'], d['value'], d['e2e']['value']), k['tile_models'], k['ctas_per_sm'], k['smem_bytes'], 'v%d'%k['kernel_variant'])
"; }
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
run
run --opt prefetch=0
run --opt prefetch=0 --opt tile_models=34
run --opt prefetch=0 --opt tile_models=36
run --opt prefetch=0 --opt tile_models=38
run
