#!/bin/bash
# quick A/B: python bench.py with debug flags, prints variant, evals/s, ms/step, e2e, frac, ctas, tile
for args in "$@"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu $args 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$args', '| v',c['kernel_variant'], '%.3e'%d['value'], '%.2f ms'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'frac %.4f'%d['roofline']['frac'], 'ctas',c['ctas_per_sm'],'M',c['tile_models'],'smem',c['smem_bytes'],'thr',c['threads'])"
done
