# end-to-end config 2 through dff_batch by waves of tiles per host chunk (default: 3)
for w in 0 2 4 6 9 12 0; do
  python bench.py --steps 8 --warmup 3 --no-cpu --opt chunk_waves=$w 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('chunk_waves=$w', 'resident %.4f ms  e2e %.4f ms %.3e evals/s  pageable %.4f ms %.3e'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['pageable']['ms_per_step'], d['e2e']['pageable']['value']))
"
done
