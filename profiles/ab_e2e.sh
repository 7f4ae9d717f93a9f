# e2e (host buffers through dff_batch) against the host pipeline's chunk size / compute streams
for o in "--opt chunk_models=9472" "--opt chunk_models=18944" "--opt chunk_models=37888" "" "--opt chunk_models=151552" "--opt comp_streams=1" ; do
  python bench.py --steps 10 --warmup 3 --no-cpu $o | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RES', '$o', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
