python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/other_configs.py --steps 200 --warmup 10 2>/dev/null | tee gpurun_out/other_configs_r02.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('  %-40s %.4f ms %.3e evals/s frac %.4f M %d grid %d ctas %d'%(d['shape'][:40], d['ms_per_step'], d['evals_per_s'], d['roofline_frac'], d['tile_models'], d['grid'], d['ctas_per_sm']))
"
python profiles/config4_chains.py --sweeps 30 --moves 20 | cut -c1-400
