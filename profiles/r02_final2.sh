# final evidence of round 2 (one B200): GPU suite, ncu captures, traffic.json, parity soak, bench both arms
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash profiles/capture_r02.sh
python profiles/make_traffic.py gpurun_out/prof_r02_c2.ncu-rep ${1:-61964fa} profiles/traffic.json gpurun_out/traffic.json > /dev/null
for n in c2 c3u c4; do python profiles/extract_metrics.py gpurun_out/prof_r02_$n.ncu-rep > gpurun_out/r02_${n}_ncu_full.txt; done
python profiles/parity_soak.py --seconds 120 --seed 71 > gpurun_out/parity_soak_r02.json 2> gpurun_out/parity_soak_r02.err; cut -c1-400 gpurun_out/parity_soak_r02.json
python profiles/other_configs.py --steps 50 --warmup 10 > gpurun_out/other_configs_r02.jsonl 2>/dev/null
bash profiles/r02_final.sh
