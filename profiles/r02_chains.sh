python profiles/config4_chains.py --sweeps 30 --moves 20 | tee gpurun_out/c4chains_graph.json
python profiles/config4_chains.py --sweeps 30 --moves 20 --eager | tee gpurun_out/c4chains_eager.json
python profiles/config4_chains.py --sweeps 30 --moves 20 --replicas 64 | tee gpurun_out/c4chains_graph_64.json
