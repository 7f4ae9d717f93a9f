python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python profiles/sanitize_small.py 2>&1 | tail -3
bash profiles/r02_v1_ab.sh 2>&1 | tee gpurun_out/v1_ab.txt
bash profiles/r02_small3.sh 2>&1 | tee gpurun_out/small3.txt
python examples/invert_test1.py --chains 256 --iters 200 2>&1 | tail -1 | cut -c1-600
python profiles/parity_soak.py --seconds 150 --seed 7 | tee gpurun_out/parity_soak_r02.json
