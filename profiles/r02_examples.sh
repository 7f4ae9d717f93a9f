# the examples, end to end on a B200
python examples/test1_dff.py 2>&1 | tail -3
python examples/replica_sweep.py 2>&1 | tail -2
python examples/invert_test1.py --chains 256 --iters 200 2>&1 | tail -2
gcc -O2 -Iinclude examples/dff_batch_example.c -o /tmp/dff_example -Lraytracerfortran_b200 -lraytrace_b200 -Wl,-rpath,$PWD/raytracerfortran_b200 -lm && /tmp/dff_example | tail -3
python profiles/parity_soak.py --seconds 100 --seed 31 | cut -c1-300
