# usage: r02_ab_tests.sh <lib name> : parity + fuzz tests against build/ab/lib_<name>.so, then A/B on config 2
export RTB200_LIB=$PWD/build/ab/lib_$1.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
unset RTB200_LIB
bash profiles/r02_ab_generic.sh default $1 default $1
