#!/bin/bash
# ab2.sh <lib.so>: config-2 bench line + config-5 slice ms for an experimental build
lib=$1
RTB200_LIB=$PWD/$lib bash profiles/quick_bench.sh "--variant 1" | sed "s|^|$lib |"
RTB200_LIB=$PWD/$lib python profiles/other_configs.py --only config5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   config5 slice ms', d['ms_per_step'], 'ctas', d['ctas_per_sm'])"
