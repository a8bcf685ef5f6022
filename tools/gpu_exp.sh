#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.txt
run() { echo "== B=$B $*" | tee -a gpurun_out/exp.txt; env "$@" timeout 200 python tools/ncu_target.py --batch $B --frames 100 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/exp.txt; }
B=32
run X=0
run CSM_ATT_PS=32
B=16
run X=0
B=8
run X=0
run CSM_ATT_PS=32
B=4
run X=0
B=3
run X=0
