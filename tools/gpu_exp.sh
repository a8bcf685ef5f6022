#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.txt
run() { echo "== B=$B $*" | tee -a gpurun_out/exp.txt; env "$@" timeout 200 python tools/ncu_target.py --batch $B --frames 100 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/exp.txt; }
B=1
run CSM_L2_KEEP_LAYERS=0
run CSM_L2_KEEP_LAYERS=1
run CSM_L2_KEEP_LAYERS=2
run CSM_L2_KEEP_LAYERS=3
run CSM_L2_KEEP_LAYERS=2 CSM_L2_AHEAD_KB=512
B=8
run CSM_L2_KEEP_LAYERS=0
run CSM_L2_KEEP_LAYERS=2
