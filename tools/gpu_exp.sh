#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "decode_frames or invariance or bench_config_batched or stress" 2>&1 | tail -3 | tee -a gpurun_out/exp.txt
run() { echo "== B=$B $*" | tee -a gpurun_out/exp.txt; env "$@" timeout 200 python tools/ncu_target.py --batch $B --frames 100 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/exp.txt; }
B=8
run X=0
run CSM_ATT_NSUB=2
B=16
run X=0
B=32
run X=0
for b in 8 32; do timeout 200 python tools/phase_profile.py --batch $b 2>&1 | grep -v Warning > gpurun_out/phase_b${b}.txt; done
