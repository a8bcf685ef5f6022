#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "decode_frames or invariance or bench_config_batched or stress or errors or golden" 2>&1 | tail -3 | tee -a gpurun_out/exp.txt
run() { echo "== B=$B $*" | tee -a gpurun_out/exp.txt; env "$@" timeout 200 python tools/ncu_target.py --batch $B --frames 100 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/exp.txt; }
B=8
run X=0
B=32
run X=0
timeout 200 python tools/phase_profile.py --batch 8 2>&1 | grep -E "attn_dec|decode frame" | tee -a gpurun_out/exp.txt
