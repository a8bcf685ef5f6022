#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.txt
run() { echo "== B=$B $*" | tee -a gpurun_out/exp.txt; env "$@" timeout 200 python tools/phase_profile.py --batch $B 2>&1 | grep -E "decode frame|attn_bb|K-stream" | tee -a gpurun_out/exp.txt; }
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "bench_config or attention or invariance or padded" 2>&1 | tail -4 | tee -a gpurun_out/exp.txt
B=32
run X=1
run CSM_A_TPC=16
B=16
run X=1
B=8
run X=1
run CSM_A_TPC=32
run CSM_A_TPC=16
