#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.txt
run() { echo "== B=$B $*" | tee -a gpurun_out/exp.txt; env "$@" timeout 200 python tools/ncu_target.py --batch $B --frames 60 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/exp.txt; }
for B in 8; do
run CSM_PAIR=1
run CSM_PAIR=0
run CSM_PAIR=0 CSM_LIB=$PWD/gpurun_variants/lib_oldattn.so
run CSM_PAIR=0 CSM_LIB=$PWD/gpurun_variants/lib_nopair.so
run CSM_PAIR=0 CSM_LIB=$PWD/gpurun_variants/lib_both.so
run CSM_PAIR=1 CSM_LIB=$PWD/gpurun_variants/lib_oldattn.so
done
B=32
run CSM_PAIR=1 CSM_LIB=$PWD/gpurun_variants/lib_oldattn.so
run CSM_PAIR=0 CSM_LIB=$PWD/gpurun_variants/lib_both.so
