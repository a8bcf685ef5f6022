#!/bin/bash
# scratch experiment runner (one gpurun call): correctness subset, timing of library variants, ncu captures
mkdir -p gpurun_out; rm -f gpurun_out/exp.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "decisive or decode_frames or batch_invariance or families_agree or batched_vs_reference or padded or protocol_stress or rmsnorm" 2>&1 | tail -5 | tee gpurun_out/exp_pytest.log
run() { echo "== B=$B $*" | tee -a gpurun_out/exp.txt; env "$@" timeout 200 python tools/ncu_target.py --batch $B --frames 100 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/exp.txt; }
for B in 32 16 8 4; do
  run X=0
  run CSM_LIB=$PWD/gpurun_variants/lib_pieces0.so
done
for b in 8 32; do
  CSM_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_batch_kernel -s 2 -c 1 -o gpurun_out/ncu_batch_b${b}_nopair -f \
    python tools/ncu_target.py --batch $b --frames 5 > gpurun_out/ncu_full_b${b}.log 2>&1
  tail -3 gpurun_out/ncu_full_b${b}.log
done
timeout 300 python tools/phase_profile.py --batch 32 2>&1 | grep -v Warning > gpurun_out/phase_b32_exp.txt
ls -la gpurun_out/*.ncu-rep
