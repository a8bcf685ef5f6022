#!/bin/bash
# general-kernel subset of the GPU tests + decode ms/frame and per-phase profiles at 8 / 32 sequences; library variants
mkdir -p gpurun_out; rm -f gpurun_out/decode_ms.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "${TESTS:-not stress and not sampling and not linear}" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for b in ${BATCHES:-8 32}; do
  timeout 200 python tools/ncu_target.py --batch $b --frames 60 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/decode_ms.txt
  for f in gpurun_variants/lib_*.so; do
    [ -f "$f" ] || continue
    CSM_LIB=$PWD/$f timeout 200 python tools/ncu_target.py --batch $b --frames 60 --reps 2 2>&1 | tail -1 | sed "s|^|$f |" | tee -a gpurun_out/decode_ms.txt
  done
  timeout 200 python tools/phase_profile.py --batch $b 2>&1 | grep -v Warning > gpurun_out/phase_b${b}.txt
done
