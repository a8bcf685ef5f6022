#!/bin/bash
# GPU tests + decode ms/frame and per-phase profiles at 8 / 32 sequences (general kernels), optional variants via env
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for b in ${BATCHES:-8 32}; do
  timeout 200 python tools/ncu_target.py --batch $b --frames 60 --reps 2 2>&1 | tail -1 | tee -a gpurun_out/decode_ms.txt
  timeout 200 python tools/phase_profile.py --batch $b 2>&1 | grep -v Warning > gpurun_out/phase_b${b}.txt
done
