#!/bin/bash
# Warp-per-unit backbone attention: parity tests, then decode ms/frame and per-phase profile at 8 and 32 sequences.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for b in 8 32; do
  timeout 200 python tools/ncu_target.py --batch $b --frames 60 --reps 2 2>&1 | tail -1
  timeout 200 python tools/phase_profile.py --batch $b > gpurun_out/phase_b${b}_v9.txt 2>&1
  head -8 gpurun_out/phase_b${b}_v9.txt
done
