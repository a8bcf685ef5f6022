#!/bin/bash
# Parity tests + per-phase profile + decode timing over 100 frames (one gpurun call).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for b in ${BATCHES:-1}; do
  timeout 300 python tools/ncu_target.py --batch $b --frames 100 --reps 2
  timeout 300 python tools/phase_profile.py --batch $b > gpurun_out/phase_b$b.txt 2>&1
  head -${LINES_PER:-18} gpurun_out/phase_b$b.txt
done
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
