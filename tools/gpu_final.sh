#!/bin/bash
# Round-end check of the build in the tree: all GPU parity tests, the default bench line, decode ms/frame at 32 and 8 sequences.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
cat gpurun_out/bench_b1.json
timeout 120 python tools/ncu_target.py --batch 32 --frames 40 --reps 1 2>&1 | tail -1
timeout 120 python tools/ncu_target.py --batch 8 --frames 60 --reps 1 2>&1 | tail -1
