#!/bin/bash
# One gpurun call: parity tests, bench lines (B=1 default, B=8, B=32), per-phase profiles, ncu launch list and one full
# capture of the frame kernel.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
cat gpurun_out/bench_b1.json
timeout 200 python bench.py --steps 2 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err
timeout 200 python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err
cat gpurun_out/bench_b8.json gpurun_out/bench_b32.json
tail -2 gpurun_out/bench_b32.err
for b in 1 8 32; do timeout 120 python tools/phase_profile.py --batch $b > gpurun_out/phase_b$b.txt 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_b1.csv \
    python tools/ncu_target.py --batch 1 --frames 12 > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csm_stream_kernel -s 4 -c 1 -f -o gpurun_out/stream_b1 \
    python tools/ncu_target.py --batch 1 --frames 8 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
