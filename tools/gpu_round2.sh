#!/bin/bash
# One gpurun call: parity tests, bench line, per-phase profiles, ncu launch list and one full capture of the frame kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
cat gpurun_out/bench_b1.json
timeout 300 python tools/phase_profile.py --batch 1 > gpurun_out/phase_b1.txt 2>&1
CSM_FUSE_ATTN=1 timeout 300 python tools/phase_profile.py --batch 1 > gpurun_out/phase_b1_fuse.txt 2>&1
timeout 300 python tools/phase_profile.py --batch 8 > gpurun_out/phase_b8.txt 2>&1
timeout 300 python tools/phase_profile.py --batch 32 > gpurun_out/phase_b32.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_b1.csv \
    python tools/ncu_target.py --batch 1 --frames 12 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csm_stream_kernel -s 4 -c 1 -f -o gpurun_out/stream_b1 \
    python tools/ncu_target.py --batch 1 --frames 8 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
head -30 gpurun_out/phase_b1.txt; head -12 gpurun_out/phase_b1_fuse.txt; head -30 gpurun_out/phase_b8.txt; head -4 gpurun_out/phase_b32.txt
