#!/bin/bash
# ncu evidence: launch lists (device time per launch, serialised, cold cache: compare SHARES) of a short generate() at
# batch 1 and 8, and one --set full capture each of the frame kernels and of the tcgen05 GEMM
mkdir -p gpurun_out
for b in 1 8; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_b$b.csv \
      python tools/ncu_target.py --batch $b --frames 4 > gpurun_out/ncu_launch_b$b.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_stream_kernel -s 2 -c 1 -o gpurun_out/ncu_stream_b1 -f \
    python tools/ncu_target.py --batch 1 --frames 5 > gpurun_out/ncu_full_b1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_batch_kernel -s 2 -c 1 -o gpurun_out/ncu_batch_b8 -f \
    python tools/ncu_target.py --batch 8 --frames 5 > gpurun_out/ncu_full_b8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_gemm_kernel -s 70 -c 4 -o gpurun_out/ncu_gemm_b8 -f \
    python tools/ncu_target.py --batch 8 --frames 2 > gpurun_out/ncu_full_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
