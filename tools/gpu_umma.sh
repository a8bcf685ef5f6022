#!/bin/bash
# tcgen05 GEMM prototype self-check (tools/micro/umma_gemm.cu) at three shapes, each under its own timeout.
mkdir -p gpurun_out
for shape in "256 256 256" "1024 2048 2048" "2048 8192 2048" "2048 2048 8192"; do
  timeout 40 tools/micro/umma_gemm $shape 2>&1 | tail -2; echo "umma_gemm $shape rc=${PIPESTATUS[0]}"
done | tee gpurun_out/umma_gemm.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee -a gpurun_out/umma_gemm.txt
