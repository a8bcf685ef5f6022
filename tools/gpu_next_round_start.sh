#!/bin/bash
# First GPU call of the next round: run the two experiments this round left compile-checked but unrun, next to the
# default build.  Before the call (on the CPU box):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/micro/umma_gemm tools/micro/umma_gemm.cu
#   python -c "from csm_hf_b200 import build; [build.build(defines=['CSM_ATT_RING=%d' % n], out='gpurun_variants/lib_attring%d.so' % n) for n in (2, 3)]"
# Every step runs under its own timeout: a hung tcgen05 kernel must not hold the box until gpurun's limit.
mkdir -p gpurun_out
if [ -x tools/micro/umma_gemm ]; then
  for shape in "256 256 256" "1024 2048 2048" "2048 8192 2048"; do
    timeout 40 tools/micro/umma_gemm $shape 2>&1 | tail -2; echo "umma_gemm $shape rc=$?"
  done | tee gpurun_out/umma_gemm.txt
fi
BATCHES="8 32" bash tools/gpu_attring.sh 2>&1 | tee gpurun_out/attring.txt
