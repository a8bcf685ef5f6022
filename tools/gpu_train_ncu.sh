#!/bin/bash
# ncu --set full of one flash backward and one flash forward launch of the training step (csm-1b, 4096 tokens)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_bwd_kernel -s 26 -c 1 -o gpurun_out/ncu_flash_bwd -f \
   python tools/train_bench.py --seq 4096 --batch 1 --steps 1 --warmup 1 > gpurun_out/ncu_fb.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_fwd_kernel -s 20 -c 1 -o gpurun_out/ncu_flash_fwd -f \
   python tools/train_bench.py --seq 4096 --batch 1 --steps 1 --warmup 1 > gpurun_out/ncu_ff.log 2>&1
ls -la gpurun_out/ncu_flash_*.ncu-rep
