#!/bin/bash
# Round-end evidence in one call: all GPU tests, the bench line as the driver runs it (+ the reference arm), per-phase
# profiles, ncu launch lists and --set full captures of the kernels that matter (generation and training).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
BENCH_VERBOSE=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for b in 1 8 32; do timeout 300 python tools/phase_profile.py --batch $b 2>&1 | grep -v Warning > gpurun_out/phase_b${b}.txt; done
for b in 1 8; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_b$b.csv \
      python tools/ncu_target.py --batch $b --frames 4 > gpurun_out/ncu_launch_b$b.log 2>&1
done
# (ncu cannot replay the cooperative CTA-pair launch: CSM_PAIR=0 for the capture; DRAM traffic is the same)
CSM_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_batch_kernel -s 2 -c 1 -o gpurun_out/ncu_batch_b8 -f \
    python tools/ncu_target.py --batch 8 --frames 5 > gpurun_out/ncu_full_b8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_gemm_kernel -s 8 -c 8 -o gpurun_out/ncu_gemm_b8 -f \
    python tools/ncu_target.py --batch 8 --frames 2 > gpurun_out/ncu_full_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_flash_tc_kernel -s 3 -c 1 -o gpurun_out/ncu_flash_tc_b8 -f \
    python tools/ncu_target.py --batch 8 --frames 2 > gpurun_out/ncu_full_flash.log 2>&1
# training step
timeout 400 python tools/train_bench.py --seq 4096 --batch 1 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/train_bench_b1.json
bash tools/gpu_train_prof.sh > gpurun_out/train_prof.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_flash_tc_bwd_kernel -s 17 -c 1 -o gpurun_out/ncu_flash_tc_bwd -f \
    python tools/train_bench.py --seq 4096 --batch 1 --steps 1 --warmup 1 > gpurun_out/ncu_ftcb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csm_gemm_kernel -s 300 -c 12 -o gpurun_out/ncu_gemm_train -f \
    python tools/train_bench.py --seq 4096 --batch 1 --steps 1 --warmup 1 > gpurun_out/ncu_gemm_train.log 2>&1
timeout 200 python tools/flash_tc_bench.py --seq 4096 > gpurun_out/flash_tc_bench.txt 2>&1
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'verified', (j['verified'] or {}).get('ok'), 'launches', j['gpu_launches'])
for p in j['config']['points']: print(p['batch_per_gpu'], round(p['value'],1), 'ms/frame', round(p['decode_ms_per_frame'],3), 'frac', round(p['roofline']['frac'],3), 'prefill ms', round(p['prefill']['ms_incl_first_frame'],2), 'tf frac', round(p['prefill']['frac'],3))
print(j['config'].get('training'))
print(j.get('cpu_baseline'))
print(open('gpurun_out/bench_ref.json').read()[-700:])
print(open('gpurun_out/train_bench_b1.json').read()[:400])
print(open('gpurun_out/flash_tc_bench.txt').read())
PY
head -14 gpurun_out/train_prof.txt
ls -la gpurun_out/*.ncu-rep
