#!/bin/bash
# 2-GPU check of the sharded bench path (NCCL all-gather of the frame tokens; points: batch 1 and 8 per GPU)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
   bench.py --gpus 2 --steps 3 --warmup 3 --points 1,8 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'verified', (j['verified'] or {}).get('ok'))
for p in j['config']['points']: print(' batch/gpu', p['batch_per_gpu'], 'global', p['global_batch'], round(p['value'],1), 'frames/s', round(p['decode_ms_per_frame'],3), 'ms/frame')
PY
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('training point:', j['config'].get('training'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
   tools/train_bench.py --seq 4096 --batch 1 --steps 5 --warmup 2 2>&1 | tail -1 | tee gpurun_out/train_bench_n2.json
