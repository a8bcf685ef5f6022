#!/bin/bash
# N-GPU bench as the driver launches it (N = $1): generation (batch 1 per GPU headline, 8 per GPU point = BASELINE
# config #4 at N = 8) and the data-parallel training point (= BASELINE config #5 at N = 8)
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
   bench.py --gpus $N --steps 3 --warmup 3 --points 1,8 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n=sys.argv[1]
j=json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
print('N', n, 'value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'verified', (j['verified'] or {}).get('ok'))
for p in j['config']['points']: print(' batch/gpu', p['batch_per_gpu'], 'global', p['global_batch'], round(p['value'],1), 'frames/s', round(p['decode_ms_per_frame'],3), 'ms/frame')
print('training point:', j['config'].get('training'))
PY
