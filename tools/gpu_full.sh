#!/bin/bash
# full GPU test suite + the default bench line (all three batch points, verification, CPU baseline)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
BENCH_VERBOSE=1 timeout 1500 python bench.py ${BENCH_ARGS:---steps 5 --warmup 3} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'verified', j['verified'])
for p in j['config']['points']: print(p['batch_per_gpu'], round(p['value'],1), 'ms/frame', round(p['decode_ms_per_frame'],3), 'frac', round(p['roofline']['frac'],3), 'prefill ms', round(p['prefill']['ms_incl_first_frame'],2), 'tf frac', round(p['prefill']['frac'],3))
print(j.get('cpu_baseline'))
PY
