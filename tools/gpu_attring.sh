#!/bin/bash
# Round-2 experiment: the cp.async K/V ring form of the tensor-core backbone attention (csm_stream.inl, CSM_ATT_RING) against
# the default build -- parity tests that exercise it, then decode ms/frame at 8 and 32 sequences.  Build the variants
# BEFORE the gpurun call (nvcc cross-compiles without a GPU; gpurun_variants/ travels with the snapshot):
#   python -c "from csm_hf_b200 import build; [build.build(defines=['CSM_ATT_RING=%d' % n], out='gpurun_variants/lib_attring%d.so' % n) for n in (2, 3)]"
mkdir -p gpurun_out
for f in "" gpurun_variants/lib_attring2.so gpurun_variants/lib_attring3.so; do
  [ -n "$f" ] && [ ! -f "$f" ] && continue
  echo "=== ${f:-default}"
  export CSM_LIB=${f:+$PWD/$f}
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 \
      -k "backbone_attention or batch_invariance or families or long_context or decode_frames" 2>&1 | tail -3
  for b in ${BATCHES:-8 32}; do timeout 120 python tools/ncu_target.py --batch $b --frames 60 --reps 1 2>&1 | tail -1; done
done
