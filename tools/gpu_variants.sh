# time library variants built with different -D knobs (csm_hf_b200/build.py build(defines=..., out=...))
for f in "" gpurun_variants/lib_*.so; do
  echo "=== ${f:-default}"; CSM_LIB=${f:+$PWD/$f} python tools/ncu_target.py --batch ${B:-1} --frames 100 --reps 2 2>&1 | tail -1
done
