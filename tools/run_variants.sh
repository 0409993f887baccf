# usage: bash tools/run_variants.sh NAME...   (libraries _variants/lib_NAME.so: parity subset, then sustained bench lines against the default build)
for v in "$@"; do
  echo "=== $v"
  QCDGPU_LIB=_variants/lib_$v.so python -m pytest tests/test_gpu_parity.py -x -q -k "tma_staged or fused_prng_equals" 2>&1 | tail -1
done
for v in "$@"; do bash tools/ab.sh qcdgpu_b200/lib/libqcdgpu_b200.so _variants/lib_$v.so 1 6; done
