#!/bin/bash
# TMA ring depth (links in flight per warp) x parked samples: tools/ring.sh  -- needs _variants/lib_k4.so (tools/variant.sh k4 "-DQCD_TMA_K=4").
# Three blocks per SM need (parked + ring + 1) KB x 3 <= 227 KB: K = 3 fits with 12 parked samples, K = 4 with 10 or fewer.
line() { python bench.py --steps ${S:-6} --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms per sweep sustained (sm %s MHz, %s W), %.1f us per pass profiled' % (d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d['roofline']['avg_launch_us']))"; }
QCDGPU_LIB=_variants/lib_k4.so QCDGPU_RLX_KEEP=10 python -m pytest tests/test_gpu_parity.py -x -q -k "tma_staged or fused_prng_equals" 2>&1 | tail -1
for i in 1 2; do
  for cfg in "qcdgpu_b200/lib/libqcdgpu_b200.so 12" "qcdgpu_b200/lib/libqcdgpu_b200.so 10" "_variants/lib_k4.so 10" "_variants/lib_k4.so 8" "_variants/lib_k4.so 12"; do
    set -- $cfg
    echo -n "$(basename $1) keep=$2: "
    QCDGPU_LIB=$1 QCDGPU_RLX_KEEP=$2 line
  done
done
for skip in 1 2; do for cfg in "qcdgpu_b200/lib/libqcdgpu_b200.so 12" "_variants/lib_k4.so 10"; do
  set -- $cfg
  echo -n "$(basename $1) keep=$2 QCDGPU_DEBUG_SKIP=$skip: "
  QCDGPU_LIB=$1 QCDGPU_RLX_KEEP=$2 QCDGPU_DEBUG_SKIP=$skip S=3 line
done; done
