#!/bin/bash
# A/B of update-kernel variants on ONE box: tools/ab.sh libA.so libB.so [rounds] [steps] -- alternates bench runs (boxes of the pool
# differ by a few per cent, so only numbers from the same call compare).  Prints the sustained sweep time (the bench value; long runs sit
# at the board's power cap, so it also measures energy per sweep) and the update kernel's time per launch from the profiling pass.
A=$1; B=$2; N=${3:-3}; S=${4:-8}
for i in $(seq $N); do for L in "$A" "$B"; do
  echo -n "$(basename $L): "
  QCDGPU_LIB=$L python bench.py --steps $S --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms per sweep sustained (sm %s MHz, %s W), %.1f us per pass profiled' % (d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d['roofline']['avg_launch_us']))"
done; done
