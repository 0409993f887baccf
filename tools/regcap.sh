# the TMA-staged kernel compiled with a lower register cap (still three blocks per SM): physics only / fused, us per pass
for L in qcdgpu_b200/lib/libqcdgpu_b200.so "$@"; do for skip in 1 0; do
  echo -n "$(basename $L) QCDGPU_DEBUG_SKIP=$skip: "
  QCDGPU_RLX_JUMP=0 QCDGPU_LIB=$L QCDGPU_DEBUG_SKIP=$skip python bench.py --steps 4 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['avg_launch_us'], d['details']['update_kernel_registers'], d['ms_per_step'])"
done; done
