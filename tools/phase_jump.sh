for jump in 1 0; do for skip in 0 1 2; do
  echo -n "JUMP=$jump QCDGPU_DEBUG_SKIP=$skip (1: physics only, 2: generator only): us per pass "
  QCDGPU_RLX_JUMP=$jump QCDGPU_DEBUG_SKIP=$skip python bench.py --steps 3 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['avg_launch_us'], d['roofline'].get('one_by_one_us'), d['ms_per_step'])"
done; done
