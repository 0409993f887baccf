# resident blocks per SM of the two-phase update kernel (3 by registers; unused shared memory lowers it) x phase
for pad in 0 60 120; do for skip in 0 1 2; do
  echo -n "smem pad ${pad} KB, QCDGPU_DEBUG_SKIP=$skip (1: physics only, 2: generator only): us per pass "
  QCDGPU_SMEM_PAD_KB=$pad QCDGPU_DEBUG_SKIP=$skip python bench.py --steps 3 --warmup 1 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['avg_launch_us'])"
done; done
