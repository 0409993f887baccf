# more resident warps by smaller register caps x other block sizes (libraries built with -DQCD_UPDATE_THREADS / -DQCD_RLX_MAXNREG), 10 parked samples
echo -n "shipped (128 threads, 168 regs, 12 warps): "; python bench.py --steps 6 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"
for L in "$@"; do
  QCDGPU_RLX_KEEP=10 QCDGPU_LIB=$L timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "tma_staged" 2>&1 | tail -1
  for j in 1 0; do
  echo -n "$(basename $L) jump=$j: "; QCDGPU_RLX_JUMP=$j QCDGPU_RLX_KEEP=10 QCDGPU_LIB=$L timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"
  done
  QCDGPU_RLX_KEEP=10 QCDGPU_LIB=$L ncu --metrics launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__block_size --clock-control none -k regex:update_kernel_rlx -s 16 -c 1 python bench.py --steps 1 --warmup 2 --no-cpu 2>&1 | grep -E "registers|occupancy|warps_active|duration|block_size" | awk '{printf "%s=%s ", $1, $NF} END{print ""}'
done
echo -n "shipped again: "; python bench.py --steps 6 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"
