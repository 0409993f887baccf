# the two-kernel pass (QCDGPU_SPLIT=1) against the fused kernel: parity subset, sustained sweeps, per-kernel times under ncu
QCDGPU_SPLIT=1 python -m pytest tests/test_gpu_parity.py -x -q -k "tma_staged or mixed_precision" 2>&1 | tail -1
for i in 1 2; do
echo -n "fused  : "; python bench.py --steps 8 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"
echo -n "split  : "; QCDGPU_SPLIT=1 python bench.py --steps 8 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"
for L in "$@"; do echo -n "split $L: "; QCDGPU_LIB=$L QCDGPU_SPLIT=1 python bench.py --steps 8 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"; done
done
QCDGPU_SPLIT=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:"staple_kernel|heatbath_kernel" -s 16 -c 2 python bench.py --steps 1 --warmup 2 --no-cpu 2>&1 | grep -E "void|duration|dram|issue|warps_active|registers" | cut -c1-120
