# the persistent symmetric kernel (QCDGPU_SYM=1) against the two-phase kernel: parity subset, then sustained sweeps
QCDGPU_SYM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tma_staged or mixed_precision or two_phase_kernel_overflow" 2>&1 | tail -3
for i in 1 2; do
echo -n "two-phase      : "; timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['details']['plaquette_after_run'])"
echo -n "sym stepped    : "; QCDGPU_SYM=1 timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['details']['plaquette_after_run'])"
echo -n "sym jump       : "; QCDGPU_SYM=1 QCDGPU_SYM_JUMP=1 timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'], d['details']['plaquette_after_run'])"
done
