# the persistent symmetric kernel (-DQCD_SYM_EXPERIMENT library, QCDGPU_SYM=1): fused / physics only / generator only, ms per sweep
L=${1:-_variants/lib_sym.so}
for skip in 0 1 2; do
echo -n "two-phase SKIP=$skip: "; QCDGPU_DEBUG_SKIP=$skip timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"
echo -n "sym       SKIP=$skip: "; QCDGPU_LIB=$L QCDGPU_SYM=1 QCDGPU_DEBUG_SKIP=$skip timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['sm_mhz'])"
done
