#!/bin/bash
# tools/halo8.sh N: on N GPUs of one box -- (1) link tables of all ranks, halo slices included, with the NCCL exchange and with the exchange
# fused into the update kernels (left and right neighbours differ from N = 3 on): equal hashes; (2) bench lines of both; (3) the fused
# scaling lines below N.  Output under gpurun_out/r02c_*.
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for m in 0 1; do
  QCDGPU_HALO_FUSED=$m QCD_DIMS=$((4 * N))x64x4x4 QCD_GROUP=3 QCD_PREC=1 run $N $((29900 + m)) tools/halo_check.py 2> gpurun_out/halo8_check_$m.err | grep -o "HASH [0-9]* [0-9a-f]* [a-z]*" | sort > gpurun_out/halo8_hash_$m.txt
done
cut -d" " -f2,3 gpurun_out/halo8_hash_0.txt > /tmp/h0; cut -d" " -f2,3 gpurun_out/halo8_hash_1.txt > /tmp/h1
echo "ranks: $(wc -l < /tmp/h0) nccl ($(cut -d' ' -f4 gpurun_out/halo8_hash_0.txt | sort -u | tr '\n' ' ')), $(wc -l < /tmp/h1) fused ($(cut -d' ' -f4 gpurun_out/halo8_hash_1.txt | sort -u | tr '\n' ' ')); tables equal: $(cmp -s /tmp/h0 /tmp/h1 && echo yes || echo NO)"
QCDGPU_HALO_FUSED=0 run $N 29910 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/halo8_b0.err | grep "^{" > gpurun_out/r02c_scale_${N}gpu_nccl.json
QCDGPU_HALO_FUSED=1 run $N 29911 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/halo8_b1.err | grep "^{" > gpurun_out/r02c_scale_${N}gpu.json
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_scale_1gpu.json 2> gpurun_out/s1.err
for n in 2 4; do [ $n -lt $N ] && run $n $((29920 + n)) bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/s$n.err | grep "^{" > gpurun_out/r02c_scale_${n}gpu.json; done
python - <<PY
import json
def line(f):
    try: return json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: return None
base = line("gpurun_out/r02c_scale_1gpu.json")
for n, f in [(1, "1gpu"), (2, "2gpu"), (4, "4gpu"), ($N, "${N}gpu"), ($N, "${N}gpu_nccl")]:
    d = line(f"gpurun_out/r02c_scale_{f}.json")
    if d: print(f"{f}: {d['value'] / 1e9:.3f} G link updates/s, {d['ms_per_step']:.2f} ms per sweep, e2e {d['e2e']['value'] / 1e9:.3f} G, efficiency {100 * d['value'] / (n * base['value']):.1f} %")
PY
