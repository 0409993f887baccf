#!/bin/bash
# tools/halo_ab.sh N [steps]: bench.py on N GPUs with the NCCL exchange and with the exchange fused into the update kernels, alternating
N=${1:-2}; S=${2:-10}
for i in $(seq ${3:-2}); do for m in 0 1; do
  echo -n "N=$N QCDGPU_HALO_FUSED=$m: "
  QCDGPU_HALO_FUSED=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + 10 * i + m)) bench.py --gpus $N --steps $S --warmup 3 2> gpurun_out/halo_ab.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.2f ms per sweep, %.3f G link updates/s, e2e %.3f G, launches %d, %s' % (d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['gpu_launches'], d['details']['halo'][:40]))"
done; done
