#!/bin/bash
# tools/scale.sh [steps] [warmup]: the 1 / 2 / 4 / 8-GPU weak-scaling lines of bench.py back to back on one box, launched the way the
# driver launches them; JSON lines into gpurun_out/r02b_scale_{1,2,4,8}gpu.json
S=${1:-20}; W=${2:-5}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps $S --warmup $W --no-cpu > gpurun_out/r02b_scale_1gpu.json 2> gpurun_out/s1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps $S --warmup $W 2> gpurun_out/s$n.err | grep "^{" > gpurun_out/r02b_scale_${n}gpu.json
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r02b_scale_{n}gpu.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    if n == 1: base = d["value"]
    print(f"{n} GPU: {d['value'] / 1e9:.3f} G link updates/s, {d['ms_per_step']:.2f} ms per sweep, e2e {d['e2e']['value'] / 1e9:.3f} G, "
          f"efficiency {100 * d['value'] / (n * base):.1f} %" if base else d["value"])
PY
