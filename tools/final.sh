#!/bin/bash
# tools/final.sh TAG: what the end of a session leaves under gpurun_out/ for profiles/ -- the GPU test suite, the default bench line, the
# reference arm, and (after those exited without ncu) the launch list of a short bench run.
T=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${T}_pytest_gpu.txt
python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_64x4.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_pytest_gpu.txt; cut -c1-600 gpurun_out/${T}_bench_1gpu.json; cut -c1-400 gpurun_out/${T}_bench_reference_arm.json
