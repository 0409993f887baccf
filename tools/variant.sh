#!/bin/bash
# tools/variant.sh NAME "NVCC EXTRA FLAGS" [NAME2 "FLAGS2" ...]: build the library with extra nvcc flags into _variants/lib_NAME.so
# (for tools/ab.sh), then restore the default build.
set -e
cd "$(dirname "$0")/.."
mkdir -p _variants
while [ $# -ge 2 ]; do
  QCD_NVCC_EXTRA="$2" python -c "from qcdgpu_b200 import build as b; b.build(force=True)"
  cp qcdgpu_b200/lib/libqcdgpu_b200.so _variants/lib_$1.so
  echo "_variants/lib_$1.so"
  shift 2
done
python -c "from qcdgpu_b200 import build as b; b.build()"
