"""Build libqcdgpu_b200.so (sm_100a) in-tree with nvcc.

The kernels are split into instantiation units so that the 8 host cores compile them in
parallel: update_inst.cu x (SUN, precision, direction) = 16 units, measure_inst.cu x
(SUN, precision) = 4 units, misc_inst.cu, dispatch.cu, backend.cu.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libqcdgpu_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
# experiment knobs: QCD_NVCC_EXTRA="-DQCD_BUF_MINBLOCKS=4 ..." (part of the build digest)
NVCC_FLAGS += os.environ.get("QCD_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: qcdgpu_b200 has no prebuilt or CPU fallback")


def _units():
    u = []
    for sun in (2, 3):
        for prec in (0, 1):
            for mu in range(4):
                u.append((f"update_{sun}_{prec}_{mu}.o", "update_inst.cu",
                          [f"-DINST_SUN={sun}", f"-DINST_PREC={prec}", f"-DINST_MU={mu}"]))
            u.append((f"measure_{sun}_{prec}.o", "measure_inst.cu", [f"-DINST_SUN={sun}", f"-DINST_PREC={prec}"]))
    u.append(("misc.o", "misc_inst.cu", []))
    u.append(("on_model.o", "on_inst.cu", []))
    u.append(("dispatch.o", "dispatch.cu", []))
    u.append(("backend.o", "backend.cu", []))
    # heaviest first
    u.sort(key=lambda t: (0 if "update_3_1" in t[0] else 1 if "update_3" in t[0] else 2))
    return u


def _sources_digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode()); h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libqcdgpu_b200.stamp")
    digest = _sources_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    units = _units()

    def compile_one(u):
        obj, src, defs = u
        cmd = [nvcc, *NVCC_FLAGS, *defs, "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        if verbose:
            cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {obj}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    jobs = jobs or max(1, (os.cpu_count() or 2))
    logs = {}
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for obj, err in ex.map(compile_one, units):
            logs[obj] = err
    objs = [os.path.join(OBJ, u[0]) for u in units]
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            for k, v in logs.items():
                f.write(f"==== {k}\n{v}\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
