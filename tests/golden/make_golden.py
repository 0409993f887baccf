#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs oracle/_ref/,
which `make -C oracle ref` compiles from /root/reference).  The JSON files it
writes are committed; nothing at test time needs the reference tree.

  ref_cpurun_*.json : the reference's CPU_RUN program run on an init.dat
                      (plaquette/action table, "% 16.13e" columns of its result file)
  ref_prng_*.json   : instance-0 streams of the reference's PRNG host replicas
                      (RL_produce_one_CPU, PM_produce_one_CPU, XOR7_ / RANECU_ / RANMAR_produce_one_CPU) as raw float bits
"""
import glob
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")

CPURUN_FIXTURES = {
    # deterministic INTS=2 start: no PRNG involved, pins indexing + staple + heat-bath arithmetic
    "su3_4x4x4x4_gid": dict(GROUP=3, L1=4, L2=4, L3=4, L4=4, ITER=3, NITER=1, NHIT=10, BETA=6.0, NAV=0, INTS=2),
    "su3_4x6x2x8_gid": dict(GROUP=3, L1=4, L2=6, L3=2, L4=8, ITER=3, NITER=1, NHIT=10, BETA=5.7, NAV=0, INTS=2),
    "su2_4x4x4x4_gid": dict(GROUP=2, L1=4, L2=4, L3=4, L4=4, ITER=3, NITER=1, NHIT=10, BETA=2.5, NAV=0, INTS=2),
    "su2_6x4x8x2_gid": dict(GROUP=2, L1=6, L2=4, L3=8, L4=2, ITER=4, NITER=2, NHIT=5, BETA=2.3, NAV=1, INTS=2),
    # stochastic runs: ensemble references (the CPU_RUN program draws PRNs in a different order
    # than the OpenCL path, so only averages are comparable)
    "su3_8x8x8x8_cold_ens": dict(GROUP=3, L1=8, L2=8, L3=8, L4=8, ITER=41, NITER=1, NHIT=10, BETA=6.0, NAV=40, INTS=1,
                                 PRNG="RANLUX3", RANDSERIES=54),
    "su2_8x8x8x8_cold_ens": dict(GROUP=2, L1=8, L2=8, L3=8, L4=8, ITER=41, NITER=1, NHIT=10, BETA=2.5, NAV=40, INTS=1,
                                 PRNG="RANLUX3", RANDSERIES=54),
}

PRNG_FIXTURES = [("RANLUX0", 54, 200), ("RANLUX1", 54, 200), ("RANLUX2", 54, 400), ("RANLUX3", 54, 1000),
                 ("RANLUX4", 54, 1000), ("RANLUX3", 1420193282, 400), ("PM", 54, 400), ("PM", 7, 100),
                 ("XOR7", 54, 400), ("XOR7", 7, 100), ("RANECU", 54, 400), ("RANECU", 7, 100), ("RANMAR", 54, 400), ("RANMAR", 7, 300)]


def run_cpurun(params):
    with tempfile.TemporaryDirectory() as d:
        lines = ["PLATFORM = 0", "DEVICE= 0", "FINISHPATH=./", "OUTPUTPATH=./", "ND = 4"]
        for k, v in params.items():
            lines.append(f"{k} = {v:.13e}" if isinstance(v, float) else f"{k} = {v}")
        lines.append("GETRETRACE = 1")
        with open(os.path.join(d, "init.dat"), "w") as f:
            f.write("\n".join("\t" + l for l in lines) + "\n")
        subprocess.run([os.path.join(REF, "QCDGPU_cpu"), "init.dat"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL)
        out = glob.glob(os.path.join(d, "su*-*.txt"))
        assert len(out) == 1, out
        rows, cols, in_table = [], None, False
        for line in open(out[0]):
            if line.strip().startswith("#,"):
                cols = line.replace("#,", "").split()
                continue
            if cols is not None and line.strip() and line.strip()[0].isdigit():
                rows.append([float(x) for x in line.split()[1:]])
        return {"params": params, "columns": cols, "rows": rows}


def main():
    if not os.path.exists(os.path.join(REF, "QCDGPU_cpu")):
        sys.exit("oracle/_ref missing: run `make -C oracle ref` in the container that has /root/reference")
    for name, params in CPURUN_FIXTURES.items():
        res = run_cpurun(params)
        with open(os.path.join(HERE, f"ref_cpurun_{name}.json"), "w") as f:
            json.dump(res, f, indent=1)
        print(name, len(res["rows"]), "rows")
    for gen, rs, n in PRNG_FIXTURES:
        out = subprocess.run([os.path.join(REF, "prng_dump"), gen, str(rs), str(n)], check=True,
                             capture_output=True, text=True).stdout.split()
        with open(os.path.join(HERE, f"ref_prng_{gen}_{rs}.json"), "w") as f:
            json.dump({"generator": gen, "randseries": rs, "bits": out}, f)
        print(gen, rs, len(out))


if __name__ == "__main__":
    main()
