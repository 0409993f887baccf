#!/usr/bin/env python3
"""Generate the O(N)-branch fixtures tests/golden/ref_on_check_*.json and ref_prng_XOR128_*.json
from the UNMODIFIED reference (oracle/_ref/on_check = O(N)/suncl/suncpu.cpp + O(N)/random/random.cpp
behind oracle/ref_harness/on_check.cpp).  TEST INFRASTRUCTURE; runs only where /root/reference exists.

  ref_on_check_*.json   : a seeded site table (raw little-endian hex) and what the reference's CPU
                          re-measurement SU::lattice_check_cpu returns for it: per-site means of the action,
                          the physical field, its square, the variance and the two correlators
  ref_prng_XOR128_*.json: the reference's XOR128 host stream (XOR128_produce_one_CPU) as float bits
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

FIXTURES = {
    # name: dims, precision, model parameters (None = the defaults of O(N)/QCDGPU.cpp:119-126), correlator step, seed
    "4x4x4x4_f64": dict(dims=(4, 4, 4, 4), precision="f64", params={}, step=(1, 1, 1, 1), seed=1),
    "4x6x2x8_f64": dict(dims=(4, 6, 2, 8), precision="f64", params={}, step=(1, 2, 1, 3), seed=2),
    "6x4x4x2_f32": dict(dims=(6, 4, 4, 2), precision="f32", params={}, step=(4, 1, 2, 1), seed=3),
    "4x4x6x4_f64_eta0": dict(dims=(4, 4, 6, 4), precision="f64", params=dict(eta=0.0, b=4.0, z=0.2, lam=0.5, zeta=0.5, skgroup=2),
                             step=(2, 2, 2, 2), seed=4),
}
XOR128_FIXTURES = [(54, 400), (7, 100)]


def main():
    if not os.path.exists(O.ON_CHECK_BIN):
        sys.exit("oracle/_ref/on_check missing: run `make -C oracle ref` in the container that has /root/reference")
    for name, fx in FIXTURES.items():
        L = O.OnLattice(fx["dims"], precision=fx["precision"], **fx["params"])
        rng = np.random.default_rng(fx["seed"])
        L.table[:L.sites] = (rng.uniform(0.01, 0.95, L.sites) * float(L.params[22])).astype(L.dtype)
        ref = L.reference_check(fx["step"])
        out = {"dims": fx["dims"], "precision": fx["precision"], "params": fx["params"], "step": fx["step"],
               "table_hex": L.table[:L.sites].tobytes().hex(),
               "columns": ["S", "phi", "phi2", "variance", "corr_step", "corr_diag"], "reference": ref}
        with open(os.path.join(HERE, f"ref_on_check_{name}.json"), "w") as f:
            json.dump(out, f)
        print(name, ref)
    for rs, n in XOR128_FIXTURES:
        bits = subprocess.run([O.ON_CHECK_BIN, "-", *["1"] * 15, "XOR128", str(rs), str(n)], check=True,
                              capture_output=True, text=True).stdout.split()
        with open(os.path.join(HERE, f"ref_prng_XOR128_{rs}.json"), "w") as f:
            json.dump({"generator": "XOR128", "randseries": rs, "bits": bits}, f)
        print("XOR128", rs, len(bits))


if __name__ == "__main__":
    main()
