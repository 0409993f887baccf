"""CPU-side checks: the C ABI library loads and exports every symbol include/qcdgpu_b200.h declares,
host logic (init.dat, option strings, sizes) behaves like the reference's, and the product fails loudly
without a CUDA device (no CPU fallback, no oracle on the product path)."""
import ast
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from qcdgpu_b200 import capi
    L = capi.load()
    header = open(os.path.join(ROOT, "include", "qcdgpu_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(qgpu_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 45
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/qcdgpu_b200.h but not exported"
    assert set(capi.EXPORTS) == set(declared)
    assert b"sm_100a" in L.qgpu_version()


def test_product_never_touches_the_oracle():
    """Nothing under qcdgpu_b200/, include/ may import, link or execute oracle/."""
    for base in ("qcdgpu_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep) or "lib" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "oracle_lib" not in txt and "liboracle" not in txt and "qcd_oracle" not in txt, f


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qcdgpu_b200 import Model, QcdGpuError
    m = Model()
    with pytest.raises(QcdGpuError, match="no CPU fallback"):
        m.lattice_init()


def test_program_options_are_validated_like_compile_errors():
    from qcdgpu_b200 import GPU, QcdGpuError
    g = GPU()
    ok = g.program_create(None, "-Werror -D SUN=3 -D ND=4 -D N1=4 -D N2=4 -D N3=4 -D N4=4 -D ROWSIZE=256 -D PRECISION=1 -D NHIT=10")
    assert ok == 1
    assert g.program_create(None, "-D RL_skip=223") == 2
    with pytest.raises(QcdGpuError, match="N2 must be even"):
        g.program_create(None, "-D SUN=3 -D ND=4 -D N1=4 -D N2=5 -D N3=4 -D N4=4")
    with pytest.raises(QcdGpuError, match="SUN=4"):
        g.program_create(None, "-D SUN=4 -D ND=4 -D N1=4 -D N2=4 -D N3=4 -D N4=4")
    with pytest.raises(QcdGpuError, match="not implemented"):
        g.kernel_init("lattice_action_diff_x", 1, [64], None)
    with pytest.raises(QcdGpuError, match="RL_skip=100"):
        g.program_create(None, "-D RL_skip=100")
        g.kernel_init("ranlux", 1, [64], None)
    # GPU::buffer_size_align (clinterface.cpp:1587-1590)
    assert g.buffer_size_align(65, 64) == 128 and g.buffer_size_align(64, 64) == 64 and g.buffer_size_align(1, 64) == 64


def test_init_dat_keys(tmp_path):
    """The reference's sample init.dat (SU(N)/init.dat) keys, `KEY = value # comment` syntax."""
    from qcdgpu_b200.model import Model
    p = tmp_path / "init.dat"
    p.write_text("\tPLATFORM = 0\n\tDEVICE= 0\n\tFINISHPATH=./\n\tOUTPUTPATH=/tmp/\n\tGROUP = 2\n\tND = 4\n\tLT = 16\n"
                 "\tLS = 12 # spatial\n\tL1 = 8\n\tITER = 50\n\tNITER = 10\n\tNHIT = 7\n\tBETA = 3.5000000000000e+00\n"
                 "\tNAV = 50\n\tINTS = 0\n\tPRNG = XOR128\n\tRANDSERIES = 45\n\tPL_LEVEL = 2\n\tOMEGA = 0.0\n\tPHI = 3.0e+00\n")
    m = Model()
    m.lattice_get_init_file(str(p))
    assert m.lattice_group == 2 and m.fprefix == "su2-"
    assert m.lattice_full_size == [8, 12, 12, 16]
    assert (m.ITER, m.NITER, m.NHIT, m.NAV, m.ints) == (50, 10, 7, 50, 0)
    assert m.BETA == 3.5 and m.PHI == 3.0 and m.OMEGA == 0.0
    assert m.PRNG0.PRNG_generator == "XOR128" and m.PRNG0.PRNG_randseries == 45
    assert m.path == "/tmp/" and m.PL_level == 2
    assert "-D TBC" in (m.__setattr__("lattice_domain_size", m.lattice_full_size) or m.__setattr__("lattice_table_row_size", 1) or m.options_common())


def test_host_rand_matches_glibc():
    """Seed tables are defined through glibc srand/rand (random.cpp:235, 462-465)."""
    import ctypes
    from qcdgpu_b200.capi import host_rand_fill, host_srand
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(54)
    want = [libc.rand() for _ in range(6)]
    host_srand(54)
    got = host_rand_fill(6)
    assert list(got) == want
    host_srand(54)
    assert list(host_rand_fill(3, 2147483647, 1)) == [w % 2147483647 + 1 for w in want[:3]]


def test_on_model_host_logic(tmp_path):
    """O(N) branch host mirror: parameter derivation (O(N)/suncl/suncl.cpp:1244-1253) against the checker's restatement,
    init.dat keys (:249-300) and the option string (:1454-1482).  No device call."""
    import numpy as np
    import oracle_lib as O
    from qcdgpu_b200.on_model import OnModel
    m = OnModel()
    szl, max_u = m.on_parameters()
    p = O.on_params(m.ON_z, m.ON_lambda, m.ON_zeta, m.ON_eta, m.ON_b)
    assert szl == p[21] and abs(max_u - p[22]) < 1e-15
    m.ON_eta, m.ON_b = 0.0, 4.0
    assert abs(m.on_parameters()[1] - O.on_params(m.ON_z, m.ON_lambda, m.ON_zeta, 0.0, 4.0)[22]) < 1e-15
    f = tmp_path / "init.dat"
    f.write_text("\tGROUP = 1\n\tLS = 32\n\tLT = 32\n\tNHIT = 10\n\tPRNG = XOR128\n\tRANDSERIES = 45\n\tSKGROUP = 4\n"
                 "\tON_Z = 2.0e-01\n\tON_LAMBDA = 5.0e-01 # quartic\n\tON_ZETA = 0.5\n\tON_ETA = 0.0\n\tON_B = 4.0\n\tINTS = 0\n")
    m = OnModel()
    m.lattice_get_init_file(str(f))
    assert m.lattice_full_size == [32, 32, 32, 32] and (m.ON_z, m.ON_lambda, m.ON_zeta, m.ON_eta, m.ON_b) == (0.2, 0.5, 0.5, 0.0, 4.0)
    assert m.PRNG0.PRNG_generator == "XOR128" and m.ints == 0 and m.ON_SK_group == 4
    m.lattice_table_row_size = 32 ** 4
    o = m.options_common()
    assert "-D ON_MODEL" in o and "-D SUN=1" in o and "-D SKGROUP=4" in o and "-D N1EXACT=32" in o and "-D ACC_RATE" in o
