"""CPU-side checks: the C ABI library loads and exports every symbol include/qcdgpu_b200.h declares,
host logic (init.dat, option strings, sizes) behaves like the reference's, and the product fails loudly
without a CUDA device (no CPU fallback, no oracle on the product path)."""
import ast
import ctypes as C
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
        g.kernel_init("lattice_measurement_WLx0", 1, [64], None)      # BIGLAT Wilson-loop family
    with pytest.raises(QcdGpuError, match="RL_skip=100"):
        g.program_create(None, "-D RL_skip=100")
        g.kernel_init("ranlux", 1, [64], None)
    # GPU::buffer_size_align (clinterface.cpp:1587-1590)
    assert g.buffer_size_align(65, 64) == 128 and g.buffer_size_align(64, 64) == 64 and g.buffer_size_align(1, 64) == 64


def test_biglat_programs_follow_the_reference_n1_convention():
    """The reference's BIGLAT host gives its update program N1 = Nx + 2 (halo slices included, suncl.cpp:2422) and its measuring
    programs N1 = Nx (suncl.cpp:2597, 2889), both with -D BIGLAT; the update program is the one with -D NHIT.  Seen from outside
    through the range check of qgpu_kernel_set_peer_halo (the source slice must be a real one: 1 <= x < stored N1 - 1) -- no device
    is needed -- and through the kernel names of the rest of init.dat's PRNG key."""
    from qcdgpu_b200 import GPU, QcdGpuError
    g = GPU()
    common = "-D SUN=3 -D ND=4 -D N2=4 -D N3=4 -D N4=4 -D ROWSIZE=512 -D PRECISION=1 -D BIGLAT -D FULL_SITES=1024 -D LEFT_SITES=0"
    g.program_create(None, common + " -D N1=6 -D NHIT=10 -D NHITPar=1 -D PRNGSTEP=256")          # update program: 4 real slices
    k = g.kernel_init("update_odd_X", 1, [128], None)
    g.kernel_set_peer_halo(k, 1, 0x1000, 4, 0, 512)
    with pytest.raises(QcdGpuError, match="real source slice"):
        g.kernel_set_peer_halo(k, 1, 0x1000, 5, 0, 512)
    g.program_create(None, common + " -D N1=6")                                                 # measuring program: 6 real slices
    k2 = g.kernel_init("overrelax_odd_X", 1, [128], None)
    g.kernel_set_peer_halo(k2, 1, 0x1000, 6, 0, 512)
    with pytest.raises(QcdGpuError, match="real source slice"):
        g.kernel_set_peer_halo(k2, 1, 0x1000, 7, 0, 512)
    with pytest.raises(QcdGpuError, match="no peer halo"):
        g.kernel_set_peer_halo(g.kernel_init("lattice_measurement", 1, [128], [64]), 0, 0x1000, 1, 0, 512)
    g.program_create(None, "")
    for name in ("xor7", "ranecu", "ranmar", "rmproduce", "ranmar_init", "rmseed", "pm", "xor128", "mrg32k3a"):
        assert g.kernel_init(name, 1, [64], None) > 0
    g.close()


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


# ----------------------------------------------------------------------------- RANLUX skip-ahead (csrc/ranlux_jump*.{cuh,h})
def _load_jump_generator():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_ranlux_jump", os.path.join(ROOT, "tools", "gen_ranlux_jump.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ranlux_jump_header_is_what_the_checked_generator_prints():
    """csrc/ranlux_jump.cuh is generated: tools/gen_ranlux_jump.py holds the algorithm as a list of 32-bit operations with PTX carry-flag
    semantics, runs that list against the plain subtract-with-borrow recurrence (random.cpp:504-574) on random and degenerate states,
    and prints the same list as CUDA.  Re-run the model here and require the committed header to be its output."""
    gen = _load_jump_generator()
    prog = gen.build()
    assert gen.self_test(prog, trials=120, seed=11) == 120
    with open(os.path.join(ROOT, "qcdgpu_b200", "csrc", "ranlux_jump.cuh")) as f:
        assert f.read() == gen.to_cuda(prog)


def _replay_pass(left, samples, keep, skip, from_start):
    """RanluxPk::pass (prng.cuh) step by step: subtract-with-borrow steps behind the parked samples, `left` after the pass"""
    k, steps, told = 0, 0, from_start
    while k < samples:
        if left == 0:
            steps += skip if told else 0
            left = 24
        j = min(left >> 2, samples - k)
        steps += 4 * j if told else 0
        k += j
        left -= 4 * j
        if not told and k >= keep:
            told = True
    return steps, left


@pytest.mark.parametrize("samples,keep,from_start", [(33, 12, 0), (33, 8, 0), (12, 12, 0), (21, 0, 0), (33, 0, 1), (6, 0, 1), (33 * 8 * 1000, 0, 1),
                                                     (2**40 + 5, 0, 1)])
def test_ranlux_jump_table_of_the_library(samples, keep, from_start):
    """qgpu_ranlux_jump_table (host code of the library, no device): multipliers against python's pow(2^24, -(steps - 48), m) and the
    pass plan against a step-by-step replay of the kernel's loop."""
    from qcdgpu_b200 import capi
    L = capi.load()
    out = np.zeros(7 * 20, dtype=np.uint32)
    L.qgpu_ranlux_jump_table(samples, keep, 223, from_start, out.ctypes.data_as(C.c_void_p))
    b, m = 1 << 24, (1 << 576) - (1 << 240) + 1
    for e in range(7):
        if samples < 10**7:
            steps, left_after = _replay_pass(4 * e, samples, keep, 223, bool(from_start))
        else:       # closed form from a sample boundary: whole periods of 6 samples = 24 steps + a 223-step skip in front of each
            assert from_start
            first = min(e, samples)                  # samples before the first skip
            rest = samples - first
            periods, tail = divmod(rest, 6)
            steps = 4 * first + periods * (223 + 24) + ((223 + 4 * tail) if tail else 0)
            left_after = (4 * e - 4 * first) if rest == 0 else ((24 - 4 * tail) if tail else 0)
        entry = out[20 * e:20 * e + 20]
        if steps < 48:
            assert entry[19] == 0
            continue
        assert entry[19] == 1 and entry[18] == left_after
        got = sum(int(w) << (32 * i) for i, w in enumerate(entry[:18]))
        assert got == pow(b, -(steps - 48), m), (e, steps)
