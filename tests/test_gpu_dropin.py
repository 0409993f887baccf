"""Drop-in check: the UNMODIFIED reference host program (QCDGPU.cpp main, suncl.cpp run loop, random.cpp,
su{2,3}cpu.cpp LatticeCheckCPU, data_analysis.cpp) linked against integration/clinterface_cuda.cpp +
libqcdgpu_b200.so (integration/Makefile; built where the reference tree exists, the binary travels).
The reference verifies itself: it re-measures the last configuration on the CPU and compares with the
device values at 1e-11 (double) / 5e-5 (single) -- data_analysis.cpp:40-101."""
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_build", "QCDGPU_b200")

INIT = """\tPLATFORM = 0
\tDEVICE= 0
\tFINISHPATH=./
\tOUTPUTPATH=./
\tGROUP = {group}
\tND = 4
\tLS = {ls}
\tLT = {lt}
\tITER = {iter}
\tNITER = 1
\tNHIT = 10
\tBETA = {beta:.13e}
\tNAV = {nav}
\tINTS = {ints}
\tPRNG = {prng}
\tRANDSERIES = 54
\tPL_LEVEL = {pl}
\tPHI = {phi:.13e}
{extra}"""


def run_reference_host(tmp_path, **kw):
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/QCDGPU_b200 not built (needs the reference tree: make -C integration)")
    cfg = dict(group=3, ls=8, lt=8, iter=4, beta=6.0, nav=2, ints=1, prng="RANLUX3", phi=0.0, pl=2, extra="")
    cfg.update(kw)
    cfg["extra"] = "".join(f"\t{line}\n" for line in cfg["extra"].split(";") if line)
    (tmp_path / "init.dat").write_text(INIT.format(**cfg))
    r = subprocess.run([BIN, "init.dat"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = glob.glob(str(tmp_path / "su*-*.txt"))
    assert len(out) == 1, (out, r.stdout[-2000:])
    return open(out[0]).read(), r.stdout


def last_values(text):
    vals = {}
    for kind, name, v in re.findall(r"^ (CPU|GPU) last (\S+)\s*:\s*(\S+)", text, flags=re.M):
        vals.setdefault(name, {})[kind] = float(v)
    return vals


@pytest.mark.parametrize("group,ints,beta", [(3, 1, 6.0), (3, 0, 5.7), (2, 0, 2.5)])
def test_reference_host_runs_on_cuda_backend_and_verifies_itself(tmp_path, group, ints, beta):
    text, stdout = run_reference_host(tmp_path, group=group, ints=ints, beta=beta)
    assert "Verification successfully passed" in text, text[-3000:]
    vals = last_values(text)
    for name in ("S_total", "Plq_total", "Polyakov_loop"):
        assert name in vals, (name, sorted(vals))
        c, g = vals[name]["CPU"], vals[name]["GPU"]
        assert abs(c) > 1e-3
        assert abs(1.0 - g / c) < 1e-11, (name, c, g)
    # the per-iteration table is there and physical
    rows = [l.split() for l in text.splitlines() if re.match(r"^\s+\d+\s+[-\d]", l)]
    assert len(rows) == 4
    plq_total = [float(r[3]) for r in rows]
    if ints == 1:
        assert plq_total[0] == float(group)


@pytest.mark.parametrize("prng", ["XOR7", "RANECU", "RANMAR", "PM", "XOR128"])
def test_reference_host_with_the_other_generators(tmp_path, prng):
    """init.dat's PRNG key (random.cpp:146-153): the reference's host builds the seed tables and asks for the kernels by its own names
    ("xor7", "ranecu", "rmseed" / "rmproduce" -- the latter two do not even exist in its .cl file, which defines ranmar_init / ranmar)."""
    text, stdout = run_reference_host(tmp_path, group=3, ints=0, beta=6.0, prng=prng, ls=6, lt=4)
    assert "Verification successfully passed" in text, text[-3000:]
    if prng != "PM":        # print_generator has no line for Park-Miller (random.cpp:160-169)
        assert re.search(r"PRN generator\s*:\s*" + prng, text), text[:3000]
    rows = [l.split() for l in text.splitlines() if re.match(r"^\s+\d+\s+[-\d]", l)]
    plq = [float(r[3]) for r in rows]
    assert len(plq) == 4 and all(-0.5 < p < 3.0 for p in plq) and len(set(plq)) == 4 and plq[-1] > plq[0], plq


def test_reference_host_twisted_run_verifies(tmp_path):
    """PHI != 0 (init.dat sample value 3.0): twisted loads in staples, measurements and the CPU check."""
    text, _ = run_reference_host(tmp_path, group=3, ints=0, beta=6.0, phi=3.0, ls=6, lt=4)
    assert "Verification successfully passed" in text, text[-3000:]


@pytest.mark.parametrize("group,beta,extra", [
    (2, 3.5, "GETWILSON = 1;WILSONR = 3;WILSONT = 2;FMUNU = 1"),          # the reference's sample init.dat, on 8^4
    (3, 6.0, "GETWILSON = 1;WILSONR = 2;WILSONT = 3;FMUNU = 1"),
    (3, 6.0, "F0MU = 1;FMUNU1 = 1;FMUNU5 = 1"),
])
def test_reference_sample_configuration_wilson_fmunu_polyakov_slices(tmp_path, group, beta, extra):
    """SU(N)/init.dat enables GETWILSON, PL_LEVEL = 3, FMUNU and the twist PHI = 3: every kernel that configuration
    initialises runs, and the reference's LatticeCheckCPU (Wilson loop su3cpu.cpp:1435-1520, Fmunu 1293-1345) agrees."""
    text, stdout = run_reference_host(tmp_path, group=group, ints=0, beta=beta, phi=3.0, pl=3, extra=extra)
    assert "Verification successfully passed" in text, text[-4000:]
    vals = last_values(text)
    names = ["S_total", "Plq_total", "Polyakov_loop"] + (["Wilson_loop"] if "GETWILSON" in extra else [])
    for name in names:
        assert name in vals, (name, sorted(vals))
        c, g = vals[name]["CPU"], vals[name]["GPU"]
        assert abs(c) > 1e-6 and abs(1.0 - g / c) < 1e-11, (name, c, g)
    fm = [k for k in vals if k.startswith("Fmunu_") and "abs" not in k]
    assert len(fm) >= 12, sorted(vals)
    for k in fm:
        c, g = vals[k].get("CPU"), vals[k].get("GPU")
        if c is not None and g is not None and abs(c) > 1e-9:
            assert abs(1.0 - g / c) < 1e-9, (k, c, g)
    # slice-resolved Polyakov loop table is written and adds up to the total
    assert "Differentiated Polyakov loop data" in text
    row = [l for l in text.splitlines() if l.startswith("PL             ")][0].split()[1:]
    assert len(row) == 8


BIN_BIGLAT = os.path.join(ROOT, "integration", "_build", "QCDGPU_b200_biglat")
INIT_BIGLAT = """\tPLATFORM = {{0, 0}}
\tDEVICE= {{0, 0}}
\tFINISHPATH=./
\tOUTPUTPATH=./
\tGROUP = {group}
\tND = 4
\tLT = {lt}
\tLS = {ls}
\tL1 = {l1}
\tN1 = {{{n1a}, {n1b}}}
\tITER = {iter}
\tNITER = 1
\tNHIT = 10
\tBETA = {beta:.13e}
\tNAV = {nav}
\tINTS = {ints}
\tPRNG = RANLUX3
\tRANDSERIES = {{54, 1054}}
\tPL_LEVEL = {pl}
"""


def _table_rows(text):
    return np.array([[float(v) for v in l.split()[1:]] for l in text.splitlines() if re.match(r"^\s+\d+\s+[-\d]", l)])


@pytest.mark.parametrize("group,beta,n1,pl", [(3, 6.0, (8, 4), 0), (2, 2.5, (6, 6), 0), (3, 5.7, (2, 10), 0), (3, 6.0, (4, 8), 2), (2, 2.5, (8, 4), 1)])
def test_reference_biglat_host_runs_on_cuda_backend(tmp_path, group, beta, n1, pl):
    """The reference's BIGLAT build (Makefile:45-48; -D BIGLAT -D NPARTS=2 -D NDEV=2, both parts on device 0 like its init_2.dat): its own
    partition (SubLattice, N1 = {a, b}), per-part programs (-D BIGLAT -D N1 -D LEFT_SITES -D FULL_SITES), per-part kernel launches and
    model::UpdateEdges -- halo slices copied through mapped ranges of the link tables (buffer_map(id, offset, size) / buffer_unmap) --
    over this library.  With the deterministic INTS = 2 start the table must equal the single-lattice build's, which verifies itself
    against the reference's CPU re-measurement."""
    if not os.path.exists(BIN_BIGLAT):
        pytest.skip("integration/_build/QCDGPU_b200_biglat not built (needs the reference tree: make -C integration)")
    cfg = dict(group=group, ls=6, lt=4, l1=sum(n1), n1a=n1[0], n1b=n1[1], iter=4, beta=beta, nav=0, ints=2, pl=pl)
    d = tmp_path / "big"
    d.mkdir()
    (d / "init.dat").write_text(INIT_BIGLAT.format(**cfg))
    r = subprocess.run([BIN_BIGLAT, "init.dat"], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = glob.glob(str(d / "su*-*.txt"))
    assert len(out) == 1, (out, r.stdout[-2000:])
    big = _table_rows(open(out[0]).read())
    s = tmp_path / "single"
    s.mkdir()
    single_cfg = dict(group=group, ls=6, lt=4, iter=4, beta=beta, nav=0, ints=2, prng="RANLUX3", phi=0.0, pl=pl, extra="\tL1 = %d\n" % sum(n1))
    (s / "init.dat").write_text(INIT.format(**single_cfg))
    r1 = subprocess.run([BIN, "init.dat"], cwd=s, capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stdout[-3000:] + r1.stderr[-3000:]
    text = open(glob.glob(str(s / "su*-*.txt"))[0]).read()
    assert "Verification successfully passed" in text
    one = _table_rows(text)
    # the Python driver reads the same BIGLAT file (python -m qcdgpu_b200.biglattice init.dat): same partition, same table
    m = tmp_path / "mirror"
    m.mkdir()
    (m / "init.dat").write_text(INIT_BIGLAT.format(**cfg))
    rm = subprocess.run([sys.executable, "-m", "qcdgpu_b200.biglattice", "init.dat"], cwd=m, capture_output=True, text=True, timeout=600,
                        env=dict(os.environ, PYTHONPATH=ROOT))
    assert rm.returncode == 0, rm.stdout[-3000:] + rm.stderr[-3000:]
    assert f"[{n1[0]}, {n1[1]}]" in rm.stdout
    mir = _table_rows(open(glob.glob(str(m / "su*-*.txt"))[0]).read())
    assert mir.shape == big.shape, (mir.shape, big.shape)
    np.testing.assert_allclose(mir, big, rtol=1e-11, atol=1e-15)
    assert big.shape == one.shape and big.shape[0] == 4 and big.shape[1] >= (6 if pl == 0 else 8), (big.shape, one.shape)
    np.testing.assert_allclose(big, one, rtol=1e-11, atol=1e-15)          # pl > 0: the Polyakov-loop columns too
