"""State files (.qcg) and continuing a run from one (suncl.cpp:1857-2079, 5349-5354).

The reference's writer and reader are reached through integration/_build/QCDGPU_b200_ckpt: the reference's unmodified model class
(run loop, lattice_save_state / lattice_load_state, PRNG fast-forward) over the C ABI, behind a main of ours that switches saving
on -- the stock main hard-codes turnoff_config_save = true (QCDGPU.cpp:203) and no init.dat key sets INIT = 0.  Pinned here:
  * a file written by the Python mirror equals the reference-written file of the same run byte for byte;
  * a run continued from a state file (by the reference's code and by the mirror) ends in the same file as the uninterrupted run:
    same link table, same results, same generator position (PRNG_counter produce() calls replayed -- in fused mode as one jump)."""
import glob
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_build", "QCDGPU_b200_ckpt")

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
"""
ITER_WORD, INIT_WORD = 13, 4        # dwords of the header (suncl.cpp:1950-1959)


def reference_run(tmp, cfg, state=None):
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/QCDGPU_b200_ckpt not built (needs the reference tree: make -C integration)")
    os.makedirs(tmp, exist_ok=True)
    with open(os.path.join(tmp, "init.dat"), "w") as f:
        f.write(INIT.format(**cfg))
    args = [BIN, "init.dat"] + ([state] if state else [])
    r = subprocess.run(args, cwd=tmp, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "state driver completed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    # the reference saves after every iteration (suncl.cpp:5342) under a name with one-second resolution: the newest file is the final state
    files = sorted((f for f in glob.glob(os.path.join(tmp, "su*-*.qcg")) if os.path.basename(f) != (state or "")),
                   key=lambda f: (os.path.getmtime(f), f))
    assert files, os.listdir(tmp)
    return np.fromfile(files[-1], dtype=np.uint8)


def mirror_model(cfg, fused=False):
    from qcdgpu_b200 import Model
    m = Model()
    m.lattice_group = cfg["group"]
    m.lattice_full_size = [cfg["ls"]] * 3 + [cfg["lt"]]
    m.fprefix = f"su{cfg['group']}-"
    m.ITER, m.NAV, m.NITER, m.NHIT = cfg["iter"], cfg["nav"], 1, 10
    m.BETA, m.PHI, m.ints, m.PL_level = cfg["beta"], cfg["phi"], cfg["ints"], cfg["pl"]
    m.PRNG0.PRNG_generator, m.PRNG0.PRNG_randseries = cfg["prng"], 54
    m.fused_prng = fused
    return m


def mirror_run(tmp, cfg, state=None, fused=False):
    os.makedirs(tmp, exist_ok=True)
    m = mirror_model(cfg, fused)
    m.path = str(tmp)
    if state:
        m.INIT, m.fstate = 0, state
    m.lattice_init()
    m.lattice_simulate()
    out = os.path.join(tmp, "mirror.qcg")
    m.lattice_save_state(out)
    counter = m.PRNG0.PRNG_counter
    m.close()
    return np.fromfile(out, dtype=np.uint8), counter


def with_iter(data, iters, path):
    """the state at the end of a run of fewer iterations = the state of the longer run at that point, once the header says so"""
    d = data.copy()
    d.view(np.uint32)[ITER_WORD] = iters
    d.tofile(path)
    return path


def same_but_init_word(a, b):
    assert a.size == b.size
    ha, hb = a.view(np.uint32)[:64].copy(), b.view(np.uint32)[:64].copy()
    ha[INIT_WORD] = hb[INIT_WORD] = 0
    np.testing.assert_array_equal(ha, hb)
    np.testing.assert_array_equal(a[256:], b[256:])


@pytest.mark.parametrize("group,ints,prng,beta,phi,pl", [(3, 1, "RANLUX3", 6.0, 0.0, 2), (3, 0, "RANLUX3", 5.7, 3.0, 3), (2, 0, "XOR128", 2.5, 0.0, 1)])
def test_mirror_writes_the_reference_file(tmp_path, group, ints, prng, beta, phi, pl):
    cfg = dict(group=group, ls=6, lt=4, iter=5, beta=beta, nav=2, ints=ints, prng=prng, phi=phi, pl=pl)
    ref = reference_run(str(tmp_path / "ref"), cfg)
    mine, _ = mirror_run(str(tmp_path / "mirror"), cfg)
    assert ref.size == mine.size
    np.testing.assert_array_equal(ref.view(np.uint32)[:64], mine.view(np.uint32)[:64])
    np.testing.assert_array_equal(ref, mine)


@pytest.mark.parametrize("group,ints,prng,beta", [(3, 1, "RANLUX3", 6.0), (3, 0, "RANLUX3", 5.7), (2, 0, "PM", 2.5)])
def test_continued_run_equals_uninterrupted_run(tmp_path, group, ints, prng, beta):
    cfg = dict(group=group, ls=6, lt=4, iter=9, beta=beta, nav=1, ints=ints, prng=prng, phi=0.0, pl=2)
    full = reference_run(str(tmp_path / "full"), cfg)
    part = reference_run(str(tmp_path / "part"), dict(cfg, iter=4))
    # the reference continues its own file ...
    os.makedirs(tmp_path / "cont")
    with_iter(part, 9, str(tmp_path / "cont" / "state.qcg"))
    cont = reference_run(str(tmp_path / "cont"), cfg, state="state.qcg")
    same_but_init_word(full, cont)
    # ... and so does the mirror, buffered and with the generators in the update kernel (the replayed produce() calls become a jump)
    for fused in (False, True):
        d = tmp_path / f"mirror{int(fused)}"
        os.makedirs(d)
        if fused:       # the fused run has its own instance count, so it continues its own file and is compared with its own full run
            mfull, counter = mirror_run(str(d / "full"), cfg, fused=True)
            mpart, _ = mirror_run(str(d / "part"), dict(cfg, iter=4), fused=True)
            with_iter(mpart, 9, str(d / "state.qcg"))
            mcont, counter2 = mirror_run(str(d), cfg, state=str(d / "state.qcg"), fused=True)
            same_but_init_word(mfull, mcont)
            assert counter == counter2
        else:
            with_iter(part, 9, str(d / "state.qcg"))
            mcont, _ = mirror_run(str(d), cfg, state=str(d / "state.qcg"))
            same_but_init_word(full, mcont)


def test_action_slices_through_the_reference_host(tmp_path):
    """model::get_actions_diff is one more switch of the reference without an init.dat key (QCDGPU.cpp:218); the state driver sets it from
    the environment.  The reference's host code then sizes the buffers, launches lattice_action_diff_{x,y,z} / reduce_action_diff_* and
    prints "Differentiated S data" with its own analysis (suncl.cpp:1376-1481, 1753-1791): its per-slice means must be the mirror's."""
    import re
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/QCDGPU_b200_ckpt not built")
    cfg = dict(group=3, ls=8, lt=4, iter=5, beta=6.0, nav=1, ints=0, prng="RANLUX3", phi=0.0, pl=1)
    d = tmp_path / "ref"
    os.makedirs(d)
    (d / "init.dat").write_text(INIT.format(**cfg))
    r = subprocess.run([BIN, "init.dat"], cwd=d, capture_output=True, text=True, timeout=600, env=dict(os.environ, QCDGPU_ACTIONS_DIFF="1"))
    assert r.returncode == 0 and "state driver completed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    out = sorted(glob.glob(str(d / "su3-*.txt")))
    assert out
    text = open(out[-1]).read()
    block = text[text.index("Differentiated S data"):]
    ref = {}
    for axis in "XYZ":
        mm = re.search(rf"\n{axis} .*\nS\s+(.*)\nS_variance\s+(.*)\n", block)
        ref[axis] = (np.array(mm.group(1).split(), dtype=float), np.array(mm.group(2).split(), dtype=float))
    m = mirror_model(cfg)
    m.get_actions_diff = True
    m.lattice_init()
    m.lattice_simulate()
    A = m.lattice_analysis()
    for axis in "XYZ":
        mean = np.array([x["mean_value"] for x in A[f"S_diff_{axis}"]])
        var = np.array([x["variance"] for x in A[f"S_diff_{axis}"]])
        assert ref[axis][0].size == 8
        np.testing.assert_allclose(mean, ref[axis][0], rtol=1e-11)
        np.testing.assert_allclose(var, ref[axis][1], rtol=1e-7, atol=1e-16)
    m.close()
