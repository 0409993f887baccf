"""BigLattice (x-slab split + halo exchange) on a B200.  With the PRNG-free INTS=2 start a partitioned run must
reproduce the unpartitioned one exactly (the reference only promises statistical agreement once PRNGs are
involved, suncl.cpp:377-386); parts can share one device, as in the reference's init_2.dat."""
import json
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _configure(sun, beta, ints, iters, prec=1, fused=False, prng="RANLUX3"):
    def f(p):
        p.lattice_group, p.precision, p.ints, p.BETA, p.ITER = sun, prec, ints, beta, iters
        p.fused_prng = fused
        p.PRNG0.PRNG_generator = prng
        if not fused:
            p.GPU0.L.qgpu_set_max_memory_width(p.GPU0.h, 64)
    return f


@pytest.mark.parametrize("sun,dims,nparts,or_steps", [(3, (4, 6, 2, 8), 1, 0), (3, (4, 6, 2, 8), 2, 0), (3, (12, 4, 2, 4), 3, 0), (2, (8, 4, 8, 2), 2, 0),
                                                     (3, (12, 4, 2, 4), 3, 2), (2, (8, 4, 8, 2), 2, 1)])
def test_gid_start_partitioned_equals_single(sun, dims, nparts, or_steps, monkeypatch):
    """or_steps > 0: overrelaxation passes after each heat-bath sweep go through the same boundary-first schedule and exchange."""
    from qcdgpu_b200 import Model
    from qcdgpu_b200.biglattice import BigLattice
    beta = 5.7 if sun == 3 else 2.3
    m = Model()
    m.lattice_group, m.lattice_full_size, m.precision, m.ints, m.BETA, m.ITER, m.PL_level = sun, list(dims), 1, 2, beta, 4, 0
    m.or_steps = or_steps
    m.PL_level = 2
    m.lattice_init()
    m.lattice_simulate()
    m.lattice_analysis()
    cols = ("Plq_spat", "Plq_temp", "S_spat", "S_temp", "Polyakov_loop", "Polyakov_loop_im", "Polyakov_loop_P2", "Polyakov_loop_P4")
    want = np.array([[m.Analysis[c]["data"][i] for c in cols] for i in range(4)])
    m.close()
    for overlap, fused_halo in ((False, "1"), (True, "1"), (True, "0")):
        # fused_halo: boundary links stored into the neighbours' halos by the update kernels (events between the parts' streams), or
        # copy kernels over peer memory behind the pass
        monkeypatch.setenv("QCDGPU_HALO_FUSED", fused_halo)
        def conf(p, base=_configure(sun, beta, 2, 4)):
            base(p)
            p.or_steps = or_steps
            p.PL_level = 2
        B = BigLattice(dims, nparts, configure=conf)
        B.overlap = overlap
        B.lattice_init()
        B.lattice_simulate()
        assert B._inproc == (fused_halo == "1")
        A = B.lattice_analysis()
        got = np.array([[A[c]["data"][i] for c in cols] for i in range(4)])
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)     # Polyakov loops of the parts add up to the single lattice's
        B.close()


def test_gid_biglattice_matches_reference_golden():
    gold = json.load(open(os.path.join(GOLD, "ref_cpurun_su3_4x6x2x8_gid.json")))
    p = gold["params"]
    from qcdgpu_b200.biglattice import BigLattice

    def conf(q):
        _configure(3, p["BETA"], 2, p["ITER"])(q)
        q.turnoff_gramschmidt = True
    B = BigLattice((p["L1"], p["L2"], p["L3"], p["L4"]), 2, configure=conf)
    B.lattice_init()
    B.lattice_simulate()
    A = B.lattice_analysis()
    got = np.array([[A[c]["data"][i] for c in gold["columns"]] for i in range(p["ITER"])])
    np.testing.assert_allclose(got, np.array(gold["rows"]), rtol=5e-6)
    B.close()


@pytest.mark.parametrize("fused", [False, True])
def test_overlap_schedule_is_bit_identical(fused):
    """Boundary-slices-first + interior (3 launches per pass) vs one launch per pass: same links, hot start, PRNGs on."""
    from qcdgpu_b200.biglattice import BigLattice
    tables = []
    for overlap in (False, True):
        B = BigLattice((12, 4, 4, 4), 2, configure=_configure(3, 6.0, 0, 3, fused=fused))
        B.overlap = overlap
        B.lattice_init()
        B.lattice_simulate()
        tables.append([p.table_numpy() for p in B.parts.values()])
        B.close()
    for a, b in zip(*tables):
        np.testing.assert_array_equal(a.view(np.uint8), b.view(np.uint8))


def test_tma_staged_links_in_slabs():
    """N2 = 64 slabs: the x-range launches of the two-phase RANLUX kernel stage their links by TMA (halo slices included);
    the same generators with every produce() materialised use the buffered kernel with plain loads -- same links bit for bit."""
    from qcdgpu_b200.biglattice import BigLattice
    tables = []
    for fused in (True, False):
        B = BigLattice((8, 64, 2, 2), 2, configure=_configure(3, 5.7, 0, 3, fused=True))
        B.overlap = True
        B.lattice_init()
        if not fused:
            for part in B.parts.values():
                part.GPU0.set_fused_prng(False)
        B.lattice_simulate()
        tables.append([p.table_numpy() for p in B.parts.values()])
        B.close()
    for a, b in zip(*tables):
        np.testing.assert_array_equal(a.view(np.uint8), b.view(np.uint8))


def test_two_devices_in_one_process_n2_64(monkeypatch):
    """Several parts in ONE process on different GPUs (devices=[0, 1], P2P halo copies), N2 = 64: every device launches the kernels
    that need more than 48 KB of dynamic shared memory (TMA-staged update, 63 KB), and the opt-in is per device
    (update_inst.cu, ensure_smem).  Same links as with both parts on device 0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from qcdgpu_b200.biglattice import BigLattice
    tables = []
    for devices, fused_halo in (([0, 0], "1"), ([0, 1], "1"), ([1, 0], "1"), ([0, 1], "0")):
        # fused_halo "1": the update kernels store boundary links into the other device's halo slices (peer stores, events);
        # "0": copy kernels reading peer memory behind the pass
        monkeypatch.setenv("QCDGPU_HALO_FUSED", fused_halo)
        B = BigLattice((8, 64, 2, 2), 2, configure=_configure(3, 5.7, 0, 3, fused=True), devices=devices)
        B.overlap = True
        B.lattice_init()
        B.lattice_simulate()
        assert B._inproc == (fused_halo == "1")
        tables.append([p.table_numpy() for p in B.parts.values()])
        B.close()
    for t in tables[1:]:
        for a, b in zip(tables[0], t):
            np.testing.assert_array_equal(a.view(np.uint8), b.view(np.uint8))


def test_partitioned_ensemble_agrees():
    """With PRNGs the parts use their own series: plaquette after thermalisation agrees statistically with the
    reference's CPU_RUN series (8^4, beta = 6)."""
    gold = json.load(open(os.path.join(GOLD, "ref_cpurun_su3_8x8x8x8_cold_ens.json")))
    ref = np.array(gold["rows"])[15:, 2]
    from qcdgpu_b200.biglattice import BigLattice

    def conf(q):
        _configure(3, 6.0, 1, 61, fused=True)(q)
        q.NAV = 40
    B = BigLattice((8, 8, 8, 8), 2, configure=conf)
    B.lattice_init()
    B.lattice_simulate()
    A = B.lattice_analysis()
    mine = A["Plq_total"]["data"][1:]
    se = 3 * np.sqrt(ref.var() / len(ref) + mine.var() / len(mine))
    assert abs(mine.mean() - ref.mean()) < 5 * se, (mine.mean(), ref.mean(), se)
    B.close()


NCCL_WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["QCD_ROOT"])
from qcdgpu_b200.biglattice import BigLattice
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
def conf(p):
    p.lattice_group, p.precision, p.ints, p.BETA, p.ITER = 3, 1, 2, 5.7, 3
    p.turnoff_gramschmidt = True
    p.desired_device = rank
B = BigLattice((4, 6, 2, 8), world, dist=dist, configure=conf, devices=[rank])
B.lattice_init(); B.lattice_simulate(); A = B.lattice_analysis()
if rank == 0:
    print("ROWS " + json.dumps([[float(A[c]["data"][i]) for c in ("Plq_spat", "Plq_temp", "Plq_total", "S_spat", "S_temp", "S_total")] for i in range(3)]))
B.close(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("direct", ["1", "0", "fused"])
def test_nccl_two_gpus_gid(tmp_path, direct):
    """Two processes, NCCL send/recv of the boundary slices after every pass -- straight from / into the link tables (direct) or
    staged through packed buffers -- or the exchange fused into the update kernels (the default), against the reference's CPU_RUN
    table of the same deterministic (INTS = 2) run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    gold = json.load(open(os.path.join(GOLD, "ref_cpurun_su3_4x6x2x8_gid.json")))
    script = tmp_path / "w.py"
    script.write_text(NCCL_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, env=dict(os.environ, QCD_ROOT=ROOT, QCDGPU_HALO_DIRECT="1" if direct == "1" else "0", QCDGPU_HALO_FUSED="1" if direct == "fused" else "0"), timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rows = json.loads([l for l in r.stdout.splitlines() if l.startswith("ROWS ")][0][5:])
    np.testing.assert_allclose(np.array(rows), np.array(gold["rows"]), rtol=5e-6)


FUSED_WORKER = open(os.path.join(ROOT, "tools", "halo_check.py")).read()


@pytest.mark.parametrize("group,prec,dims,orx", [(3, 1, "12x64x4x4", 0), (3, 1, "4x6x2x8", 0), (2, 0, "8x8x4x6", 0), (3, 1, "8x64x2x4", 1)])
def test_fused_halo_exchange_two_gpus(tmp_path, group, prec, dims, orx):
    """QCDGPU_HALO_FUSED=1: the boundary-slice update kernels store their links into the neighbours' halo slices over peer memory
    (no pack / NCCL / unpack), a flag word per neighbour carries the pass number.  Same seeds, hot start, three sweeps: every
    rank's link table -- halo slices included -- must equal the NCCL exchange's bit for bit (TMA-staged kernels, the unsplit launch
    of two-slice parts, SU(2) fp32 and overrelaxation passes all take the peer path)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "w.py"
    script.write_text(FUSED_WORKER)
    out = {}
    for mode in ("0", "1"):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        env = dict(os.environ, QCD_ROOT=ROOT, QCDGPU_HALO_FUSED=mode, QCD_DIMS=dims, QCD_GROUP=str(group), QCD_PREC=str(prec), QCD_OR=str(orx))
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", str(port), str(script)], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        h = sorted(list(x) for x in re.findall(r"HASH (\d+) ([0-9a-f]{64}) (peer|nccl) (\d+);", r.stdout))     # the ranks' lines may run into each other
        assert len(h) == 2 and all(x[2] == ("peer" if mode == "1" else "nccl") for x in h), h
        out[mode] = ([x[1] for x in h], re.findall(r"PLQ (\[[^\]]*\])", r.stdout)[0], [int(x[3]) for x in h])
    assert out["0"][0] == out["1"][0], (out["0"], out["1"])
    assert out["0"][1] == out["1"][1]
    assert out["1"][2][0] < out["0"][2][0]          # fewer launches: no pack / unpack kernels
