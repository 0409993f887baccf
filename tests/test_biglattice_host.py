"""Host-side logic of the BigLattice path on CPU: slab partition, ring topology, and the process-to-process
halo exchange (the same `exchange_packed` the NCCL path uses) run with world_size 2 and 3 on gloo."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_and_ring():
    from qcdgpu_b200.biglattice import ring_neighbours, split_slabs
    assert split_slabs(64, 8) == [8] * 8
    assert split_slabs(12, 2) == [6, 6]
    assert split_slabs(13, 3) == [5, 4, 4] and sum(split_slabs(13, 3)) == 13
    with pytest.raises(ValueError):
        split_slabs(2, 3)
    assert ring_neighbours(0, 4) == (3, 1) and ring_neighbours(3, 4) == (2, 0)
    assert ring_neighbours(0, 2) == (1, 1) and ring_neighbours(0, 1) == (0, 0)


def test_sublattice_geometry_matches_reference_options():
    """-D option string of a part (suncl.cpp:2410-2446): N1 = Nx + 2, FULL_SITES, LEFT_SITES."""
    from qcdgpu_b200.biglattice import SubLattice, split_slabs
    nx = split_slabs(12, 2)
    p = SubLattice(1, 2, (12, 6, 6, 4), nx)
    p.lattice_domain_size = [p.Nx + 2, 6, 6, 4]
    p.lattice_full_site = 12 * 6 * 6 * 4
    p.lattice_table_row_size = 1216
    o = p.options_common()
    assert "-D N1=8 " in o and "-D BIGLAT" in o
    assert f"-D FULL_SITES={12 * 6 * 6 * 4}" in o and f"-D LEFT_SITES={6 * 6 * 6 * 4}" in o


def test_biglat_init_file(tmp_path):
    """The BIGLAT build's init.dat (SU(N)/init_2.dat): per-part values in braces, everything else as in the single-lattice file."""
    from qcdgpu_b200.biglattice import from_init_file, init_file_vectors
    f = tmp_path / "init_2.dat"
    f.write_text("\tPLATFORM = {0, 0}\n\tDEVICE= {0, 1}\n\tGROUP = 3\n\tND = 4\n\tLT = 4\n\tLS = 6\n\tL1 = 12\n\tN1 = {8, 4}\n\tITER = 5\n"
                 "\tNHIT = 7\n\tBETA = 6.3690000000000e+00\n\tNAV = 3\n\tINTS = 2\n\tPRNG = RANLUX4\n\tRANDSERIES = {1420193282, 45}\n"
                 "\tPL_LEVEL = 3\n\tFMUNU = 1   # measured for the single lattice only\n")
    assert init_file_vectors(str(f)) == {"PLATFORM": [0, 0], "DEVICE": [0, 1], "N1": [8, 4], "RANDSERIES": [1420193282, 45]}
    B = from_init_file(str(f))
    assert B.full_size == [12, 6, 6, 4] and B.Nx_all == [8, 4] and B.nparts == 2
    p0, p1 = B.parts[0], B.parts[1]
    assert (p0.Nx, p0.left_x, p0.desired_device, p0.PRNG0.PRNG_randseries) == (8, 0, 0, 1420193282)
    assert (p1.Nx, p1.left_x, p1.desired_device, p1.PRNG0.PRNG_randseries) == (4, 8, 1, 45)
    for p in (p0, p1):
        assert (p.lattice_group, p.ITER, p.NHIT, p.NAV, p.ints, p.PL_level) == (3, 5, 7, 3, 2, 2) and abs(p.BETA - 6.369) < 1e-12
        assert p.PRNG0.PRNG_generator == "RANLUX4" and p.lattice_full_size == [12, 6, 6, 4] and not p.get_Fmunu
    B.close()
    f.write_text("\tGROUP = 3\n\tLS = 6\n\tLT = 4\n\tL1 = 12\n\tN1 = {7, 5}\n")
    from qcdgpu_b200.capi import QcdGpuError
    with pytest.raises(QcdGpuError):
        from_init_file(str(f))                  # odd slabs: the parts' checkerboards would not line up
    f.write_text("\tGROUP = 3\n\tLS = 6\n\tLT = 4\n\tL1 = 12\n\tN1 = {8, 6}\n")
    with pytest.raises(QcdGpuError):
        from_init_file(str(f))                  # 8 + 6 != L1


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["QCD_ROOT"])
from qcdgpu_b200.biglattice import exchange_packed, ring_neighbours, split_slabs
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
nx = split_slabs(4 * world + 1, world)
n234, planes = 24, 3
# each rank's table: (x incl. halos, plane, site) filled with a value that encodes (owner, x, plane, site)
def table(k):
    t = np.zeros((nx[k] + 2, planes, n234))
    for x in range(1, nx[k] + 1):
        for pl in range(planes):
            t[x, pl] = 1e6 * k + 1e3 * x + 10 * pl + np.arange(n234) / 100.0
    return t
mine = table(rank)
left, right = ring_neighbours(rank, world)
for sweep in range(3):
    s_first = torch.tensor(mine[1].copy()); s_last = torch.tensor(mine[nx[rank]].copy())
    r_left = torch.empty_like(s_first); r_right = torch.empty_like(s_first)
    for w in exchange_packed(dist, s_first, s_last, r_left, r_right, left, right):
        w.wait()
    mine[0] = r_left.numpy(); mine[nx[rank] + 1] = r_right.numpy()
    L, R = table(left), table(right)
    assert np.array_equal(mine[0], L[nx[left]]), "left halo must hold the left neighbour's LAST real slice"
    assert np.array_equal(mine[nx[rank] + 1], R[1]), "right halo must hold the right neighbour's FIRST real slice"
dist.barrier()
if rank == 0:
    print("HALO-OK", world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, QCD_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"HALO-OK {world}" in r.stdout


FUSED_PROTOCOL_WORKER = r'''
"""The host side of the fused halo exchange (BigLattice._setup_fused_halo / _update_pass_fused / lattice_sweep) on gloo, with a stand-in
for the device: link tables and flag words live in shared memory (one value per x-slice: the number of the pass that wrote it), the
"kernels" run synchronously in Python.  What is checked is the protocol: which slice goes where in a ring whose left and right
neighbours differ, which flag word is written, and that a boundary launch of pass n finds pass n - 1 or n in its halos, never anything else."""
import os, sys, time
import numpy as np
from multiprocessing import shared_memory
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["QCD_ROOT"])
import qcdgpu_b200.biglattice as BL
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
NX = 4

class FakeGPU:
    def __init__(self):
        self.shm, self.arr, self.opened, self.peer, self.xr, self.log = {}, {}, {}, {}, {}, []
        self.pass_no = 0
    def buffer_init(self, kind, size, host, size_of):
        bid = len(self.shm) + 100
        self._make(bid, size * size_of)
        return bid
    def _make(self, bid, nbytes):
        m = shared_memory.SharedMemory(create=True, size=max(nbytes, 64))
        self.shm[bid] = m
        self.arr[bid] = np.ndarray((max(nbytes, 64) // 8,), dtype=np.float64 if bid == TABLE else np.uint64, buffer=m.buf)
        self.arr[bid][:] = 0
    def buffer_ipc_export(self, bid):
        return self.shm[bid].name.encode().ljust(64, b"\0")
    def ipc_open(self, handle):
        if os.environ.get("QCD_FAIL_OPEN") == str(rank):
            raise BL.QcdGpuError("ipc_open: peer access not available (simulated)")
        m = shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode())
        key = (len(self.opened) + 1) << 32
        self.opened[key] = m
        return key
    def ipc_close(self, ptr):
        self.opened.pop(ptr).close()
    def _peer_array(self, ptr, dtype):
        m = self.opened[ptr & ~0xffffffff]
        return np.ndarray((m.size // 8,), dtype=dtype, buffer=m.buf), (ptr & 0xffffffff) // 8
    def kernel_set_peer_halo(self, kid, side, table, src_x, dst_x, rowsize):
        self.peer[(kid, side)] = (table, src_x, dst_x)
    def kernel_set_x_range(self, kid, first, count, stride=1):
        self.xr[kid] = (first, count, stride)
    def kernel_run(self, kid):
        if kid == "gs":
            return
        first, count, stride = self.xr[kid]
        xs = [1 + first + i * stride for i in range(NX if count < 0 else count)]
        t = self.arr[TABLE]
        n = B._pass_no
        if 1 in xs or NX in xs:          # a boundary launch reads its halos: they hold the neighbours' previous pass (n - 2 behind a
            ok = (n - 1, n) if (n % 9) != 2 else (n - 2, n)       # sweep's fence, which stores no links), or already this one
            assert t[0] in ok and t[NX + 1] in ok, (rank, n, t.tolist())
            assert self.arr[B._flags][0] >= n - 1 and self.arr[B._flags][1] >= n - 1
        for x in xs:
            t[x] = n
            for side in (0, 1):
                table, src_x, dst_x = self.peer[(kid, side)]
                if x == src_x:
                    a, off = self._peer_array(table, np.float64)
                    a[off + dst_x] = n
        self.log.append((n, tuple(xs)))
    def halo_signal(self, a, b, value):
        for ptr in (a, b):
            arr, off = self._peer_array(ptr, np.uint64)
            arr[off] = value
    def halo_wait(self, flags, value, timeout_s=5.0):
        t0 = time.time()
        f = self.arr[flags]
        while f[0] < value or f[1] < value:
            assert time.time() - t0 < 20, ("wait timed out", rank, value, f[:3].tolist())
            time.sleep(0.0005)
    def halo_bytes(self, program, d): return 1
    def buffer_map(self, bid, dtype): return self.arr[bid]
    def synchronize(self): pass
    def close(self): pass

TABLE = 7
class FakePRNG:
    def produce(self): pass
class FakePart:
    pass
B = BL.BigLattice.__new__(BL.BigLattice)
B.nparts, B.dist, B.local, B.Nx_all = world, dist, [rank], [NX] * world
B._torch, B._stream, B._peer, B._pass_no, B.fused_halo, B.halo_bytes_per_pass, B._inproc = torch, None, None, 0, True, 0, False
p = FakePart()
p.k, p.Nx, p.desired_device, p.lattice_table, p.lattice_table_row_size, p.program_update = rank, NX, 0, TABLE, 64, 1
p.x_first_real, p.x_last_real, p.or_steps, p.turnoff_gramschmidt, p.sun_GramSchmidt_id = 1, NX, 0, False, "gs"
p.PRNG0 = FakePRNG()
p.sun_update_id = {(par, d): f"{par}{d}" for par in ("odd", "even") for d in BL.DIRS}
g = p.GPU0 = FakeGPU()
g._make(TABLE, (NX + 2) * 8)
B.parts = {rank: p}
B._setup_fused_halo()
if os.environ.get("QCD_FAIL_OPEN"):
    # one rank cannot open its neighbours' tables: ALL ranks must end up on the NCCL path, with nothing left open
    assert B._peer is None and B.fused_halo is False and not g.opened, (rank, B._peer, g.opened)
    dist.barrier()
    for m in g.shm.values():
        m.close(); m.unlink()
    if rank == 0:
        print("FUSED-PROTOCOL-OK", world)
    dist.destroy_process_group()
    sys.exit(0)
assert B._peer is not None
left, right = BL.ring_neighbours(rank, world)
for sweep in range(3):
    if rank == (sweep % world):
        time.sleep(0.05)            # one rank lags: the others must wait for it, not run ahead by more than a pass
    B.lattice_sweep()
    n = B._pass_no
    assert n == 9 * (sweep + 1)         # a fence + 8 passes
    t = g.arr[TABLE].copy()
    time.sleep(0.02)                    # "measurement": the neighbours may be done with theirs -- they must not touch my halos yet
    assert np.array_equal(t, g.arr[TABLE])
    assert t[0] == n and t[NX + 1] == n and all(t[1:NX + 1] == n), t.tolist()     # halos complete when the sweep returns
# every pass: boundary slices (1 and NX) in one launch first, then the interior
assert [x for _, x in g.log[:2]] == [(1, NX), tuple(range(2, NX))] and g.log[0][0] == 2
dist.barrier()
B._peer and [g.ipc_close(k) for k in list(g.opened)]
dist.barrier()
for m in g.shm.values():
    m.close(); m.unlink()
if rank == 0:
    print("FUSED-PROTOCOL-OK", world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world,fail", [(2, ""), (3, ""), (3, "1")])
def test_fused_halo_protocol_gloo(world, fail, tmp_path):
    """QCDGPU_HALO_FUSED host logic without a GPU: ring of 2 (left == right) and 3 (left != right) processes; fail: that rank cannot
    open peer memory and everybody falls back to the NCCL exchange."""
    script = tmp_path / "worker.py"
    script.write_text(FUSED_PROTOCOL_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, QCD_ROOT=ROOT, OMP_NUM_THREADS="1", QCD_FAIL_OPEN=fail)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"FUSED-PROTOCOL-OK {world}" in r.stdout


ANALYSIS_WORKER = r'''
"""BigLattice.lattice_analysis across processes (gloo): per-part raw sums of action, plaquettes and Polyakov loops are packed into one
all-reduce and normalised by the FULL lattice -- checked against the sums a single process would see."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["QCD_ROOT"])
import qcdgpu_b200.biglattice as BL
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ITER, OFF = 5, 8
def raw(k):                       # what part k's device buffers hold: (ITER, 2) doubles per observable, Polyakov |P|^2 / |P|^4 behind OFF
    r = np.random.default_rng(100 + k)
    pl = np.zeros((2 * OFF, 2)); pl[:ITER] = r.normal(size=(ITER, 2)); pl[OFF:OFF + ITER] = r.uniform(size=(ITER, 2))
    return {"lattice_energies": r.uniform(size=(ITER + 3, 2)), "lattice_energies_plq": r.uniform(size=(ITER + 3, 2)), "lattice_polyakov_loop": pl}
class FakeGPU:
    def __init__(self, k): self.bufs = raw(k)
    def buffer_map(self, name, dtype): return self.bufs[name].reshape(-1)
class Part: pass
p = Part()
p.ITER, p.PL_level, p.lattice_polyakov_loop_size, p.lattice_group, p.desired_device = ITER, 2, OFF, 3, 0
p.lattice_energies, p.lattice_energies_plq, p.lattice_polyakov_loop = "lattice_energies", "lattice_energies_plq", "lattice_polyakov_loop"
p.GPU0 = FakeGPU(rank)
B = BL.BigLattice.__new__(BL.BigLattice)
B.full_size, B.nparts, B.dist, B._torch, B.parts = [4 * world, 6, 2, 8], world, dist, torch, {rank: p}
A = B.lattice_analysis()
tot = {k: sum(raw(j)[k] for j in range(world)) for k in raw(0)}
V, v3 = float(np.prod(B.full_size)), float(np.prod(B.full_size[:3]))
np.testing.assert_allclose(A["S_spat"]["data"], tot["lattice_energies"][:ITER, 0] / (3 * V), rtol=1e-14)
np.testing.assert_allclose(A["Plq_temp"]["data"], tot["lattice_energies_plq"][:ITER, 1] / (3 * V), rtol=1e-14)
np.testing.assert_allclose(A["Polyakov_loop"]["data"], tot["lattice_polyakov_loop"][:ITER, 0] / (v3 * 3), rtol=1e-13)
np.testing.assert_allclose(A["Polyakov_loop_im"]["data"], tot["lattice_polyakov_loop"][:ITER, 1] / (v3 * 3), rtol=1e-13)
np.testing.assert_allclose(A["Polyakov_loop_P2"]["data"], tot["lattice_polyakov_loop"][OFF:OFF + ITER, 0] / (v3 * 9), rtol=1e-13)
np.testing.assert_allclose(A["Polyakov_loop_P4"]["data"], tot["lattice_polyakov_loop"][OFF:OFF + ITER, 1] / (v3 * 81), rtol=1e-13)
dist.barrier()
if rank == 0:
    print("ANALYSIS-OK", world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_biglattice_analysis_all_reduce_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(ANALYSIS_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, QCD_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"ANALYSIS-OK {world}" in r.stdout
