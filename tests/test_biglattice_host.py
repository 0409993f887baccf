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
