"""Under torchrun: a short BigLattice run (hot start, three sweeps) whose every rank prints the SHA-256 of its link table, halo slices
included -- run once with QCDGPU_HALO_FUSED=0 and once with 1, the lines must be equal (tests/test_gpu_biglattice.py, tools/halo8.sh).
QCD_DIMS=12x64x4x4 QCD_GROUP=3 QCD_PREC=1 [QCD_OR=1]"""
import os, sys, json, hashlib
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("QCD_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcdgpu_b200.biglattice import BigLattice
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
dims = tuple(int(x) for x in os.environ["QCD_DIMS"].split("x"))
def conf(p):
    p.lattice_group, p.precision, p.ints, p.BETA, p.ITER, p.NAV = int(os.environ["QCD_GROUP"]), int(os.environ["QCD_PREC"]), 0, 5.7, 3, 1
    p.fused_prng = True
    p.or_steps = int(os.environ.get("QCD_OR", "0"))
    p.desired_device = rank
B = BigLattice(dims, world, dist=dist, configure=conf, devices=[rank])
B.lattice_init(); B.lattice_simulate(); A = B.lattice_analysis()
p = B.parts[rank]
t = p.table_numpy()
print("HASH %d %s %s %d;" % (rank, hashlib.sha256(t.tobytes()).hexdigest(), "peer" if B._peer is not None else "nccl", int(p.GPU0.launch_count())), flush=True)
if rank == 0:
    print("PLQ " + json.dumps([float(x) for x in A["Plq_total"]["data"]]))
B.close(); dist.destroy_process_group()
