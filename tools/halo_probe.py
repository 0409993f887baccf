"""Under torchrun (N >= 2): what the pieces of the fused halo exchange cost on one rank -- the boundary launch with and without the
peer stores, the interior launch, a signal + wait pair.  QCDGPU_HALO_FUSED=1 python -m torch.distributed.run ... tools/halo_probe.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcdgpu_b200.biglattice import BigLattice
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
def conf(p):
    p.lattice_group, p.precision, p.ints, p.BETA, p.NHIT, p.ITER = 3, 1, 1, 6.0, 10, 64
    p.fused_prng = True
    p.desired_device = rank
B = BigLattice((64 * world, 64, 64, 64), world, dist=dist, configure=conf, devices=[rank])
B.lattice_init(); B.lattice_start()
for _ in range(2):
    B.lattice_sweep()
B.synchronize(); dist.barrier()
p = B.parts[rank]; g = p.GPU0
kid = p.sun_update_id[("odd", "X")]
def timed(fn, n=10):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(B._stream):
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n
def boundary():
    p.PRNG0.produce(); g.kernel_set_x_range(kid, 0, 2, p.Nx - 1); g.kernel_run(kid)
def interior():
    p.PRNG0.produce(); g.kernel_set_x_range(kid, 1, p.Nx - 2); g.kernel_run(kid)
res = {}
res["boundary, peer stores"] = timed(boundary)
res["interior"] = timed(interior)
peers = B._peer
if peers is not None:
    def sigwait():
        B._pass_no += 1
        g.halo_signal(peers["flag_left"], peers["flag_right"], B._pass_no)
        g.halo_wait(B._flags, B._pass_no)
    res["signal + wait"] = timed(sigwait)
    def full_pass():
        B._update_pass_fused("odd", "X")
    res["whole fused pass"] = timed(full_pass, 16)
    for side in (0, 1):
        g.kernel_set_peer_halo(kid, side, 0, 0, 0, 0)
    res["boundary, no peer stores"] = timed(boundary)
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank}: " + ", ".join(f"{k} {v * 1e3:.0f} us" for k, v in res.items()), flush=True)
B._peer = None if peers is None else B._peer
dist.barrier()
B.close(); dist.destroy_process_group()
