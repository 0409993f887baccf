#!/usr/bin/env python3
"""Where the multi-GPU overhead of a pass goes (run under torchrun, one rank per GPU):
sweep time (max over ranks) with both boundary slices in one launch, in two launches, and without the boundary-first split, alternating.   torchrun --nproc-per-node 2 tools/halo_timeline.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from qcdgpu_b200.biglattice import BigLattice

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
N = int(os.environ.get("SLAB", "64"))


def configure(p):
    p.lattice_group, p.precision, p.ints, p.BETA, p.NHIT = 3, 1, 1, 6.0, 10
    p.PRNG0.PRNG_generator = "RANLUX3"
    p.fused_prng = True
    p.ITER = 64
    p.desired_device = lr


B = BigLattice((N * world, 64, 64, 64), world, dist=dist, configure=configure, devices=[lr])
B.lattice_init(); B.lattice_start()
p = B.parts[rank]; g = p.GPU0
steps = int(os.environ.get("STEPS", "10"))
for rnd in range(int(os.environ.get("ROUNDS", "3"))):
    for name, overlap, pair in (("boundary pair + interior", True, True), ("two boundary launches + interior", True, False), ("one launch, exchange behind it", False, True)):
        B.overlap, B.pair_launch = overlap, pair
        for _ in range(2):
            B.lattice_sweep()
        B.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(B._stream):
            e0.record()
            for _ in range(steps):
                B.lattice_sweep()
            e1.record()
        B.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{lr}")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name:34s}: {float(ms.item()) / steps:.2f} ms per sweep")
B.close(); dist.destroy_process_group()
