"""BigLattice: 1-D slab decomposition along x over several GPUs with NVLink halo exchange.

Mirror of the reference's BIGLAT build (SubLattice, model::UpdateEdges, lattice_simulate,
SU(N)/suncl/suncl.cpp:278-287, 2388-2561, 4005-4062, 4804-5210): the lattice is cut along x = N1, the
slowest storage index (what north_star calls the "t-split"); every part stores
Nx real slices plus one halo slice on each side (N1 = Nx + 2, kernels built with -D BIGLAT -D FULL_SITES
-D LEFT_SITES), parts form a ring, and after each of the 8 (parity, direction) passes the direction-mu
links of the two boundary slices go to the neighbours' halos.  Each part has its own PRNG and RANDSERIES
(suncl.cpp:377-386), so a partitioned run agrees with the single-device run statistically, and exactly
for the PRNG-free INTS=2 start.

What is new here:
  * one process per GPU (torch.distributed / NCCL over NVLink) or several parts in one process (P2P
    `qgpu_halo_copy` kernels reading peer memory, also usable with all parts on one device, like the
    reference's init_2.dat);
  * the reference stages halos through mapped host memory after the whole pass (clEnqueueMapBuffer +
    memcpy, fully serialised); here the two boundary slices are updated first, their links start
    travelling (NCCL send/recv on its own stream, or a P2P copy kernel), and the interior slices are
    updated meanwhile;
  * optionally (QCDGPU_HALO_DIRECT=1) NCCL sends the boundary slices straight out of the link table and receives straight into the
    halo slices (a slice of a vec4 plane is contiguous): no pack / unpack kernels, no staging buffers -- measured slightly slower
    than the staged exchange, which stays the default;
  * QCDGPU_HALO_FUSED=1 (one process per GPU): no exchange step at all.  Every part opens its neighbours' link tables as peer
    memory (CUDA IPC), and the update kernel of a boundary slice stores each updated link twice: into its own table and, over NVLink,
    into the neighbour's halo slice (Geom::peer_*, lattice.cuh).  A flag word per neighbour carries the pass number: a part waits for
    "pass n - 1 written" from both sides before its boundary launch of pass n and announces n behind it (two one-warp kernels);
    inside a pass the checkerboard already keeps writers and readers apart, across parts as within one; what reads the halos between
    sweeps (measurements) is fenced by a pass number of its own at the head of each sweep;
  * observables are summed over parts with one small all-reduce.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from .capi import GPU, QcdGpuError
from .model import DIRS, Model, model_start_hot


def split_slabs(L1: int, nparts: int) -> list[int]:
    """Nx per part: as equal as possible, larger parts first (the reference takes N1 = {a, b, ...} from init.dat)."""
    if nparts < 1 or L1 < nparts:
        raise ValueError(f"cannot cut {L1} x-slices into {nparts} parts")
    base, rem = divmod(L1, nparts)
    nx = [base + (1 if k < rem else 0) for k in range(nparts)]
    return nx


def parity_consistent(nx: list[int]) -> bool:
    """The kernels classify even/odd sites from the part-local x (sun_common.cl:383-417): the checkerboards of
    neighbouring parts only line up when every part owns an even number of slices."""
    return all(n % 2 == 0 for n in nx)


def ring_neighbours(k: int, nparts: int) -> tuple[int, int]:
    """(left, right) part indices; k_next = (k + 1) % Nparts (suncl.cpp:4027)."""
    return (k - 1) % nparts, (k + 1) % nparts


def exchange_packed(dist, send_first, send_last, recv_left, recv_right, left: int, right: int):
    """Ring exchange of packed boundary slices between processes (model::UpdateEdges across devices):
    my LAST real slice goes to the right neighbour's left halo, my FIRST real slice to the left neighbour's
    right halo.  Returns the work handles.  Send/recv order matters when left == right (two parts): messages
    between one pair of ranks match in posting order, so "to the right" is posted before "to the left" and
    "from the left" before "from the right"."""
    ops = [dist.P2POp(dist.isend, send_last, right), dist.P2POp(dist.isend, send_first, left),
           dist.P2POp(dist.irecv, recv_left, left), dist.P2POp(dist.irecv, recv_right, right)]
    return dist.batch_isend_irecv(ops)


class _DeviceBytes:
    """A device allocation of the backend as a __cuda_array_interface__ object (torch.as_tensor maps it without a copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2, "strides": None}


def exchange_direct(dist, send_last, send_first, recv_left, recv_right, left: int, right: int):
    """exchange_packed without staging buffers: lists of tensors that alias the link tables (one contiguous chunk per vec4 plane of
    the boundary / halo slice).  Same posting order as exchange_packed, chunk by chunk."""
    ops = ([dist.P2POp(dist.isend, t, right) for t in send_last] + [dist.P2POp(dist.isend, t, left) for t in send_first] +
           [dist.P2POp(dist.irecv, t, left) for t in recv_left] + [dist.P2POp(dist.irecv, t, right) for t in recv_right])
    return dist.batch_isend_irecv(ops)


class SubLattice(Model):
    """One part: a Model built with -D BIGLAT, owning Nx real x-slices (SubLattice, suncl.h:470-700)."""

    def __init__(self, k: int, nparts: int, full_size, Nx_all: list[int], device: int = 0):
        super().__init__()
        self.k, self.nparts = k, nparts
        self.lattice_full_size = list(full_size)
        self.biglat = True
        self.Nx = Nx_all[k]
        self.left_x = sum(Nx_all[:k])
        self.desired_device = device
        self.PL_level = 0            # 1, 2: Polyakov loops per part, summed in BigLattice.lattice_analysis (3, the slice-resolved loops: single lattice only)

    # x index (in the stored part, halos included) of the slices involved in the exchange
    @property
    def x_first_real(self): return 1
    @property
    def x_last_real(self): return self.Nx
    @property
    def x_left_halo(self): return 0
    @property
    def x_right_halo(self): return self.Nx + 1


class BigLattice:
    """The driver: parts of this process + transport to the ring neighbours."""

    def __init__(self, full_size, nparts: int, local_parts: list[int] | None = None, devices: list[int] | None = None,
                 dist=None, configure=None, randseries: list[int] | None = None, Nx_all: list[int] | None = None):
        """`dist`: a torch.distributed module with an initialised process group when parts live in other
        processes (then this process owns exactly part `dist.get_rank()`), else None (all parts local).
        `Nx_all`: the x-slices of every part (init.dat's N1 = {a, b, ...}); default: as equal as possible."""
        self.full_size = list(full_size)
        self.nparts = nparts
        self.Nx_all = list(Nx_all) if Nx_all is not None else split_slabs(full_size[0], nparts)
        if len(self.Nx_all) != nparts or sum(self.Nx_all) != full_size[0] or min(self.Nx_all) < 1:
            raise QcdGpuError(f"x-slabs {self.Nx_all} do not cut {full_size[0]} x-slices into {nparts} parts")
        if nparts > 1 and not parity_consistent(self.Nx_all):
            raise QcdGpuError(f"x-slabs {self.Nx_all}: every part needs an even number of x-slices for a consistent checkerboard")
        self.dist = dist
        if dist is not None:
            local_parts = [dist.get_rank()]
        elif local_parts is None:
            local_parts = list(range(nparts))
        self.local = local_parts
        devices = devices or [0] * len(local_parts)
        self.parts: dict[int, SubLattice] = {}
        for k, dev in zip(local_parts, devices):
            p = SubLattice(k, nparts, full_size, self.Nx_all, dev)
            p.PRNG0.PRNG_randseries = (randseries[k] if randseries else 54 + 1000 * k)   # per-part series (init_2.dat)
            if configure:
                configure(p)
            self.parts[k] = p
        self.overlap = True
        self._torch = None
        self._stream = None
        self._pack = {}
        self._tables, self._chunks = {}, {}
        # QCDGPU_HALO_DIRECT=1: NCCL straight from / into the link tables (no pack / unpack kernels).  Measured on 2 B200: 30.1-30.3 ms
        # per 64^4-slab sweep against 29.9 staged through packed buffers (twelve chunks per exchange instead of four), so staging stays
        # the default
        self.direct_halo = os.environ.get("QCDGPU_HALO_DIRECT", "0") == "1"
        self.pair_launch = os.environ.get("QCDGPU_HALO_PAIR", "1") != "0"     # both boundary slices in one launch
        # halo exchange fused into the update kernels (peer stores + flags); needs one process per GPU and equal slabs
        # (default in that case since it was measured: 28.25 against 29.67 ms per 64^4-slab sweep on 2 B200; QCDGPU_HALO_FUSED=0: NCCL)
        self.fused_halo = os.environ.get("QCDGPU_HALO_FUSED", "1") != "0" and dist is not None and nparts > 1 and len(set(self.Nx_all)) == 1
        self._peer = None
        self._pass_no = 0
        # the same peer stores for parts of THIS process (one or several devices): events instead of flags
        self.fused_inproc = os.environ.get("QCDGPU_HALO_FUSED", "1") != "0" and dist is None
        self._inproc = False
        self.halo_bytes_per_pass = 0

    # ------------------------------------------------------------------ set-up
    def lattice_init(self):
        for p in self.parts.values():
            p.lattice_init()
        if self.dist is not None:
            import torch
            self._torch = torch
            p = self.parts[self.local[0]]
            torch.cuda.set_device(p.desired_device)
            self._stream = torch.cuda.Stream(device=p.desired_device)
            p.GPU0.set_stream(self._stream.cuda_stream)      # kernels and NCCL ops are ordered through this stream

    def lattice_start(self):
        for p in self.parts.values():
            p.lattice_create_buffers()
            p.lattice_make_programs()
            if p.ints == model_start_hot:
                p.PRNG0.produce()
                for i, k in enumerate(p.sun_init_hot_id):
                    p.GPU0.kernel_run(k)
                    if i < 3:
                        p.PRNG0.produce()
            else:
                p.GPU0.kernel_run(p.sun_init_id)
        self.finish_edges(self.update_edges(-1))          # a hot start fills halos with unrelated links: take the neighbours' real ones
        if self.fused_halo:
            # the start-up exchange went through NCCL, whose receive side may still be copying into my halos when a neighbour's first
            # boundary launch stores into them: everybody finishes it before anybody starts a pass
            self.synchronize()
            self.dist.barrier()
            self._setup_fused_halo()
        if self.dist is None and len(self.parts) > 1:
            # This exchange copies ALL planes of the neighbours' boundary slices on the READERS' streams; the owners' first update pass
            # writes direction X there.  Per-pass exchanges are ordered by plane separation, this one needs the readers to be done.
            for p in self.parts.values():
                p.GPU0.synchronize()
        if self.fused_inproc and len(self.parts) == self.nparts:
            self._setup_fused_inprocess()
        self.measure()
        for p in self.parts.values():
            p.ITER_counter += 1

    # ------------------------------------------------------------------ halo exchange (model::UpdateEdges)
    def _packed(self, p: SubLattice, name: str, nbytes: int):
        t = self._pack.get((p.k, name))
        if t is None or t.numel() < nbytes:
            t = self._torch.empty(nbytes, dtype=self._torch.uint8, device=f"cuda:{p.desired_device}")
            self._pack[(p.k, name)] = t
        return t

    def _slice_chunks(self, p: SubLattice, x: int, d: int):
        """uint8 tensors over the link table: x-slice `x` of the vec4 planes of direction d (all directions if d < 0)"""
        key = (p.k, x, d)
        t = self._chunks.get(key)
        if t is None:
            table = self._tables.get(p.k)
            if table is None:
                mem = _DeviceBytes(p.GPU0.buffer_device_ptr(p.lattice_table), p.GPU0.buffer_size_bytes(p.lattice_table))
                table = self._tables[p.k] = self._torch.as_tensor(mem, device=f"cuda:{p.desired_device}")
            vec = 32 if p.precision == 1 else 16
            n234, row = p.lattice_domain_n2n3n4, p.lattice_table_row_size
            planes = [dd + 4 * k for dd in ((d,) if d >= 0 else range(4)) for k in range(p.lattice_group_elements // 4)]
            t = self._chunks[key] = [table[(pl * row + x * n234) * vec: (pl * row + (x + 1) * n234) * vec] for pl in planes]
        return t

    def _setup_fused_halo(self):
        """Open the neighbours' link tables and flag words as peer memory and tell every update kernel where its boundary links go."""
        dist, p = self.dist, self.parts[self.local[0]]
        g = p.GPU0
        self._flags = g.buffer_init(GPU.buffer_type_Global, 4, np.zeros(4, dtype=np.uint64), 8)
        try:
            mine = (g.buffer_ipc_export(p.lattice_table), g.buffer_ipc_export(self._flags), int(p.lattice_table_row_size), int(p.Nx))
        except QcdGpuError:
            mine = None
        everyone = [None] * self.nparts
        dist.all_gather_object(everyone, mine)
        left, right = ring_neighbours(p.k, self.nparts)
        opened = {}
        try:
            if any(e is None for e in everyone):
                raise QcdGpuError("a part could not export its link table")
            for k in {left, right}:
                th, fh, row, nx = everyone[k]
                opened[k] = (g.ipc_open(th), g.ipc_open(fh), row, nx)
            ok = 1
        except QcdGpuError:
            ok = 0
        # all parts or none: a part without peer access to a neighbour (or neighbours in other containers) sends everybody back to NCCL
        agreed = self._torch.tensor([ok], device=f"cuda:{p.desired_device}" if self._torch.cuda.is_available() else "cpu")
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        if int(agreed.item()) == 0:
            for t, f, _, _ in opened.values():
                g.ipc_close(t)
                g.ipc_close(f)
            self.fused_halo = False
            return
        (lt, lf, lrow, lnx), (rt, rf, rrow, _) = opened[left], opened[right]
        kids = list(p.sun_update_id.values()) + list(getattr(p, "sun_overrelax_id", {}).values())
        for kid in kids:
            g.kernel_set_peer_halo(kid, 0, lt, p.x_first_real, lnx + 1, lrow)     # my first real slice -> left neighbour's right halo
            g.kernel_set_peer_halo(kid, 1, rt, p.x_last_real, 0, rrow)            # my last real slice  -> right neighbour's left halo
        # the word I write at each neighbour: its "from my right" (word 1) at the left one, its "from my left" (word 0) at the right one
        self._peer = {"opened": opened, "flag_left": lf + 8, "flag_right": rf}
        self.halo_bytes_per_pass = 2 * g.halo_bytes(p.program_update, 0)
        self.synchronize()
        dist.barrier()

    def _setup_fused_inprocess(self):
        """All parts in this process: every update kernel gets its neighbours' tables (plain device pointers; the part's own table
        when it is its own neighbour), the writer's device is given access to the neighbour's."""
        for p in self.parts.values():
            p.GPU0.synchronize()
        for k, p in self.parts.items():
            left, right = ring_neighbours(k, self.nparts)
            L, R = self.parts[left], self.parts[right]
            for nb in (L, R):
                if nb is not p:
                    p.GPU0.enable_peer_access(nb.GPU0)
            lt, rt = L.GPU0.buffer_device_ptr(L.lattice_table), R.GPU0.buffer_device_ptr(R.lattice_table)
            for kid in list(p.sun_update_id.values()) + list(getattr(p, "sun_overrelax_id", {}).values()):
                p.GPU0.kernel_set_peer_halo(kid, 0, lt, p.x_first_real, L.Nx + 1, L.lattice_table_row_size)
                p.GPU0.kernel_set_peer_halo(kid, 1, rt, p.x_last_real, 0, R.lattice_table_row_size)
            self.halo_bytes_per_pass = 2 * p.GPU0.halo_bytes(p.program_update, 0)
        self._inproc = True

    def _order_neighbours(self):
        """every part's stream waits for what its neighbours' streams hold so far"""
        for k, p in self.parts.items():
            for nb in set(ring_neighbours(k, self.nparts)):
                if nb != k:
                    p.GPU0.halo_order(self.parts[nb].GPU0)

    def _check_halo_timeouts(self):
        if self._peer is not None:
            p = self.parts[self.local[0]]
            lost = int(p.GPU0.buffer_map(self._flags, np.uint64)[2])
            if lost:
                raise QcdGpuError(f"fused halo exchange: {lost} wait(s) for a neighbour's pass flag timed out")

    def update_edges(self, d: int):
        """Direction-d planes (all planes if d < 0) of every part's two boundary slices -> neighbours' halos."""
        if self.nparts == 1 and self.dist is None:
            p = self.parts[0]
            # periodic in x through the part's own halos
            p.GPU0.halo_copy(p.program_update, p.lattice_table, p.x_left_halo, p.GPU0, p.program_update, p.lattice_table, p.x_last_real, d)
            p.GPU0.halo_copy(p.program_update, p.lattice_table, p.x_right_halo, p.GPU0, p.program_update, p.lattice_table, p.x_first_real, d)
            self.halo_bytes_per_pass = 2 * p.GPU0.halo_bytes(p.program_update, d)
            return None
        if self.dist is None:
            for k, p in self.parts.items():
                left, right = ring_neighbours(k, self.nparts)
                L, R = self.parts[left], self.parts[right]
                # my left halo <- left neighbour's last real slice; my right halo <- right neighbour's first real slice
                p.GPU0.halo_copy(p.program_update, p.lattice_table, p.x_left_halo, L.GPU0, L.program_update, L.lattice_table, L.x_last_real, d)
                p.GPU0.halo_copy(p.program_update, p.lattice_table, p.x_right_halo, R.GPU0, R.program_update, R.lattice_table, R.x_first_real, d)
            self.halo_bytes_per_pass = 2 * p.GPU0.halo_bytes(p.program_update, d)
            return None
        # one part per process, everything ordered on self._stream
        torch, dist = self._torch, self.dist
        p = self.parts[self.local[0]]
        left, right = ring_neighbours(p.k, self.nparts)
        nb = p.GPU0.halo_bytes(p.program_update, d)
        self.halo_bytes_per_pass = 2 * nb
        if self.direct_halo:
            # NCCL reads the boundary slices and writes the halo slices IN the link table: a slice of one vec4 plane is contiguous
            # (element gid + plane * ROWSIZE, x the slowest site index), so no pack / unpack kernels and no staging buffers.  The
            # interior update that runs meanwhile reads the boundary slices too and touches no halo slice.
            with torch.cuda.stream(self._stream):
                works = exchange_direct(dist, self._slice_chunks(p, p.x_last_real, d), self._slice_chunks(p, p.x_first_real, d),
                                        self._slice_chunks(p, p.x_left_halo, d), self._slice_chunks(p, p.x_right_halo, d), left, right)
            return (works, None, None, d)
        # pack -> NCCL send/recv -> unpack
        with torch.cuda.stream(self._stream):
            s_first, s_last = self._packed(p, "s_first", nb), self._packed(p, "s_last", nb)
            r_left, r_right = self._packed(p, "r_left", nb), self._packed(p, "r_right", nb)
            p.GPU0.halo_pack(p.program_update, p.lattice_table, p.x_first_real, d, s_first.data_ptr())
            p.GPU0.halo_pack(p.program_update, p.lattice_table, p.x_last_real, d, s_last.data_ptr())
            works = exchange_packed(dist, s_first[:nb], s_last[:nb], r_left[:nb], r_right[:nb], left, right)
        return (works, r_left, r_right, d)

    def finish_edges(self, pending):
        if pending is None:
            return
        works, r_left, r_right, d = pending
        p = self.parts[self.local[0]]
        with self._torch.cuda.stream(self._stream):
            for w in works:
                w.wait()                 # stream-level wait: the unpack kernels run after the NCCL transfer
            if r_left is None:           # direct exchange: the halos are in place
                return
            p.GPU0.halo_unpack(p.program_update, p.lattice_table, p.x_left_halo, d, r_left.data_ptr())
            p.GPU0.halo_unpack(p.program_update, p.lattice_table, p.x_right_halo, d, r_right.data_ptr())

    # ------------------------------------------------------------------ sweeps (lattice_simulate BIGLAT, suncl.cpp:4804-5210)
    def _update_pass_fused(self, par: str, d: str, overrelax: bool = False):
        """One pass with the exchange inside the kernels: wait for the neighbours' previous pass, boundary slices (stored here AND in
        the neighbours' halos), announce the pass, interior slices."""
        p = self.parts[self.local[0]]
        g, kid = p.GPU0, (p.sun_overrelax_id if overrelax else p.sun_update_id)[(par, d)]
        self._pass_no += 1
        n = self._pass_no
        if not overrelax:
            p.PRNG0.produce()
        g.halo_wait(self._flags, n - 1)
        if p.Nx >= 3:
            g.kernel_set_x_range(kid, 0, 2, p.Nx - 1)
            g.kernel_run(kid)
            g.halo_signal(self._peer["flag_left"], self._peer["flag_right"], n)
            g.kernel_set_x_range(kid, 1, p.Nx - 2)
            g.kernel_run(kid)
        else:
            g.kernel_set_x_range(kid, 0, -1)
            g.kernel_run(kid)
            g.halo_signal(self._peer["flag_left"], self._peer["flag_right"], n)

    def _halo_fence(self):
        """Between sweeps a part reads its halo slices outside the passes (measurements, a real reunitarisation launch).  A neighbour
        that is already in its next sweep must not store into them meanwhile: what this part has enqueued so far is announced as a pass of
        its own, and the neighbours' first boundary launch of the sweep waits for it like for any other pass."""
        self._pass_no += 1
        self.parts[self.local[0]].GPU0.halo_signal(self._peer["flag_left"], self._peer["flag_right"], self._pass_no)

    def _update_pass(self, par: str, d: str, overrelax: bool = False):
        """One (parity, direction) pass of all local parts + the exchange of what it changed; overrelax: the PRNG-free overrelaxation
        pass (`or_steps`) through the same schedule."""
        if self._peer is not None:
            return self._update_pass_fused(par, d, overrelax)
        di = DIRS.index(d)
        split = self.overlap and all(p.Nx >= 3 for p in self.parts.values())
        kids = {k: (p.sun_overrelax_id if overrelax else p.sun_update_id)[(par, d)] for k, p in self.parts.items()}
        for p in self.parts.values():
            if not overrelax:
                p.PRNG0.produce()
            kid = kids[p.k]
            if split:
                # boundary slices first so that their links can travel while the interior is updated: both in ONE launch (slices 0 and
                # Nx - 1 = a range of two with stride Nx - 1; as two launches each pays its own ramp-up and tail, ~0.1 ms at 64^3)
                if self.pair_launch:
                    p.GPU0.kernel_set_x_range(kid, 0, 2, p.Nx - 1)
                    p.GPU0.kernel_run(kid)
                else:
                    p.GPU0.kernel_set_x_range(kid, 0, 1)
                    p.GPU0.kernel_run(kid)
                    p.GPU0.kernel_set_x_range(kid, p.Nx - 1, 1)
                    p.GPU0.kernel_run(kid)
            else:
                p.GPU0.kernel_set_x_range(kid, 0, -1)
                p.GPU0.kernel_run(kid)
        if self._inproc:
            # the launches above stored the boundary links into the neighbours' halos as well: what is left of the exchange is order --
            # nothing of a part's stream that follows (its next pass reads the halos) before the neighbours' launches of this pass
            self._order_neighbours()
            if split:
                for p in self.parts.values():
                    kid = kids[p.k]
                    p.GPU0.kernel_set_x_range(kid, 1, p.Nx - 2)
                    p.GPU0.kernel_run(kid)
            return
        if self.dist is not None:
            pending = self.update_edges(di)
            if split:
                for p in self.parts.values():
                    kid = kids[p.k]
                    p.GPU0.kernel_set_x_range(kid, 1, p.Nx - 2)
                    p.GPU0.kernel_run(kid)
            self.finish_edges(pending)
        else:
            if split:
                for p in self.parts.values():
                    kid = kids[p.k]
                    p.GPU0.kernel_set_x_range(kid, 1, p.Nx - 2)
                    p.GPU0.kernel_run(kid)
            self.update_edges(di)

    def lattice_sweep(self):
        if self._peer is not None:
            self._halo_fence()
        if self._inproc:
            self._order_neighbours()           # the same fence with events: the neighbours' measurements are done with my boundary slices' copies
        for par in ("odd", "even"):
            for d in DIRS:
                self._update_pass(par, d)
        for _ in range(next(iter(self.parts.values())).or_steps):
            for par in ("odd", "even"):
                for d in DIRS:
                    self._update_pass(par, d, overrelax=True)
        if self.dist is None and len(self.parts) > 1:
            for p in self.parts.values():      # halo copies read neighbours' tables on the readers' streams
                p.GPU0.synchronize()
        if self._peer is not None:             # my halos hold the neighbours' last pass before anything else reads them
            self.parts[self.local[0]].GPU0.halo_wait(self._flags, self._pass_no)
        for p in self.parts.values():
            if not p.turnoff_gramschmidt:
                p.GPU0.kernel_run(p.sun_GramSchmidt_id)

    # ------------------------------------------------------------------ measurements
    def measure(self):
        for p in self.parts.values():
            p._measure()

    def synchronize(self):
        for p in self.parts.values():
            p.GPU0.synchronize()
        if self._stream is not None:
            self._stream.synchronize()
        self._check_halo_timeouts()

    def lattice_simulate(self):
        self.lattice_start()
        p0 = next(iter(self.parts.values()))
        for _ in range(p0.NAV):
            self.lattice_sweep()
        for _ in range(1, p0.ITER):
            for _ in range(p0.NITER):
                self.lattice_sweep()
            self.measure()
            for p in self.parts.values():
                p.ITER_counter += 1
        self.synchronize()

    def lattice_analysis(self) -> dict:
        """Per-part sums -> global observables (lattice_analysis_SLtoL, suncl.cpp:909-959): the parts' raw sums
        add up; denominators are those of the full lattice."""
        p0 = next(iter(self.parts.values()))
        ITER = p0.ITER
        sums = {}
        names = [("S", "lattice_energies", 0), ("Plq", "lattice_energies_plq", 0)]
        if p0.PL_level > 0:      # Polyakov loops close inside a part (the t direction is not cut): per-part sums of Re, Im (and |P|^2, |P|^4) add up
            names.append(("PL", "lattice_polyakov_loop", 0))
            if p0.PL_level > 1:
                names.append(("PL24", "lattice_polyakov_loop", p0.lattice_polyakov_loop_size))
        for name, buf, off in names:
            acc = np.zeros((ITER, 2))
            for p in self.parts.values():
                acc += p.GPU0.buffer_map(getattr(p, buf), np.float64).reshape(-1, 2)[off: off + ITER]
            sums[name] = acc
        if self.dist is not None:
            torch = self._torch
            flat = torch.tensor(np.concatenate([sums[n].ravel() for n, _, _ in names]),
                                device=f"cuda:{p0.desired_device}" if torch.cuda.is_available() else "cpu")
            self.dist.all_reduce(flat)
            flat = flat.cpu().numpy()
            for i, (n, _, _) in enumerate(names):
                sums[n] = flat[2 * ITER * i: 2 * ITER * (i + 1)].reshape(ITER, 2)
        den = float(np.prod(self.full_size) * 3)
        A = {}
        for name in ("S", "Plq"):
            A[f"{name}_spat"] = Model._series(sums[name][:, 0].copy(), den)
            A[f"{name}_temp"] = Model._series(sums[name][:, 1].copy(), den)
            A[f"{name}_total"] = Model._joint(A[f"{name}_spat"], A[f"{name}_temp"])
        if p0.PL_level > 0:      # model::lattice_analysis, suncl.cpp:1440-1480: normalised by the full spatial volume and powers of N
            v3, N = float(np.prod(self.full_size[:3])), float(p0.lattice_group)
            A["Polyakov_loop"] = Model._series(sums["PL"][:, 0].copy(), v3 * N)
            A["Polyakov_loop_im"] = Model._series(sums["PL"][:, 1].copy(), v3 * N)
            if p0.PL_level > 1:
                A["Polyakov_loop_P2"] = Model._series(sums["PL24"][:, 0].copy(), v3 * N ** 2)
                A["Polyakov_loop_P4"] = Model._series(sums["PL24"][:, 1].copy(), v3 * N ** 4)
        self.Analysis = A
        return A

    def close(self):
        if self._peer is not None:
            self.synchronize()
            self.dist.barrier()                 # nobody unmaps a table a neighbour may still be storing into
            g = self.parts[self.local[0]].GPU0
            for t, f, _, _ in self._peer["opened"].values():
                g.ipc_close(t)
                g.ipc_close(f)
            self._peer = None
        for p in self.parts.values():
            p.close()


# ---------------------------------------------------------------------------------------------- bench (N > 1)
def bench_main(args, workload):
    """bench.py --gpus N under torchrun: weak scaling, one 64-slice slab of SU(3) fp64 per GPU."""
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, measured_peaks, workload_config

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dims = tuple(int(x) for x in args.dims.split("x")) if args.dims else workload["dims"]
    full = (dims[0] * world, dims[1], dims[2], dims[3])

    def configure(p):
        p.lattice_group, p.precision, p.ints = workload["group"], workload["precision"], workload["ints"]
        p.BETA, p.NHIT = workload["beta"], workload["nhit"]
        p.PRNG0.PRNG_generator = workload["prng"]
        p.fused_prng = not args.buffered
        p.ITER = 4096
        p.desired_device = local_rank

    B = BigLattice(full, world, dist=dist, configure=configure, devices=[local_rank])
    B.lattice_init()
    B.lattice_start()
    p = B.parts[rank]
    g = p.GPU0
    for _ in range(args.warmup):
        B.lattice_sweep()
    B.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = g.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(B._stream):
        ev0.record()
        for _ in range(args.steps):
            B.lattice_sweep()
        ev1.record()
    ev1.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=f"cuda:{local_rank}")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = g.launch_count() - l0
    V_total = int(np.prod(full))
    links = 4 * V_total
    value = links * args.steps / (ms * 1e-3)
    # e2e through the host-facing loop: per step parameter upload, sweep, measurement, all-reduced result read back
    t0 = time.perf_counter()
    params = g.buffer_map(p.lattice_parameters, np.uint8)
    for _ in range(args.steps):
        g._chk(g.L.qgpu_buffer_unmap(g.h, p.lattice_parameters, params.ctypes.data), "unmap")
        B.lattice_sweep()
        g.kernel_run(p.sun_measurement_plq_id)
        g.kernel_init_constant_reset(p.sun_measurement_plq_reduce_id, 1, p.argument_plq_index)
        g.kernel_run(p.sun_measurement_plq_reduce_id)
        e = g.buffer_map(p.lattice_energies_plq, np.float64)[2:4].copy()
        tt = torch.tensor(e, device=f"cuda:{local_rank}")
        dist.all_reduce(tt)
        plq = tt.cpu().numpy()
    dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local_rank}")
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = links * args.steps / float(e2e_s.item())
    ck = clocks.stop() if rank == 0 else None
    peak, peak_kind = measured_peaks()
    if rank == 0:
        bytes_per_s = value * 864.0 / 1e9 / world
        out = {
            "metric": "SU(3) link updates/sec", "value": value, "unit": "link-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, dims),
            "details": {"sweeps_per_s": args.steps / (ms * 1e-3), "halo_bytes_per_pass_per_gpu": B.halo_bytes_per_pass,
                        "plaquette_after_run": float(plq.sum() / (6.0 * V_total)),
                        "halo": ("fused into the update kernels: boundary links are stored into the neighbours' halo slices over NVLink peer memory (CUDA IPC), a flag word per neighbour carries the pass number"
                                 if B._peer is not None else
                                 "NCCL send/recv straight from / into the link tables (contiguous slice per vec4 plane), overlapped with the interior update"
                                 if B.direct_halo else "NCCL send/recv of packed boundary slices, overlapped with the interior update")},
            "roofline": {"bound": "hbm", "achieved": bytes_per_s, "peak": peak, "unit": "GB/s", "frac": bytes_per_s / peak,
                         "traffic": None, "peak_source": peak_kind, "note": "per GPU, whole sweep (update + halo + reunitarisation)"},
            "e2e": {"value": e2e, "unit": "link-updates/s", "h2d_bytes_per_step": int(params.size) * world, "d2h_bytes_per_step": 16 * world},
            "gpu_launches": int(launches), "clocks": ck,
        }
        print(json.dumps(out))
    B.close()
    dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- init.dat of the BIGLAT build
def init_file_vectors(path: str) -> dict:
    """The keys a BIGLAT init.dat gives one value per part, `KEY = {a, b, ...}` (clinterface.cpp:2122-2142, suncl.cpp:307-386):
    PLATFORM, DEVICE, N1, RANDSERIES.  A plain number counts as a one-element list."""
    from .model import get_init_file
    out = {}
    for key, ival, _, text in get_init_file(path):
        if key in ("PLATFORM", "DEVICE", "N1", "RANDSERIES"):
            t = text.strip()
            if t.startswith("{"):
                out[key] = [int(x) for x in t.strip("{}").replace(",", " ").split()]
            else:
                out[key] = [ival]
    return out


def from_init_file(path: str, dist=None) -> "BigLattice":
    """A BigLattice from the reference's BIGLAT init.dat (its sample: SU(N)/init_2.dat): N1 = {a, b, ...} are the parts' x-slices (their
    sum is L1), DEVICE = {...} their devices (all parts in this process, like the reference's single host program; several parts may
    share a device), RANDSERIES = {...} their random series.  Everything else is read by every part as model::lattice_get_init_file does."""
    from .model import Model
    v = init_file_vectors(path)
    probe = Model.__new__(Model)
    Model.__init__(probe)
    probe.lattice_get_init_file(path)
    full = list(probe.lattice_full_size)
    probe.close()
    nx = v.get("N1")
    if not nx:
        raise QcdGpuError(f"{path}: a BIGLAT init file names the parts' x-slices, N1 = {{a, b, ...}}")
    nparts = len(nx)
    devices = (v.get("DEVICE") or [0]) * nparts if len(v.get("DEVICE") or [0]) == 1 else v["DEVICE"]
    series = v.get("RANDSERIES")
    if series is not None and len(series) == 1:
        series = [series[0] + 1000 * k for k in range(nparts)]

    def configure(p):
        dev, k, rs = p.desired_device, p.k, p.PRNG0.PRNG_randseries
        p.lattice_get_init_file(path)                 # scalar keys; the per-part ones come back as 0 and are restored
        p.desired_device, p.desired_platform = dev, 0
        p.PRNG0.PRNG_randseries = rs
        p.lattice_full_size = list(full)
        if p.PL_level > 2:
            p.PL_level = 2                            # the slice-resolved loops exist for the single lattice only
        # measured for the single lattice only: a BIGLAT file that asks for them still runs, without those columns
        p.get_wilson_loop = p.get_Fmunu = p.get_F0mu = p.get_actions_diff = False

    return BigLattice(full, nparts, devices=devices[:nparts], dist=dist, configure=configure, randseries=series, Nx_all=nx)


def main(argv=None):
    """The BIGLAT build's main (QCDGPU.cpp:60-266): init.dat -> init -> simulate -> analysis -> table."""
    import sys
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        raise SystemExit("usage: python -m qcdgpu_b200.biglattice init.dat   (BIGLAT format: N1 = {a, b}, DEVICE = {0, 1}, ...)")
    t0 = time.time()
    B = from_init_file(argv[0])
    B.lattice_init()
    p0 = next(iter(B.parts.values()))
    print(p0.header)
    print(f" parts (x-slices)            : {B.Nx_all}   devices: {[p.desired_device for p in B.parts.values()]}")
    B.lattice_simulate()
    A = B.lattice_analysis()
    cols = [c for c in p0.columns() if c in A]
    print(" #     " + " ".join(f"{c:>14s}" for c in cols))
    for i in range(p0.ITER):
        print(f"[{i:5d}]: " + " ".join(f"{A[c]['data'][i]:14.8f}" for c in cols))
    p0.Analysis, p0.elapsed = A, time.time() - t0
    out = p0.lattice_write_results()
    print(f"\nresults: {out}\nEXECUTION TIME: {time.time() - t0:f} sec\n")
    B.close()


if __name__ == "__main__":
    main()
