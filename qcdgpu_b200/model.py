"""Host run loop: mirror of model_CL::model (SU(N)/suncl/suncl.{h,cpp}) for the gauge-update path.

Same field names, same init.dat keys, same buffer sizes, same kernel wiring and the same call
order as the reference's `lattice_init` / `lattice_simulate` / `lattice_analysis` /
`lattice_write_results`, talking to the device only through the GPU_CL::GPU-shaped C ABI
(capi.GPU).  The two knobs that are new are `fused_prng` (PRNG_instances = PRNGSTEP and the
generator state kept in registers of the update kernel) and `or_steps` (overrelaxation passes
after each heat-bath sweep; the reference has none).

State files (.qcg, suncl.cpp:1857-2079) are written and read in the reference's byte layout (lattice_save_state /
lattice_load_state; INIT = 0 continues a run, fast-forwarding the generators by PRNG_counter passes).
Not here (out of the hot path): lattice_action_diff_*, the start.txt/finish.txt handshake.
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass, field

import numpy as np

from .capi import GPU, QcdGpuError
from .prng import PRNG

VERSION = "1.5"

model_start_hot, model_start_cold, model_start_gid = 0, 1, 2
model_precision_single, model_precision_double, model_precision_mixed = 0, 1, 2   # suncl.h:73

X, Y, Z, T = 0, 1, 2, 3
DIRS = "XYZT"


def get_init_file(path: str) -> list[tuple[str, int, float, str]]:
    """GPU::get_init_file (clinterface.cpp:2060-2154): `KEY = value  # comment` lines."""
    out = []
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if "=" not in line:
                continue
            key, val = line.split("=", 1)
            key, val = key.strip(), val.strip()
            try:
                ival = int(float(val)) if val and val[0] != "{" else 0
            except ValueError:
                ival = 0
            try:
                fval = float(val)
            except ValueError:
                fval = 0.0
            out.append((key, ival, fval, val))
    return out


MODEL_energies_size = 7     # suncl.h:222: plaquette sums + 6 Fmunu projections
BIN_HEADER_SIZE = 64        # suncl.cpp:60: dwords in front of the data of a .qcg state file
# PRNG::convert_generator_to_uint (random.cpp:110-122); MRG32K3A has no number there (the reference maps unknown numbers to RANLUX3)
GENERATOR_CODES = {"XOR128": 1, "RANLUX0": 2, "RANLUX1": 3, "RANLUX2": 4, "RANLUX3": 5, "RANLUX4": 6, "RANLUX": 7, "RANMAR": 8,
                   "PM": 9, "XOR7": 10, "RANECU": 11, "MRG32K3A": 12}


def convert_str_uint(text: str, offset: int) -> int:
    """model::convert_str_uint (suncl.cpp:709-716): four characters from `offset`, little end first."""
    return sum(ord(text[offset + i]) << (8 * i) for i in range(4) if offset + i < len(text))


def getK(n1: int, n2: int, ws: int) -> int:
    """suncl.cpp:2081-2091: smallest i >= n2 with n1 * i a multiple of the work-group size."""
    i = n2
    while i < n2 * ws:
        if n1 * i % ws == 0:
            break
        i += 1
    return i


class Model:
    """model_CL::model for one device (the non-BIGLAT build)."""

    def __init__(self):
        self.GPU0 = GPU()
        self.PRNG0 = PRNG()
        # QCDGPU.cpp:61-220 defaults
        self.GPU0.GPU_limit_max_workgroup_size = 64
        self.desired_platform = 0
        self.desired_device = 0
        self.PRNG0.PRNG_generator = "RANLUX3"
        self.PRNG0.PRNG_randseries = 54
        self.lattice_nd = 4
        self.lattice_group = 3
        self.lattice_full_size = [32, 32, 32, 32]
        self.precision = model_precision_double
        self.INIT = 1
        self.NAV = 0
        self.ITER = 2
        self.NITER = 1
        self.NHIT = 10
        self.NHITPar = 1
        self.BETA = 6.5
        self.PHI = 0.0
        self.OMEGA = 0.0
        self.ints = model_start_cold
        self.PL_level = 2           # 3 adds the slice-resolved lattice_polyakov_diff_{x,y,z} (the reference program's default)
        self.get_wilson_loop = False      # GETWILSON, WILSONR, WILSONT (QCDGPU.cpp:197-209)
        self.wilson_R = 7
        self.wilson_T = 2
        self.get_Fmunu = False            # FMUNU: chromomagnetic projections of the spatial plaquettes
        self.get_F0mu = False             # F0MU: chromoelectric projections of the temporal plaquettes
        self.Fmunu_gen1 = 0               # FMUNU1/2/4 (0 = lambda_3 / sigma_3)
        self.Fmunu_gen2 = 0               # FMUNU5/6/7 (0 = lambda_8; SU(2): always sigma_2)
        self.turnoff_updates = False
        self.turnoff_prns = False
        self.turnoff_gramschmidt = False
        self.get_actions_avr = True
        self.get_plaquettes_avr = True
        self.get_actions_diff = False     # action per x / y / z slice (model::get_actions_diff: QCDGPU.cpp:218, no init.dat key in the reference)
        self.path = "./"
        self.fprefix = "su3-"
        self.fstate = ""            # state file to continue from when INIT = 0 (model::fstate, QCDGPU.cpp:234)
        self.turnoff_config_save = True   # QCDGPU.cpp:203
        self.write_lattice_state_every_secs = 2.0     # QCDGPU.cpp:215
        self.PRNG_counter = 0       # produce() calls of the run that wrote the loaded state file (suncl.h:172)
        self.LOAD_state = 0
        # BIGLAT part description (biglattice.py): this Model then owns Nx real x-slices plus one halo slice each side
        self.biglat = False
        self.Nx = 0                 # real x-slices of the part (SubLat[k].Nx, QCDGPU.cpp:117-118)
        self.left_x = 0             # x-slices owned by the parts to the left
        # extensions
        self.fused_prng = False
        self.async_run = True
        self.or_steps = 0
        # counters
        self.NAV_counter = 0
        self.ITER_counter = 0
        self.Analysis: dict[str, dict] = {}
        self.header = ""
        self.elapsed = 0.0

    # ------------------------------------------------------------------ init.dat (suncl.cpp:289-464)
    def lattice_get_init_file(self, path: str):
        for key, ival, fval, text in get_init_file(path):
            if key == "PLATFORM": self.desired_platform = ival
            elif key == "DEVICE": self.desired_device = ival
            elif key == "GROUP":
                self.lattice_group = ival
                self.fprefix = f"su{ival}-"
            elif key == "ND": self.lattice_nd = ival
            elif key in ("L1", "L2", "L3", "L4"): self.lattice_full_size[int(key[1]) - 1] = ival
            elif key == "LS": self.lattice_full_size[0:3] = [ival, ival, ival]
            elif key == "LT": self.lattice_full_size[self.lattice_nd - 1] = ival
            elif key == "ITER": self.ITER = ival
            elif key == "NITER": self.NITER = ival
            elif key == "NHIT": self.NHIT = ival
            elif key == "BETA": self.BETA = fval
            elif key == "PHI": self.PHI = fval
            elif key == "OMEGA": self.OMEGA = fval
            elif key == "NAV": self.NAV = ival
            elif key == "INTS": self.ints = ival if ival in (0, 1, 2) else model_start_cold
            elif key == "OUTPUTPATH": self.path = text
            elif key == "GETRETRACE": self.get_plaquettes_avr = True
            elif key == "PL_LEVEL": self.PL_level = ival
            elif key == "GETWILSON": self.get_wilson_loop = True
            elif key == "WILSONR": self.wilson_R = ival
            elif key == "WILSONT": self.wilson_T = ival
            elif key == "TURNOFFFMUNU": self.get_Fmunu = self.get_F0mu = False
            elif key == "FMUNU": self.get_Fmunu, self.get_F0mu = True, False
            elif key == "F0MU": self.get_Fmunu, self.get_F0mu = False, True
            elif key in ("FMUNU1", "FMUNU2", "FMUNU4"): self.Fmunu_gen1 = int(key[-1])
            elif key in ("FMUNU5", "FMUNU6", "FMUNU7"): self.Fmunu_gen2 = int(key[-1])
            # extensions (not reference keys): PRECISION 0/1/2, NHITPAR, FUSEDPRNG, ORSTEPS
            elif key == "PRECISION": self.precision = ival
            elif key == "NHITPAR": self.NHITPar = ival
            elif key == "FUSEDPRNG": self.fused_prng = bool(ival)
            elif key == "ORSTEPS": self.or_steps = ival
            # state files: the reference has the fields (INIT, fstate, turnoff_config_save) but no init.dat keys for them
            elif key == "ACTIONSDIFF": self.get_actions_diff = bool(ival)      # model::get_actions_diff has no key in the reference
            elif key == "INIT": self.INIT = ival
            elif key == "FSTATE": self.fstate = text
            elif key == "SAVESTATE": self.turnoff_config_save = not bool(ival)
            self.PRNG0.parameters_setup(key, ival, text)

    # ------------------------------------------------------------------ lattice_init (suncl.cpp:672-706)
    def lattice_init(self):
        g = self.GPU0
        g.device_select(self.desired_platform, self.desired_device)
        g.device_initialize()
        g.set_async(self.async_run)
        g.set_fused_prng(self.fused_prng)
        if self.INIT == 0:
            self.lattice_load_state()      # the header: run parameters of the interrupted run (suncl.cpp:695)
        self.model_lattice_init()

    # model_lattice_init (suncl.cpp:2212-2384)
    def model_lattice_init(self):
        g = self.GPU0
        if self.lattice_group not in (2, 3):
            raise QcdGpuError("GROUP must be 2 or 3")
        if self.biglat and (self.PL_level > 2 or self.get_wilson_loop or self.get_Fmunu or self.get_F0mu or self.get_actions_diff):
            raise QcdGpuError("PL_LEVEL 3, Wilson loops, Fmunu and the slice-resolved action are implemented for the single-lattice build only")
        full = self.lattice_full_size
        if self.biglat:
            # suncl.cpp:2388-2446: the part is stored with N1 = Nx + 2 (one halo slice on each side)
            n = [self.Nx + 2, full[1], full[2], full[3]]
            real_n1 = self.Nx
        else:
            n = list(full)
            real_n1 = full[0]
        self.lattice_domain_size = list(n)
        self.lattice_domain_n1 = n[0]
        self.lattice_domain_n2n3 = n[1] * n[2]
        self.lattice_domain_exact_n1n2n3 = real_n1 * n[1] * n[2]
        self.lattice_domain_n2n3n4 = n[1] * n[2] * n[3]
        self.lattice_domain_site = n[0] * n[1] * n[2] * n[3]
        self.lattice_domain_exact_site = real_n1 * n[1] * n[2] * n[3]
        self.lattice_full_site = full[0] * full[1] * full[2] * full[3]
        self.lattice_full_n1n2n3 = full[0] * full[1] * full[2]
        self.lattice_domain_link = self.lattice_nd * self.lattice_domain_site
        local_size_intel = g.GPU_limit_max_workgroup_size or 0
        workgroup_factor = local_size_intel if local_size_intel else 32
        ge = 12 if self.lattice_group == 3 else 4
        self.lattice_group_elements = ge
        al = g.buffer_size_align
        self.lattice_parameters_size = al(8)
        self.lattice_energies_size = al(self.ITER)
        self.lattice_energies_offset = self.lattice_energies_size
        self.lattice_table_row_size = al(self.lattice_domain_site)
        self.lattice_table_row_size_half = al(self.lattice_domain_exact_site // 2)
        self.lattice_table_exact_row_size = al(self.lattice_domain_exact_site)
        self.lattice_table_exact_row_size_half = al(self.lattice_domain_exact_site // 2)
        self.lattice_table_size = al(self.lattice_table_row_size * self.lattice_nd * ge // 4)
        self.lattice_measurement_size = al(int(math.ceil(self.lattice_table_exact_row_size / workgroup_factor)))
        fm = self.get_Fmunu or self.get_F0mu
        self.lattice_measurement_size_F = self.lattice_measurement_size * (MODEL_energies_size if fm else 1)   # suncl.cpp:2281, 2296-2299
        if self.PL_level > 2:      # suncl.cpp:2284-2285: one transverse extent padded so that a work-group stays inside a slice
            self.PLK = getK(n[0], n[1], local_size_intel)
            self.lattice_polyakov_size = al(self.lattice_domain_n2n3 * self.PLK)
        else:
            self.PLK = 0
            self.lattice_polyakov_size = al(self.lattice_domain_exact_n1n2n3)
        self.lattice_polyakov_loop_size = al(self.ITER)
        self.lattice_polyakov_loop_offset = al(int(math.ceil(self.lattice_polyakov_size / workgroup_factor)))
        if self.get_actions_diff:  # suncl.cpp:2291-2294: grid of all action kernels = the index space padded to PLK
            self.lattice_action_size = al(self.lattice_domain_n2n3n4 * getK(n[0], n[1], local_size_intel))
            # (the reference warns "buffer plattice_measurement should be resized" instead, suncl.cpp:3511-3515)
            self.lattice_measurement_size = max(self.lattice_measurement_size, al(int(math.ceil(self.lattice_action_size / workgroup_factor))))
            self.lattice_measurement_size_F = self.lattice_measurement_size * (MODEL_energies_size if fm else 1)
        else:
            self.lattice_action_size = self.lattice_table_exact_row_size
        # PRNG preparation (suncl.cpp:2301-2311)
        P = self.PRNG0
        half = self.lattice_table_row_size_half
        P.PRNG_instances = half if self.fused_prng else 0       # field is user-settable: "0 for autoselect"
        if self.ints == model_start_hot:
            P.PRNG_samples = al(self.lattice_table_row_size * self.lattice_nd + 3 * half * (self.NHIT + 1))
        else:
            P.PRNG_samples = al(self.NHITPar * (3 * half * (self.NHIT + 1)))
        if P.PRNG_instances:   # explicit instance count: samples per instance, as PRNG::initialize does for autoselect
            P.PRNG_samples = 1 + (P.PRNG_samples - 1) // P.PRNG_instances
        P.GPU0 = g
        self.prngstep = half
        P.initialize()
        self.header = self.lattice_make_header()

    def lattice_make_header(self) -> str:
        n = self.lattice_full_size
        prec = {0: "single", 1: "double", 2: "mixed"}[self.precision]
        h = (f" GPU SU({self.lattice_group}) simulator {VERSION} (qcdgpu_b200)\n\n"
             f" Monte Carlo simulation of {self.lattice_nd}D SU({self.lattice_group}) LGT\n\n"
             " ***************************************************\n"
             f" Active platform          : CUDA (sm_{self.GPU0.GPU_info.cc_major}{self.GPU0.GPU_info.cc_minor})\n"
             f" Active device            : {self.GPU0.GPU_info.device_name.decode()}\n"
             " ***************************************************\n"
             f" lattice size                : {n[0]:3d} x {n[1]:3d} x {n[2]:3d} x {n[3]:3d}\n"
             f" init                        : {self.INIT}\n"
             f" ints (0=hot, 1=cold, 2=gid) : {self.ints}\n"
             + self.PRNG0.print_generator() +
             f" nav                         : {self.NAV}\n"
             f" niter                       : {self.NITER}\n"
             f" iter (# of samples)         : {self.ITER}\n"
             f" nhit                        : {self.NHIT}\n"
             f" nhitPar                     : {self.NHITPar}\n"
             f" precision                   : {prec}\n"
             " ***************************************************\n"
             f" BETA                        : {self.BETA: 16.13e}\n"
             f" PHI   (lambda_3)            : {self.PHI: 16.13e}\n"
             f" OMEGA (lambda_8)            : {self.OMEGA: 16.13e}\n"
             " ***************************************************\n")
        return h

    # ------------------------------------------------------------------ lattice_create_buffers (suncl.cpp:3877-3991)
    def lattice_create_buffers(self):
        g = self.GPU0
        double = self.precision == model_precision_double
        # NB: the reference allocates double4 tables for anything != single while its kernels treat mixed as
        # single; here mixed == float4 storage, as the kernels define it.
        rdt = np.float64 if double else np.float32
        vec = 32 if double else 16
        self.size_lattice_table = self.lattice_table_size
        self.size_lattice_measurement = max(self.lattice_measurement_size_F,
                                            2 * self.lattice_polyakov_loop_offset if self.PL_level > 1 else 0)
        self.size_lattice_energies = self.lattice_energies_size
        if self.get_actions_diff:  # suncl.cpp:3892-3896
            self.size_lattice_energies *= 2 * getK(self.lattice_domain_n1, self.lattice_domain_size[1], g.GPU_limit_max_workgroup_size or 0)
        self.size_lattice_energies_plq = self.lattice_energies_offset * MODEL_energies_size      # suncl.cpp:3890
        if self.PL_level > 2:      # suncl.cpp:3898-3901
            self.size_lattice_polyakov_loop = (self.lattice_domain_n1 + 1) * self.lattice_polyakov_loop_size * max(self.PLK, 1)
        else:
            self.size_lattice_polyakov_loop = 2 * self.lattice_polyakov_loop_size * max(self.PL_level, 1)
        params = np.zeros(self.lattice_parameters_size, dtype=rdt)
        params[0] = self.BETA / self.lattice_group
        params[1] = self.PHI
        params[2] = self.OMEGA
        params[3] = self.wilson_R          # suncl.cpp:3942-3953
        params[4] = self.wilson_T
        self.lattice_table = g.buffer_init(GPU.buffer_type_IO, self.size_lattice_table, None, vec)
        self.lattice_parameters = g.buffer_init(GPU.buffer_type_Constant, self.lattice_parameters_size, params, params.itemsize)
        self.lattice_measurement = g.buffer_init(GPU.buffer_type_IO, self.size_lattice_measurement, None, 16)
        lds_size = int(g.GPU_info.local_memory_size) // 4 // 4 // 2
        self.lattice_lds = g.buffer_init(GPU.buffer_type_LDS, lds_size, None, 16)
        self.lattice_energies = g.buffer_init(GPU.buffer_type_IO, self.size_lattice_energies, None, 16)
        self.lattice_energies_plq = g.buffer_init(GPU.buffer_type_IO, self.size_lattice_energies_plq, None, 16)
        if self.get_actions_diff:  # suncl.cpp:3973-3978
            self.lattice_action_diff = {a: g.buffer_init(GPU.buffer_type_IO, self.size_lattice_energies, None, 16) for a in "xyz"}
        if self.PL_level > 0:
            self.lattice_polyakov_loop = g.buffer_init(GPU.buffer_type_IO, self.size_lattice_polyakov_loop, None, 16)
        if self.PL_level > 2:
            self.lattice_polyakov_loop_diff = {a: g.buffer_init(GPU.buffer_type_IO, self.size_lattice_polyakov_loop, None, 16) for a in "xyz"}
        if self.get_wilson_loop:
            self.lattice_wilson_loop = g.buffer_init(GPU.buffer_type_IO, self.lattice_energies_size, None, 8)
        if self.INIT == 0:
            self.lattice_load_state()      # the data (suncl.cpp:3956; the reference fills its host arrays before buffer_init copies them)

    # ------------------------------------------------------------------ state files (suncl.cpp:1857-2079)
    @property
    def lattice_energies_size_F(self) -> int:
        return self.lattice_energies_size * (MODEL_energies_size if (self.get_Fmunu or self.get_F0mu) else 1)   # suncl.cpp:2268, 2296-2297

    def _state_sections(self):
        """(buffer id, bytes) of the data sections of a state file, in file order (suncl.cpp:1889-1901)."""
        double = self.precision == model_precision_double
        sec = [(self.lattice_measurement, 16 * self.lattice_measurement_size_F),
               (self.lattice_energies, 16 * self.lattice_energies_size)]
        if self.get_plaquettes_avr or self.get_Fmunu or self.get_F0mu:
            sec.append((self.lattice_energies_plq, 16 * self.lattice_energies_size_F))
        if self.get_wilson_loop:
            sec.append((self.lattice_wilson_loop, 8 * self.lattice_energies_size))
        if self.PL_level > 0:
            sec.append((self.lattice_polyakov_loop, 16 * self.lattice_polyakov_loop_size))
        # the reference writes double4 for anything but single; mixed precision is float4 storage here
        sec.append((self.lattice_table, (32 if double else 16) * self.lattice_table_size))
        return sec

    def lattice_make_bin_header(self) -> np.ndarray:
        """model::lattice_make_bin_header (suncl.cpp:1939-1990): 64 dwords."""
        head = np.zeros(BIN_HEADER_SIZE, dtype=np.uint32)
        for i in range(2):
            head[i] = convert_str_uint("QCDGPU", 4 * i)
        head[2] = convert_str_uint(VERSION, 0)
        lo_hi = lambda x: np.array([x], dtype=np.float64).view(np.uint32)       # convert_to_uint_LOW / _HIGH
        i1, i2 = self.Fmunu_indices()
        P = self.PRNG0
        fields = [self.INIT, self.ints, P.PRNG_randseries, GENERATOR_CODES.get(P.PRNG_generator, 5), P.RL_nskip, P.PRNG_counter,
                  self.NAV, self.NAV_counter, self.NITER, self.ITER, self.ITER_counter, self.NHIT, self.wilson_R, self.wilson_T,
                  {model_precision_single: 1, model_precision_double: 2, model_precision_mixed: 3}[self.precision],
                  *lo_hi(self.BETA), *lo_hi(self.PHI), *lo_hi(self.OMEGA),
                  self.PL_level, i1, i2, int(self.get_Fmunu), int(self.get_actions_avr), int(self.get_plaquettes_avr), int(self.get_wilson_loop),
                  self.lattice_measurement_size_F, self.lattice_energies_size, self.lattice_energies_size_F, self.lattice_polyakov_loop_size,
                  self.lattice_group, self.lattice_nd, *self.lattice_full_size[:self.lattice_nd], *self.lattice_domain_size[:self.lattice_nd]]
        head[4:4 + len(fields)] = np.array([int(f) & 0xFFFFFFFF for f in fields], dtype=np.uint32)
        return head

    def lattice_load_bin_header(self, head: np.ndarray) -> bool:
        """model::lattice_load_bin_header (suncl.cpp:1992-2041): run parameters from a state file; False if it is not one."""
        if [int(head[0]), int(head[1])] != [convert_str_uint("QCDGPU", 0), convert_str_uint("QCDGPU", 4)] or int(head[2]) != convert_str_uint(VERSION, 0):
            return False
        it = iter(int(v) for v in head[5:])            # head[4] (INIT of the writer) is skipped
        dbl = lambda: float(np.array([next(it), next(it)], dtype=np.uint32).view(np.float64)[0])
        P = self.PRNG0
        v = next(it)
        self.ints = v if v in (0, 1, 2) else model_start_hot            # convert_uint_to_start
        P.PRNG_randseries = next(it)
        code = next(it)
        P.PRNG_generator = {v: k for k, v in GENERATOR_CODES.items()}.get(code, "RANLUX3")
        P.RL_nskip = next(it)
        self.PRNG_counter = next(it)
        self.NAV, self.NAV_counter, self.NITER, self.ITER, self.ITER_counter, self.NHIT = (next(it) for _ in range(6))
        self.wilson_R, self.wilson_T = next(it), next(it)
        self.precision = {1: model_precision_single, 2: model_precision_double, 3: model_precision_mixed}.get(next(it), model_precision_single)
        self.BETA, self.PHI, self.OMEGA = dbl(), dbl(), dbl()
        self.PL_level = next(it)
        i1, i2 = next(it), next(it)
        self.get_Fmunu, self.get_actions_avr, self.get_plaquettes_avr, self.get_wilson_loop = (bool(next(it)) for _ in range(4))
        self._loaded_sizes = [next(it) for _ in range(4)]      # recomputed by model_lattice_init; checked when the data is read
        self.lattice_group, self.lattice_nd = next(it), next(it)
        self.lattice_full_size = [next(it) for _ in range(self.lattice_nd)]
        self._loaded_domain = [next(it) for _ in range(self.lattice_nd)]
        if self.lattice_group == 3:
            self.Fmunu_gen1, self.Fmunu_gen2 = (0 if i1 == 3 else i1), (0 if i2 == 8 else i2)
        else:
            self.Fmunu_gen1 = 1 if i1 == 1 else 0
        self.fprefix = f"su{self.lattice_group}-"
        return True

    def lattice_save_state(self, filename: str | None = None) -> str:
        """model::lattice_save_state (suncl.cpp:1857-1937): header, result buffers, link table; the reference's name
        <path><prefix>yy-Mon-dd-hh-mm-ss.qcg unless one is given."""
        g = self.GPU0
        if filename is None:
            filename = os.path.join(self.path, self.fprefix + time.strftime("%y-%b-%d-%H-%M-%S") + ".qcg")
        with open(filename, "wb") as f:
            f.write(self.lattice_make_bin_header().tobytes())
            for bid, nbytes in self._state_sections():
                f.write(g.buffer_map(bid, np.uint8)[:nbytes].tobytes())
        return filename

    def lattice_load_state(self, filename: str | None = None):
        """model::lattice_load_state (suncl.cpp:2043-2079): first call = header (before the sizes are computed), second call = data
        (once the buffers exist)."""
        g = self.GPU0
        if filename is None:
            filename = self.fstate if os.path.isabs(self.fstate) else os.path.join(self.path, self.fstate)
        with open(filename, "rb") as f:
            head = np.frombuffer(f.read(4 * BIN_HEADER_SIZE), dtype=np.uint32)
            if self.LOAD_state == 0:
                if head.size != BIN_HEADER_SIZE or not self.lattice_load_bin_header(head):
                    raise QcdGpuError(f"{filename}: not a QCDGPU {VERSION} state file")
            elif self.LOAD_state == 1:
                sizes = [self.lattice_measurement_size_F, self.lattice_energies_size, self.lattice_energies_size_F, self.lattice_polyakov_loop_size]
                if sizes != self._loaded_sizes:
                    raise QcdGpuError(f"{filename}: buffer sizes {self._loaded_sizes} of the writer differ from {sizes} here "
                                      "(other work-group size or alignment)")
                for bid, nbytes in self._state_sections():
                    data = np.frombuffer(f.read(nbytes), dtype=np.uint8)
                    if data.size != nbytes:
                        raise QcdGpuError(f"{filename}: truncated")
                    cur = g.buffer_map(bid, np.uint8).copy()
                    cur[:nbytes] = data
                    g.buffer_write_from(bid, cur)
        self.LOAD_state += 1

    # ------------------------------------------------------------------ lattice_make_programs (suncl.cpp:3265-3648)
    def options_common(self, measuring: bool = False) -> str:
        """measuring: the option string of the measurement / Polyakov / Wilson programs.  In the BIGLAT build the reference gives those
        N1 = Nx, the real slices (suncl.cpp:2597, 2889), and only the update program N1 = Nx + 2 (suncl.cpp:2422)."""
        n = list(self.lattice_domain_size)
        if self.biglat and measuring:
            n[0] -= 2
        o = "-Werror"
        if self.ints == model_start_gid:
            o += " -D GID_UPD"
        o += f" -D SUN={self.lattice_group} -D ND={self.lattice_nd} -D N1={n[0]} -D N2={n[1]} -D N3={n[2]} -D N4={n[3]}"
        if not (self.PHI == 0.0 and self.OMEGA == 0.0):
            o += " -D TBC"
        o += f" -D ROWSIZE={self.lattice_table_row_size} -D PRECISION={self.precision}"
        if not self.biglat:     # suncl.cpp:3292
            o += f" -D PLK={getK(n[0], n[1], self.GPU0.GPU_limit_max_workgroup_size or 0)}"
        if self.biglat:   # suncl.cpp:2427-2441
            n234 = n[1] * n[2] * n[3]
            o += f" -D BIGLAT -D FULL_SITES={self.lattice_full_site} -D LEFT_SITES={self.left_x * n234}"
        return o

    def lattice_make_programs(self):
        g = self.GPU0
        oc = self.options_common()
        options = oc + f" -D NHIT={self.NHIT} -D NHITPar={self.NHITPar} -D PRNGSTEP={self.lattice_table_row_size_half}"
        self.program_update = g.program_create(None, options)
        init_gs = [self.lattice_table_exact_row_size]
        init_hot_gs = [self.lattice_group_elements * self.lattice_table_row_size]
        monte_gs = [self.lattice_table_exact_row_size_half]
        wg = int(g.GPU_info.max_workgroup_size)
        if self.ints == model_start_gid:
            self.sun_init_id = g.kernel_init("lattice_init_gid", 1, init_gs, None)
            g.kernel_init_buffer(self.sun_init_id, self.lattice_table)
        elif self.ints == model_start_cold:
            self.sun_init_id = g.kernel_init("lattice_init_cold", 1, init_gs, None)
            g.kernel_init_buffer(self.sun_init_id, self.lattice_table)
        else:
            self.sun_init_hot_id = []
            for d in DIRS:
                k = g.kernel_init(f"lattice_init_hot_{d}", 1, init_hot_gs, None)
                g.kernel_init_buffer(k, self.lattice_table)
                g.kernel_init_buffer(k, self.PRNG0.PRNG_randoms_id)
                self.sun_init_hot_id.append(k)
        self.sun_GramSchmidt_id = g.kernel_init("lattice_GramSchmidt", 1, init_gs, None)
        g.kernel_init_buffer(self.sun_GramSchmidt_id, self.lattice_table)
        g.kernel_init_buffer(self.sun_GramSchmidt_id, self.lattice_parameters)
        self.sun_update_id = {}
        self.sun_overrelax_id = {}
        for par in ("odd", "even"):
            for d in DIRS:
                k = g.kernel_init(f"update_{par}_{d}", 1, monte_gs, None)
                g.kernel_init_buffer(k, self.lattice_table)
                g.kernel_init_buffer(k, self.lattice_parameters)
                g.kernel_init_buffer(k, self.PRNG0.PRNG_randoms_id)
                self.sun_update_id[(par, d)] = k
                if self.or_steps:
                    k = g.kernel_init(f"overrelax_{par}_{d}", 1, monte_gs, None)
                    g.kernel_init_buffer(k, self.lattice_table)
                    g.kernel_init_buffer(k, self.lattice_parameters)
                    self.sun_overrelax_id[(par, d)] = k
        # measurements
        om = oc = self.options_common(measuring=True)              # suncl.cpp:3398-3408; from here on the measuring programs
        if self.get_Fmunu: om += " -D FMUNU"
        if self.get_F0mu: om += " -D F0MU"
        if self.get_Fmunu or self.get_F0mu:
            if self.Fmunu_gen1: om += f" -D FMUNU{self.Fmunu_gen1}"
            if self.Fmunu_gen2: om += f" -D FMUNU{self.Fmunu_gen2}"
        self.program_measurements = g.program_create(None, om)
        meas_gs = [self.lattice_action_size]
        self.sun_clear_measurement_id = g.kernel_init("clear_measurement", 1, [self.size_lattice_measurement], None)
        g.kernel_init_buffer(self.sun_clear_measurement_id, self.lattice_measurement)
        self.sun_measurement_id = g.kernel_init("lattice_measurement", 1, meas_gs, [wg])
        for b in (self.lattice_table, self.lattice_measurement, self.lattice_parameters, self.lattice_lds):
            g.kernel_init_buffer(self.sun_measurement_id, b)
        size_reduce = int(math.ceil(self.lattice_table_exact_row_size / g.kernel_get_worksize(self.sun_measurement_id)))
        self.sun_measurement_reduce_id = g.kernel_init("reduce_measurement_double2", 1, [wg], [wg])
        for b in (self.lattice_measurement, self.lattice_energies, self.lattice_lds):
            g.kernel_init_buffer(self.sun_measurement_reduce_id, b)
        self.argument_measurement_index = g.kernel_init_constant(self.sun_measurement_reduce_id, size_reduce)
        self.sun_measurement_plq_id = g.kernel_init("lattice_measurement_plq", 1, meas_gs, [wg])
        ws = g.kernel_get_worksize(self.sun_measurement_plq_id)
        offset_plq = g.buffer_size_align(int(math.ceil(self.lattice_table_exact_row_size / ws)), ws)
        for b in (self.lattice_table, self.lattice_measurement, self.lattice_parameters, self.lattice_lds):
            g.kernel_init_buffer(self.sun_measurement_plq_id, b)
        g.kernel_init_constant(self.sun_measurement_plq_id, offset_plq)
        size_plq = int(math.ceil(self.lattice_table_exact_row_size / ws))
        self.sun_measurement_plq_reduce_id = g.kernel_init("reduce_measurement_plq_double2", 1, [wg], [wg])
        for b in (self.lattice_measurement, self.lattice_energies_plq, self.lattice_lds):
            g.kernel_init_buffer(self.sun_measurement_plq_reduce_id, b)
        self.argument_plq_index = g.kernel_init_constant(self.sun_measurement_plq_reduce_id,
                                                         [size_plq, offset_plq, self.lattice_energies_offset, 0])
        # Wilson loop (suncl.cpp:3518-3546)
        if self.get_wilson_loop:
            self.program_wilson = g.program_create(None, oc)
            self.sun_measurement_wilson_id = g.kernel_init("lattice_measurement_wilson", 1, meas_gs, [wg])
            for b in (self.lattice_table, self.lattice_measurement, self.lattice_parameters, self.lattice_lds):
                g.kernel_init_buffer(self.sun_measurement_wilson_id, b)
            size_w = int(math.ceil(self.lattice_domain_exact_site / g.kernel_get_worksize(self.sun_measurement_wilson_id)))
            self.sun_wilson_loop_reduce_id = g.kernel_init("reduce_wilson_double2", 1, [wg], [wg])
            for b in (self.lattice_measurement, self.lattice_wilson_loop, self.lattice_lds):
                g.kernel_init_buffer(self.sun_wilson_loop_reduce_id, b)
            self.argument_wilson_index = g.kernel_init_constant(self.sun_wilson_loop_reduce_id, size_w)
        # Polyakov loop
        if self.PL_level > 0:
            self.program_polyakov = g.program_create(None, oc + f" -D PL={self.PL_level}")
            self.sun_polyakov_id = g.kernel_init("lattice_polyakov", 1, [self.lattice_polyakov_size], [wg])
            ws = g.kernel_get_worksize(self.sun_polyakov_id)
            offset_p = g.buffer_size_align(int(math.ceil(self.lattice_polyakov_size / ws)), ws)
            for b in (self.lattice_table, self.lattice_measurement, self.lattice_parameters, self.lattice_lds):
                g.kernel_init_buffer(self.sun_polyakov_id, b)
            g.kernel_init_constant(self.sun_polyakov_id, offset_p)
            size_p = int(math.ceil(self.lattice_polyakov_size / ws))
            self.sun_polyakov_reduce_id = g.kernel_init("reduce_polyakov_double2", 1, [wg], [wg])
            for b in (self.lattice_measurement, self.lattice_polyakov_loop, self.lattice_lds):
                g.kernel_init_buffer(self.sun_polyakov_reduce_id, b)
            self.argument_polyakov_index = g.kernel_init_constant(
                self.sun_polyakov_reduce_id, [size_p, offset_p, self.lattice_polyakov_loop_size, 0])
        if self.get_actions_diff:  # suncl.cpp:3440-3490
            self.sun_action_diff_id, self.sun_action_diff_reduce_id, self.argument_action_diff_index = {}, {}, {}
            for a in "xyz":
                k = g.kernel_init(f"lattice_action_diff_{a}", 1, [self.lattice_action_size], [wg])
                for b in (self.lattice_table, self.lattice_measurement, self.lattice_parameters, self.lattice_lds):
                    g.kernel_init_buffer(k, b)
                r = g.kernel_init(f"reduce_action_diff_{a}_double2", 1, [wg], [wg])
                for b in (self.lattice_measurement, self.lattice_action_diff[a], self.lattice_lds):
                    g.kernel_init_buffer(r, b)
                sparam = [int(math.ceil(self.lattice_action_size / g.kernel_get_worksize(self.sun_measurement_id))), self.lattice_energies_size, 0, 0]
                self.argument_action_diff_index[a] = g.kernel_init_constant(r, sparam)
                self.sun_action_diff_id[a], self.sun_action_diff_reduce_id[a] = k, r
        if self.PL_level > 2:      # suncl.cpp:3590-3644
            self.sun_polyakov_diff_id, self.sun_polyakov_diff_reduce_id, self.argument_polyakov_diff_index = {}, {}, {}
            for a in "xyz":
                k = g.kernel_init(f"lattice_polyakov_diff_{a}", 1, [self.lattice_polyakov_size], [wg])
                ws = g.kernel_get_worksize(k)
                offset_d = g.buffer_size_align(int(math.ceil(self.lattice_polyakov_size / ws)), ws)
                for b in (self.lattice_table, self.lattice_measurement, self.lattice_parameters, self.lattice_lds):
                    g.kernel_init_buffer(k, b)
                g.kernel_init_constant(k, offset_d)
                size_d = int(math.ceil(self.lattice_polyakov_size / ws))
                r = g.kernel_init(f"reduce_polyakov_diff_{a}_double2", 1, [wg], [wg])
                for b in (self.lattice_measurement, self.lattice_polyakov_loop_diff[a], self.lattice_lds):
                    g.kernel_init_buffer(r, b)
                self.argument_polyakov_diff_index[a] = g.kernel_init_constant(r, [size_d, offset_d, self.lattice_polyakov_loop_size, 0])
                self.sun_polyakov_diff_id[a], self.sun_polyakov_diff_reduce_id[a] = k, r

    # ------------------------------------------------------------------ run loop (suncl.cpp:5212-5505)
    def _measure(self):
        """Measurements of the current configuration into slot ITER_counter, in the reference's order (suncl.cpp:5416-5482)."""
        g = self.GPU0
        if self.get_wilson_loop:
            g.kernel_run(self.sun_measurement_wilson_id)
            g.kernel_init_constant_reset(self.sun_wilson_loop_reduce_id, self.ITER_counter, self.argument_wilson_index)
            g.kernel_run(self.sun_wilson_loop_reduce_id)
        if self.get_actions_avr:
            g.kernel_run(self.sun_measurement_id)
            g.kernel_init_constant_reset(self.sun_measurement_reduce_id, self.ITER_counter, self.argument_measurement_index)
            g.kernel_run(self.sun_measurement_reduce_id)
        if self.get_actions_diff:  # suncl.cpp:5284-5310
            for a in "xyz":
                g.kernel_run(self.sun_action_diff_id[a])
                g.kernel_init_constant_reset(self.sun_action_diff_reduce_id[a], self.ITER_counter, self.argument_action_diff_index[a])
                g.kernel_run(self.sun_action_diff_reduce_id[a])
        if self.PL_level > 0:
            g.kernel_run(self.sun_clear_measurement_id)
            g.kernel_run(self.sun_polyakov_id)
            g.kernel_init_constant_reset(self.sun_polyakov_reduce_id, self.ITER_counter, self.argument_polyakov_index)
            g.kernel_run(self.sun_polyakov_reduce_id)
        if self.PL_level > 2:
            g.kernel_run(self.sun_clear_measurement_id)
            for a in "xyz":
                g.kernel_run(self.sun_polyakov_diff_id[a])
                g.kernel_init_constant_reset(self.sun_polyakov_diff_reduce_id[a], self.ITER_counter, self.argument_polyakov_diff_index[a])
                g.kernel_run(self.sun_polyakov_diff_reduce_id[a])
        if self.get_plaquettes_avr or self.get_Fmunu or self.get_F0mu:
            g.kernel_run(self.sun_clear_measurement_id)
            g.kernel_run(self.sun_measurement_plq_id)
            g.kernel_init_constant_reset(self.sun_measurement_plq_reduce_id, self.ITER_counter, self.argument_plq_index)
            g.kernel_run(self.sun_measurement_plq_reduce_id)

    def lattice_sweep(self):
        """One sweep: 8 x (produce, update) + reunitarisation  (suncl.cpp:5357-5377)."""
        g, P = self.GPU0, self.PRNG0
        for par in ("odd", "even"):
            for d in DIRS:
                if not self.turnoff_prns:
                    P.produce()
                if not self.turnoff_updates:
                    g.kernel_run(self.sun_update_id[(par, d)])
        for _ in range(self.or_steps):
            for par in ("odd", "even"):
                for d in DIRS:
                    g.kernel_run(self.sun_overrelax_id[(par, d)])
        if not self.turnoff_gramschmidt:
            g.kernel_run(self.sun_GramSchmidt_id)

    def lattice_start(self):
        """Buffers, programs and the initial configuration + its measurement (slot 0)."""
        g, P = self.GPU0, self.PRNG0
        self.lattice_create_buffers()
        self.lattice_make_programs()
        if self.ints == model_start_hot:
            P.produce()
        if self.INIT == 0:
            # continue from a state file (suncl.cpp:5349-5354): the generators were re-seeded from RANDSERIES; replay the produce() calls
            # of the interrupted run.  Back-to-back produce() calls nobody consumes cost one jump per generator in fused mode.
            while self.PRNG_counter > P.PRNG_counter:
                P.produce()
            return
        if self.ints == model_start_hot:
            for i, k in enumerate(self.sun_init_hot_id):
                g.kernel_run(k)
                if i < 3 and not self.turnoff_prns:
                    P.produce()
        else:
            g.kernel_run(self.sun_init_id)
        self._measure()
        self.ITER_counter += 1
        if not self.turnoff_config_save:
            self.lattice_save_state()      # suncl.cpp:5342

    def _timed_save(self):
        """suncl.cpp:5383-5387, 5488-5492: a state file every write_lattice_state_every_secs seconds"""
        if not self.turnoff_config_save and time.time() - self._last_save > self.write_lattice_state_every_secs:
            self.lattice_save_state()
            self._last_save = time.time()

    def lattice_simulate(self):
        g = self.GPU0
        t0 = self._last_save = time.time()
        self.lattice_start()
        for _ in range(self.NAV_counter, self.NAV):
            self.lattice_sweep()
            self.NAV_counter += 1
            self._timed_save()
        for _ in range(self.ITER_counter, self.ITER):
            for _ in range(self.NITER):
                self.lattice_sweep()
            self._measure()
            self.ITER_counter += 1
            self._timed_save()
        g.synchronize()
        self.elapsed = time.time() - t0
        if not self.turnoff_config_save:
            self.lattice_save_state()      # suncl.cpp:5498-5500

    # ------------------------------------------------------------------ analysis (suncl.cpp:1328-1675, data_analysis.cpp:107-135)
    @staticmethod
    def _series(values: np.ndarray, denominator: float) -> dict:
        data = values / denominator
        last = len(data) - 1
        mean = float(data[1:].sum() / last) if last > 0 else 0.0
        var = float(((data[1:] - mean) ** 2).sum() / last) if last > 0 else 0.0
        return {"data": data, "mean_value": mean, "variance": var, "GPU_last_value": float(data[last])}

    @staticmethod
    def _joint(a: dict, b: dict) -> dict:
        data = 0.5 * (a["data"] + b["data"])
        last = len(data) - 1
        mean = float(data[1:].sum() / last) if last > 0 else 0.0
        var = float(((data[1:] - mean) ** 2).sum() / last) if last > 0 else 0.0
        return {"data": data, "mean_value": mean, "variance": var, "GPU_last_value": float(data[last])}

    def lattice_analysis(self):
        g = self.GPU0
        A = self.Analysis
        den = float(self.lattice_full_site * 3)
        if self.get_actions_avr:
            e = g.buffer_map(self.lattice_energies, np.float64).reshape(-1, 2)[: self.ITER]
            A["S_spat"] = self._series(e[:, 0].copy(), den)
            A["S_temp"] = self._series(e[:, 1].copy(), den)
            A["S_total"] = self._joint(A["S_spat"], A["S_temp"])
        if self.get_plaquettes_avr:
            e = g.buffer_map(self.lattice_energies_plq, np.float64).reshape(-1, 2)[: self.ITER]
            A["Plq_spat"] = self._series(e[:, 0].copy(), den)
            A["Plq_temp"] = self._series(e[:, 1].copy(), den)
            A["Plq_total"] = self._joint(A["Plq_spat"], A["Plq_temp"])
        if self.PL_level > 0:
            e = g.buffer_map(self.lattice_polyakov_loop, np.float64).reshape(-1, 2)
            pden = float(self.lattice_full_n1n2n3 * self.lattice_group)
            A["Polyakov_loop"] = self._series(e[: self.ITER, 0].copy(), pden)
            A["Polyakov_loop_im"] = self._series(e[: self.ITER, 1].copy(), pden)
            if self.PL_level > 1:
                o = self.lattice_polyakov_loop_size
                A["Polyakov_loop_P2"] = self._series(e[o: o + self.ITER, 0].copy(), float(self.lattice_full_n1n2n3 * self.lattice_group ** 2))
                A["Polyakov_loop_P4"] = self._series(e[o: o + self.ITER, 1].copy(), float(self.lattice_full_n1n2n3 * self.lattice_group ** 4))
        if self.PL_level > 2:      # suncl.cpp:1500-1560: per-slice series, result r of slice i at [r + i * lattice_polyakov_loop_size]
            # the reference divides every axis by full_n2n3 * group (suncl.cpp:1516, 1544): exact for cubic spatial volumes
            pden_slice = lambda n: float(self.lattice_full_size[1] * self.lattice_full_size[2] * self.lattice_group)
            for ax, a in enumerate("xyz"):
                n = self.lattice_full_size[ax]
                e = g.buffer_map(self.lattice_polyakov_loop_diff[a], np.float64).reshape(-1, 2)
                o = self.lattice_polyakov_loop_size
                A[f"PL_{a.upper()}"] = [self._series(e[i * o: i * o + self.ITER, 0].copy(), pden_slice(n)) for i in range(n)]
                A[f"PL_{a.upper()}_im"] = [self._series(e[i * o: i * o + self.ITER, 1].copy(), pden_slice(n)) for i in range(n)]
        if self.get_actions_diff:  # suncl.cpp:1376-1481: slice i of an axis at [i * lattice_energies_size + iteration], / (transverse volume * 3)
            full = self.lattice_full_size
            vol = full[0] * full[1] * full[2] * full[3]
            for ax, a in enumerate("xyz"):
                e = g.buffer_map(self.lattice_action_diff[a], np.float64).reshape(-1, 2)
                o, den_a = self.lattice_energies_size, float(vol // full[ax] * 3)
                s_ = [self._series(e[i * o: i * o + self.ITER, 0].copy(), den_a) for i in range(full[ax])]
                t_ = [self._series(e[i * o: i * o + self.ITER, 1].copy(), den_a) for i in range(full[ax])]
                A[f"S_diff_{a.upper()}_s"], A[f"S_diff_{a.upper()}_t"] = s_, t_
                A[f"S_diff_{a.upper()}"] = [self._joint(x, y) for x, y in zip(s_, t_)]
        if self.get_wilson_loop:   # suncl.cpp:1606-1612
            w = g.buffer_map(self.lattice_wilson_loop, np.float64)[: self.ITER]
            A["Wilson_loop"] = self._series(w.copy(), den)
        if self.get_Fmunu or self.get_F0mu:     # suncl.cpp:1613-1672
            e = g.buffer_map(self.lattice_energies_plq, np.float64).reshape(-1, 2)
            o = self.lattice_energies_offset
            i1, i2 = self.Fmunu_indices()
            planes = ("xy", "xz", "yz") if self.get_Fmunu else ("xt", "yt", "zt")
            for q, pl in enumerate(planes):
                for c, reim in enumerate(("re", "im")):
                    A[f"Fmunu_{pl}_{i1}_{reim}"] = self._series(e[(1 + q) * o: (1 + q) * o + self.ITER, c].copy(), float(self.lattice_full_site))
                    A[f"Fmunu_{pl}_{i2}_{reim}"] = self._series(e[(4 + q) * o: (4 + q) * o + self.ITER, c].copy(), float(self.lattice_full_site))
        return A

    def Fmunu_indices(self) -> tuple[int, int]:
        """Generator numbers used in the data names (model::Fmunu_defaults, suncl.cpp:2094-2127)."""
        if self.lattice_group == 2:
            return (1 if self.Fmunu_gen1 == 1 else 3), 2
        return (self.Fmunu_gen1 or 3), (self.Fmunu_gen2 or 8)

    # lattice_write_results (suncl.cpp:1679-1854)
    def columns(self) -> list[str]:
        cols = []
        if self.get_plaquettes_avr: cols += ["Plq_spat", "Plq_temp", "Plq_total"]
        if self.get_wilson_loop: cols += ["Wilson_loop"]
        if self.get_actions_avr: cols += ["S_spat", "S_temp", "S_total"]
        if self.PL_level > 0: cols += ["Polyakov_loop", "Polyakov_loop_im"]
        if self.PL_level > 1: cols += ["Polyakov_loop_P2", "Polyakov_loop_P4"]
        if self.get_Fmunu or self.get_F0mu:
            i1, i2 = self.Fmunu_indices()
            planes = ("xy", "xz", "yz") if self.get_Fmunu else ("xt", "yt", "zt")
            for i in (i1, i2):
                for pl in planes:
                    cols += [f"Fmunu_{pl}_{i}_re", f"Fmunu_{pl}_{i}_im"]
        return cols

    def lattice_write_results(self, filename: str | None = None) -> str:
        A = self.Analysis
        if filename is None:
            filename = os.path.join(self.path, self.fprefix + time.strftime("%y-%b-%d-%H-%M-%S") + ".txt")
        cols = self.columns()
        with open(filename, "w") as f:
            f.write(self.header)
            f.write(f" Elapsed time             : {self.elapsed:.3f} s\n")
            f.write(" ***************************************************\n")
            for c in ("S_total", "Plq_total", "Polyakov_loop", "Polyakov_loop_im", "Polyakov_loop_P2", "Polyakov_loop_P4"):
                if c in A:
                    f.write(f" Mean {c:20s}: {A[c]['mean_value']: 16.13e}\n")
                    f.write(f" Variance {c:16s}: {A[c]['variance']: 16.13e}\n")
            if self.get_actions_diff:      # suncl.cpp:1753-1791
                f.write(" ***************************************************\n")
                f.write("Differentiated S data (#, S_total, S_variance):\n")
                for ax, a in enumerate("XYZ"):
                    n = self.lattice_full_size[ax]
                    f.write(f"\n{a}              " + "".join(f"{i:2d}                   " for i in range(n)))
                    f.write("\nS             " + "".join(f"{A[f'S_diff_{a}'][i]['mean_value']: 16.13e} " for i in range(n)))
                    f.write("\nS_variance    " + "".join(f"{A[f'S_diff_{a}'][i]['variance']: 16.13e} " for i in range(n)) + "\n")
            f.write(" ***************************************************\n")
            f.write(" Data fields:\n")
            f.write("    #,  " + "".join(f"{c:<21s} " for c in cols) + "\n")
            f.write(" ***************************************************\n")
            for i in range(self.ITER):
                f.write(f"{i:5d}" + "".join(f" {A[c]['data'][i]: 16.13e}" for c in cols) + "\n")
        return filename

    # ------------------------------------------------------------------ helpers for tests / tools
    def table_numpy(self) -> np.ndarray:
        """Link table as (planes, ROWSIZE, 4) in the storage precision (a copy)."""
        rdt = np.float64 if self.precision == model_precision_double else np.float32
        t = self.GPU0.buffer_map(self.lattice_table, rdt)
        planes = self.lattice_group_elements
        return t[: planes * self.lattice_table_row_size * 4].reshape(planes, self.lattice_table_row_size, 4).copy()

    def set_table(self, table: np.ndarray):
        rdt = np.float64 if self.precision == model_precision_double else np.float32
        full = self.GPU0.buffer_map(self.lattice_table, rdt).copy()
        flat = np.ascontiguousarray(table, dtype=rdt).reshape(-1)
        full[: flat.size] = flat
        self.GPU0.buffer_write_from(self.lattice_table, full)

    def close(self):
        self.GPU0.close()


def main(argv=None):
    """QCDGPU.cpp main(): defaults, optional init.dat, init -> simulate -> analysis -> write."""
    import sys
    argv = sys.argv[1:] if argv is None else argv
    m = Model()
    if argv:
        m.lattice_get_init_file(argv[0])
    t0 = time.time()
    m.lattice_init()
    print(m.header)
    m.lattice_simulate()
    m.lattice_analysis()
    out = m.lattice_write_results()
    cols = m.columns()
    print(" #     " + " ".join(f"{c:>14s}" for c in cols))
    for i in range(m.ITER):
        print(f"[{i:5d}]: " + " ".join(f"{m.Analysis[c]['data'][i]:14.8f}" for c in cols))
    print(f"\nresults: {out}\nEXECUTION TIME: {time.time() - t0:f} sec\n")
    m.close()


if __name__ == "__main__":
    main()
