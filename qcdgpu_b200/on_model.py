"""Host run loop of the O(N) scalar branch: mirror of the reference's O(N)/suncl/suncl.{h,cpp} `model`
(MODEL_ON == 1, one lattice part) driven by O(N)/QCDGPU.cpp and O(N)/suncl/biglattice.cpp.

Same run-parameter names and init.dat keys (`SKGROUP ON_Z ON_LAMBDA ON_ZETA ON_ETA ON_B`, O(N)/suncl/suncl.cpp:249-300),
same sizes (model_lattice_init, :1241-1345), same option string, kernel names and argument order
(lattice_make_programs, :1454-1610) and the same call order (lattice_simulate / lattice_update / lattice_measure,
:2084-2330), talking to the device only through the GPU_CL::GPU-shaped C ABI (capi.GPU).  New: `fused_prng`
(PRNG_instances = PRNGSTEP, XOR128 / PM / MRG32k3a state kept in registers of the update kernel).

Not here: several lattice parts (BL red/green scheduling), boundary extraction, the .qcg / .cnf files.
"""
from __future__ import annotations

import math
import time

import numpy as np

from .capi import GPU, QcdGpuError
from .model import get_init_file, model_precision_double, model_precision_single, model_start_cold, model_start_hot
from .prng import PRNG

MODEL_parameter_size = 6      # O(N)/suncl/suncl.h:223 (aligned to the work-group size, so [16..22] exist)


class OnModel:
    """model_CL::model of the O(N) tree for one device and one part."""

    def __init__(self):
        # defaults of O(N)/QCDGPU.cpp:104-160
        self.GPU0 = GPU()
        self.PRNG0 = PRNG()
        self.PRNG0.PRNG_generator = "RANLUX3"
        self.PRNG0.PRNG_randseries = 7
        self.desired_platform = self.desired_device = 0
        self.lattice_nd = 4
        self.lattice_group = 1
        self.lattice_full_size = [18, 18, 18, 4]
        self.ON_SK_group = 4
        self.ON_z = 0.1561573990981998
        self.ON_lambda = 0.25
        self.ON_zeta = 0.25
        self.ON_eta = 0.1
        self.ON_b = 9.661840260273483
        self.correlator = [4, 4, 4, 4]
        self.get_correlators = True
        self.get_actions_avr = True
        self.get_plaquettes_avr = True
        self.get_acceptance_rate = True
        self.precision = model_precision_double
        self.ints = model_start_cold
        self.NAV, self.ITER, self.NITER, self.NHIT = 50, 200, 10, 10
        self.turnoff_updates = self.turnoff_prns = False
        self.fused_prng = False
        self.async_run = True
        self.NAV_counter = self.ITER_counter = 0
        self.Analysis: dict[str, dict] = {}
        self.elapsed = 0.0

    # ------------------------------------------------------------------ parameters_setup (O(N)/suncl/suncl.cpp:249-300)
    def lattice_get_init_file(self, path: str):
        for key, ival, fval, text in get_init_file(path):
            if key == "PLATFORM": self.desired_platform = ival
            elif key == "DEVICE": self.desired_device = ival
            elif key == "GROUP": self.lattice_group = ival
            elif key == "ND": self.lattice_nd = ival
            elif key in ("L1", "N1"): self.lattice_full_size[0] = ival
            elif key in ("L2", "N2"): self.lattice_full_size[1] = ival
            elif key in ("L3", "N3"): self.lattice_full_size[2] = ival
            elif key in ("L4", "N4"): self.lattice_full_size[3] = ival
            elif key in ("LS", "NS"): self.lattice_full_size[0:3] = [ival] * 3
            elif key in ("LT", "NT"): self.lattice_full_size[3] = ival
            elif key == "SKGROUP": self.ON_SK_group = ival
            elif key == "ITER": self.ITER = ival
            elif key == "NITER": self.NITER = ival
            elif key == "NHIT": self.NHIT = ival
            elif key == "NAV": self.NAV = ival
            elif key == "ON_Z": self.ON_z = fval
            elif key == "ON_ZETA": self.ON_zeta = fval
            elif key == "ON_ETA": self.ON_eta = fval
            elif key == "ON_B": self.ON_b = fval
            elif key == "ON_LAMBDA": self.ON_lambda = fval
            elif key == "INTS": self.ints = ival
            elif key == "GETRETRACE": self.get_plaquettes_avr = True
            self.PRNG0.parameters_setup(key, ival, text)

    # ------------------------------------------------------------------ model_lattice_init (O(N)/suncl/suncl.cpp:1241-1345)
    def on_parameters(self) -> tuple[float, float]:
        szl = 0.25 * math.sqrt(self.ON_z / (self.ON_lambda * self.ON_zeta))
        if self.ON_eta > 0.0:
            r = math.sqrt(1.0 + 4.0 * math.sqrt(2.0) * self.ON_eta / math.sqrt(self.ON_b))
            max_u = 1.0 - math.exp((1.0 - r) / self.ON_eta)
            if max_u < 0.0:
                max_u = 1.0 - math.exp((1.0 + r) / self.ON_eta)
        else:
            max_u = 1.0 - math.exp(-1.9999999999 / math.sqrt(self.ON_b))
        return szl, max_u

    def lattice_init(self):
        g = self.GPU0
        g.device_select(self.desired_platform, self.desired_device)
        g.device_initialize()
        g.set_async(self.async_run)
        g.set_fused_prng(self.fused_prng)
        if self.lattice_group != 1 or self.lattice_nd != 4:
            raise QcdGpuError("the O(N) branch stores one real per site (GROUP = 1) on a 4-d lattice")
        n = self.lattice_full_size
        if n[1] % 2:
            raise QcdGpuError("N2 must be even for the checkerboard")
        self.ON_sqrt_z_lambda_zeta, self.ON_max_U = self.on_parameters()
        self.lattice_full_site = n[0] * n[1] * n[2] * n[3]
        al = g.buffer_size_align
        local_size_intel = g.GPU_limit_max_workgroup_size or 0
        workgroup_factor = local_size_intel if local_size_intel else 32
        self.local_size_intel = local_size_intel
        self.lattice_parameters_size = al(MODEL_parameter_size)
        if self.lattice_parameters_size < 23:
            raise QcdGpuError("work-group size below 23: lattice_parameters[16..22] would not exist")
        self.lattice_energies_size = al(self.ITER)
        self.lattice_table_row_size = al(self.lattice_full_site)
        self.lattice_table_row_size_half = al(self.lattice_full_site // 2)
        self.lattice_table_exact_row_size = al(self.lattice_full_site)
        self.lattice_table_exact_row_size_half = al(self.lattice_full_site // 2)
        self.lattice_table_size = al(self.lattice_table_row_size)
        self.lattice_measurement_size = al(int(math.ceil(self.lattice_table_exact_row_size / workgroup_factor)))
        # PRNG preparation (:1334-1345): 2 * ceil(NHIT / 4) vec4 per update thread
        P = self.PRNG0
        half = self.lattice_table_row_size_half
        P.PRNG_instances = half if self.fused_prng else 0
        P.PRNG_samples = al(int(math.ceil(half * 2 * math.ceil(0.25 * self.NHIT))))
        if P.PRNG_instances:
            P.PRNG_samples = 1 + (P.PRNG_samples - 1) // P.PRNG_instances
        P.GPU0 = g
        self.prngstep = half
        P.initialize()

    # ------------------------------------------------------------------ lattice_create_buffers (:1921-2082)
    def lattice_create_buffers(self):
        g = self.GPU0
        double = self.precision == model_precision_double
        rdt = np.float64 if double else np.float32
        params = np.zeros(self.lattice_parameters_size, dtype=rdt)
        params[16] = self.ON_z
        params[17] = self.ON_lambda
        params[18] = self.ON_zeta
        params[19] = self.ON_eta
        params[20] = self.ON_b
        params[21] = self.ON_sqrt_z_lambda_zeta
        params[22] = self.ON_max_U
        self.lattice_table = g.buffer_init(GPU.buffer_type_IO, self.lattice_table_size, None, params.itemsize)
        self.lattice_parameters = g.buffer_init(GPU.buffer_type_Constant, self.lattice_parameters_size, params, params.itemsize)
        self.lattice_measurement = g.buffer_init(GPU.buffer_type_IO, self.lattice_measurement_size, None, 16)
        lds_size = int(g.GPU_info.local_memory_size) // 4 // 4 // 2
        self.lattice_lds = g.buffer_init(GPU.buffer_type_LDS, lds_size, None, 16)
        self.lattice_energies = g.buffer_init(GPU.buffer_type_IO, self.lattice_energies_size, None, 16)
        self.lattice_energies_plq = g.buffer_init(GPU.buffer_type_IO, self.lattice_energies_size, None, 16)
        self.lattice_acceptance_rate = g.buffer_init(GPU.buffer_type_IO, self.lattice_energies_size, None, 16)
        self.lattice_correlators = g.buffer_init(GPU.buffer_type_IO, self.lattice_energies_size, None, 16)

    # ------------------------------------------------------------------ lattice_make_programs (:1454-1610)
    def options_common(self) -> str:
        n = self.lattice_full_size
        o = "-Werror -D ON_MODEL"
        if self.get_acceptance_rate:
            o += " -D ACC_RATE"
        o += f" -D SUN={self.lattice_group} -D ND={self.lattice_nd} -D N1EXACT={n[0]} -D N1={n[0]} -D N2={n[1]} -D N3={n[2]} -D N4={n[3]}"
        o += f" -D SKGROUP={self.ON_SK_group} -D ROWSIZE={self.lattice_table_row_size} -D PRECISION={self.precision}"
        return o

    def lattice_make_programs(self):
        g, P = self.GPU0, self.PRNG0
        common = self.options_common()
        self.program_update = g.program_create(None, common + f" -D NHIT={self.NHIT} -D PRNGSTEP={self.lattice_table_row_size_half}")
        full, half = [self.lattice_table_exact_row_size], [self.lattice_table_exact_row_size_half]
        if self.ints == model_start_cold:
            self.sun_init_id = g.kernel_init("lattice_init_cold", 1, full, None)
            g.kernel_init_buffer(self.sun_init_id, self.lattice_table)
        else:
            self.sun_init_id = g.kernel_init("lattice_init_hot", 1, full, None)
            g.kernel_init_buffer(self.sun_init_id, self.lattice_table)
            g.kernel_init_buffer(self.sun_init_id, P.PRNG_randoms_id)
            g.kernel_init_buffer(self.sun_init_id, self.lattice_parameters)
        self.sun_update_id = {}
        sizes = []
        for par in ("odd", "even"):
            k = g.kernel_init(f"update_{par}", 1, half, None)
            g.kernel_init_buffer(k, self.lattice_table)
            g.kernel_init_buffer(k, self.lattice_parameters)
            g.kernel_init_buffer(k, P.PRNG_randoms_id)
            if self.get_acceptance_rate:
                g.kernel_init_buffer(k, self.lattice_lds)
                g.kernel_init_buffer(k, self.lattice_measurement)
                sizes.append(int(math.ceil(half[0] / g.kernel_get_worksize(k))))
            self.sun_update_id[par] = k
        reduce_global = [int(g.GPU_info.max_workgroup_size)]
        if self.get_acceptance_rate:
            k = g.kernel_init("reduce_acceptance_rate", 1, reduce_global, None)
            g.kernel_init_buffer(k, self.lattice_measurement)
            g.kernel_init_buffer(k, self.lattice_acceptance_rate)
            g.kernel_init_buffer(k, self.lattice_lds)
            self.argument_acceptance_rate_index = g.kernel_init_constant(k, max(sizes))
            self.sun_reduce_acceptance_rate_id = k
        # measurements
        self.program_measurements = g.program_create(None, common)
        local = [self.local_size_intel] if self.local_size_intel else None
        k = g.kernel_init("clear_measurement", 1, [self.lattice_measurement_size], None)
        g.kernel_init_buffer(k, self.lattice_measurement)
        self.sun_clear_measurement_id = k

        def measurement(name, result_buffer, reducer, extra=None):
            k = g.kernel_init(name, 1, full, local)
            g.kernel_init_buffer(k, self.lattice_table)
            g.kernel_init_buffer(k, self.lattice_measurement)
            g.kernel_init_buffer(k, self.lattice_parameters)
            g.kernel_init_buffer(k, self.lattice_lds)
            if extra is not None:
                g.kernel_init_constant(k, extra)
            size = int(math.ceil(full[0] / g.kernel_get_worksize(k)))
            r = g.kernel_init(reducer, 1, reduce_global, None)
            g.kernel_init_buffer(r, self.lattice_measurement)
            g.kernel_init_buffer(r, result_buffer)
            g.kernel_init_buffer(r, self.lattice_lds)
            return k, r, g.kernel_init_constant(r, size)

        self.sun_measurement_id, self.sun_measurement_reduce_id, self.argument_measurement_index = \
            measurement("lattice_measurement", self.lattice_energies, "reduce_measurement_double2")
        if self.get_plaquettes_avr:
            self.sun_measurement_plq_id, self.sun_measurement_plq_reduce_id, self.argument_plq_index = \
                measurement("lattice_measurement_plq", self.lattice_energies_plq, "reduce_measurement_plq_double2")
        if self.get_correlators:
            step = np.array(self.correlator, dtype=np.uint32)
            self.sun_measurement_corr_id, self.sun_measurement_corr_reduce_id, self.argument_correlators_index = \
                measurement("lattice_measurement_correlator", self.lattice_correlators, "reduce_measurement_plq_double2", step)

    # ------------------------------------------------------------------ run loop (:2084-2330)
    def lattice_update(self):
        g, P = self.GPU0, self.PRNG0
        if self.get_acceptance_rate:
            g.kernel_run(self.sun_clear_measurement_id)
        for par in ("odd", "even"):
            if not self.turnoff_prns:
                P.produce()
            if not self.turnoff_updates:
                g.kernel_run(self.sun_update_id[par])

    def lattice_measure(self):
        g, i = self.GPU0, self.ITER_counter
        if self.get_acceptance_rate:
            g.kernel_init_constant_reset(self.sun_reduce_acceptance_rate_id, i, self.argument_acceptance_rate_index)
            if self.NAV_counter == self.NAV:
                g.kernel_run(self.sun_reduce_acceptance_rate_id)
        if self.get_actions_avr:
            g.kernel_run(self.sun_measurement_id)
            g.kernel_init_constant_reset(self.sun_measurement_reduce_id, i, self.argument_measurement_index)
            g.kernel_run(self.sun_measurement_reduce_id)
        if self.get_plaquettes_avr:
            g.kernel_run(self.sun_measurement_plq_id)
            g.kernel_init_constant_reset(self.sun_measurement_plq_reduce_id, i, self.argument_plq_index)
            g.kernel_run(self.sun_measurement_plq_reduce_id)
        if self.get_correlators:
            g.kernel_run(self.sun_measurement_corr_id)
            g.kernel_init_constant_reset(self.sun_measurement_corr_reduce_id, i, self.argument_correlators_index)
            g.kernel_run(self.sun_measurement_corr_reduce_id)

    def lattice_start(self):
        """Buffers, programs, the initial configuration and its measurement (slot 0): lattice_simulate_start + first lattice_measure."""
        self.lattice_create_buffers()
        self.lattice_make_programs()
        if self.ints == model_start_hot:
            self.PRNG0.produce()
        self.GPU0.kernel_run(self.sun_init_id)
        self.lattice_measure()
        self.ITER_counter += 1

    def lattice_simulate(self):
        t0 = time.time()
        self.lattice_start()
        for _ in range(self.NAV_counter, self.NAV):
            self.lattice_update()
            self.NAV_counter += 1
        for _ in range(self.ITER_counter, self.ITER):
            for _ in range(self.NITER):
                self.lattice_update()
            self.lattice_measure()
            self.ITER_counter += 1
        self.GPU0.synchronize()
        self.elapsed = time.time() - t0

    # ------------------------------------------------------------------ lattice_analysis (:680-900): every denominator is the site count
    @staticmethod
    def _series(values: np.ndarray, denominator: float) -> dict:
        data = values / denominator
        last = len(data) - 1
        mean = float(data[1:].sum() / last) if last > 0 else 0.0
        var = float(((data[1:] - mean) ** 2).sum() / last) if last > 0 else 0.0
        return {"data": data, "mean_value": mean, "variance": var, "GPU_last_value": float(data[last])}

    def lattice_analysis(self):
        g, A, V = self.GPU0, self.Analysis, float(self.lattice_full_site)
        n = max(self.ITER_counter, 1)

        def two(buf):
            return g.buffer_map(buf, np.float64).reshape(-1, 2)[:n]
        if self.get_actions_avr:
            A["S"] = self._series(two(self.lattice_energies)[:, 0].copy(), V)
        if self.get_plaquettes_avr:
            e = two(self.lattice_energies_plq)
            A["Phi"] = self._series(e[:, 0].copy(), V)              # "Plq_spat": mean physical field
            A["Phi2"] = self._series(e[:, 1].copy(), V)             # "Plq_temp": mean square
            A["Phi_variance"] = {"data": A["Phi2"]["data"] - A["Phi"]["data"] ** 2}
        if self.get_correlators:
            e = two(self.lattice_correlators)
            A["Correlator1"] = self._series(e[:, 0].copy(), V)
            A["Correlator2"] = self._series(e[:, 1].copy(), V)
        if self.get_acceptance_rate:
            e = two(self.lattice_acceptance_rate)
            den = float(self.lattice_full_site * self.NHIT * self.NITER)
            A["AR_even"] = self._series(e[:, 0].copy(), den)
            A["AR_odd"] = self._series(e[:, 1].copy(), den)
        return A

    # ------------------------------------------------------------------ helpers for tests / bench
    def table_numpy(self) -> np.ndarray:
        dt = np.float64 if self.precision == model_precision_double else np.float32
        return self.GPU0.buffer_map(self.lattice_table, dt)[: self.lattice_full_site].copy()

    def set_table(self, table: np.ndarray):
        dt = np.float64 if self.precision == model_precision_double else np.float32
        full = np.zeros(self.lattice_table_size, dtype=dt)
        full[: self.lattice_full_site] = table
        self.GPU0.buffer_write_from(self.lattice_table, full)

    def close(self):
        self.GPU0.close()
