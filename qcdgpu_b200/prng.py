"""Host side of the PRNG subsystem: mirror of PRNG_CL::PRNG (SU(N)/random/random.{h,cpp}).

Seeds are drawn exactly as the reference draws them: `srand(RANDSERIES)` then glibc `rand()`
in the reference's order (random.cpp:235, 462-465, 508, 651-676, 866-895) -- glibc's rand
sequence is part of the stream definition, so libc is called through ctypes.
Device work goes through the GPU_CL::GPU-shaped C ABI (capi.GPU).
"""
from __future__ import annotations

import numpy as np

from .capi import GPU, host_rand_fill, host_srand

RL_NSKIP = {"RANLUX0": 24, "RANLUX1": 48, "RANLUX2": 97, "RANLUX3": 223, "RANLUX4": 389}
GENERATORS = tuple(RL_NSKIP) + ("XOR128", "PM", "MRG32K3A", "XOR7", "RANECU", "RANMAR")
# convert_generator_to_uint (random.cpp:108-121); MRG32K3A has no number in the reference
GENERATOR_IDS = {"XOR128": 1, "RANLUX0": 2, "RANLUX1": 3, "RANLUX2": 4, "RANLUX3": 5, "RANLUX4": 6,
                 "RANLUX": 7, "RANMAR": 8, "PM": 9, "XOR7": 10, "RANECU": 11, "MRG32K3A": 12}


class PRNG:
    """PRNG_CL::PRNG: initialize() builds seed tables + kernels, produce() = one producer pass."""

    def __init__(self):
        self.PRNG_generator = "RANLUX3"
        self.PRNG_instances = 0          # 0: autoselect = GPU_info.max_memory_width (random.cpp:219-223)
        self.PRNG_samples = 0            # set by the model; per-instance after initialize()
        self.PRNG_randseries = 0
        self.PRNG_srandtime = 0
        self.PRNG_counter = 0
        self.RL_nskip = 24
        self.GPU0: GPU | None = None
        self.PRNG_seed_id = 0
        self.PRNG_seed_table_id = 0
        self.PRNG_randoms_id = 0
        self.PRNG_randoms_seed_id = 0
        self.PRNG_randoms_kernel_id = 0
        self.randoms_produced = 0
        self.seeds = None

    def parameters_setup(self, parameter: str, ivalue: int, text_value: str):
        # random.cpp:139-156
        if parameter == "RANDSERIES":
            self.PRNG_randseries = ivalue
        if parameter == "PRNG" and text_value in GENERATORS + ("RANLUX",):
            self.PRNG_generator = text_value

    def print_generator(self) -> str:
        return (f" random series               : {self.PRNG_srandtime}\n"
                f" PRN generator               : {self.PRNG_generator}\n")

    def initialize(self):
        gpu = self.GPU0
        autoselect = False
        if self.PRNG_instances == 0:
            autoselect = True
            self.PRNG_instances = int(gpu.GPU_info.max_memory_width)
            self.PRNG_samples = 1 + (self.PRNG_samples - 1) // self.PRNG_instances
        self.autoselect = autoselect
        import time
        self.PRNG_srandtime = int(time.time()) & 0xFFFFFFFF
        if not self.PRNG_randseries:
            self.PRNG_randseries = self.PRNG_srandtime
        else:
            self.PRNG_srandtime = self.PRNG_randseries
        host_srand(self.PRNG_srandtime)
        g = self.PRNG_generator
        if g in RL_NSKIP or g == "RANLUX":
            if g in RL_NSKIP:
                self.RL_nskip = RL_NSKIP[g]
            self._ranlux_initialize()
        elif g == "XOR128":
            self._small_initialize("xor128", 0, aligned=True)
        elif g == "PM":
            self._small_initialize("pm", 2147483647, aligned=False)
        elif g == "MRG32K3A":
            self._mrg_initialize()
        elif g == "XOR7":
            self._xor7_initialize()
        elif g == "RANECU":
            self._ranecu_initialize()
        elif g == "RANMAR":
            self._ranmar_initialize()
        else:
            raise ValueError(f"PRNG {g} is not supported (RANLUX0-4, XOR128, PM, MRG32K3A, XOR7, RANECU, RANMAR)")
        self.produce()     # random.cpp:276

    # RL_initialize_CPU + RL_initialize: random.cpp:425-522
    def _ranlux_initialize(self):
        gpu, I = self.GPU0, self.PRNG_instances
        seeds = np.empty(I, dtype=np.uint32)
        seeds[0:1] = host_rand_fill(1, 2147483647)     # RL_jseed = rand() % 2147483647 (random.cpp:508)
        seeds[1:] = host_rand_fill(I - 1)              # the other instances: rand() (random.cpp:464-465)
        self.seeds = seeds
        gpu.program_create(None, f"-D RL_skip={self.RL_nskip}")
        seed_table_size = I * 7                       # float4 x 7 per instance (112 B; 20 words are used: the packed generator, prng.cuh)
        randoms_size = I * self.PRNG_samples
        self.PRNG_seed_id = gpu.buffer_init(GPU.buffer_type_Constant, I, seeds, 4)
        self.PRNG_seed_table_id = gpu.buffer_init(GPU.buffer_type_IO, seed_table_size, None, 16)
        self.PRNG_randoms_id = gpu.buffer_init(GPU.buffer_type_IO, randoms_size, None, 16)
        self.PRNG_randoms_seed_id = gpu.kernel_init("ranlux_init", 1, [I], None)
        gpu.kernel_init_buffer(self.PRNG_randoms_seed_id, self.PRNG_seed_id)
        gpu.kernel_init_buffer(self.PRNG_randoms_seed_id, self.PRNG_seed_table_id)
        gpu.kernel_run(self.PRNG_randoms_seed_id)
        self.PRNG_randoms_kernel_id = gpu.kernel_init("ranlux", 1, [I], None)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_seed_table_id)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_randoms_id)
        gpu.kernel_init_constant(self.PRNG_randoms_kernel_id, int(self.PRNG_samples))

    # XOR128_initialize / PM_initialize: random.cpp:636-676, 860-895
    def _small_initialize(self, kernel: str, modulo: int, aligned: bool):
        gpu, I = self.GPU0, self.PRNG_instances
        table = gpu.buffer_size_align(I) if aligned else I
        randoms = gpu.buffer_size_align(I * self.PRNG_samples) if aligned else I * self.PRNG_samples
        seeds = host_rand_fill(4 * table, modulo)
        self.seeds = seeds
        gpu.program_create(None, "")
        self.PRNG_seed_table_id = gpu.buffer_init(GPU.buffer_type_IO, table, seeds, 16)
        self.PRNG_randoms_id = gpu.buffer_init(GPU.buffer_type_IO, randoms, None, 16)
        self.PRNG_randoms_kernel_id = gpu.kernel_init(kernel, 1, [I], None)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_seed_table_id)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_randoms_id)
        gpu.kernel_init_constant(self.PRNG_randoms_kernel_id, int(self.PRNG_samples))

    # no host code in the reference; table = uint4 seed1[I] then uint4 seed2[I]
    def _mrg_initialize(self):
        gpu, I = self.GPU0, self.PRNG_instances
        seeds = host_rand_fill(8 * I, 2147483647, 1)
        self.seeds = seeds
        gpu.program_create(None, "")
        self.PRNG_seed_table_id = gpu.buffer_init(GPU.buffer_type_IO, 2 * I, seeds, 16)
        self.PRNG_randoms_id = gpu.buffer_init(GPU.buffer_type_IO, I * self.PRNG_samples, None, 16)
        self.PRNG_randoms_kernel_id = gpu.kernel_init("mrg32k3a", 1, [I], None)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_seed_table_id)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_randoms_id)
        gpu.kernel_init_constant(self.PRNG_randoms_kernel_id, int(self.PRNG_samples))

    def _two_row_kernel(self, kernel: str, seeds: np.ndarray):
        """seed table = uint4 row 1 [I] then uint4 row 2 [I]; producer kernel(seed_table, randoms, samples)."""
        gpu, I = self.GPU0, self.PRNG_instances
        self.seeds = seeds
        gpu.program_create(None, "")
        self.PRNG_seed_table_id = gpu.buffer_init(GPU.buffer_type_IO, 2 * I, seeds, 16)
        self.PRNG_randoms_id = gpu.buffer_init(GPU.buffer_type_IO, I * self.PRNG_samples, None, 16)
        self.PRNG_randoms_kernel_id = gpu.kernel_init(kernel, 1, [I], None)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_seed_table_id)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_randoms_id)
        gpu.kernel_init_constant(self.PRNG_randoms_kernel_id, int(self.PRNG_samples))

    # XOR7_initialize_CPU + XOR7_initialize: random.cpp:921-974.  Eight rand() for the host replica's state, then rows 1 .. 2I-1 of
    # the table (row 0 is skipped, row I is drawn and then overwritten); instance 0 = the host replica.
    def _xor7_initialize(self):
        I = self.PRNG_instances
        state = host_rand_fill(8)
        seeds = np.zeros(8 * I, dtype=np.uint32)
        seeds[4:] = host_rand_fill(8 * I - 4)
        seeds[0:4] = state[0:4]
        seeds[4 * I:4 * I + 4] = state[4:8]
        self._two_row_kernel("xor7", seeds)

    # RANECU_initialize_CPU + RANECU_initialize: random.cpp:1003-1048.  Lane x of instance 0 = the host replica.
    def _ranecu_initialize(self):
        I = self.PRNG_instances
        jseed = host_rand_fill(2, 2147483647)
        seeds = host_rand_fill(8 * I, 2147483647)
        seeds[0] = jseed[0]
        seeds[4 * I] = jseed[1]
        self._two_row_kernel("ranecu", seeds)

    # RANMAR_initialize_CPU + RANMAR_initialize: random.cpp:708-799.  Two seeds per lane; all four lanes of instance 0 carry the host
    # replica's pair; per further instance four first seeds (% 31328) then four second seeds (% 30081).  The reference's host names
    # the kernels "rmseed" / "rmproduce" (its .cl file: ranmar_init / ranmar); the backend binds both spellings.
    def _ranmar_initialize(self):
        gpu, I = self.GPU0, self.PRNG_instances
        rm = np.array([host_rand_fill(1, 31328)[0], host_rand_fill(1, 30081)[0]], dtype=np.uint32)
        seeds = np.zeros(8 * I, dtype=np.uint32)
        seeds[0:4] = rm[0]
        seeds[4 * I:4 * I + 4] = rm[1]
        if I > 1:
            raw = host_rand_fill(8 * (I - 1)).reshape(I - 1, 8)      # per instance: four first seeds, then four second seeds
            seeds[4:4 * I] = (raw[:, :4] % 31328).reshape(-1)
            seeds[4 * I + 4:] = (raw[:, 4:] % 30081).reshape(-1)
        self.seeds = seeds
        gpu.program_create(None, "")
        seeds_size = 2 * I
        seed_table_size = seeds_size * (97 + 1)       # random.cpp:719: twice what the kernels touch
        self.PRNG_seed_id = gpu.buffer_init(GPU.buffer_type_Constant, seeds_size, seeds, 16)
        self.PRNG_seed_table_id = gpu.buffer_init(GPU.buffer_type_IO, seed_table_size, None, 16)
        self.PRNG_randoms_id = gpu.buffer_init(GPU.buffer_type_IO, I * self.PRNG_samples, None, 16)
        self.PRNG_randoms_seed_id = gpu.kernel_init("rmseed", 1, [I], None)
        gpu.kernel_init_buffer(self.PRNG_randoms_seed_id, self.PRNG_seed_id)
        gpu.kernel_init_buffer(self.PRNG_randoms_seed_id, self.PRNG_seed_table_id)
        gpu.kernel_run(self.PRNG_randoms_seed_id)
        self.PRNG_randoms_kernel_id = gpu.kernel_init("rmproduce", 1, [I], None)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_seed_table_id)
        gpu.kernel_init_buffer(self.PRNG_randoms_kernel_id, self.PRNG_randoms_id)
        gpu.kernel_init_constant(self.PRNG_randoms_kernel_id, int(self.PRNG_samples))

    def produce(self):
        # random.cpp:278-284
        self.GPU0.kernel_run(self.PRNG_randoms_kernel_id)
        self.randoms_produced += self.PRNG_samples * 4
        self.PRNG_counter += 1

    def randoms(self) -> np.ndarray:
        """buffer_map_float4(PRNG_randoms_id): the PRN buffer as (n, 4) float32."""
        return self.GPU0.buffer_map_float4(self.PRNG_randoms_id)
