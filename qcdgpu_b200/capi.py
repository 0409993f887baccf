"""ctypes binding of libqcdgpu_b200.so (include/qcdgpu_b200.h).

`GPU` mirrors the reference class GPU_CL::GPU (SU(N)/clinterface/clinterface.h:54-415) method for
method, so host code written against the reference reads the same here.  There is no CPU
fallback: if the CUDA library cannot be built/loaded, or no sm_100 device is present, calls fail
loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

# GPU::GPU_buffer_types (clinterface.h:73-82)
buffer_type_None, buffer_type_Global, buffer_type_Input, buffer_type_IO = 0, 1, 2, 3
buffer_type_Output, buffer_type_Constant, buffer_type_LDS, buffer_type_UAV = 4, 5, 6, 7

# launch.h PrnKind
PRN_BUFFERED, PRN_GID, PRN_RANLUX3, PRN_XOR128, PRN_PM, PRN_MRG32K3A = 0, 1, 2, 3, 4, 5


class DeviceInfo(C.Structure):
    _fields_ = [("device_name", C.c_char * 256), ("local_memory_size", C.c_uint64),
                ("max_constant_size", C.c_uint64), ("max_memory_size", C.c_uint64),
                ("max_memory_width", C.c_size_t), ("max_workgroup_size", C.c_size_t),
                ("memory_align_factor", C.c_size_t), ("sm_count", C.c_int),
                ("cc_major", C.c_int), ("cc_minor", C.c_int)]


class QcdGpuError(RuntimeError):
    pass


_lib = None

EXPORTS = [
    "qgpu_version", "qgpu_create", "qgpu_destroy", "qgpu_device_select", "qgpu_device_initialize",
    "qgpu_device_finalize", "qgpu_get_device_info", "qgpu_set_limit_max_workgroup_size",
    "qgpu_set_max_memory_width", "qgpu_set_debug", "qgpu_set_error_mode", "qgpu_last_error",
    "qgpu_program_create", "qgpu_program_set_active", "qgpu_program_get_active", "qgpu_kernel_init",
    "qgpu_kernel_init_buffer", "qgpu_kernel_init_constant_int", "qgpu_kernel_init_constant_uint4",
    "qgpu_kernel_init_constant_float", "qgpu_kernel_init_constant_double",
    "qgpu_kernel_init_constant_reset", "qgpu_kernel_run", "qgpu_kernel_get_worksize",
    "qgpu_kernel_get_execution_time", "qgpu_buffer_init", "qgpu_buffer_write", "qgpu_buffer_map",
    "qgpu_buffer_map_range", "qgpu_buffer_unmap", "qgpu_buffer_kill", "qgpu_buffer_size_bytes",
    "qgpu_buffer_size_align", "qgpu_buffer_size_align_step", "qgpu_set_async", "qgpu_synchronize",
    "qgpu_set_fused_prng", "qgpu_set_tile_x", "qgpu_set_fold_reunitarisation", "qgpu_profile_collect", "qgpu_set_stream", "qgpu_kernel_set_x_range", "qgpu_kernel_set_x_range_strided", "qgpu_buffer_device_ptr", "qgpu_stream", "qgpu_timer_start",
    "qgpu_timer_stop", "qgpu_launch_count", "qgpu_update_kernel_registers", "qgpu_halo_copy",
    "qgpu_halo_pack", "qgpu_halo_unpack", "qgpu_halo_bytes", "qgpu_host_srand", "qgpu_host_rand_fill",
    "qgpu_ranlux_jump_table", "qgpu_buffer_ipc_export", "qgpu_ipc_open", "qgpu_ipc_close", "qgpu_kernel_set_peer_halo",
    "qgpu_halo_signal", "qgpu_halo_wait", "qgpu_enable_peer_access", "qgpu_halo_order",
]


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """dlopen the CUDA library (building it in-tree first if the sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("QCDGPU_LIB"):       # experiments: a library built with other QCD_NVCC_EXTRA flags (e.g. under _variants/)
        path, build_if_missing = os.environ["QCDGPU_LIB"], False
    if build_if_missing:
        try:
            path = _build.build()
        except Exception as e:  # no nvcc on this box: use the shipped .so if it is there
            if not os.path.exists(path):
                raise QcdGpuError(f"libqcdgpu_b200.so is missing and cannot be built: {e}") from e
    if not os.path.exists(path):
        raise QcdGpuError(f"{path} not found: run `python -m qcdgpu_b200.build` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, ci, cu, sz = C.c_void_p, C.c_int, C.c_uint, C.c_size_t
    L.qgpu_version.restype = C.c_char_p
    L.qgpu_create.restype = vp
    L.qgpu_destroy.argtypes = [vp]
    L.qgpu_device_select.argtypes = [vp, cu, cu]
    L.qgpu_device_initialize.argtypes = [vp]
    L.qgpu_device_finalize.argtypes = [vp, ci]
    L.qgpu_get_device_info.argtypes = [vp, C.POINTER(DeviceInfo)]
    L.qgpu_set_limit_max_workgroup_size.argtypes = [vp, cu]
    L.qgpu_set_max_memory_width.argtypes = [vp, sz]
    L.qgpu_set_debug.argtypes = [vp, ci, ci, ci]
    L.qgpu_set_error_mode.argtypes = [vp, ci]
    L.qgpu_last_error.argtypes = [vp]
    L.qgpu_last_error.restype = C.c_char_p
    L.qgpu_program_create.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.qgpu_program_set_active.argtypes = [vp, ci]
    L.qgpu_program_get_active.argtypes = [vp]
    L.qgpu_kernel_init.argtypes = [vp, C.c_char_p, cu, C.POINTER(sz), C.POINTER(sz)]
    L.qgpu_kernel_init_buffer.argtypes = [vp, ci, ci]
    L.qgpu_kernel_init_constant_int.argtypes = [vp, ci, C.POINTER(ci)]
    L.qgpu_kernel_init_constant_uint4.argtypes = [vp, ci, C.POINTER(C.c_uint32)]
    L.qgpu_kernel_init_constant_float.argtypes = [vp, ci, C.POINTER(C.c_float)]
    L.qgpu_kernel_init_constant_double.argtypes = [vp, ci, C.POINTER(C.c_double)]
    L.qgpu_kernel_init_constant_reset.argtypes = [vp, ci, C.POINTER(ci), ci]
    L.qgpu_kernel_run.argtypes = [vp, ci]
    L.qgpu_kernel_get_worksize.argtypes = [vp, ci]
    L.qgpu_kernel_get_execution_time.argtypes = [vp, ci, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ci)]
    L.qgpu_buffer_init.argtypes = [vp, ci, sz, vp, sz]
    L.qgpu_buffer_write.argtypes = [vp, ci]
    L.qgpu_buffer_map.argtypes = [vp, ci]
    L.qgpu_buffer_map.restype = vp
    L.qgpu_buffer_map_range.argtypes = [vp, ci, sz, sz]
    L.qgpu_buffer_map_range.restype = vp
    L.qgpu_buffer_unmap.argtypes = [vp, ci, vp]
    L.qgpu_buffer_kill.argtypes = [vp, ci]
    L.qgpu_buffer_size_bytes.argtypes = [vp, ci]
    L.qgpu_buffer_size_bytes.restype = sz
    L.qgpu_buffer_size_align.argtypes = [vp, cu]
    L.qgpu_buffer_size_align.restype = cu
    L.qgpu_buffer_size_align_step.argtypes = [vp, cu, cu]
    L.qgpu_buffer_size_align_step.restype = cu
    L.qgpu_set_async.argtypes = [vp, ci]
    L.qgpu_synchronize.argtypes = [vp]
    L.qgpu_set_fused_prng.argtypes = [vp, ci]
    L.qgpu_set_tile_x.argtypes = [vp, ci]
    L.qgpu_set_fold_reunitarisation.argtypes = [vp, ci]
    L.qgpu_profile_collect.argtypes = [vp]
    L.qgpu_set_stream.argtypes = [vp, vp]
    L.qgpu_kernel_set_x_range.argtypes = [vp, ci, ci, ci]
    L.qgpu_kernel_set_x_range_strided.argtypes = [vp, ci, ci, ci, ci]
    L.qgpu_buffer_device_ptr.argtypes = [vp, ci]
    L.qgpu_buffer_device_ptr.restype = vp
    L.qgpu_stream.argtypes = [vp]
    L.qgpu_stream.restype = vp
    L.qgpu_timer_start.argtypes = [vp]
    L.qgpu_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.qgpu_launch_count.argtypes = [vp]
    L.qgpu_launch_count.restype = C.c_uint64
    L.qgpu_update_kernel_registers.argtypes = [vp, ci, ci, ci]
    L.qgpu_halo_copy.argtypes = [vp, ci, ci, cu, vp, ci, ci, cu, ci]
    L.qgpu_halo_pack.argtypes = [vp, ci, ci, cu, ci, vp, C.POINTER(sz)]
    L.qgpu_halo_unpack.argtypes = [vp, ci, ci, cu, ci, vp]
    L.qgpu_halo_bytes.argtypes = [vp, ci, ci]
    L.qgpu_buffer_ipc_export.argtypes = [vp, ci, vp]
    L.qgpu_ipc_open.argtypes = [vp, vp]
    L.qgpu_ipc_open.restype = vp
    L.qgpu_ipc_close.argtypes = [vp, vp]
    L.qgpu_kernel_set_peer_halo.argtypes = [vp, ci, ci, vp, cu, cu, sz]
    L.qgpu_halo_signal.argtypes = [vp, vp, vp, C.c_ulonglong]
    L.qgpu_enable_peer_access.argtypes = [vp, vp]
    L.qgpu_halo_order.argtypes = [vp, vp]
    L.qgpu_halo_wait.argtypes = [vp, ci, C.c_ulonglong, C.c_double]
    L.qgpu_halo_bytes.restype = sz
    L.qgpu_host_srand.argtypes = [cu]
    L.qgpu_host_rand_fill.argtypes = [vp, sz, C.c_uint32, C.c_uint32]
    L.qgpu_ranlux_jump_table.argtypes = [C.c_uint64, C.c_uint64, cu, ci, vp]
    L.qgpu_ranlux_jump_table.restype = None
    _lib = L
    return L


def host_srand(seed: int):
    load().qgpu_host_srand(seed & 0xFFFFFFFF)


def host_rand_fill(n: int, modulo: int = 0, add: int = 0) -> np.ndarray:
    """n consecutive glibc rand() values (optionally % modulo, + add) as uint32."""
    out = np.empty(n, dtype=np.uint32)
    load().qgpu_host_rand_fill(out.ctypes.data_as(C.c_void_p), n, modulo, add)
    return out


class GPU:
    """Python face of GPU_CL::GPU.  Method names and argument meaning follow clinterface.h."""

    buffer_type_None, buffer_type_Global, buffer_type_Input, buffer_type_IO = 0, 1, 2, 3
    buffer_type_Output, buffer_type_Constant, buffer_type_LDS, buffer_type_UAV = 4, 5, 6, 7

    def __init__(self):
        self.L = load()
        self.h = self.L.qgpu_create()
        self.L.qgpu_set_error_mode(self.h, 1)      # Python raises instead of exit(-1)
        self.GPU_limit_max_workgroup_size = 0
        self.GPU_info = None
        self._keep = []                            # host arrays that must outlive their buffers

    # -- helpers
    def _chk(self, r, what=""):
        if r is None or (isinstance(r, int) and r < 0):
            raise QcdGpuError(f"{what}: {self.L.qgpu_last_error(self.h).decode()}")
        return r

    # -- device
    def device_select(self, platform_id: int, device_id: int) -> bool:
        return self._chk(self.L.qgpu_device_select(self.h, platform_id, device_id), "device_select") > 0

    def device_initialize(self) -> int:
        if self.GPU_limit_max_workgroup_size:
            self.L.qgpu_set_limit_max_workgroup_size(self.h, self.GPU_limit_max_workgroup_size)
        n = self._chk(self.L.qgpu_device_initialize(self.h), "device_initialize")
        info = DeviceInfo()
        self._chk(self.L.qgpu_get_device_info(self.h, C.byref(info)), "get_device_info")
        self.GPU_info = info
        return n

    def device_finalize(self, error_code: int = 0) -> int:
        return self.L.qgpu_device_finalize(self.h, error_code)

    def set_max_memory_width(self, w: int):
        self.L.qgpu_set_max_memory_width(self.h, w)
        if self.GPU_info is not None:
            self.GPU_info.max_memory_width = w

    def set_debug(self, profiling=False, brief_report=False, show_stage=False):
        self.L.qgpu_set_debug(self.h, int(profiling), int(brief_report), int(show_stage))

    def profile_collect(self) -> int:
        """profiling = 2: wait for the launches recorded so far and add their times to the kernels' statistics."""
        return self._chk(self.L.qgpu_profile_collect(self.h), "profile_collect")

    # -- programs
    def program_create(self, source, options: str = "") -> int:
        return self._chk(self.L.qgpu_program_create(self.h, None, options.encode()), "program_create")

    def program_set_active(self, pid: int) -> int:
        return self._chk(self.L.qgpu_program_set_active(self.h, pid), "program_set_active")

    # -- kernels
    def kernel_init(self, name: str, work_dimensions: int, global_size, local_size=None) -> int:
        g = (C.c_size_t * 3)(*(list(global_size) + [1, 1, 1])[:3])
        l = None if local_size is None else (C.c_size_t * 3)(*(list(local_size) + [0, 0, 0])[:3])
        return self._chk(self.L.qgpu_kernel_init(self.h, name.encode(), work_dimensions, g, l), f"kernel_init({name})")

    def kernel_init_buffer(self, kid: int, bid: int) -> int:
        return self._chk(self.L.qgpu_kernel_init_buffer(self.h, kid, bid), "kernel_init_buffer")

    def kernel_init_constant(self, kid: int, value) -> int:
        if isinstance(value, (list, tuple, np.ndarray)):
            v = (C.c_uint32 * 4)(*[int(x) for x in value])
            return self._chk(self.L.qgpu_kernel_init_constant_uint4(self.h, kid, v), "kernel_init_constant")
        if isinstance(value, float):
            return self._chk(self.L.qgpu_kernel_init_constant_double(self.h, kid, C.byref(C.c_double(value))), "kernel_init_constant")
        return self._chk(self.L.qgpu_kernel_init_constant_int(self.h, kid, C.byref(C.c_int(int(value)))), "kernel_init_constant")

    def kernel_init_constant_reset(self, kid: int, value: int, argument_id: int) -> int:
        return self._chk(self.L.qgpu_kernel_init_constant_reset(self.h, kid, C.byref(C.c_int(int(value))), argument_id),
                         "kernel_init_constant_reset")

    def kernel_run(self, kid: int) -> int:
        return self._chk(self.L.qgpu_kernel_run(self.h, kid), "kernel_run")

    def kernel_get_worksize(self, kid: int) -> int:
        return self._chk(self.L.qgpu_kernel_get_worksize(self.h, kid), "kernel_get_worksize")

    def kernel_get_execution_time(self, kid: int):
        m, d, n = C.c_double(), C.c_double(), C.c_int()
        self.L.qgpu_kernel_get_execution_time(self.h, kid, C.byref(m), C.byref(d), C.byref(n))
        return m.value, d.value, n.value

    # -- buffers
    def buffer_init(self, buffer_type: int, size: int, host_ptr, size_of: int) -> int:
        ptr = None
        if host_ptr is not None:
            arr = np.ascontiguousarray(host_ptr)
            assert arr.nbytes >= size * size_of, (arr.nbytes, size, size_of)
            self._keep.append(arr)
            ptr = arr.ctypes.data_as(C.c_void_p)
        return self._chk(self.L.qgpu_buffer_init(self.h, buffer_type, size, ptr, size_of), "buffer_init")

    def buffer_map(self, bid: int, dtype=np.uint32) -> np.ndarray:
        """Device -> host; returns a numpy view of the backend-owned (pinned) shadow."""
        p = self.L.qgpu_buffer_map(self.h, bid)
        if not p:
            raise QcdGpuError(f"buffer_map: {self.L.qgpu_last_error(self.h).decode()}")
        n = self.L.qgpu_buffer_size_bytes(self.h, bid)
        buf = (C.c_char * n).from_address(p)
        return np.frombuffer(buf, dtype=dtype)

    def buffer_map_float4(self, bid: int) -> np.ndarray:
        return self.buffer_map(bid, np.float32).reshape(-1, 4)

    def buffer_write_from(self, bid: int, arr: np.ndarray):
        """Extension: overwrite a device buffer from a host array (tests, .qcg load)."""
        view = self.buffer_map(bid, np.uint8)
        src = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        assert src.size == view.size, (src.size, view.size)
        view[:] = src
        p = view.ctypes.data_as(C.c_void_p)
        self._chk(self.L.qgpu_buffer_unmap(self.h, bid, p), "buffer_unmap")

    def buffer_kill(self, bid: int):
        return self._chk(self.L.qgpu_buffer_kill(self.h, bid), "buffer_kill")

    def buffer_size_align(self, size: int, step: int | None = None) -> int:
        if step is None:
            return self.L.qgpu_buffer_size_align(self.h, size)
        return self.L.qgpu_buffer_size_align_step(self.h, size, step)

    def buffer_device_ptr(self, bid: int) -> int:
        return self.L.qgpu_buffer_device_ptr(self.h, bid) or 0

    def buffer_size_bytes(self, bid: int) -> int:
        return self.L.qgpu_buffer_size_bytes(self.h, bid)

    # -- extensions
    def set_async(self, on: bool):
        self.L.qgpu_set_async(self.h, int(on))

    def set_fused_prng(self, on: bool):
        self.L.qgpu_set_fused_prng(self.h, int(on))

    def set_tile_x(self, tile: int):
        self.L.qgpu_set_tile_x(self.h, int(tile))

    def set_fold_reunitarisation(self, on: bool):
        self.L.qgpu_set_fold_reunitarisation(self.h, int(on))

    def set_stream(self, cuda_stream: int):
        self._chk(self.L.qgpu_set_stream(self.h, C.c_void_p(cuda_stream)), "set_stream")

    def kernel_set_x_range(self, kid: int, x_first: int, x_count: int, x_stride: int = 1):
        self._chk(self.L.qgpu_kernel_set_x_range_strided(self.h, kid, x_first, x_count, x_stride), "kernel_set_x_range")

    def synchronize(self):
        self._chk(self.L.qgpu_synchronize(self.h), "synchronize")

    def timer_start(self):
        self._chk(self.L.qgpu_timer_start(self.h), "timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._chk(self.L.qgpu_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def launch_count(self) -> int:
        return int(self.L.qgpu_launch_count(self.h))

    def update_kernel_registers(self, program: int, direction: int, prn_kind: int) -> int:
        return self.L.qgpu_update_kernel_registers(self.h, program, direction, prn_kind)

    def halo_copy(self, dst_program, dst_buffer, x_dst, src: "GPU", src_program, src_buffer, x_src, direction):
        self._chk(self.L.qgpu_halo_copy(self.h, dst_program, dst_buffer, x_dst, src.h, src_program, src_buffer,
                                        x_src, direction), "halo_copy")

    def halo_bytes(self, program: int, direction: int) -> int:
        return self.L.qgpu_halo_bytes(self.h, program, direction)

    def halo_pack(self, program, buffer, x_src, direction, packed_ptr: int):
        n = C.c_size_t()
        self._chk(self.L.qgpu_halo_pack(self.h, program, buffer, x_src, direction, C.c_void_p(packed_ptr), C.byref(n)), "halo_pack")
        return n.value

    def halo_unpack(self, program, buffer, x_dst, direction, packed_ptr: int):
        self._chk(self.L.qgpu_halo_unpack(self.h, program, buffer, x_dst, direction, C.c_void_p(packed_ptr)), "halo_unpack")

    # -- halo exchange fused into the update kernels (peer memory through CUDA IPC)
    def buffer_ipc_export(self, buffer: int) -> bytes:
        h = (C.c_ubyte * 64)()
        self._chk(self.L.qgpu_buffer_ipc_export(self.h, buffer, h), "buffer_ipc_export")
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = self.L.qgpu_ipc_open(self.h, h)
        if not p:
            self._chk(-1, "ipc_open")
        return int(p)

    def ipc_close(self, ptr: int):
        self._chk(self.L.qgpu_ipc_close(self.h, C.c_void_p(ptr)), "ipc_close")

    def kernel_set_peer_halo(self, kernel: int, side: int, peer_table: int, src_x: int, dst_x: int, peer_rowsize: int):
        self._chk(self.L.qgpu_kernel_set_peer_halo(self.h, kernel, side, C.c_void_p(peer_table or None), src_x, dst_x, peer_rowsize), "kernel_set_peer_halo")

    def enable_peer_access(self, other: "GPU"):
        self._chk(self.L.qgpu_enable_peer_access(self.h, other.h), "enable_peer_access")

    def halo_order(self, signaller: "GPU"):
        """this handle's stream waits for what `signaller`'s stream holds so far"""
        self._chk(self.L.qgpu_halo_order(self.h, signaller.h), "halo_order")

    def halo_signal(self, flag_a: int, flag_b: int, value: int):
        self._chk(self.L.qgpu_halo_signal(self.h, C.c_void_p(flag_a or None), C.c_void_p(flag_b or None), value), "halo_signal")

    def halo_wait(self, flags_buffer: int, value: int, timeout_s: float | None = None):
        """timeout_s: how long the stream may be held before the wait gives up and is counted (env QCDGPU_HALO_TIMEOUT_S, default 30 s:
        long enough for a neighbour that writes a state file between two sweeps, short enough that a lost neighbour does not hang the GPU)."""
        if timeout_s is None:
            timeout_s = float(os.environ.get("QCDGPU_HALO_TIMEOUT_S", "30"))
        self._chk(self.L.qgpu_halo_wait(self.h, flags_buffer, value, timeout_s), "halo_wait")

    def close(self):
        if self.h:
            self.L.qgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
