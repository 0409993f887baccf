"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module.  The product package (qcdgpu_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "QCDGPU_cpu")

PRNG_IDS = {"XOR128": 1, "RANLUX0": 2, "RANLUX1": 3, "RANLUX2": 4, "RANLUX3": 5,
            "RANLUX4": 6, "RANMAR": 8, "PM": 9, "XOR7": 10, "RANECU": 11, "MRG32K3A": 12}


class OrcCfg(C.Structure):
    _fields_ = [("sun", C.c_int32), ("n1", C.c_uint32), ("n2", C.c_uint32), ("n3", C.c_uint32),
                ("n4", C.c_uint32), ("rowsize", C.c_uint32), ("prngstep", C.c_uint32),
                ("nhit", C.c_int32), ("nhitpar", C.c_int32), ("tbc", C.c_int32),
                ("gid_upd", C.c_int32), ("biglat", C.c_int32), ("left_sites", C.c_uint32),
                ("full_sites", C.c_uint32), ("pl", C.c_int32)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB_PATH)
            for f in ("qcd_oracle.c", "qcd_oracle_lattice.inc", "qcd_oracle.h", "on_oracle.c", "on_oracle.inc", "on_oracle.h")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_prng_create.restype = C.c_void_p
        _lib.orc_prng_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.orc_prng_destroy.argtypes = [C.c_void_p]
        _lib.orc_prng_produce.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.orc_prng_seeds.restype = C.POINTER(C.c_uint32)
        _lib.orc_prng_seeds.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        _lib.orc_prng_host_stream.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
        for f in ("orc_even_gid", "orc_odd_gid"):
            getattr(_lib, f).restype = C.c_uint32
            getattr(_lib, f).argtypes = [C.POINTER(OrcCfg), C.c_uint32]
        _lib.orc_neighbour.restype = C.c_uint32
        _lib.orc_neighbour.argtypes = [C.POINTER(OrcCfg), C.c_uint32, C.c_int, C.c_int]
        for sfx in ("f32", "f64"):
            getattr(_lib, f"orc_link_retrace_ustaple_{sfx}").restype = C.c_double
    return _lib


def align(n: int, step: int) -> int:
    """GPU::buffer_size_align (clinterface.cpp:1580-1590)."""
    return (1 + (n - 1) // step) * step


@dataclass
class Lattice:
    """Oracle-side lattice with the device SoA layout."""
    sun: int
    dims: tuple  # (N1, N2, N3, N4) = (x, y, z, t)
    precision: str = "f64"   # "f64" | "f32"
    nhit: int = 10
    nhitpar: int = 1
    wg: int = 64             # alignment quantum (GPU_info.max_workgroup_size)
    tbc: bool = False
    gid_upd: int = 0
    biglat: bool = False
    left_sites: int = 0
    full_sites: int = 0

    def __post_init__(self):
        n1, n2, n3, n4 = self.dims
        self.sites = n1 * n2 * n3 * n4
        self.rowsize = align(self.sites, self.wg)
        real_sites = (n1 - 2) * n2 * n3 * n4 if self.biglat else self.sites
        self.prngstep = align(real_sites // 2, self.wg)
        self.planes = 12 if self.sun == 3 else 4
        self.dtype = np.float64 if self.precision == "f64" else np.float32
        self.table = np.zeros((self.planes, self.rowsize, 4), dtype=self.dtype)
        self.cfg = OrcCfg(self.sun, n1, n2, n3, n4, self.rowsize, self.prngstep, self.nhit,
                          self.nhitpar, int(self.tbc), self.gid_upd, int(self.biglat),
                          self.left_sites, self.full_sites, 2)
        self.params = np.zeros(8, dtype=self.dtype)

    def set_params(self, beta: float, phi: float = 0.0, omega: float = 0.0):
        self.params[0] = beta / self.sun
        self.params[1] = phi
        self.params[2] = omega

    def _fn(self, name):
        return getattr(lib(), f"{name}_{self.precision}")

    def _t(self):
        return self.table.ctypes.data_as(C.c_void_p)

    def _p(self):
        return self.params.ctypes.data_as(C.c_void_p)

    def init_cold(self):
        self._fn("orc_init_cold")(C.byref(self.cfg), self._t())

    def init_gid(self):
        self._fn("orc_init_gid")(C.byref(self.cfg), self._t())

    def init_hot(self, prns: np.ndarray, d: int):
        prns = np.ascontiguousarray(prns, dtype=np.float32)
        self._fn("orc_init_hot")(C.byref(self.cfg), self._t(), prns.ctypes.data_as(C.c_void_p), d)

    def gramschmidt(self):
        self._fn("orc_gramschmidt")(C.byref(self.cfg), self._t())

    def update(self, prns, odd: int, d: int):
        if prns is None:
            ptr = None
        else:
            prns = np.ascontiguousarray(prns, dtype=np.float32)
            ptr = prns.ctypes.data_as(C.c_void_p)
        self._fn("orc_update")(C.byref(self.cfg), self._t(), self._p(), ptr, odd, d)

    def overrelax(self, odd: int, d: int):
        self._fn("orc_overrelax")(C.byref(self.cfg), self._t(), odd, d)

    def measure_action(self):
        out = (C.c_double * 2)()
        self._fn("orc_measure_action")(C.byref(self.cfg), self._t(), self._p(), out)
        return out[0], out[1]

    def measure_plq(self):
        out = (C.c_double * 2)()
        self._fn("orc_measure_plq")(C.byref(self.cfg), self._t(), self._p(), out)
        return out[0], out[1]

    def polyakov(self):
        out = (C.c_double * 4)()
        self._fn("orc_polyakov")(C.byref(self.cfg), self._t(), self._p(), out)
        return tuple(out)

    def retrace_ustaple(self, gindex: int, d: int) -> float:
        return self._fn("orc_link_retrace_ustaple")(C.byref(self.cfg), self._t(), self._p(),
                                                    C.c_uint32(gindex), d)

    def even_gid(self, g):
        return lib().orc_even_gid(C.byref(self.cfg), g)

    def odd_gid(self, g):
        return lib().orc_odd_gid(C.byref(self.cfg), g)


class Prng:
    def __init__(self, generator: str, randseries: int, instances: int, table_size: int = 0):
        self.gen = PRNG_IDS[generator]
        self.instances = instances
        self.h = lib().orc_prng_create(self.gen, randseries, instances, table_size)
        if not self.h:
            raise ValueError(generator)

    def produce(self, samples: int) -> np.ndarray:
        out = np.empty((samples * self.instances, 4), dtype=np.float32)
        lib().orc_prng_produce(self.h, samples, out.ctypes.data_as(C.c_void_p))
        return out

    def seeds(self) -> np.ndarray:
        n = C.c_uint32()
        p = lib().orc_prng_seeds(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def __del__(self):
        try:
            lib().orc_prng_destroy(self.h)
        except Exception:
            pass


def host_stream(generator: str, randseries: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    lib().orc_prng_host_stream(PRNG_IDS[generator], randseries, n, out.ctypes.data_as(C.c_void_p))
    return out


# ---------------------------------------------------------------------------------------------- O(N) scalar branch
ON_CHECK_BIN = os.path.join(ORACLE_DIR, "_ref", "on_check")
# defaults of the reference's O(N)/QCDGPU.cpp:119-126
ON_DEFAULTS = dict(z=0.1561573990981998, lam=0.25, zeta=0.25, eta=0.1, b=9.661840260273483, skgroup=4)


class OrcOnCfg(C.Structure):
    _fields_ = [("n1", C.c_uint32), ("n2", C.c_uint32), ("n3", C.c_uint32), ("n4", C.c_uint32),
                ("prngstep", C.c_uint32), ("nhit", C.c_uint32), ("skgroup", C.c_uint32)]


def on_params(z, lam, zeta, eta, b, dtype=np.float64, size=64) -> np.ndarray:
    p = np.zeros(size, dtype=np.float64)
    lib().orc_on_params(C.c_double(z), C.c_double(lam), C.c_double(zeta), C.c_double(eta), C.c_double(b),
                        p.ctypes.data_as(C.c_void_p))
    return p.astype(dtype)


@dataclass
class OnLattice:
    """Oracle-side O(N) lattice: one real per site, index gid = y + z N2 + t N2 N3 + x N2 N3 N4."""
    dims: tuple
    precision: str = "f64"
    nhit: int = 10
    wg: int = 64
    z: float = ON_DEFAULTS["z"]
    lam: float = ON_DEFAULTS["lam"]
    zeta: float = ON_DEFAULTS["zeta"]
    eta: float = ON_DEFAULTS["eta"]
    b: float = ON_DEFAULTS["b"]
    skgroup: int = ON_DEFAULTS["skgroup"]

    def __post_init__(self):
        n1, n2, n3, n4 = self.dims
        self.sites = n1 * n2 * n3 * n4
        self.prngstep = align(self.sites // 2, self.wg)
        self.dtype = np.float64 if self.precision == "f64" else np.float32
        self.table = np.zeros(align(self.sites, self.wg), dtype=self.dtype)
        self.cfg = OrcOnCfg(n1, n2, n3, n4, self.prngstep, self.nhit, self.skgroup)
        self.params = on_params(self.z, self.lam, self.zeta, self.eta, self.b, self.dtype)

    def _fn(self, name):
        return getattr(lib(), f"{name}_{self.precision}")

    def _t(self):
        return self.table.ctypes.data_as(C.c_void_p)

    def _p(self):
        return self.params.ctypes.data_as(C.c_void_p)

    def init_cold(self):
        self._fn("orc_on_init_cold")(C.byref(self.cfg), self._t())

    def init_hot(self, prns):
        prns = np.ascontiguousarray(prns, dtype=np.float32)
        self._fn("orc_on_init_hot")(C.byref(self.cfg), self._t(), prns.ctypes.data_as(C.c_void_p), self._p())

    def update(self, prns, odd: int) -> float:
        prns = np.ascontiguousarray(prns, dtype=np.float32)
        acc = C.c_double()
        self._fn("orc_on_update")(C.byref(self.cfg), self._t(), self._p(), prns.ctypes.data_as(C.c_void_p), odd, C.byref(acc))
        return acc.value

    def _two(self, name, *extra):
        out = (C.c_double * 2)()
        self._fn(name)(C.byref(self.cfg), self._t(), self._p(), *extra, out)
        return out[0], out[1]

    def measure_action(self):
        return self._two("orc_on_measure_action")

    def measure_plq(self):
        return self._two("orc_on_measure_plq")

    def measure_corr(self, step):
        s = (C.c_uint32 * 4)(*step)
        return self._two("orc_on_measure_corr", s)

    def reference_check(self, step=(1, 1, 1, 1)):
        """Run the UNMODIFIED reference checker (oracle/_ref/on_check) on this table: per-site means
        (S, phi, phi^2, variance, corr1, corr2) as SU::lattice_check_cpu stores them."""
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            f.write(self.table[:self.sites].tobytes())
            f.flush()
            args = [ON_CHECK_BIN, f.name, *map(str, self.dims), "1" if self.precision == "f64" else "0",
                    repr(self.z), repr(self.lam), repr(self.zeta), repr(self.eta), repr(self.b), str(self.skgroup),
                    *map(str, step)]
            out = subprocess.check_output(args, text=True)
        return [float(v) for v in out.split()]
