"""The oracle against the reference's own outputs (tests/golden/, made by make_golden.py
from the UNMODIFIED reference CPU_RUN build and PRNG host replicas)."""
import glob
import json
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def _observables(L):
    V = L.sites
    ps, pt = L.measure_plq()
    ss, st = L.measure_action()
    return [ps / (3 * V), pt / (3 * V), (ps + pt) / (6 * V), ss / (3 * V), st / (3 * V), (ss + st) / (6 * V)]


def _sweep(L, prn_fn=None):
    for odd in (1, 0):
        for d in range(4):
            L.update(prn_fn() if prn_fn else None, odd, d)


@pytest.mark.parametrize("name", ["su3_4x4x4x4_gid", "su3_4x6x2x8_gid", "su2_4x4x4x4_gid", "su2_6x4x8x2_gid"])
def test_gid_start_matches_reference_cpurun(name):
    """INTS=2 deterministic run: indexing, staple, CM/KP arithmetic, plaquette and action.
    Tolerance 2e-12 relative: the reference prints 14 significant digits."""
    g = _load(f"ref_cpurun_{name}.json")
    p = g["params"]
    L = O.Lattice(p["GROUP"], (p["L1"], p["L2"], p["L3"], p["L4"]), "f64", nhit=p["NHIT"], gid_upd=2)
    L.set_params(p["BETA"])
    L.init_gid()
    rows = [_observables(L)]
    for _ in range(p["NAV"]):
        _sweep(L)
    for _ in range(1, p["ITER"]):
        for _ in range(p["NITER"]):
            _sweep(L)
        rows.append(_observables(L))
    np.testing.assert_allclose(np.array(rows), np.array(g["rows"]), rtol=2e-12, atol=0)


def test_gid_start_opencl_flavour_is_close_to_cpurun_flavour():
    """The OpenCL path keeps the closed-form pseudo-randoms in working precision, CPU_RUN rounds
    them to float (suncpp/su2/update_su2.cpp:71-76): both flavours agree to ~1e-7."""
    g = _load("ref_cpurun_su3_4x4x4x4_gid.json")
    L = O.Lattice(3, (4, 4, 4, 4), "f64", gid_upd=1)
    L.set_params(6.0)
    L.init_gid()
    _sweep(L)
    np.testing.assert_allclose(_observables(L), g["rows"][1], rtol=1e-7)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ref_prng_*.json"))))
def test_prng_streams_bit_exact(path):
    g = _load(os.path.basename(path))
    want = np.array([int(b, 16) for b in g["bits"]], dtype=np.uint32)
    n = len(want)
    host = O.host_stream(g["generator"], g["randseries"], n).view(np.uint32)
    np.testing.assert_array_equal(host, want)
    if g["generator"].startswith("RANLUX"):
        # the device-layout producer: instance 0 of a multi-instance table, float4 samples
        P = O.Prng(g["generator"], g["randseries"], instances=8)
        buf = P.produce(n // 4).reshape(n // 4, 8, 4)
        np.testing.assert_array_equal(buf[:, 0, :].reshape(-1).view(np.uint32), want[: 4 * (n // 4)])
        # produce() in two calls continues the same streams
        P2 = O.Prng(g["generator"], g["randseries"], instances=8)
        a = P2.produce(5).reshape(5, 8, 4)
        b = P2.produce(n // 4 - 5).reshape(n // 4 - 5, 8, 4)
        np.testing.assert_array_equal(np.concatenate([a, b]), buf)
    if g["generator"] == "PM":
        # integer stream bit-exact; the device kernel divides by float(2^31) (prngcl_pm.cl:96,119)
        # while the host replica divides by 2147483647.0 in double (random.cpp:906): same integers
        P = O.Prng("PM", g["randseries"], instances=4)
        buf = P.produce(n).reshape(n, 4, 4)[:, 0, 0]
        ints_dev = np.round(buf.astype(np.float64) * 2147483648.0)
        ints_host = np.round(host.view(np.float32).astype(np.float64) * 2147483647.0)
        assert np.max(np.abs(ints_dev - ints_host)) <= 128  # float32 has 24 bits
    if g["generator"] == "XOR7":
        # four consecutive numbers per float4: instance 0 of the device-layout producer IS the host replica's stream
        P = O.Prng("XOR7", g["randseries"], instances=8)
        buf = P.produce(n // 4).reshape(n // 4, 8, 4)
        np.testing.assert_array_equal(buf[:, 0, :].reshape(-1).view(np.uint32), want[: 4 * (n // 4)])
    if g["generator"] in ("RANECU", "RANMAR"):
        # four independent lanes per instance: lane x of instance 0 is the host replica (i_quads = 1, random.cpp:360-362)
        P = O.Prng(g["generator"], g["randseries"], instances=8)
        buf = P.produce(n).reshape(n, 8, 4)
        np.testing.assert_array_equal(buf[:, 0, 0].view(np.uint32), want)
        P2 = O.Prng(g["generator"], g["randseries"], instances=8)
        a = P2.produce(7).reshape(7, 8, 4)
        b = P2.produce(n - 7).reshape(n - 7, 8, 4)
        np.testing.assert_array_equal(np.concatenate([a, b]), buf)
        if g["generator"] == "RANMAR":       # all four lanes of instance 0 carry the replica's seeds
            for lane in range(1, 4):
                np.testing.assert_array_equal(buf[:, 0, lane], buf[:, 0, 0])
        assert np.all(buf >= 0.0) and np.all(buf < 1.0)


def test_xor128_known_answer():
    """Marsaglia's xorshift128 from the canonical state must give the published first outputs."""
    import ctypes as C
    s = [123456789, 362436069, 521288629, 88675123]
    outs = []
    for _ in range(3):
        t = (s[0] ^ (s[0] << 11)) & 0xFFFFFFFF
        s = [s[1], s[2], s[3], ((s[3] ^ (s[3] >> 19)) ^ (t ^ (t >> 8))) & 0xFFFFFFFF]
        outs.append(s[3])
    assert outs == [3701687786, 458299110, 2500872618]
    P = O.Prng("XOR128", 54, instances=4, table_size=64)
    seeds = P.seeds().reshape(-1, 4)
    buf = P.produce(3).reshape(3, 4, 4)
    st = [int(x) for x in seeds[2]]
    exp = []
    for _ in range(12):
        t = (st[0] ^ (st[0] << 11)) & 0xFFFFFFFF
        st = [st[1], st[2], st[3], ((st[3] ^ (st[3] >> 19)) ^ (t ^ (t >> 8))) & 0xFFFFFFFF]
        exp.append(np.float32(st[0]) / np.float32(4294967296.0))
    np.testing.assert_array_equal(buf[:, 2, :].reshape(-1), np.array(exp, dtype=np.float32))
    host = O.host_stream("XOR128", 54, 12)
    np.testing.assert_array_equal(host, buf[:, 0, :].reshape(-1))


def test_mrg32k3a_recurrence():
    """MRG32k3a has no host replica in the reference (only prngcl_mrg32k3a.cl): parity UNPINNED by
    the reference.  The oracle is checked against an exact big-integer evaluation of L'Ecuyer's
    recurrence (a12=1403580, a13=-810728 mod m1=2^32-209; a21=527612, a23=-1370589 mod m2=2^32-22853).
    Hand-derived first value from the all-12345 state: p1=3023790853, p2=2478282264."""
    m1, m2 = 4294967087, 4294944443
    assert (1403580 * 12345 - 810728 * 12345) % m1 == 3023790853
    assert (527612 * 12345 - 1370589 * 12345) % m2 == 2478282264
    P = O.Prng("MRG32K3A", 1, instances=1)
    # and the oracle's producer follows the same recurrence from its own seed table
    seeds = P.seeds().reshape(2, 4)
    buf = P.produce(2).reshape(-1)
    x1 = [int(v) for v in seeds[0][:3]]
    x2 = [int(v) for v in seeds[1][:3]]
    exp = []
    for _ in range(8):
        p1 = (1403580 * x1[1] - 810728 * x1[0]) % m1
        x1 = [x1[1], x1[2], p1]
        p2 = (527612 * x2[2] - 1370589 * x2[0]) % m2
        x2 = [x2[1], x2[2], p2]
        exp.append(np.float32(((p1 - p2) if p1 > p2 else (p1 - p2 + m1)) * 2.3283065492957276882397381481401e-10))
    np.testing.assert_array_equal(buf, np.array(exp, dtype=np.float32))


def test_cold_start_values_and_invariants():
    for sun, prec in ((3, "f64"), (2, "f64"), (3, "f32"), (2, "f32")):
        L = O.Lattice(sun, (4, 4, 6, 4), prec)
        L.set_params(6.0 if sun == 3 else 2.5)
        L.init_cold()
        ps, pt = L.measure_plq()
        assert ps == 3 * sun * L.sites and pt == 3 * sun * L.sites
        assert L.measure_action() == (0.0, 0.0)
        re, im, p2, p4 = L.polyakov()
        assert re == sun * 4 * 4 * 6 and im == 0.0


def test_hot_start_unitary_and_heatbath_preserves_unitarity():
    P = O.Prng("RANLUX3", 54, instances=64)
    for sun, prec, tol in ((3, "f64", 1e-6), (2, "f32", 1e-5)):
        L = O.Lattice(sun, (4, 4, 4, 4), prec)
        L.set_params(6.0 if sun == 3 else 2.5)
        n = (3 * L.rowsize + 63) // 64
        for d in range(4):
            L.init_hot(P.produce(max(n, 33 * L.prngstep // 64 + 1)), d)
        t = L.table.astype(np.float64)
        if sun == 3:
            # rows u (uv1.xyz + i uv2.xyz) and v are orthonormal (float-precision PRNs -> 1e-6)
            u = t[0:4, : L.sites, :3] + 1j * t[4:8, : L.sites, :3]
            v = np.stack([t[8:12, : L.sites, 0] + 1j * t[8:12, : L.sites, 2],
                          t[8:12, : L.sites, 1] + 1j * t[8:12, : L.sites, 3],
                          t[0:4, : L.sites, 3] + 1j * t[4:8, : L.sites, 3]], axis=-1)
            assert np.max(np.abs(np.sum(np.abs(u) ** 2, -1) - 1)) < tol
            assert np.max(np.abs(np.sum(np.abs(v) ** 2, -1) - 1)) < tol
            assert np.max(np.abs(np.sum(u * np.conj(v), -1))) < tol
        else:
            assert np.max(np.abs(np.sum(t[:, : L.sites, :] ** 2, -1) - 1)) < tol
        before = L.measure_plq()
        for odd in (1, 0):
            for d in range(4):
                L.update(P.produce(33 * L.prngstep // 64 + 1), odd, d)
        L.gramschmidt()
        after = L.measure_plq()
        assert after[0] > before[0] and after[1] > before[1]   # heat bath at large beta orders a hot lattice


def test_overrelaxation_preserves_local_action():
    """Overrelaxation is not in the reference: pinned by its invariant,
    Re Tr(U*staple) unchanged for every updated link, and by leaving total action unchanged."""
    P = O.Prng("RANLUX3", 11, instances=64)
    for sun in (3, 2):
        L = O.Lattice(sun, (4, 4, 4, 4), "f64")
        L.set_params(6.0)
        for d in range(4):
            L.init_hot(P.produce(3 * L.rowsize // 64 + 1), d)
        s0 = sum(L.measure_plq())
        sites = [L.even_gid(g) for g in range(0, L.sites // 2, 7)]
        before = [L.retrace_ustaple(g, 2) for g in sites]
        L.overrelax(0, 2)
        after = [L.retrace_ustaple(g, 2) for g in sites]
        np.testing.assert_allclose(after, before, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(sum(L.measure_plq()), s0, rtol=1e-12)
        changed = np.abs(L.table).sum()
        assert changed > 0


def test_ensemble_average_matches_reference_cpurun():
    """Stochastic agreement with the reference's own CPU_RUN run (different PRN consumption order
    => only ensemble averages are comparable).  8^4, 40 thermalisation + 40 measured sweeps;
    tolerance = 5 sigma of the two time series' standard errors with a x3 autocorrelation allowance."""
    for name, sun, beta in (("su3_8x8x8x8_cold_ens", 3, 6.0),):
        g = _load(f"ref_cpurun_{name}.json")
        ref = np.array(g["rows"])[1:, 2]
        L = O.Lattice(sun, (8, 8, 8, 8), "f64")
        L.set_params(beta)
        L.init_cold()
        P = O.Prng("RANLUX3", 99, instances=L.prngstep)
        mine = []
        for it in range(30):
            for odd in (1, 0):
                for d in range(4):
                    L.update(P.produce(33), odd, d)
            L.gramschmidt()
            if it >= 12:
                ps, pt = L.measure_plq()
                mine.append((ps + pt) / (6 * L.sites))
        mine = np.array(mine)
        se = 3 * np.sqrt(ref[10:].var() / len(ref[10:]) + mine.var() / len(mine))
        assert abs(mine.mean() - ref[10:].mean()) < 5 * se, (mine.mean(), ref[10:].mean(), se)
