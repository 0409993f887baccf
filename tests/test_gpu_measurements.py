"""Wilson loops, Fmunu projections and the slice-resolved Polyakov loop (PL_LEVEL = 3) against an independent numpy
evaluation of the same definitions on the configuration read back from the device.

Reference definitions: wilson_loop.cl:54-452, su{2,3}_measurements_cl.cl (trace_lambda*/trace_sigma*,
lattice_retrace_plaquette{2,3}_F), polyakov.cl:143-569, host analysis suncl.cpp:1500-1672.  The reference's own CPU
check of the same quantities (su3cpu.cpp:1293-1345, 1435-1520) is exercised through the drop-in binary in
test_gpu_dropin.py.  Tolerance: 1e-11 relative (data_analysis.cpp:40), fp32 storage 5e-5.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GELLMANN = {
    1: [[0, 1, 0], [1, 0, 0], [0, 0, 0]], 2: [[0, -1j, 0], [1j, 0, 0], [0, 0, 0]], 3: [[1, 0, 0], [0, -1, 0], [0, 0, 0]],
    4: [[0, 0, 1], [0, 0, 0], [1, 0, 0]], 5: [[0, 0, -1j], [0, 0, 0], [1j, 0, 0]], 6: [[0, 0, 0], [0, 0, 1], [0, 1, 0]],
    7: [[0, 0, 0], [0, 0, -1j], [0, 1j, 0]], 8: (np.diag([1, 1, -2]) / np.sqrt(3.0)).tolist(),
}
PAULI = {1: [[0, 1], [1, 0]], 2: [[0, -1j], [1j, 0]], 3: [[1, 0], [0, -1]]}


def make_model(sun, dims, precision=1, phi=0.0, omega=0.0, **kw):
    from qcdgpu_b200 import Model
    m = Model()
    m.lattice_group = sun
    m.lattice_full_size = list(dims)
    m.precision = precision
    m.ints = 0                      # hot start: generic links
    m.BETA, m.PHI, m.OMEGA = (6.0 if sun == 3 else 2.5), phi, omega
    m.ITER, m.NAV, m.NITER = 2, 0, 1
    m.PRNG0.PRNG_generator = "RANLUX3"
    for k, v in kw.items():
        setattr(m, k, v)
    m.lattice_init()
    return m


def links(m):
    """U[x, y, z, t, dir] as complex N x N matrices, twist applied (what lattice_table_{2,3} hand to the kernels)."""
    n1, n2, n3, n4 = m.lattice_full_size
    V = n1 * n2 * n3 * n4
    t = m.table_numpy().astype(np.float64)              # (planes, ROWSIZE, 4)
    N = m.lattice_group
    U = np.zeros((V, 4, N, N), dtype=np.complex128)
    for d in range(4):
        if N == 3:
            a, b, c = t[d, :V], t[d + 4, :V], t[d + 8, :V]
            u = np.stack([a[:, 0] + 1j * b[:, 0], a[:, 1] + 1j * b[:, 1], a[:, 2] + 1j * b[:, 2]], axis=1)
            v = np.stack([c[:, 0] + 1j * c[:, 2], c[:, 1] + 1j * c[:, 3], a[:, 3] + 1j * b[:, 3]], axis=1)
            w = np.conj(np.cross(u, v))
            U[:, d] = np.stack([u, v, w], axis=1)
        else:
            a = t[d, :V]
            u1, u2 = a[:, 0] + 1j * a[:, 2], a[:, 1] + 1j * a[:, 3]
            U[:, d, 0, 0], U[:, d, 0, 1], U[:, d, 1, 0], U[:, d, 1, 1] = u1, u2, -np.conj(u2), np.conj(u1)
    # gid = y + z*N2 + t*N2N3 + x*N2N3N4
    U = U.reshape(n1, n4, n3, n2, 4, N, N).transpose(0, 3, 2, 1, 4, 5, 6).copy()
    if m.PHI != 0.0 or m.OMEGA != 0.0:
        if N == 3:
            D = np.diag(np.exp(1j * np.array([(m.PHI + m.OMEGA) / 2, (m.OMEGA - m.PHI) / 2, -m.OMEGA])))
        else:
            D = np.diag(np.exp(1j * np.array([m.PHI / 2, -m.PHI / 2])))
        U[n1 - 1, :, :, :, 1] = D @ U[n1 - 1, :, :, :, 1]
    return U


def shift(A, d, k=1):
    return np.roll(A, -k, axis=d)


def line(Ud, d, n):
    """ordered product of n links of direction d starting at every site"""
    out = Ud.copy()
    for k in range(1, n):
        out = out @ shift(Ud, d, k)
    return out


def dag(A):
    return np.conj(np.swapaxes(A, -1, -2))


def rel(a, b, tol):
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-300)


@pytest.mark.parametrize("sun,dims,precision,phi,omega", [
    (3, (4, 4, 4, 6), 1, 0.0, 0.0), (3, (8, 8, 8, 4), 1, 0.7, 0.3), (2, (8, 8, 8, 4), 1, 3.0, 0.0), (3, (4, 4, 4, 4), 0, 0.0, 0.0)])
def test_polyakov_slices_match_numpy(sun, dims, precision, phi, omega):
    m = make_model(sun, dims, precision, phi, omega, PL_level=3)
    m.lattice_simulate()
    A = m.lattice_analysis()
    U = links(m)
    P = line(U[..., 3, :, :], 3, dims[3])[:, :, :, 0]            # loops starting at t = 0
    tr = np.trace(P, axis1=-2, axis2=-1)
    den = dims[1] * dims[2] * sun
    tol = 1e-11 if precision == 1 else 5e-5
    last = m.ITER - 1
    for ax, name in enumerate("XYZ"):
        per = tr.sum(axis=tuple(a for a in range(3) if a != ax)) / den
        got_re = np.array([s["data"][last] for s in A[f"PL_{name}"]])
        got_im = np.array([s["data"][last] for s in A[f"PL_{name}_im"]])
        scale = np.abs(per).max()
        assert np.abs(got_re - per.real).max() <= tol * scale * 10, (name, got_re, per.real)
        if sun == 3:
            assert np.abs(got_im - per.imag).max() <= tol * scale * 10, (name, got_im, per.imag)
    # the slices add up to the plain Polyakov loop of the same configuration
    total = A["Polyakov_loop"]["data"][last] * dims[0] * dims[1] * dims[2] * sun
    assert rel(sum(s["data"][last] for s in A["PL_X"]) * den, total, 100 * tol)
    assert rel(total, tr.sum().real, 100 * tol)
    m.close()


@pytest.mark.parametrize("sun,dims,R,T,phi", [(3, (4, 4, 4, 6), 2, 3, 0.0), (3, (6, 6, 6, 4), 3, 2, 1.1), (2, (8, 8, 8, 4), 3, 2, 3.0), (2, (4, 4, 4, 4), 1, 1, 0.0)])
def test_wilson_loop_matches_numpy(sun, dims, R, T, phi):
    m = make_model(sun, dims, 1, phi, 0.0, get_wilson_loop=True, wilson_R=R, wilson_T=T)
    m.lattice_simulate()
    A = m.lattice_analysis()
    U = links(m)
    Tl = line(U[..., 3, :, :], 3, T)
    w = 0.0
    for d in range(3):
        Rl = line(U[..., d, :, :], d, R)
        loop = Tl @ shift(Rl, 3, T) @ dag(shift(Tl, d, R)) @ dag(Rl)
        w += np.trace(loop, axis1=-2, axis2=-1).real.sum()
    want = w / (np.prod(dims) * 3)
    got = A["Wilson_loop"]["data"][m.ITER - 1]
    assert rel(got, want, 1e-11), (got, want)
    if R == 1 and T == 1:       # W(1,1) is the mean temporal plaquette
        assert rel(got, A["Plq_temp"]["data"][m.ITER - 1], 1e-11)
    m.close()


@pytest.mark.parametrize("sun,dims,mode,g1,g2,phi", [
    (3, (4, 4, 4, 4), "FMUNU", 0, 0, 2.0), (3, (4, 4, 6, 4), "F0MU", 0, 0, 0.0), (3, (4, 4, 4, 4), "FMUNU", 1, 5, 1.0),
    (3, (4, 4, 4, 4), "FMUNU", 2, 6, 0.0), (3, (4, 4, 4, 4), "F0MU", 4, 7, 0.5), (2, (4, 4, 4, 6), "FMUNU", 0, 0, 3.0),
    (2, (4, 4, 4, 4), "F0MU", 1, 0, 0.0)])
def test_fmunu_matches_numpy(sun, dims, mode, g1, g2, phi):
    m = make_model(sun, dims, 1, phi, 0.0, get_Fmunu=(mode == "FMUNU"), get_F0mu=(mode == "F0MU"), Fmunu_gen1=g1, Fmunu_gen2=g2)
    m.lattice_simulate()
    A = m.lattice_analysis()
    U = links(m)
    i1, i2 = m.Fmunu_indices()
    gens = GELLMANN if sun == 3 else PAULI
    planes = [(0, 1, "xy"), (0, 2, "xz"), (1, 2, "yz")] if mode == "FMUNU" else [(0, 3, "xt"), (1, 3, "yt"), (2, 3, "zt")]
    V = float(np.prod(dims))
    last = m.ITER - 1
    for a, b, name in planes:
        W = U[..., a, :, :] @ shift(U[..., b, :, :], a) @ dag(shift(U[..., a, :, :], b)) @ dag(U[..., b, :, :])
        for idx in (i1, i2):
            f = (-0.5j * np.trace(W @ np.array(gens[idx], dtype=np.complex128), axis1=-2, axis2=-1)).sum() / V
            got = A[f"Fmunu_{name}_{idx}_re"]["data"][last] + 1j * A[f"Fmunu_{name}_{idx}_im"]["data"][last]
            assert abs(got - f) <= 1e-11 * max(abs(f), 1e-3), (name, idx, got, f)
    # the plaquette sums that share the kernel are unchanged by the projection outputs
    spat = sum(np.trace(U[..., a, :, :] @ shift(U[..., b, :, :], a) @ dag(shift(U[..., a, :, :], b)) @ dag(U[..., b, :, :]),
                        axis1=-2, axis2=-1).real.sum() for a, b in [(0, 1), (0, 2), (1, 2)]) / (V * 3)
    assert rel(A["Plq_spat"]["data"][last], spat, 1e-11)
    m.close()


def test_measurement_kernels_write_every_slot_their_reducer_sums():
    """The host sizes the reducer from its own aligned grid (ceil(row size / work-group size), like the reference), which on a small
    lattice is more slots than the sites fill (4^4: 8 against 2).  The measurement kernels must write the surplus slots too (zeros):
    with the partial buffer full of garbage and no clear_measurement in front -- lattice_measurement has none in the run loop --
    action and plaquette must come out as on a clean buffer."""
    m = make_model(3, (4, 4, 4, 4), PL_level=0)
    m.lattice_start()
    g = m.GPU0
    m._measure()
    m.ITER_counter += 1
    clean = (g.buffer_map(m.lattice_energies, np.float64).copy(), g.buffer_map(m.lattice_energies_plq, np.float64).copy())
    junk = np.full(g.buffer_map(m.lattice_measurement, np.float64).shape, 1.0e30)
    m.ITER_counter -= 1
    g.buffer_write_from(m.lattice_measurement, junk)
    g.kernel_run(m.sun_measurement_id)
    g.kernel_init_constant_reset(m.sun_measurement_reduce_id, m.ITER_counter, m.argument_measurement_index)
    g.kernel_run(m.sun_measurement_reduce_id)
    g.buffer_write_from(m.lattice_measurement, junk)
    g.kernel_run(m.sun_measurement_plq_id)
    g.kernel_init_constant_reset(m.sun_measurement_plq_reduce_id, m.ITER_counter, m.argument_plq_index)
    g.kernel_run(m.sun_measurement_plq_reduce_id)
    got = (g.buffer_map(m.lattice_energies, np.float64).copy(), g.buffer_map(m.lattice_energies_plq, np.float64).copy())
    for a, b in zip(clean, got):
        np.testing.assert_array_equal(a, b)
    m.close()


@pytest.mark.parametrize("sun,dims,precision,phi,omega", [
    (3, (8, 8, 8, 4), 1, 0.0, 0.0), (3, (6, 6, 6, 4), 1, 0.7, 0.3), (2, (8, 8, 8, 6), 1, 3.0, 0.0), (3, (4, 4, 4, 4), 0, 0.0, 0.0)])
def test_action_slices_match_numpy(sun, dims, precision, phi, omega):
    """lattice_action_diff_{x,y,z} + reduce_action_diff_*_double2 (sun_measurements_cl.cl:186-706): the site action
    beta/N (N - Re Tr plaquette) of the six plaquettes that start at a site, summed per x / y / z slice, against numpy on the same links;
    6^3 x 4 pads a transverse extent from 6 to PLK = 32."""
    m = make_model(sun, dims, precision, phi, omega, PL_level=0, get_actions_diff=True)
    m.lattice_simulate()
    A = m.lattice_analysis()
    U = links(m)
    N, bet = sun, m.BETA / sun
    s_spat = np.zeros(dims)
    s_temp = np.zeros(dims)
    for a in range(4):
        for b in range(a + 1, 4):
            P = U[..., a, :, :] @ shift(U[..., b, :, :], a) @ dag(shift(U[..., a, :, :], b)) @ dag(U[..., b, :, :])
            s = bet * (N - np.trace(P, axis1=-2, axis2=-1).real)
            if b == 3:
                s_temp += s
            else:
                s_spat += s
    tol = 1e-11 if precision == 1 else 5e-5
    last = m.ITER - 1
    vol = dims[0] * dims[1] * dims[2] * dims[3]
    for ax, name in enumerate("XYZ"):
        den = vol // dims[ax] * 3
        other = tuple(a for a in range(4) if a != ax)
        for kind, field in (("s", s_spat), ("t", s_temp)):
            per = field.sum(axis=other) / den
            got = np.array([x["data"][last] for x in A[f"S_diff_{name}_{kind}"]])
            assert np.abs(got - per).max() <= 10 * tol * np.abs(per).max(), (name, kind, got, per)
        # the slices add up to the plain action of the same configuration (lattice_measurement, denominator full_site * 3)
        total = sum(x["data"][last] for x in A[f"S_diff_{name}_s"]) * den
        assert rel(total, A["S_spat"]["data"][last] * vol * 3, 100 * tol)
    m.close()
