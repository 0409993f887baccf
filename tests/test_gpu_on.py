"""O(N) scalar branch on the GPU, through the C ABI (capi.GPU + on_model.OnModel), against
  * the UNMODIFIED reference's CPU re-measurement (tests/golden/ref_on_check_*.json), tolerance = the reference's own
    (1e-11 double, 5e-5 single, data_analysis.cpp:40-43),
  * the C restatement (oracle/on_oracle.c) pass by pass on the same PRN buffers.
Metropolis accept/reject is a data-dependent branch: in double a differing last bit of exp/log flips a decision with
probability ~1e-15 per proposal, in single ~1e-7; the single-precision trajectories therefore tolerate isolated sites."""
import glob
import json
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_model(dims, precision="f64", nhit=10, ints=1, fused=False, prng="XOR128", iters=8, **par):
    from qcdgpu_b200.on_model import OnModel
    m = OnModel()
    m.lattice_full_size = list(dims)
    m.precision = 1 if precision == "f64" else 0
    m.NHIT, m.ints, m.ITER, m.NITER, m.NAV = nhit, ints, iters, 1, 0
    m.fused_prng = fused
    m.PRNG0.PRNG_generator, m.PRNG0.PRNG_randseries = prng, 54
    for k, v in par.items():
        setattr(m, {"z": "ON_z", "lam": "ON_lambda", "zeta": "ON_zeta", "eta": "ON_eta", "b": "ON_b", "skgroup": "ON_SK_group"}[k], v)
    if not fused:
        m.GPU0.L.qgpu_set_max_memory_width(m.GPU0.h, 256)
    m.lattice_init()
    return m


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ref_on_check_*.json"))))
def test_measurements_against_the_reference_checker(path):
    g = json.load(open(path))
    m = make_model(g["dims"], g["precision"], **g["params"])
    m.correlator = list(g["step"])
    m.lattice_start()
    dt = np.float64 if g["precision"] == "f64" else np.float32
    m.set_table(np.frombuffer(bytes.fromhex(g["table_hex"]), dtype=dt))
    m.lattice_measure()
    m.ITER_counter += 1
    A = m.lattice_analysis()
    got = [A["S"]["data"][1], A["Phi"]["data"][1], A["Phi2"]["data"][1], A["Phi_variance"]["data"][1],
           A["Correlator1"]["data"][1], A["Correlator2"]["data"][1]]
    np.testing.assert_allclose(got, g["reference"], rtol=1e-11 if g["precision"] == "f64" else 5e-5)
    # slot 0 holds the cold start: U = 0.25 everywhere
    L = O.OnLattice(tuple(g["dims"]), precision=g["precision"], **g["params"])
    L.init_cold()
    np.testing.assert_allclose(A["S"]["data"][0], L.measure_action()[0] / L.sites, rtol=1e-12 if g["precision"] == "f64" else 1e-5)
    m.close()


def _oracle_for(m, precision, nhit, **par):
    L = O.OnLattice(tuple(m.lattice_full_size), precision=precision, nhit=nhit, wg=int(m.GPU0.GPU_info.max_workgroup_size), **par)
    assert L.prngstep == m.prngstep
    np.testing.assert_array_equal(L.params[16:23], m.GPU0.buffer_map(m.lattice_parameters, L.dtype)[16:23])
    return L


@pytest.mark.parametrize("precision,nhit,ints,prng", [("f64", 10, 1, "XOR128"), ("f64", 7, 0, "RANLUX3"), ("f32", 10, 0, "XOR128"),
                                                      ("f64", 1, 1, "PM"), ("f64", 4, 0, "MRG32K3A")])
def test_update_trajectory_equals_the_oracle(precision, nhit, ints, prng):
    dims = (4, 6, 8, 4)
    m = make_model(dims, precision, nhit=nhit, ints=ints, prng=prng)
    m.lattice_create_buffers()
    m.lattice_make_programs()
    L = _oracle_for(m, precision, nhit)
    g, P = m.GPU0, m.PRNG0
    if ints == 0:
        P.produce()
        prns = P.randoms().copy()
        g.kernel_run(m.sun_init_id)
        L.init_hot(prns)
    else:
        g.kernel_run(m.sun_init_id)
        L.init_cold()
    np.testing.assert_array_equal(m.table_numpy(), L.table[:L.sites])
    flips = 0
    for sweep in range(3):
        g.kernel_run(m.sun_clear_measurement_id)
        accs = {}
        for par, odd in (("odd", 1), ("even", 0)):
            P.produce()
            prns = P.randoms().copy()
            g.kernel_run(m.sun_update_id[par])
            accs[par] = L.update(prns, odd)
            got = m.table_numpy()
            bad = np.flatnonzero(got != L.table[:L.sites])
            if precision == "f64":
                assert bad.size == 0, (sweep, par, bad[:8])
            else:
                assert bad.size <= 2, (sweep, par, bad.size)
                flips += bad.size
                L.table[:L.sites] = got          # follow the GPU's trajectory after an isolated differing decision
        # -D ACC_RATE: per-block sums, .x even pass / .y odd pass (oncl.cl:213-216, 353-356)
        part = g.buffer_map(m.lattice_measurement, np.float64).reshape(-1, 2)
        nb = -(-(L.sites // 2) // g.kernel_get_worksize(m.sun_update_id["odd"]))
        if flips == 0:
            assert part[:nb, 0].sum() == accs["even"] and part[:nb, 1].sum() == accs["odd"]
    tol = 1e-12 if precision == "f64" else 1e-5
    m.lattice_measure()
    m.ITER_counter += 1
    A = m.lattice_analysis()
    np.testing.assert_allclose(A["S"]["data"][0], L.measure_action()[0] / L.sites, rtol=tol)
    np.testing.assert_allclose(A["Phi"]["data"][0], L.measure_plq()[0] / L.sites, rtol=tol)
    np.testing.assert_allclose(A["Correlator1"]["data"][0], L.measure_corr(m.correlator)[0] / L.sites, rtol=tol)
    m.close()


@pytest.mark.parametrize("prng", ["XOR128", "PM", "MRG32K3A", "XOR7", "RANECU", "RANMAR"])
@pytest.mark.parametrize("nhit", [10, 5])
def test_fused_generator_equals_buffered(prng, nhit):
    """PRNG_instances == PRNGSTEP: the generator runs inside the update kernel (state in registers) and the PRN buffer is
    never written; the configuration and the generator states must be those of the buffered run, bit for bit."""
    dims = (8, 8, 8, 8)
    tables, seeds = [], []
    for fused in (False, True):
        m = make_model(dims, "f64", nhit=nhit, ints=0, fused=True, prng=prng, iters=4)
        m.GPU0.set_fused_prng(fused)
        m.NAV = 3
        m.lattice_simulate()
        tables.append(m.table_numpy())
        seeds.append(m.GPU0.buffer_map(m.PRNG0.PRNG_seed_table_id, np.uint32).copy())
        A = m.lattice_analysis()
        assert np.all(np.isfinite(A["S"]["data"])) and 0.2 < A["AR_even"]["data"][1] < 0.9
        m.close()
    np.testing.assert_array_equal(tables[0], tables[1])
    np.testing.assert_array_equal(seeds[0], seeds[1])


def test_run_loop_and_error_behaviour(tmp_path):
    from qcdgpu_b200 import QcdGpuError
    from qcdgpu_b200.on_model import OnModel
    p = tmp_path / "init.dat"
    p.write_text("\tGROUP = 1\n\tLS = 8\n\tLT = 4\n\tITER = 6\n\tNITER = 2\n\tNAV = 5\n\tNHIT = 10\n\tINTS = 1\n\tPRNG = XOR128\n"
                 "\tRANDSERIES = 45\n\tSKGROUP = 4\n\tON_ETA = 1.0e-01\n")
    m = OnModel()
    m.lattice_get_init_file(str(p))
    assert m.lattice_full_size == [8, 8, 8, 4] and m.PRNG0.PRNG_generator == "XOR128"
    m.lattice_init()
    m.lattice_simulate()
    A = m.lattice_analysis()
    # slot 0 is the uniform cold start (no kinetic term); after NAV + NITER sweeps the action sits at its thermal value
    assert A["S"]["data"].shape == (6,) and abs(A["S"]["data"][0] + 2.0) < 1e-9 and -1.0 < A["S"]["data"][-1] < 0.5
    assert np.all(A["Phi_variance"]["data"][1:] > 0.0)
    # acceptance is accumulated for the last sweep in front of a measurement: (NHIT - rejected) / (V NHIT NITER)
    assert 0.0 < A["AR_even"]["data"][-1] < 1.0 and 0.0 < A["AR_odd"]["data"][-1] < 1.0
    g = m.GPU0
    with pytest.raises(QcdGpuError, match="ON_SUB_SCHEME"):
        g.program_create(None, m.options_common() + " -D ON_SUB_SCHEME")
    with pytest.raises(QcdGpuError, match="N1EXACT"):
        g.program_create(None, m.options_common().replace("-D N1EXACT=8", "-D N1EXACT=6"))
    g.program_create(None, m.options_common())
    with pytest.raises(QcdGpuError, match="not implemented"):
        g.kernel_init("lattice_polyakov", 1, [64], None)
    m.close()
