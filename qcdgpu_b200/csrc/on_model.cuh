// on_model.cuh -- the O(N) scalar branch of the reference (tree O(N)/: one real U in (0, max_U) per site,
// Skalozub's action with SKGROUP in the measure) for sm_100a.
//
// Replaces the kernels of O(N)/oncl/oncl.cl (lattice_init_cold / lattice_init_hot, update_even / update_odd,
// reduce_acceptance_rate) and O(N)/oncl/on_measurements_cl.cl (lattice_measurement, lattice_measurement_plq,
// lattice_measurement_correlator and their reducers) with everything they inline from o1cl.cl (o1_action, o1_Phi,
// o1_to_physical_field), o1_matrix_memory.cl and kernel/neighbours.cl.  Site index and checkerboard are those of the
// gauge branch (gid = y + z N2 + t N2 N3 + x N2 N3 N4), the table is one real per site.
//
// Type semantics kept from the reference: its expressions mix hgpu_float variables with double literals, so with
// PRECISION = single they are evaluated in double and rounded on assignment, while log / exp of a float-typed
// expression are the float functions.  C++ promotes the same way, so the code below is written once for R.
//
// What is new:
//   * a proposal's action needs the four forward neighbours only through log Phi(U_k); those are computed once per
//     site, not once per proposal (the reference recomputes 4 log(1-U_k), 4 Phi and 4 divisions per proposal);
//     log(Phi_k / Phi) = log Phi_k - log Phi and log((1 - eta l) / (1 - U)) = log(1 - eta l) - l,  l = log(1 - U);
//   * the two proposals of a PRN vec4 are evaluated together (two independent chains of transcendentals per thread);
//     the first accepted one in the reference's order wins, so the result is the reference's;
//   * the generator can run inside the kernel (one XOR128 / PM / MRG32k3a instance per thread, state in registers);
//   * reductions by warp shuffles; the measurement kernels use the reference's exact expression order.
#pragma once
#include "launch.h"
#include "update.cuh"      // PRN sources (PrnBuffered, PrnSmallState), F4
#include "measure.cuh"     // block_sum

namespace qcd {

constexpr int kOnThreads = 128;

template <typename R> struct OnPar { R z, zeta, eta, b, szl, max_u; double skm1_2; };
template <typename R> __device__ __forceinline__ OnPar<R> on_par(const R* __restrict__ p, uint32_t skgroup) {
    // lattice_parameters[16..22]: O(N)/suncl/suncl.cpp:1987-1994, 2015-2022; SKGROUPM1_2: kernel/model.cl:106
    OnPar<R> q = { p[16], p[18], p[19], p[20], p[21], p[22], 0.5 * double(int(skgroup) - 1) };
    return q;
}
template <typename R> __device__ __forceinline__ R on_log(R x);
// fp64 logarithms: the positive-normal-argument log of the gauge kernels (update.cuh, Mth<double>::log_, at most 1 ulp off; zero gives
// -inf as the library does).  Only the two logs of the reference-order action tail (measurement kernel) stay with the library.
template <> __device__ __forceinline__ double on_log<double>(double x) { return Mth<double>::log_(x); }
__device__ __forceinline__ double on_dlog(double x) { return Mth<double>::log_(x); }
template <> __device__ __forceinline__ float on_log<float>(float x) { return logf(x); }
template <typename R> __device__ __forceinline__ R on_exp(R x);
template <> __device__ __forceinline__ double on_exp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float on_exp<float>(float x) { return expf(x); }

// o1_log_one_m_Ux, o1_Phi: o1cl.cl:40-52
template <typename R> __device__ __forceinline__ R on_log1m(R u) { return R(on_dlog(1.0 - double(u))); }
template <typename R> __device__ __forceinline__ R on_phi(R l, const OnPar<R>& q) {
    R a = R((-1.0 + 0.5 * double(q.eta) * double(l)) * double(l));
    return R(0.5 * double(a) * double(a) * double(q.b));
}
// o1_to_physical_field: o1cl.cl:148-156
template <typename R> __device__ __forceinline__ R on_field(R u, const OnPar<R>& q) {
    return R(sqrt(2.0 * double(on_phi(on_log1m(u), q))));
}
// potential + measure part of o1_action (o1cl.cl:134-145) for a site with l = log(1 - U), Phi, kinetic sum `term`
template <typename R, bool EXACT> __device__ __forceinline__ R on_action_tail(R u, R l, R Phi, R term, const OnPar<R>& q) {
    double s2p = sqrt(2.0 * double(Phi));
    double mn = (1.9999999999 < s2p) ? 1.9999999999 : s2p;
    R el = q.eta * l;
    double lg = EXACT ? log((1.0 - double(el)) / (1.0 - double(u))) : on_dlog(1.0 - double(el)) - double(l);
    R st = q.szl * term;
    return R(-q.skm1_2 * (EXACT ? log(2.0 - mn) : on_dlog(2.0 - mn)) - lg + double(Phi) * (double(Phi) - 1.0 + double(st)) / double(q.z));
}
// o1_action as the reference writes it (o1cl.cl:54-146, without ON_SUB_SCHEME, which the host never defines)
template <typename R> __device__ __forceinline__ R on_action_exact(R u, const R nb[4], const OnPar<R>& q) {
    R l = on_log1m(u), Phi = on_phi(l, q), term = R(0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        R lp = on_log<R>(on_phi(on_log1m(nb[k]), q) / Phi);
        if (k == 3) lp = q.zeta * lp;
        term = (k == 0) ? lp * lp : term + lp * lp;
    }
    return on_action_tail<R, true>(u, l, Phi, term, q);
}
// the same with the neighbours entering through lnb[k] = log Phi(U_k)
template <typename R> __device__ __forceinline__ R on_action_fast(R u, const R lnb[4], const OnPar<R>& q) {
    R l = on_log1m(u), Phi = on_phi(l, q), lP = on_log<R>(Phi), term = R(0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        R lp = lnb[k] - lP;
        if (k == 3) lp = q.zeta * lp;
        term = (k == 0) ? lp * lp : term + lp * lp;
    }
    return on_action_tail<R, false>(u, l, Phi, term, q);
}
template <typename R> __device__ __forceinline__ void on_neighbours(const R* __restrict__ table, const Geom& g, const Coord& p, R nb[4]) {
    nb[0] = table[coords_to_gid(g, nb_plus<0>(g, p))];     // lattice_neighbours_gid, kernel/neighbours.cl:119-146
    nb[1] = table[coords_to_gid(g, nb_plus<1>(g, p))];
    nb[2] = table[coords_to_gid(g, nb_plus<2>(g, p))];
    nb[3] = table[coords_to_gid(g, nb_plus<3>(g, p))];
}

// lattice_init_cold (oncl.cl:63-71, lattice_ground_o_1 = 0.25) / lattice_init_hot (oncl.cl:48-61, rnd.x * max_U)
template <typename R>
__global__ void __launch_bounds__(kOnThreads) on_init_kernel(R* __restrict__ table, const R* __restrict__ params, const float4* __restrict__ prns, uint32_t sites) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sites) return;
    table[g] = prns ? R(prns[g].x) * params[22] : R(0.25);
}

// update_even / update_odd (oncl.cl:73-370): multi-hit Metropolis, two proposals per PRN vec4, first accepted proposal wins.
// acc != null (-D ACC_RATE): acc[BID].x (even pass) or .y (odd pass) += sum over the block of NHIT - rejected proposals.
template <typename R, class Prn>
__global__ void __launch_bounds__(kOnThreads)
on_update_kernel(R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, uint32_t skgroup, int odd,
                 typename Prn::Args pa, double2* __restrict__ acc) {
    const uint32_t GID = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = { 0.0 };
    if (GID < g.nthreads) {
        const OnPar<R> q = on_par(params, skgroup);
        const uint32_t gindex = parity_gid(g, GID, odd != 0);
        const Coord p = gid_to_coords(g, gindex);
        Prn prn(pa, g, GID, gindex);
        R nb[4], lnb[4];
        on_neighbours(table, g, p, nb);
        const R u0 = table[gindex];
#pragma unroll
        for (int k = 0; k < 4; k++) lnb[k] = on_log<R>(on_phi(on_log1m(nb[k]), q));
        const R s_old = on_action_fast(u0, lnb, q);
        const int nhit = int(g.nhit), pairs = nhit >> 1;       // NHIT_2, kernel/model.cl:146-148
        int rejected = nhit;
        R u_new = u0;
        bool flag = false;
#pragma unroll 1
        for (int i = 0; i < pairs; i++) {
            const F4 r = prn.trial();
            const R c1 = R(r.x) * q.max_u, c2 = R(r.z) * q.max_u;
            const R d1 = on_exp<R>(s_old - on_action_fast(c1, lnb, q));
            const R d2 = on_exp<R>(s_old - on_action_fast(c2, lnb, q));
            if (R(r.y) < d1) { u_new = c1; flag = true; rejected = 2 * i; break; }
            if (R(r.w) < d2) { u_new = c2; flag = true; rejected = 2 * i + 1; break; }
        }
        if ((nhit & 1) && !flag) {                              // odd NHIT: one more proposal from a fresh vec4 (oncl.cl:187-207)
            const F4 r = prn.trial();
            const R c1 = R(r.x) * q.max_u;
            if (R(r.y) < on_exp<R>(s_old - on_action_fast(c1, lnb, q))) { u_new = c1; flag = true; rejected = nhit - 1; }
        }
        if (flag) table[gindex] = u_new;
        prn.finish();
        v[0] = double(nhit - rejected);
    }
    if (acc) {
        block_sum<1>(v);
        if (threadIdx.x == 0) { if (odd) acc[blockIdx.x].y += v[0]; else acc[blockIdx.x].x += v[0]; }
    }
}

// lattice_measurement (what = 0: action), lattice_measurement_plq (1: field, field^2),
// lattice_measurement_correlator (2: field(x) field(x + step), field(x) field(x + (1,1,1,1))): on_measurements_cl.cl:46-207
template <typename R>
__global__ void __launch_bounds__(kOnThreads)
on_measure_kernel(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, uint32_t skgroup, int what,
                  uint4 step, double2* __restrict__ out) {
    const uint32_t gindex = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = { 0.0, 0.0 };
    if (gindex < g.sites) {
        const OnPar<R> q = on_par(params, skgroup);
        const Coord p = gid_to_coords(g, gindex);
        const R u = table[gindex];
        if (what == 0) {
            R nb[4];
            on_neighbours(table, g, p, nb);
            v[0] = double(on_action_exact(u, nb, q));
        } else if (what == 1) {
            const R f = on_field(u, q);
            v[0] = double(f); v[1] = double(R(f * f));
        } else {
            Coord a = p, d = p;                                 // lattice_neighbours_step2_gid / _diagonal_gid, neighbours.cl:204-233
            a.x += step.x; if (a.x >= g.n1) a.x -= g.n1;
            a.y += step.y; if (a.y >= g.n2) a.y -= g.n2;
            a.z += step.z; if (a.z >= g.n3) a.z -= g.n3;
            a.t += step.w; if (a.t >= g.n4) a.t -= g.n4;
            d.x++; if (d.x >= g.n1) d.x = 0;
            d.y++; if (d.y >= g.n2) d.y = 0;
            d.z++; if (d.z >= g.n3) d.z = 0;
            d.t++; if (d.t >= g.n4) d.t = 0;
            const R f = on_field(u, q);
            v[0] = double(R(f * on_field(table[coords_to_gid(g, a)], q)));
            v[1] = double(R(f * on_field(table[coords_to_gid(g, d)], q)));
        }
    }
    block_sum<2>(v);
    if (threadIdx.x == 0) out[blockIdx.x] = make_double2(v[0], v[1]);
}

// reduce_measurement_double2 / reduce_measurement_plq_double2 / reduce_acceptance_rate of the O(N) tree ADD to the slot
// (on_measurements_cl.cl:112-122, 162-171; oncl.cl:358-368); fixed summation order, one block
static __global__ void __launch_bounds__(1024)
on_reduce_add_kernel(const double2* __restrict__ table, uint32_t size, double2* __restrict__ dst, uint32_t index) {
    double v[2] = { 0.0, 0.0 };
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) { double2 t = table[i]; v[0] += t.x; v[1] += t.y; }
    block_sum<2>(v);
    if (threadIdx.x == 0) { double2 d = dst[index]; d.x += v[0]; d.y += v[1]; dst[index] = d; }
}

}  // namespace qcd
