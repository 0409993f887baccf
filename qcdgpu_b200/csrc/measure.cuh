// measure.cuh -- plaquette / action / Polyakov-loop measurement with warp-shuffle reductions.
//
// Replaces lattice_measurement, lattice_measurement_plq (plain variant), lattice_polyakov and
// the LDS tree reducers reduce_*_double2 (SU(N)/suncl/sun_measurements_cl.cl:55-184, 1178-1357,
// polyakov.cl:54-141, 405-432, kernel/misc.cl:94-149).  As in the reference all traces are
// accumulated in fp64 whatever the storage precision (su3_measurements_cl.cl:292-318, 527-547).
//
// One kernel computes BOTH the plaquette sums and the action (identical loads; the reference
// runs two kernels over the same 16 links/site).  Per-block partials go to the reference's
// lattice_measurement buffer (slot BID) so the host-visible buffers keep their meaning; the
// second-stage kernel sums them in a fixed order (deterministic, no atomics).
//
// Roofs at SU(3) 64^4 fp64: 6.44 GB of links (1.0 ms at the measured HBM rate) and ~1500 FP64 operations per site (two-row products
// and the Lagrange identity for the third rows, retrace_plq3_r2; 1800 with full 3x3 products, 3.0 ms).  A three-threads-per-site split (two plaquettes each, 168 registers) was measured
// slower (3.6 ms): the extra link loads and third-row reconstructions cost more than the occupancy gains.
#pragma once
#include "lattice.cuh"
#include "update.cuh"

namespace qcd {

constexpr int kMeasThreads = 128;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum of NV doubles per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV]) {
    __shared__ double sm[NV][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) { v[k] = warp_sum(v[k]); if (lane == 0) sm[k][wid] = v[k]; }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) { double t = (lane < nw) ? sm[k][lane] : 0.0; v[k] = warp_sum(t); }
    }
    __syncthreads();
}

// promote a packed link to a full double matrix (lattice_reconstruct3_double)
template <typename R> __device__ __forceinline__ M3<double> promote3(const M3r2<R>& a) {
    M3<double> m;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) m.e[i][j] = { double(a.e[i][j].re), double(a.e[i][j].im) };
    third_row(m.e[0], m.e[1], m.e[2]);
    return m;
}
// Re Tr(U1 U2 U3^+ U4^+) = Re Tr( (U1 U2) (U4 U3)^+ )
__device__ __forceinline__ double retrace_plq3(const M3<double>& U1, const M3<double>& U2, const M3<double>& U3, const M3<double>& U4) {
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<double> a = cmul(U1.e[i][0], U2.e[0][j]); cfma(a, U1.e[i][1], U2.e[1][j]); cfma(a, U1.e[i][2], U2.e[2][j]);
            Cx<double> b = cmul(U4.e[i][0], U3.e[0][j]); cfma(b, U4.e[i][1], U3.e[1][j]); cfma(b, U4.e[i][2], U3.e[2][j]);
            tr += a.re * b.re + a.im * b.im;
        }
    return tr;
}
// The same trace from the two stored rows of the LEFT factors U1, U4 (no third rows for them): A = U1 U2 and B = U4 U3 are SU(3), so only
// their rows 0, 1 are formed and the third rows enter through (a x b) . conj(c x d) = (a.c*)(b.d*) - (a.d*)(b.c*):
//   Re Tr(A B^+) = Re[ d00 + d11 + conj(d00 d11 - d01 d10) ],  dij = A_i . conj(B_j)
// 200 instead of 234 FP64 operations per plaquette, and four third-row reconstructions less per site.
template <typename T1, typename T4>
__device__ __forceinline__ double retrace_plq3_r2(const T1& U1, const M3<double>& U2, const M3<double>& U3, const T4& U4) {
    Cx<double> A[2][3], B[2][3];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<double> a = cmul(U1.e[i][0], U2.e[0][j]); cfma(a, U1.e[i][1], U2.e[1][j]); cfma(a, U1.e[i][2], U2.e[2][j]);
            Cx<double> b = cmul(U4.e[i][0], U3.e[0][j]); cfma(b, U4.e[i][1], U3.e[1][j]); cfma(b, U4.e[i][2], U3.e[2][j]);
            A[i][j] = a; B[i][j] = b;
        }
    Cx<double> d00 = cmulc(A[0][0], B[0][0]), d11 = cmulc(A[1][0], B[1][0]), d01 = cmulc(A[0][0], B[1][0]), d10 = cmulc(A[1][0], B[0][0]);
#pragma unroll
    for (int j = 1; j < 3; j++) { cfmac(d00, A[0][j], B[0][j]); cfmac(d11, A[1][j], B[1][j]); cfmac(d01, A[0][j], B[1][j]); cfmac(d10, A[1][j], B[0][j]); }
    return d00.re + d11.re + (d00.re * d11.re - d00.im * d11.im) - (d01.re * d10.re - d01.im * d10.im);
}
// the two stored rows in double (no third row)
template <typename R> __device__ __forceinline__ M3r2<double> promote3_r2(const M3r2<R>& a) {
    M3r2<double> m;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) m.e[i][j] = { double(a.e[i][j].re), double(a.e[i][j].im) };
    return m;
}
__device__ __forceinline__ V4<double> promote2(const V4<float>& a) { return { double(a.x), double(a.y), double(a.z), double(a.w) }; }
__device__ __forceinline__ V4<double> promote2(const V4<double>& a) { return a; }
__device__ __forceinline__ double retrace_plq2(const V4<double>& U1, const V4<double>& U2, const V4<double>& U3, const V4<double>& U4) {
    V4<double> p = su2_mul(su2_mul(su2_mul(U1, U2), su2_dag(U3)), su2_dag(U4));
    return 2.0 * p.x;
}

// Thread index on the padded 4D index space of the slice-resolved action kernels -> site index, or g.sites if the thread has no site
// (lattice_gidK_{x,y,z}_to_gid, sun_common.cl:222-349): the slices of `axis` lie `per-slice size` apart, one transverse extent is
// padded to PLK so that the host's work-groups stay inside a slice.
__device__ __forceinline__ uint32_t gidk_to_gid(const Geom& g, int axis, uint32_t plk, uint32_t k) {
    uint32_t x, y, z, t;
    if (axis == 0) {            // k = t + y N4 + z N2 N4 + x N2 PLK N4
        const uint32_t per = g.n2 * plk * g.n4, z4 = k / per;
        uint32_t z1 = k - z4 * per;
        const uint32_t z3 = z1 / (g.n2 * g.n4); z1 -= z3 * g.n2 * g.n4;
        const uint32_t z2 = z1 / g.n4; z1 -= z2 * g.n4;
        if (!(z1 < g.n4 && z2 < g.n2 && z3 < g.n3 && z4 < g.n1)) return g.sites;
        x = z4; y = z2; z = z3; t = z1;
    } else if (axis == 1) {     // k = t + x N4 + z N1 N4 + y N1 PLK N4
        const uint32_t per = g.n1 * plk * g.n4, z4 = k / per;
        uint32_t z1 = k - z4 * per;
        const uint32_t z3 = z1 / (g.n1 * g.n4); z1 -= z3 * g.n1 * g.n4;
        const uint32_t z2 = z1 / g.n4; z1 -= z2 * g.n4;
        if (!(z1 < g.n4 && z2 < g.n1 && z3 < g.n3 && z4 < g.n2)) return g.sites;
        x = z2; y = z4; z = z3; t = z1;
    } else {                    // k = t + y N4 + x PLK N4 + z N1 PLK N4
        const uint32_t per = g.n1 * plk * g.n4, z4 = k / per;
        uint32_t z1 = k - z4 * per;
        const uint32_t z3 = z1 / (plk * g.n4); z1 -= z3 * plk * g.n4;
        const uint32_t z2 = z1 / g.n4; z1 -= z2 * g.n4;
        if (!(z1 < g.n4 && z2 < g.n2 && z3 < g.n1 && z4 < g.n3)) return g.sites;
        x = z3; y = z2; z = z4; t = z1;
    }
    Coord c; c.x = x; c.y = y; c.z = z; c.t = t;
    return coords_to_gid(g, c);
}

// out layout: measurement[BID] = (S_spat, S_temp) partial; measurement[BID + offset] = (Plq_spat, Plq_temp) partial
// DIFF: lattice_action_diff_{x,y,z} (sun_measurements_cl.cl:186-615): the same site action on the padded index space of `axis`
template <int SUN, typename R, bool TBC, bool DIFF = false>
__global__ void __launch_bounds__(kMeasThreads)
measure_kernel(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g,
               double2* __restrict__ out_action, double2* __restrict__ out_plq, uint32_t first_site, uint32_t nsites,
               int axis = 0, uint32_t plk = 0) {
    const R beta = params[0];
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    double v[3] = { 0.0, 0.0, 0.0 };
    uint32_t site = g.sites;
    if (DIFF) site = gidk_to_gid(g, axis, plk, k);
    else if (k < nsites) site = first_site + k;
    if (site < g.sites) {
        uint32_t gid = site;
        v[2] = 1.0;
        Coord p = gid_to_coords(g, gid);
        Coord pX = nb_plus<0>(g, p), pY = nb_plus<1>(g, p), pZ = nb_plus<2>(g, p), pT = nb_plus<3>(g, p);
        uint32_t gX = coords_to_gid(g, pX), gY = coords_to_gid(g, pY), gZ = coords_to_gid(g, pZ), gT = coords_to_gid(g, pT);
        double spat, temp;
        if (SUN == 3) {
            // the site's own links are only ever LEFT factors: two rows
            M3r2<double> ux = promote3_r2(load3<R, 0, TBC>(table, g, gid, p.x, tw));
            M3r2<double> uy = promote3_r2(load3<R, 1, TBC>(table, g, gid, p.x, tw));
            M3r2<double> uz = promote3_r2(load3<R, 2, TBC>(table, g, gid, p.x, tw));
            M3r2<double> ut = promote3_r2(load3<R, 3, TBC>(table, g, gid, p.x, tw));
            // sun_measurements_cl.cl:140-172
            spat  = retrace_plq3_r2(ux, promote3(load3<R, 1, TBC>(table, g, gX, pX.x, tw)), promote3(load3<R, 0, TBC>(table, g, gY, pY.x, tw)), uy);
            spat += retrace_plq3_r2(ux, promote3(load3<R, 2, TBC>(table, g, gX, pX.x, tw)), promote3(load3<R, 0, TBC>(table, g, gZ, pZ.x, tw)), uz);
            temp  = retrace_plq3_r2(ux, promote3(load3<R, 3, TBC>(table, g, gX, pX.x, tw)), promote3(load3<R, 0, TBC>(table, g, gT, pT.x, tw)), ut);
            spat += retrace_plq3_r2(uy, promote3(load3<R, 2, TBC>(table, g, gY, pY.x, tw)), promote3(load3<R, 1, TBC>(table, g, gZ, pZ.x, tw)), uz);
            temp += retrace_plq3_r2(uy, promote3(load3<R, 3, TBC>(table, g, gY, pY.x, tw)), promote3(load3<R, 1, TBC>(table, g, gT, pT.x, tw)), ut);
            temp += retrace_plq3_r2(uz, promote3(load3<R, 3, TBC>(table, g, gZ, pZ.x, tw)), promote3(load3<R, 2, TBC>(table, g, gT, pT.x, tw)), ut);
        } else {
            V4<double> ux = promote2(load2<R, 0, TBC>(table, g, gid, p.x, tw));
            V4<double> uy = promote2(load2<R, 1, TBC>(table, g, gid, p.x, tw));
            V4<double> uz = promote2(load2<R, 2, TBC>(table, g, gid, p.x, tw));
            V4<double> ut = promote2(load2<R, 3, TBC>(table, g, gid, p.x, tw));
            spat  = retrace_plq2(ux, promote2(load2<R, 1, TBC>(table, g, gX, pX.x, tw)), promote2(load2<R, 0, TBC>(table, g, gY, pY.x, tw)), uy);
            spat += retrace_plq2(ux, promote2(load2<R, 2, TBC>(table, g, gX, pX.x, tw)), promote2(load2<R, 0, TBC>(table, g, gZ, pZ.x, tw)), uz);
            temp  = retrace_plq2(ux, promote2(load2<R, 3, TBC>(table, g, gX, pX.x, tw)), promote2(load2<R, 0, TBC>(table, g, gT, pT.x, tw)), ut);
            spat += retrace_plq2(uy, promote2(load2<R, 2, TBC>(table, g, gY, pY.x, tw)), promote2(load2<R, 1, TBC>(table, g, gZ, pZ.x, tw)), uz);
            temp += retrace_plq2(uy, promote2(load2<R, 3, TBC>(table, g, gY, pY.x, tw)), promote2(load2<R, 1, TBC>(table, g, gT, pT.x, tw)), ut);
            temp += retrace_plq2(uz, promote2(load2<R, 3, TBC>(table, g, gZ, pZ.x, tw)), promote2(load2<R, 2, TBC>(table, g, gT, pT.x, tw)), ut);
        }
        v[0] = spat; v[1] = temp;
    }
    block_sum<3>(v);
    if (threadIdx.x == 0) {
        const double trN = (SUN == 3) ? 9.0 : 6.0;
        // number of active sites in this block (the reference adds bet*(9 - retr) per active thread)
        uint32_t base = blockIdx.x * blockDim.x;
        double act = (base < nsites) ? double(min(uint32_t(blockDim.x), nsites - base)) : 0.0;
        if (DIFF) act = v[2];
        if (out_plq) out_plq[blockIdx.x] = make_double2(v[0], v[1]);
        if (out_action) {
            double b = double(beta);
            out_action[blockIdx.x] = make_double2(b * (trN * act - v[0]), b * (trN * act - v[1]));
        }
    }
}

// ---------------------------------------------------------------- Fmunu projections  (-D FMUNU / -D F0MU)
// lattice_measurement_plq variants (sun_measurements_cl.cl:710-1175): next to the plaquette sums the spatial (FMUNU, H field) or
// temporal (F0MU, E field) plaquette matrices W = U1 U2 U3^+ U4^+ are projected on two generators, -i Tr(W lambda_a / 2)
// (su3_measurements_cl.cl:357-458, 550-589; SU(2): sigma_3 or sigma_1, and sigma_2, su2_measurements_cl.cl:89-122, 234-262).
__device__ __forceinline__ M3<double> plaquette_matrix3(const M3<double>& U1, const M3<double>& U2, const M3<double>& U3, const M3<double>& U4) {
    M3<double> A = mul_full(U1, U2), B = mul_full(U4, U3), W;     // W = (U1 U2)(U4 U3)^+
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<double> a = cmulc(A.e[i][0], B.e[j][0]); cfmac(a, A.e[i][1], B.e[j][1]); cfmac(a, A.e[i][2], B.e[j][2]);
            W.e[i][j] = a;
        }
    return W;
}
// -i Tr(W lambda_a / 2), a = 1..8
__device__ __forceinline__ Cx<double> trace_lambda(const M3<double>& W, int a) {
    Cx<double> t = { 0.0, 0.0 };
    switch (a) {
        case 1: t.re =  W.e[0][1].im + W.e[1][0].im; t.im = -W.e[0][1].re - W.e[1][0].re; break;
        case 2: t.re =  W.e[0][1].re - W.e[1][0].re; t.im =  W.e[0][1].im - W.e[1][0].im; break;
        case 3: t.re =  W.e[0][0].im - W.e[1][1].im; t.im = -W.e[0][0].re + W.e[1][1].re; break;
        case 4: t.re =  W.e[0][2].im + W.e[2][0].im; t.im = -W.e[0][2].re - W.e[2][0].re; break;
        case 5: t.re =  W.e[0][2].re - W.e[2][0].re; t.im =  W.e[0][2].im - W.e[2][0].im; break;
        case 6: t.re =  W.e[1][2].im + W.e[2][1].im; t.im = -W.e[1][2].re - W.e[2][1].re; break;
        case 7: t.re =  W.e[1][2].re - W.e[2][1].re; t.im =  W.e[1][2].im - W.e[2][1].im; break;
        default:
            t.re = ( W.e[0][0].im + W.e[1][1].im - 2.0 * W.e[2][2].im) * 0.57735026918962576450914878050196;
            t.im = (-W.e[0][0].re - W.e[1][1].re + 2.0 * W.e[2][2].re) * 0.57735026918962576450914878050196;
            break;
    }
    t.re /= 2; t.im /= 2;
    return t;
}
// SU(2): W as 4-vector (x = Re u1, y = Re u2, z = Im u1, w = Im u2); sigma_1 -> Im u2, sigma_2 -> Re u2, sigma_3 -> Im u1
__device__ __forceinline__ double trace_sigma(const V4<double>& W, int a) { return a == 1 ? W.w : a == 2 ? W.y : W.z; }

// fmode 1: F on the (x,y), (x,z), (y,z) plaquettes; fmode 2: on (x,t), (y,t), (z,t).
// out[BID] = (Plq_spat, Plq_temp) partial, out[BID + (1..3)*offset] = first generator, out[BID + (4..6)*offset] = second generator
template <int SUN, typename R, bool TBC>
__global__ void __launch_bounds__(kMeasThreads)
measure_f_kernel(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g,
                 double2* __restrict__ out, uint32_t offset, int fmode, int gen1, int gen2, uint32_t first_site, uint32_t nsites) {
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    double v[14];
#pragma unroll
    for (int i = 0; i < 14; i++) v[i] = 0.0;
    if (k < nsites) {
        uint32_t gid = first_site + k;
        Coord p = gid_to_coords(g, gid);
        Coord pn[4] = { nb_plus<0>(g, p), nb_plus<1>(g, p), nb_plus<2>(g, p), nb_plus<3>(g, p) };
        uint32_t gn[4] = { coords_to_gid(g, pn[0]), coords_to_gid(g, pn[1]), coords_to_gid(g, pn[2]), coords_to_gid(g, pn[3]) };
        int fslot = 0;
        // plane order of the reference: xy, xz, xt, yz, yt, zt
#pragma unroll 1
        for (int pl = 0; pl < 6; pl++) {
            const int a = (pl < 3) ? 0 : (pl < 5) ? 1 : 2;
            const int b = (pl == 0) ? 1 : (pl == 1 || pl == 3) ? 2 : 3;
            const bool temporal = (b == 3);
            const bool want_f = (fmode == 1 && !temporal) || (fmode == 2 && temporal);
            double tr;
            Cx<double> f1 = { 0.0, 0.0 }, f2 = { 0.0, 0.0 };
            if (SUN == 3) {
                M3<double> U1 = promote3(load3_rt<R, TBC>(table, g, gid, p.x, a, tw));
                M3<double> U2 = promote3(load3_rt<R, TBC>(table, g, gn[a], pn[a].x, b, tw));
                M3<double> U3 = promote3(load3_rt<R, TBC>(table, g, gn[b], pn[b].x, a, tw));
                M3<double> U4 = promote3(load3_rt<R, TBC>(table, g, gid, p.x, b, tw));
                M3<double> W = plaquette_matrix3(U1, U2, U3, U4);
                tr = W.e[0][0].re + W.e[1][1].re + W.e[2][2].re;
                if (want_f) { f1 = trace_lambda(W, gen1); f2 = trace_lambda(W, gen2); }
            } else {
                V4<double> U1 = promote2(load2_rt<R, TBC>(table, g, gid, p.x, a, tw));
                V4<double> U2 = promote2(load2_rt<R, TBC>(table, g, gn[a], pn[a].x, b, tw));
                V4<double> U3 = promote2(load2_rt<R, TBC>(table, g, gn[b], pn[b].x, a, tw));
                V4<double> U4 = promote2(load2_rt<R, TBC>(table, g, gid, p.x, b, tw));
                V4<double> W = su2_mul(su2_mul(su2_mul(U1, U2), su2_dag(U3)), su2_dag(U4));
                tr = 2.0 * W.x;
                if (want_f) { f1.re = trace_sigma(W, gen1); f2.re = trace_sigma(W, gen2); }
            }
            if (temporal) v[1] += tr; else v[0] += tr;
            if (want_f) {
                v[2 + 2 * fslot] = f1.re; v[3 + 2 * fslot] = f1.im;
                v[8 + 2 * fslot] = f2.re; v[9 + 2 * fslot] = f2.im;
                fslot++;
            }
        }
    }
    block_sum<14>(v);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 7; q++) out[blockIdx.x + q * offset] = make_double2(v[2 * q], v[2 * q + 1]);
    }
}

// ---------------------------------------------------------------- Wilson loop W(R,T)  (wilson_loop.cl:54-452)
// per site: sum over the three spatial directions d of Re Tr( Tline(p) . Rline_d(p+T t) . Tline(p+R d)^+ . Rline_d(p)^+ ),
// lines = ordered products of T resp. R links; R, T = lattice_parameters[3], [4].  Accumulated in fp64.
template <int SUN> struct WMat;
template <> struct WMat<3> { typedef M3<double> type; };
template <> struct WMat<2> { typedef V4<double> type; };
template <int SUN, typename R, bool TBC>
__device__ __forceinline__ typename WMat<SUN>::type wl_link(const R* table, const Geom& g, const Coord& c, int dir, const Twist<R>& tw) {
    if constexpr (SUN == 3) return promote3(load3_rt<R, TBC>(table, g, coords_to_gid(g, c), c.x, dir, tw));
    else return promote2(load2_rt<R, TBC>(table, g, coords_to_gid(g, c), c.x, dir, tw));
}
__device__ __forceinline__ M3<double> wl_mul(const M3<double>& a, const M3<double>& b) { return mul_full(a, b); }
__device__ __forceinline__ V4<double> wl_mul(const V4<double>& a, const V4<double>& b) { return su2_mul(a, b); }
// ordered product of n links in direction dir starting at c; c is left at the last site used (the reference's convention)
template <int SUN, typename R, bool TBC>
__device__ __forceinline__ typename WMat<SUN>::type wl_line(const R* table, const Geom& g, Coord& c, int dir, int n, const Twist<R>& tw) {
    typename WMat<SUN>::type m = wl_link<SUN, R, TBC>(table, g, c, dir, tw);
    for (int i = 1; i < n; i++) {
        c = nb_plus_rt(g, c, dir);
        m = wl_mul(m, wl_link<SUN, R, TBC>(table, g, c, dir, tw));
    }
    return m;
}
__device__ __forceinline__ double wl_retrace(const M3<double>& m1, const M3<double>& m2, const M3<double>& m3, const M3<double>& m4) { return retrace_plq3(m1, m2, m3, m4); }
__device__ __forceinline__ double wl_retrace(const V4<double>& m1, const V4<double>& m2, const V4<double>& m3, const V4<double>& m4) { return retrace_plq2(m1, m2, m3, m4); }
template <int SUN, typename R, bool TBC>
__global__ void __launch_bounds__(kMeasThreads)
wilson_kernel(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g,
              double2* __restrict__ out, uint32_t first_site, uint32_t nsites) {
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    // the reference counts with floats: tmp = 1.1; while (tmp < wilson_X) {...; tmp += 1.0}
    int wr = 1, wt = 1;
    { float r = float(params[3]), t = float(params[4]); for (float x = 1.1f; x < r; x += 1.0f) wr++; for (float x = 1.1f; x < t; x += 1.0f) wt++; }
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = { 0.0 };
    if (k < nsites) {
        const Coord p = gid_to_coords(g, first_site + k);
        Coord c = p;
        typename WMat<SUN>::type m1 = wl_line<SUN, R, TBC>(table, g, c, 3, wt, tw);     // bottom: T links from p
        const Coord pT = nb_plus<3>(g, c);                                               // p + T t
        double w = 0.0;
#pragma unroll 1
        for (int d = 0; d < 3; d++) {
            Coord c1 = p, c2 = pT;
            typename WMat<SUN>::type m4 = wl_line<SUN, R, TBC>(table, g, c1, d, wr, tw);  // R links from p
            typename WMat<SUN>::type m2 = wl_line<SUN, R, TBC>(table, g, c2, d, wr, tw);  // R links from p + T t
            Coord c4 = nb_plus_rt(g, c1, d);                                              // p + R d
            typename WMat<SUN>::type m3 = wl_line<SUN, R, TBC>(table, g, c4, 3, wt, tw);  // T links from p + R d
            w += wl_retrace(m1, m2, m3, m4);
        }
        v[0] = w;
    }
    block_sum<1>(v);
    if (threadIdx.x == 0) out[blockIdx.x] = make_double2(v[0], 0.0);
}

// ---------------------------------------------------------------- slice-resolved Polyakov loop (PL_LEVEL = 3)
// lattice_polyakov_diff_{x,y,z} (polyakov.cl:143-403): the thread index runs over an enlarged lattice in which one of the
// transverse extents is padded to PLK so that every work-group lies inside one slice of the resolved axis
// (lattice_gidK_{x,y,z}_to_gid_xyz, sun_common.cl:93-220); partials at out[BID].
template <int SUN, typename R, bool TBC>
__device__ __forceinline__ void polyakov_line(const R* table, const Geom& g, Coord p, const Twist<R>& tw, double& re, double& im) {
    if (SUN == 3) {
        M3<R> v0 = complete(load3<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw));
        for (uint32_t i = 1; i < g.n4; i++) {
            p = nb_plus<3>(g, p);
            M3<R> v1 = complete(load3<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw));
            v0 = mul_full(v0, v1);
        }
        re = double(v0.e[0][0].re) + double(v0.e[1][1].re) + double(v0.e[2][2].re);
        im = double(v0.e[0][0].im) + double(v0.e[1][1].im) + double(v0.e[2][2].im);
    } else {
        V4<R> v0 = load2<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw);
        for (uint32_t i = 1; i < g.n4; i++) {
            p = nb_plus<3>(g, p);
            v0 = su2_mul(v0, load2<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw));
        }
        re = 2.0 * double(v0.x); im = 0.0;
    }
}
template <int SUN, typename R, bool TBC>
__global__ void __launch_bounds__(256)
polyakov_diff_kernel(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g,
                     double2* __restrict__ out, int axis, uint32_t plk) {
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    const uint32_t gdi = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = { 0.0, 0.0 };
    // fast extent A, padded extent (PLK), slice extent
    const uint32_t A = (axis == 0) ? g.n2 : g.n1;
    const uint32_t per4 = ((axis == 2) ? g.n1 * g.n3 : g.n1 * g.n2) * plk;
    const uint32_t z4 = gdi / per4;
    uint32_t z1 = gdi - z4 * per4;
    const uint32_t z3 = z1 / (A * plk);
    z1 -= z3 * A * plk;
    const uint32_t z2 = z1 / A;
    z1 -= z2 * A;
    bool ok; Coord p;
    if (axis == 0)      { ok = z1 < g.n2 && z2 < g.n3 && z3 < g.n1 && z4 == 0; p = { z3, z1, z2, 0u }; }
    else if (axis == 1) { ok = z1 < g.n1 && z2 < g.n3 && z3 < g.n2 && z4 == 0; p = { z1, z3, z2, 0u }; }
    else                { ok = z1 < g.n1 && z2 < g.n2 && z3 < g.n3 && z4 == 0; p = { z1, z2, z3, 0u }; }
    if (ok) polyakov_line<SUN, R, TBC>(table, g, p, tw, v[0], v[1]);
    block_sum<2>(v);
    if (threadIdx.x == 0) out[blockIdx.x] = make_double2(v[0], v[1]);
}
// reduce_polyakov_diff_{x,y,z}_double2 (polyakov.cl:432-569): block i < nslices sums partials [i*size, (i+1)*size) into
// dst[index + i*stride]; block nslices (single lattice only) sums [extra_off, extra_off + extra_size) into dst[index + nslices*stride]
static __global__ void __launch_bounds__(256)
reduce_slices_kernel(const double2* __restrict__ table, uint32_t size, uint32_t nslices, double2* __restrict__ dst,
                     uint32_t index, uint32_t stride, uint32_t extra_off, uint32_t extra_size) {
    const uint32_t b = blockIdx.x;
    const uint32_t off = (b < nslices) ? b * size : extra_off, n = (b < nslices) ? size : extra_size;
    double v[2] = { 0.0, 0.0 };
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { double2 t = table[off + i]; v[0] += t.x; v[1] += t.y; }
    block_sum<2>(v);
    if (threadIdx.x == 0) dst[index + b * stride] = make_double2(v[0], v[1]);
}
// reduce_wilson_double2 (wilson_loop.cl:456-465): the result buffer holds plain doubles
static __global__ void __launch_bounds__(1024)
reduce_double_x_kernel(const double2* __restrict__ table, uint32_t size, double* __restrict__ dst, uint32_t index) {
    double v[1] = { 0.0 };
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) v[0] += table[i].x;
    block_sum<1>(v);
    if (threadIdx.x == 0) dst[index] = v[0];
}

// lattice_polyakov: thread per spatial site (x,y,z), ordered product along t in working precision
template <int SUN, typename R, bool TBC>
__global__ void __launch_bounds__(kMeasThreads)
polyakov_kernel(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g,
                double2* __restrict__ out, uint32_t offset, int pl_level, uint32_t x_first, uint32_t x_count) {
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (k < x_count * g.n23) {
        // lattice_gid_to_gid_xyz (sun_common.cl:68-90): k = y + z*N2 + x*N2N3, t = 0
        uint32_t x = k / g.n23, r = k - x * g.n23;
        uint32_t z = r / g.n2, y = r - z * g.n2;
        Coord p = { x + x_first, y, z, 0u };
        double re, im;
        if (SUN == 3) {
            M3<R> v0 = complete(load3<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw));
            for (uint32_t i = 1; i < g.n4; i++) {
                p = nb_plus<3>(g, p);
                M3<R> v1 = complete(load3<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw));
                v0 = mul_full(v0, v1);
            }
            re = double(v0.e[0][0].re) + double(v0.e[1][1].re) + double(v0.e[2][2].re);
            im = double(v0.e[0][0].im) + double(v0.e[1][1].im) + double(v0.e[2][2].im);
        } else {
            V4<R> v0 = load2<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw);
            for (uint32_t i = 1; i < g.n4; i++) {
                p = nb_plus<3>(g, p);
                v0 = su2_mul(v0, load2<R, 3, TBC>(table, g, coords_to_gid(g, p), p.x, tw));
            }
            re = 2.0 * double(v0.x); im = 0.0;   // Tr of an SU(2) matrix is real: u1 + conj(u1)
        }
        double p2 = re * re + im * im;
        v[0] = re; v[1] = im; v[2] = p2; v[3] = p2 * p2;
    }
    block_sum<4>(v);
    if (threadIdx.x == 0) {
        out[blockIdx.x] = make_double2(v[0], v[1]);
        if (pl_level > 1) out[blockIdx.x + offset] = make_double2(v[2], v[3]);
    }
}

// second stage: one block sums `size` partials starting at table[off] into dst[index]
// (reduce_measurement_double2 / reduce_measurement_plq_double2 / reduce_polyakov_double2)
static __global__ void __launch_bounds__(1024)
reduce_double2_kernel(const double2* __restrict__ table, uint32_t off, uint32_t size, double2* __restrict__ dst, uint32_t index) {
    // four independent accumulators per thread keep several loads in flight; the order is fixed, so the result is reproducible
    double a[4][2] = { { 0.0, 0.0 }, { 0.0, 0.0 }, { 0.0, 0.0 }, { 0.0, 0.0 } };
    const uint32_t n = blockDim.x;
    uint32_t i = threadIdx.x;
    for (; i + 3 * n < size; i += 4 * n) {
#pragma unroll
        for (int q = 0; q < 4; q++) { double2 t = table[off + i + q * n]; a[q][0] += t.x; a[q][1] += t.y; }
    }
    for (; i < size; i += n) { double2 t = table[off + i]; a[0][0] += t.x; a[0][1] += t.y; }
    double v[2] = { (a[0][0] + a[1][0]) + (a[2][0] + a[3][0]), (a[0][1] + a[1][1]) + (a[2][1] + a[3][1]) };
    block_sum<2>(v);
    if (threadIdx.x == 0) dst[index] = make_double2(v[0], v[1]);
}
static __global__ void clear_double2_kernel(double2* p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = make_double2(0.0, 0.0);
}

}  // namespace qcd
