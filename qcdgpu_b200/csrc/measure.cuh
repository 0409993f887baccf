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
// Roofs at SU(3) 64^4 fp64: 6.44 GB of links (1.0 ms at the measured HBM rate) and ~1800 DFMA per site (1.6 ms at the fp64
// rate); the kernel takes 3.0 ms.  A three-threads-per-site split (two plaquettes each, 168 registers) was measured
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
__device__ __forceinline__ V4<double> promote2(const V4<float>& a) { return { double(a.x), double(a.y), double(a.z), double(a.w) }; }
__device__ __forceinline__ V4<double> promote2(const V4<double>& a) { return a; }
__device__ __forceinline__ double retrace_plq2(const V4<double>& U1, const V4<double>& U2, const V4<double>& U3, const V4<double>& U4) {
    V4<double> p = su2_mul(su2_mul(su2_mul(U1, U2), su2_dag(U3)), su2_dag(U4));
    return 2.0 * p.x;
}

// out layout: measurement[BID] = (S_spat, S_temp) partial; measurement[BID + offset] = (Plq_spat, Plq_temp) partial
template <int SUN, typename R, bool TBC>
__global__ void __launch_bounds__(kMeasThreads)
measure_kernel(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g,
               double2* __restrict__ out_action, double2* __restrict__ out_plq, uint32_t first_site, uint32_t nsites) {
    const R beta = params[0];
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = { 0.0, 0.0 };
    if (k < nsites) {
        uint32_t gid = first_site + k;
        Coord p = gid_to_coords(g, gid);
        Coord pX = nb_plus<0>(g, p), pY = nb_plus<1>(g, p), pZ = nb_plus<2>(g, p), pT = nb_plus<3>(g, p);
        uint32_t gX = coords_to_gid(g, pX), gY = coords_to_gid(g, pY), gZ = coords_to_gid(g, pZ), gT = coords_to_gid(g, pT);
        double spat, temp;
        if (SUN == 3) {
            M3<double> ux = promote3(load3<R, 0, TBC>(table, g, gid, p.x, tw));
            M3<double> uy = promote3(load3<R, 1, TBC>(table, g, gid, p.x, tw));
            M3<double> uz = promote3(load3<R, 2, TBC>(table, g, gid, p.x, tw));
            M3<double> ut = promote3(load3<R, 3, TBC>(table, g, gid, p.x, tw));
            // sun_measurements_cl.cl:140-172
            spat  = retrace_plq3(ux, promote3(load3<R, 1, TBC>(table, g, gX, pX.x, tw)), promote3(load3<R, 0, TBC>(table, g, gY, pY.x, tw)), uy);
            spat += retrace_plq3(ux, promote3(load3<R, 2, TBC>(table, g, gX, pX.x, tw)), promote3(load3<R, 0, TBC>(table, g, gZ, pZ.x, tw)), uz);
            temp  = retrace_plq3(ux, promote3(load3<R, 3, TBC>(table, g, gX, pX.x, tw)), promote3(load3<R, 0, TBC>(table, g, gT, pT.x, tw)), ut);
            spat += retrace_plq3(uy, promote3(load3<R, 2, TBC>(table, g, gY, pY.x, tw)), promote3(load3<R, 1, TBC>(table, g, gZ, pZ.x, tw)), uz);
            temp += retrace_plq3(uy, promote3(load3<R, 3, TBC>(table, g, gY, pY.x, tw)), promote3(load3<R, 1, TBC>(table, g, gT, pT.x, tw)), ut);
            temp += retrace_plq3(uz, promote3(load3<R, 3, TBC>(table, g, gZ, pZ.x, tw)), promote3(load3<R, 2, TBC>(table, g, gT, pT.x, tw)), ut);
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
    block_sum<2>(v);
    if (threadIdx.x == 0) {
        const double trN = (SUN == 3) ? 9.0 : 6.0;
        // number of active sites in this block (the reference adds bet*(9 - retr) per active thread)
        uint32_t base = blockIdx.x * blockDim.x;
        uint32_t act = (base < nsites) ? min(uint32_t(blockDim.x), nsites - base) : 0u;
        if (out_plq) out_plq[blockIdx.x] = make_double2(v[0], v[1]);
        if (out_action) {
            double b = double(beta);
            out_action[blockIdx.x] = make_double2(b * (trN * act - v[0]), b * (trN * act - v[1]));
        }
    }
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
