// init_prng.cuh -- lattice start kernels, reunitarisation and the stand-alone PRN producers.
//
// Replaces lattice_init_{cold,gid,hot_X..T}, lattice_GramSchmidt (SU(N)/suncl/suncl.cl:55-396;
// lattice_random{2,3}: su2_update_cl.cl:110-147, su3_update_cl.cl:165-225) and the producer
// kernels ranlux_init / ranlux / xor128 / pm / mrg32k3a / xor7 / ranecu / ranmar_init / ranmar (SU(N)/random/prngcl_*.cl).
// The producers fill the reference-layout buffer prns[GID + i*instances] so that any
// PRNG_instances configuration of the reference can be replayed bit for bit; the fused
// update kernels (update.cuh) use the same generator structs from registers instead.
#pragma once
#include "lattice.cuh"
#include "prng.cuh"
#include "update.cuh"

namespace qcd {

template <int SUN, typename R>
__global__ void init_cold_kernel(R* __restrict__ table, const __grid_constant__ Geom g) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= g.sites) return;
    if (SUN == 3) {
        const V4<R> a = { R(1), R(0), R(0), R(0) }, b = { R(0), R(0), R(0), R(0) }, c = { R(0), R(1), R(0), R(0) };
#pragma unroll
        for (int d = 0; d < 4; d++) { st_plane(table, g, gid, d, a); st_plane(table, g, gid, d + 4, b); st_plane(table, g, gid, d + 8, c); }
    } else {
        const V4<R> a = { R(1), R(0), R(0), R(0) };
#pragma unroll
        for (int d = 0; d < 4; d++) st_plane(table, g, gid, d, a);
    }
}

// suncl.cl:189-347 -- closed-form links of the (global) site index; fp64 transcendentals as in the reference
template <int SUN, typename R>
__global__ void init_gid_kernel(R* __restrict__ table, const __grid_constant__ Geom g) {
    uint32_t G = blockIdx.x * blockDim.x + threadIdx.x;
    if (G >= g.sites) return;
    int gid = int(G), sites = int(g.sites);
    if (g.biglat) {
        gid += -1 * int(g.n234) + int(g.left_sites);
        if (gid < 0) gid += int(g.full_sites);
        if (gid >= int(g.full_sites)) gid -= int(g.full_sites);
        sites = int(g.full_sites);
    }
    const double kd[4] = { 0.001, 0.002, 0.003, 0.004 };
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const double k = kd[d];
        if (SUN == 3) {
            double f1  = sin((k + 10.0 / sites) * gid) / 2.0;
            double f2  = cos((k - 10.0 / sites) * gid) / 3.0;
            double f3  = sin((k - 18.0 / sites) * gid) / 4.0;
            double f4  = cos((k + 12.0 / sites) * gid) / 4.0;
            double f5  = sin((k - 12.0 / sites) * gid) / 3.0;
            double f6  = cos((k - 15.0 / sites) * gid) / 2.0;
            double f7  = sin((k + 15.0 / sites) * gid) / 3.0;
            double f8  = cos((k - 11.0 / sites) * gid) / 4.0;
            double f9  = sin((k - 11.0 / sites) * gid) / 3.0;
            double f10 = cos((k + 16.0 / sites) * gid) / 2.0;
            double f11 = sin((k - 16.0 / sites) * gid) / 3.0;
            double f12 = cos((k - 18.0 / sites) * gid) / 2.0;
            Cx<R> u[3] = { { R(f1), R(f2) }, { R(f3), R(f4) }, { R(f5), R(f6) } };
            Cx<R> v[3] = { { R(f7), R(f8) }, { R(f9), R(f10) }, { R(f11), R(f12) } };
            gram_schmidt3(u, v);
            V4<R> a, b, c;
            pack3(u, v, a, b, c);
            st_plane(table, g, G, d, a); st_plane(table, g, G, d + 4, b); st_plane(table, g, G, d + 8, c);
        } else {
            double f1 = sin((k + 10.0 / sites) * gid) / 2.0;
            double f2 = cos((k - 10.0 / sites) * gid) / 3.0;
            double f3 = sin((k - 18.0 / sites) * gid) / 4.0;
            double f4 = sqrt(1.0 - f1 * f1 - f2 * f2 - f3 * f3);
            V4<R> a = { R(f1), R(f3), R(f4), R(f2) };
            st_plane(table, g, G, d, a);
        }
    }
}

template <typename R> struct HotAngles { R a0, a1, a2, a3; };
template <typename R>
__device__ __forceinline__ HotAngles<R> hot_angles(float r1, float r2, float r3) {
    // Di Pierro: alpha = pi r1, phi = 2 pi r2, cos(theta) = 2 r3 - 1
    R sinal, cosal, sinph, cosph;
    if (sizeof(R) == 8) { sincospi(double(r1), (double*)&sinal, (double*)&cosal); sincospi(2.0 * double(r2), (double*)&sinph, (double*)&cosph); }
    else { sincospif(r1, (float*)&sinal, (float*)&cosal); sincospif(2.0f * r2, (float*)&sinph, (float*)&cosph); }
    R costh = R(2) * R(r3) - R(1);
    R sinth = Mth<R>::sqrt_(R(1) - costh * costh);
    R t1 = sinal * sinth;
    return { cosal, t1 * cosph, t1 * sinph, sinal * costh };
}
// lattice_init_hot_{X,Y,Z,T}: one launch per direction, PRNs at GID + {0,1,2}*ROWSIZE (SU(3)) / GID (SU(2))
template <int SUN, typename R>
__global__ void init_hot_kernel(R* __restrict__ table, const __grid_constant__ Geom g, const float4* __restrict__ prns, int dir) {
    uint32_t G = blockIdx.x * blockDim.x + threadIdx.x;
    if (G >= g.sites) return;
    if (SUN == 3) {
        float4 q1 = prns[size_t(G)], q2 = prns[size_t(G) + g.rowsize], q3 = prns[size_t(G) + 2 * size_t(g.rowsize)];
        // lanes x, y, z of the three PRN vec4 parametrise three SU(2) elements; the 2-row result is the closed form
        // the reference defines (su3_update_cl.cl:208-220) -- it IS the map PRNs -> link, so it is kept term for term
        HotAngles<R> X = hot_angles<R>(q1.x, q2.x, q3.x), Y = hot_angles<R>(q1.y, q2.y, q3.y), Z = hot_angles<R>(q1.z, q2.z, q3.z);
        const R a0x = X.a0, a1x = X.a1, a2x = X.a2, a3x = X.a3;
        const R a0y = Y.a0, a1y = Y.a1, a2y = Y.a2, a3y = Y.a3;
        const R a0z = Z.a0, a1z = Z.a1, a2z = Z.a2, a3z = Z.a3;
        V4<R> a, b, c;
        a.x = a0x * a0y - a3x * a3y;
        b.x = a0y * a3x + a0x * a3y;
        a.y = -a0x * a1y * a1z + a0z * a2x - a0x * a2y * a2z - a1z * a2y * a3x + a1y * a2z * a3x - a1x * a3z;
        b.y = a0z * a1x + a0x * a1z * a2y - a0x * a1y * a2z - a1y * a1z * a3x - a2y * a2z * a3x + a2x * a3z;
        a.z = -a1x * a1z + a0x * a0z * a2y + a2x * a2z - a0z * a1y * a3x + a0x * a1y * a3z + a2y * a3x * a3z;
        b.z = a0x * a0z * a1y + a1z * a2x + a1x * a2z + a0z * a2y * a3x - a0x * a2y * a3z + a1y * a3x * a3z;
        a.w = -a0z * a1x * a1y - a0z * a2x * a2y + a0x * a2z + a1z * a3x - a1y * a2x * a3z + a1x * a2y * a3z;
        b.w = a0x * a1z - a0z * a1y * a2x + a0z * a1x * a2y - a2z * a3x + a1x * a1y * a3z + a2x * a2y * a3z;
        c.x = -a0y * a2x - a1x * a3y;
        c.z = a0y * a1x - a2x * a3y;
        c.y = a0x * a0z + a1y * a1z * a2x - a1x * a1z * a2y + a1x * a1y * a2z + a2x * a2y * a2z + a3x * a3z;
        c.w = -a1x * a1y * a1z - a1z * a2x * a2y + a1y * a2x * a2z - a1x * a2y * a2z - a0z * a3x + a0x * a3z;
        st_plane(table, g, G, dir, a); st_plane(table, g, G, dir + 4, b); st_plane(table, g, G, dir + 8, c);
    } else {
        float4 q = prns[G];
        HotAngles<R> X = hot_angles<R>(q.x, q.y, q.z);
        V4<R> a = { X.a0, X.a2, X.a3, X.a1 };
        st_plane(table, g, G, dir, a);
    }
}

// lattice_GramSchmidt (suncl.cl:349-396): one thread per (site, direction) here, 4x the parallelism
template <int SUN, typename R>
__global__ void gramschmidt_kernel(R* __restrict__ table, const __grid_constant__ Geom g) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t d = blockIdx.y;
    if (i >= g.sites) return;
    if (SUN == 3) {
        M3r2<R> m = unpack3(ld_plane(table, g, i, d), ld_plane(table, g, i, d + 4), ld_plane(table, g, i, d + 8));
        gram_schmidt3(m.e[0], m.e[1]);
        V4<R> a, b, c;
        pack3(m.e[0], m.e[1], a, b, c);
        st_plane(table, g, i, d, a); st_plane(table, g, i, d + 4, b); st_plane(table, g, i, d + 8, c);
    } else {
        V4<R> m = ld_plane(table, g, i, d);
        R in = R(1) / Mth<R>::sqrt_(m.x * m.x + m.y * m.y + m.z * m.z + m.w * m.w);
        m.x *= in; m.y *= in; m.z *= in; m.w *= in;
        st_plane(table, g, i, d, m);
    }
}

// ---------------------------------------------------------------- PRN producers (reference-layout buffer)
static __global__ void ranlux_init_kernel(const uint32_t* __restrict__ seeds, uint32_t* __restrict__ table, uint32_t instances) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= instances) return;
    Ranlux r;
    ranlux_seed(r, int32_t(seeds[i]));
    ranlux_store(r, table, instances, i);
}
template <int NSKIP>
__global__ void __launch_bounds__(128)
ranlux_produce_kernel(uint32_t* __restrict__ table, float4* __restrict__ prns, uint32_t instances, uint32_t samples) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= instances) return;
    Ranlux r;
    ranlux_load(r, table, instances, i);
    size_t dst = i;
#pragma unroll 1
    for (uint32_t s = 0; s < samples; s++) {
        F4 v = r.template next4<NSKIP>();
        prns[dst] = make_float4(v.x, v.y, v.z, v.w);
        dst += instances;
    }
    ranlux_store(r, table, instances, i);
}
// Fast-forward (a resumed run replays PRNG_counter produce() calls, suncl.cpp:5350): every instance advanced by `samples` float4
// in one jump (ranlux_jump.cuh; table built with from_start).  An instance the table does not cover (a `left` that is no
// multiple of four) is stepped.
static __global__ void __launch_bounds__(128)
ranlux_skip_kernel(uint32_t* __restrict__ table, uint32_t instances, const uint32_t* __restrict__ jump, unsigned long long samples) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= instances) return;
    RanluxPk<1> r;
    uint32_t left = r.load(0, table, instances, i);
    const uint32_t* entry = (jump != nullptr && left <= 24u && (left & 3u) == 0u) ? jump + (left >> 2) * 20u : nullptr;
    const uint32_t flags = entry ? __ldg(entry + 19) : 0u;
    if (flags & 1u) {
        if (!ranlux_is_fixed_point(r.pk[0].w, r.carry[0])) {
            uint32_t A[18];
#pragma unroll
            for (int w = 0; w < 18; w++) A[w] = __ldg(entry + w);
            ranlux_jump(r.pk[0].w, r.carry[0], A, flags >> 1);
        }
        left = __ldg(entry + 18);
    } else {
        Ranlux s;
        ranlux_load(s, table, instances, i);
        while (samples > 0ull) {
            const uint32_t n = samples > 0x40000000ull ? 0x40000000u : (uint32_t)samples;
            s.template skip_samples<223>(n);
            samples -= n;
        }
        ranlux_store(s, table, instances, i);
        return;
    }
    r.store(0, table, instances, i, left);
}
// RANMAR (prngcl_ranmar.cl:65-191; the reference's host asks for the kernels as "rmseed" / "rmproduce", random.cpp:757-763).
// seeds: uint4[2 x instances] (first seeds, then second seeds); table: float4[98 x instances].  One thread per (instance, lane).
static __global__ void __launch_bounds__(128)
ranmar_init_kernel(const uint32_t* __restrict__ seeds, float* __restrict__ table, uint32_t instances) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 4u * instances) return;
    const uint32_t i = t >> 2, lane = t & 3u;
    RanmarLane4::seed(seeds[4u * i + lane], seeds[4u * (size_t(instances) + i) + lane], table + 4u * size_t(i) + lane, 4u * size_t(instances));
    if (lane == 0u)
        reinterpret_cast<float4*>(table)[size_t(i) + 97u * size_t(instances)] =
            make_float4(__uint_as_float(96u), __uint_as_float(32u), 362436.0f / 16777216.0f, 0.0f);     // (I97, J97, c, 0)
}
static __global__ void __launch_bounds__(128)
ranmar_produce_kernel(float4* __restrict__ table, float4* __restrict__ prns, uint32_t instances, uint32_t samples) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= instances) return;
    const float CD = 7654321.0f / 16777216.0f, CM = 16777213.0f / 16777216.0f;
    float4 idx = table[size_t(i) + 97u * size_t(instances)];
    uint32_t I97 = __float_as_uint(idx.x), J97 = __float_as_uint(idx.y);
    float c = idx.z;
    size_t dst = i;
#pragma unroll 1
    for (uint32_t s = 0; s < samples; s++) {
        const float4 a = table[size_t(i) + size_t(I97) * instances], b = table[size_t(i) + size_t(J97) * instances];
        float4 u = make_float4(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z), __fsub_rn(a.w, b.w));
        if (u.x < 0.0f) u.x = __fadd_rn(u.x, 1.0f);
        if (u.y < 0.0f) u.y = __fadd_rn(u.y, 1.0f);
        if (u.z < 0.0f) u.z = __fadd_rn(u.z, 1.0f);
        if (u.w < 0.0f) u.w = __fadd_rn(u.w, 1.0f);
        table[size_t(i) + size_t(I97) * instances] = u;
        if (I97 == 0u) I97 = 97u;
        if (J97 == 0u) J97 = 97u;
        I97--; J97--;
        c = __fsub_rn(c, CD);
        if (c < 0.0f) c = __fadd_rn(c, CM);
        u.x = __fsub_rn(u.x, c); u.y = __fsub_rn(u.y, c); u.z = __fsub_rn(u.z, c); u.w = __fsub_rn(u.w, c);
        if (u.x < 0.0f) u.x = __fadd_rn(u.x, 1.0f);
        if (u.y < 0.0f) u.y = __fadd_rn(u.y, 1.0f);
        if (u.z < 0.0f) u.z = __fadd_rn(u.z, 1.0f);
        if (u.w < 0.0f) u.w = __fadd_rn(u.w, 1.0f);
        prns[dst] = u;
        dst += instances;
    }
    idx.x = __uint_as_float(I97); idx.y = __uint_as_float(J97); idx.z = c;
    table[size_t(i) + 97u * size_t(instances)] = idx;
}
template <class P>
__global__ void __launch_bounds__(128)
small_produce_kernel(uint32_t* __restrict__ table, float4* __restrict__ prns, uint32_t instances, uint32_t samples) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= instances) return;
    typename P::Args a = { table, instances, 0u };
    Geom dummy = {};
    P p(a, dummy, i, 0u);
    size_t dst = i;
#pragma unroll 1
    for (uint32_t s = 0; s < samples; s++) {
        F4 v = p.gen.next4();
        prns[dst] = make_float4(v.x, v.y, v.z, v.w);
        dst += instances;
    }
    p.store();
}

}  // namespace qcd
