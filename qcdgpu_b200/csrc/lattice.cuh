// lattice.cuh -- geometry, SoA link table access and SU(2)/SU(3) algebra for the
// sm_100a gauge-update kernels.
//
// Layout contract (must stay bit-compatible with the reference):
//   site index   gid = y + z*N2 + t*N2*N3 + x*N2*N3*N4          (SU(N)/suncl/sun_common.cl:41-66)
//   link table   element (gid + (dir + 4*k)*ROWSIZE) of a vec4 array, k = vec4 number
//                SU(2): 1 vec4 (Re u1, Re u2, Im u1, Im u2)         (su2cl.cl:62-79)
//                SU(3): uv1=(Re u1,Re u2,Re u3,Re v3) uv2=(Im ...) uv3=(Re v1,Re v2,Im v1,Im v2)
//                                                                  (su3cl.cl:137-163, su3_matrix_memory.cl:43-55)
//   vec4 = float4 (16 B) for PRECISION 0/2, double4 (32 B = one DRAM sector) for PRECISION 1.
//
// Everything here is new code for Blackwell: fp64 planes move with single 256-bit
// LDG/STG (ld.global.v4.f64, SASS LDG.E.256), fp32 planes with 128-bit ones; the
// algebra is written for FMA contraction and for the 2-row storage (third row =
// conj(u x v)), not transcribed from the OpenCL polynomials.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qcd {

// ---------------------------------------------------------------- geometry
struct Geom {
    uint32_t n1, n2, n3, n4;      // x, y, z, t extents of the (sub)lattice as stored (BIGLAT: incl. 2 halo x-slices)
    uint32_t n23, n234, sites;    // products
    uint32_t rowsize;             // ROWSIZE  (suncl.cpp:2272)
    uint32_t prngstep;            // PRNGSTEP (suncl.cpp:2273)
    uint32_t nthreads;            // update threads: SITES/2, BIGLAT: (N1-2)*N2N3N4/2  (suncl.cl:404-418)
    uint32_t biglat;              // -D BIGLAT
    uint32_t left_sites;          // LEFT_SITES
    uint32_t full_sites;          // FULL_SITES
    uint32_t nhit, nhitpar;
    uint32_t tw_lo, tw_hi;        // BIGLAT twist test: x + left_sites/n234 == tw_lo || == tw_hi
    uint32_t tile_x;              // L2 tiling of the update passes: x-chunk width (0 = reference order)
    uint32_t debug_skip;          // diagnostics (env QCDGPU_DEBUG_SKIP): 1 = skip the generator phase, 2 = skip the physics phase
    uint32_t fold_gs;             // SU(3): reunitarise the updated link a second time before it is stored -- what the sweep's separate
                                  // lattice_GramSchmidt pass would do to it later; the backend then skips that pass (backend.cu, gs_cover)
    uint32_t upd_x0, upd_nx;      // x-range of this launch inside the updated region (halo/interior split); nthreads = upd_nx*N4*N2/2*N3
    uint32_t upd_xstride;         // distance between the launch's slices (1; Nx - 1 for the two boundary slices in one launch; tile_x == 0 then)
    // BIGLAT, halo exchange fused into the update (qgpu_kernel_set_peer_halo): an updated link of my first / last real x-slice is ALSO
    // stored into the halo slice of the left / right neighbour's link table, a pointer into peer memory (NVLink).  peer_on == 0: none.
    uint32_t peer_on;
    uint32_t peer_lo[2];          // first site index of the source slice (my first real slice, my last real slice)
    uint32_t peer_shift[2];       // added to the site index (mod 2^32): (neighbour's halo slice - source slice) * N2N3N4
    uint32_t peer_rowsize[2];     // ROWSIZE of the neighbour's table
    unsigned long long peer_table[2];   // 0: no neighbour on that side
};

struct Coord { uint32_t x, y, z, t; };

__device__ __forceinline__ Coord gid_to_coords(const Geom& g, uint32_t gdi) {
    // sun_common.cl:41-61
    Coord c;
    uint32_t z4 = gdi / g.n234;
    uint32_t z1 = gdi - z4 * g.n234;
    uint32_t z3 = z1 / g.n23;
    z1 = z1 - z3 * g.n23;
    uint32_t z2 = z1 / g.n2;
    z1 = z1 - z2 * g.n2;
    c.x = z4; c.y = z1; c.z = z2; c.t = z3;
    return c;
}
__device__ __forceinline__ uint32_t coords_to_gid(const Geom& g, const Coord& c) {
    return c.y + c.z * g.n2 + c.t * g.n23 + c.x * g.n234;
}
// checkerboard maps, sun_common.cl:351-417 (BIGLAT variants add one x-slice)
__device__ __forceinline__ uint32_t parity_gid(const Geom& g, uint32_t GID, bool odd) {
    uint32_t gde = 2u * GID + (odd ? 1u : 0u);
    Coord c = gid_to_coords(g, gde);
    uint32_t par = (c.x + c.y + c.z + c.t) & 1u;
    uint32_t gindex = odd ? gde - (par ^ 1u) : gde + par;
    if (g.biglat) gindex += g.n234;
    return gindex;
}
// L2-friendly launch order.  The reference's thread index GID enumerates sites with x slowest
// (GID = y/2 + N2/2 (z + N3 (t + N4 x))), so the x+-1 neighbours of a site are touched a whole
// N2 N3 N4 volume later -- at 64^4 that is 100 MB of links in between and they have left the L2 by then
// (measured: 1.75x the compulsory DRAM reads).  Thread k of the launch instead works on logical index
// GID = remap(k): x is cut into chunks of `tile_x` slices and inside a chunk t is the slow loop and x the
// next one, so the window that must stay cached is (tile_x + 2) x 3 (y,z)-planes.  Which thread updates a site
// changes, the site's GID (hence its PRNG instance / PRN slots) does not.
__device__ __forceinline__ uint32_t remap_gid(const Geom& g, uint32_t k) {
    const uint32_t P = (g.n2 >> 1) * g.n3;                 // threads per (x,t) plane
    if (g.tile_x == 0) {                                   // reference order, shifted to the launch's x-range
        if (g.upd_xstride == 1u) return k + g.upd_x0 * g.n4 * P;
        const uint32_t XP = g.n4 * P, xi = k / XP;
        return (k - xi * XP) + (g.upd_x0 + xi * g.upd_xstride) * XP;
    }
    const uint32_t chunk = g.tile_x * g.n4 * P;
    const uint32_t xc = k / chunk;
    uint32_t r = k - xc * chunk;
    const uint32_t x0 = xc * g.tile_x;
    const uint32_t w = min(g.tile_x, g.upd_nx - x0);
    const uint32_t wp = w * P;
    const uint32_t t = r / wp;
    r -= t * wp;
    const uint32_t xi = r / P;
    const uint32_t yz = r - xi * P;
    return yz + P * (t + g.n4 * (g.upd_x0 + x0 + xi));
}
// remap_gid + parity_gid + gid_to_coords in one go: the tiled thread index already splits into (x, t, yz), so the site's
// GID, lattice index and coordinates cost 4 integer divisions instead of 9.  Same values as the three functions above
// (2 GID + odd = (2 yh + odd) + N2 z + N2 N3 (t + N4 x); the checkerboard correction moves y by at most one inside its pair).
struct SiteId { uint32_t GID, gindex; Coord p; };
__device__ __forceinline__ SiteId thread_site(const Geom& g, uint32_t k, bool odd) {
    const uint32_t H = g.n2 >> 1, P = H * g.n3;              // threads per y-row, per (x,t) plane
    uint32_t x, t, yz;
    if (g.tile_x == 0) {
        const uint32_t XP = g.n4 * P, xi = k / XP;
        x = g.upd_x0 + xi * g.upd_xstride;
        const uint32_t r = k - xi * XP;
        t = r / P;
        yz = r - t * P;
    } else {
        const uint32_t chunk = g.tile_x * g.n4 * P;
        const uint32_t xc = k / chunk;
        uint32_t r = k - xc * chunk;
        const uint32_t x0 = xc * g.tile_x;
        const uint32_t wp = min(g.tile_x, g.upd_nx - x0) * P;
        t = r / wp;
        r -= t * wp;
        const uint32_t xi = r / P;
        yz = r - xi * P;
        x = g.upd_x0 + x0 + xi;
    }
    const uint32_t z = yz / H, yh = yz - z * H;
    SiteId s;
    s.GID = yz + P * (t + g.n4 * x);
    const uint32_t o = odd ? 1u : 0u;
    const uint32_t y0 = 2u * yh + o;
    const uint32_t par = (x + y0 + z + t) & 1u;
    const int adj = odd ? -int(par ^ 1u) : int(par);
    s.gindex = uint32_t(int(2u * s.GID + o) + adj);
    s.p.x = x; s.p.y = uint32_t(int(y0) + adj); s.p.z = z; s.p.t = t;
    if (g.biglat) { s.gindex += g.n234; s.p.x += 1u; }
    return s;
}
template <int DIR> __device__ __forceinline__ Coord nb_plus(const Geom& g, Coord c) {
    // sun_common.cl:455-483
    if (DIR == 0) { c.x++; if (!g.biglat && c.x >= g.n1) c.x = 0; }
    if (DIR == 1) { c.y++; if (c.y >= g.n2) c.y = 0; }
    if (DIR == 2) { c.z++; if (c.z >= g.n3) c.z = 0; }
    if (DIR == 3) { c.t++; if (c.t >= g.n4) c.t = 0; }
    return c;
}
template <int DIR> __device__ __forceinline__ Coord nb_minus(const Geom& g, Coord c) {
    // sun_common.cl:485-513
    if (DIR == 0) { if (c.x == 0) c.x = g.n1; c.x--; }
    if (DIR == 1) { if (c.y == 0) c.y = g.n2; c.y--; }
    if (DIR == 2) { if (c.z == 0) c.z = g.n3; c.z--; }
    if (DIR == 3) { if (c.t == 0) c.t = g.n4; c.t--; }
    return c;
}
// run-time direction variants (used where a loop over planes keeps the code size down)
__device__ __forceinline__ Coord nb_plus_rt(const Geom& g, const Coord& c, int dir) {
    switch (dir) { case 0: return nb_plus<0>(g, c); case 1: return nb_plus<1>(g, c); case 2: return nb_plus<2>(g, c); default: return nb_plus<3>(g, c); }
}
__device__ __forceinline__ Coord nb_minus_rt(const Geom& g, const Coord& c, int dir) {
    switch (dir) { case 0: return nb_minus<0>(g, c); case 1: return nb_minus<1>(g, c); case 2: return nb_minus<2>(g, c); default: return nb_minus<3>(g, c); }
}
__device__ __forceinline__ bool twisted_x(const Geom& g, uint32_t x) {
    // su3_matrix_memory.cl:57-63
    if (g.biglat) { uint32_t xg = x + g.tw_lo; return (xg == 0u) || (xg == g.tw_hi); }
    return x == g.n1 - 1u;
}

// ---------------------------------------------------------------- vec4 memory access
template <typename R> struct V4 { R x, y, z, w; };

__device__ __forceinline__ V4<double> ldv4(const double* p) {
    V4<double> v;   // one 256-bit load = one 32-byte sector (sm_100+: LDG.E.256)
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ V4<float> ldv4(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    V4<float> v = { t.x, t.y, t.z, t.w };
    return v;
}
__device__ __forceinline__ void stv4(double* p, const V4<double>& v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void stv4(float* p, const V4<float>& v) {
    *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
}
template <typename R>
__device__ __forceinline__ V4<R> ld_plane(const R* table, const Geom& g, uint32_t gindex, uint32_t plane) {
    return ldv4(table + 4 * (size_t(gindex) + size_t(plane) * g.rowsize));
}
template <typename R>
__device__ __forceinline__ void st_plane(R* table, const Geom& g, uint32_t gindex, uint32_t plane, const V4<R>& v) {
    stv4(table + 4 * (size_t(gindex) + size_t(plane) * g.rowsize), v);
}

// the same vec4 into the neighbours' halo slices when this site lies in a boundary slice (Geom::peer_*); callers test QCD_PEER_ON(g) first
#ifdef QCD_NO_PEER_STORES      // measurement knob: the update kernels without the peer-store branch
#define QCD_PEER_ON(g) false
#else
#define QCD_PEER_ON(g) ((g).peer_on != 0u)
#endif
template <typename R>
__device__ __forceinline__ void st_plane_peer(const Geom& g, uint32_t gindex, uint32_t plane, const V4<R>& v) {
#pragma unroll
    for (int s = 0; s < 2; s++)
        if (g.peer_table[s] != 0ull && (gindex - g.peer_lo[s]) < g.n234)
            stv4(reinterpret_cast<R*>(g.peer_table[s]) + 4 * (size_t(gindex + g.peer_shift[s]) + size_t(plane) * g.peer_rowsize[s]), v);
}

// ---------------------------------------------------------------- complex + matrices
template <typename R> struct Cx { R re, im; };

template <typename R> __device__ __forceinline__ Cx<R> cmul(Cx<R> a, Cx<R> b) {
    return { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
}
template <typename R> __device__ __forceinline__ Cx<R> cmulc(Cx<R> a, Cx<R> b) {   // a * conj(b)
    return { a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im };
}
template <typename R> __device__ __forceinline__ void cfma(Cx<R>& acc, Cx<R> a, Cx<R> b) {       // acc += a*b
    acc.re += a.re * b.re; acc.re -= a.im * b.im; acc.im += a.re * b.im; acc.im += a.im * b.re;
}
template <typename R> __device__ __forceinline__ void cfmac(Cx<R>& acc, Cx<R> a, Cx<R> b) {      // acc += a*conj(b)
    acc.re += a.re * b.re; acc.re += a.im * b.im; acc.im += a.im * b.re; acc.im -= a.re * b.im;
}
template <typename R> __device__ __forceinline__ void cfmacc(Cx<R>& acc, Cx<R> a, Cx<R> b) {     // acc += conj(a)*b
    acc.re += a.re * b.re; acc.re += a.im * b.im; acc.im += a.re * b.im; acc.im -= a.im * b.re;
}

// 3x3 complex matrix, rows r[0..2], fully in registers
template <typename R> struct M3 { Cx<R> e[3][3]; };
// two stored rows (what the table holds)
template <typename R> struct M3r2 { Cx<R> e[2][3]; };

// third row of an SU(3) matrix from its first two: w = conj(u x v)   (su3cl.cl:156-161)
template <typename R>
__device__ __forceinline__ void third_row(const Cx<R> u[3], const Cx<R> v[3], Cx<R> w[3]) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        // conj(u_j v_k - u_k v_j)
        w[i].re =  u[j].re * v[k].re - u[j].im * v[k].im - u[k].re * v[j].re + u[k].im * v[j].im;
        w[i].im = -u[j].re * v[k].im - u[j].im * v[k].re + u[k].re * v[j].im + u[k].im * v[j].re;
    }
}
// w += conj(u x v): the third row of a product added to the staple sum in the same fused operations
template <typename R>
__device__ __forceinline__ void third_row_add(const Cx<R> u[3], const Cx<R> v[3], Cx<R> w[3]) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        w[i].re += u[j].re * v[k].re; w[i].re -= u[j].im * v[k].im; w[i].re -= u[k].re * v[j].re; w[i].re += u[k].im * v[j].im;
        w[i].im -= u[j].re * v[k].im; w[i].im -= u[j].im * v[k].re; w[i].im += u[k].re * v[j].im; w[i].im += u[k].im * v[j].re;
    }
}
template <typename R> __device__ __forceinline__ M3<R> complete(const M3r2<R>& a) {
    M3<R> m;
#pragma unroll
    for (int j = 0; j < 3; j++) { m.e[0][j] = a.e[0][j]; m.e[1][j] = a.e[1][j]; }
    third_row(m.e[0], m.e[1], m.e[2]);
    return m;
}

// full 3x3 product
template <typename R> __device__ __forceinline__ M3<R> mul_full(const M3<R>& U, const M3<R>& S) {
    M3<R> W;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<R> acc = cmul(U.e[i][0], S.e[0][j]);
            cfma(acc, U.e[i][1], S.e[1][j]);
            cfma(acc, U.e[i][2], S.e[2][j]);
            W.e[i][j] = acc;
        }
    return W;
}

// packed <-> rows
template <typename R>
__device__ __forceinline__ M3r2<R> unpack3(const V4<R>& a, const V4<R>& b, const V4<R>& c) {
    M3r2<R> m;
    m.e[0][0] = { a.x, b.x }; m.e[0][1] = { a.y, b.y }; m.e[0][2] = { a.z, b.z };
    m.e[1][0] = { c.x, c.z }; m.e[1][1] = { c.y, c.w }; m.e[1][2] = { a.w, b.w };
    return m;
}
template <typename R>
__device__ __forceinline__ void pack3(const Cx<R> u[3], const Cx<R> v[3], V4<R>& a, V4<R>& b, V4<R>& c) {
    a = { u[0].re, u[1].re, u[2].re, v[2].re };
    b = { u[0].im, u[1].im, u[2].im, v[2].im };
    c = { v[0].re, v[1].re, v[0].im, v[1].im };
}

// SU(2) element as quaternion-like 4-vector in the reference's packing
// (x = Re u1, y = Re u2, z = Im u1, w = Im u2); product su2_matrix_memory.cl:40-53
template <typename R> __device__ __forceinline__ V4<R> su2_mul(const V4<R>& u, const V4<R>& v) {
    V4<R> a;
    a.x = -u.w * v.w + u.x * v.x - u.y * v.y - u.z * v.z;
    a.z =  u.y * v.w + u.z * v.x - u.w * v.y + u.x * v.z;
    a.y = -u.z * v.w + u.y * v.x + u.x * v.y + u.w * v.z;
    a.w =  u.x * v.w + u.w * v.x + u.z * v.y - u.y * v.z;
    return a;
}
template <typename R> __device__ __forceinline__ V4<R> su2_dag(const V4<R>& u) { return { u.x, -u.y, -u.z, -u.w }; }

// ---------------------------------------------------------------- twisted loads
template <typename R> struct Twist { R c1, s1, c2, s2; };   // cos/sin((phi+omega)/2), cos/sin((omega-phi)/2); SU(2): c1,s1 of phi/2

// twist of a Y-link at the x boundary: the reference multiplies (uv1,uv2) lanes xyz by e^{i(phi+omega)/2} and lane w (= v3)
// by e^{i(omega-phi)/2}, uv3 (= v1,v2) by e^{i(omega-phi)/2}  (su3_matrix_memory.cl:64-97)
template <typename R>
__device__ __forceinline__ void twist3(M3r2<R>& m, const Geom& g, uint32_t x, const Twist<R>& tw) {
    if (twisted_x(g, x)) {
        const Cx<R> p1 = { tw.c1, tw.s1 }, p2 = { tw.c2, tw.s2 };
#pragma unroll
        for (int j = 0; j < 3; j++) { m.e[0][j] = cmul(m.e[0][j], p1); m.e[1][j] = cmul(m.e[1][j], p2); }
    }
}
// lattice_table_3 (su3_matrix_memory.cl:43-115): rows of diag(e^{i(phi+omega)/2}, e^{i(omega-phi)/2}, e^{-i omega}) * U
template <typename R, int DIR, bool TBC>
__device__ __forceinline__ M3r2<R> load3(const R* table, const Geom& g, uint32_t gindex, uint32_t x, const Twist<R>& tw) {
    V4<R> a = ld_plane(table, g, gindex, DIR + 0);
    V4<R> b = ld_plane(table, g, gindex, DIR + 4);
    V4<R> c = ld_plane(table, g, gindex, DIR + 8);
    M3r2<R> m = unpack3(a, b, c);
    if (TBC && DIR == 1) twist3(m, g, x, tw);
    return m;
}
// lattice_table_2 (su2_matrix_memory.cl:55-96): Omega * U with Omega = (cos phi/2, 0, sin phi/2, 0)
template <typename R, int DIR, bool TBC>
__device__ __forceinline__ V4<R> load2(const R* table, const Geom& g, uint32_t gindex, uint32_t x, const Twist<R>& tw) {
    V4<R> m = ld_plane(table, g, gindex, DIR);
    if (TBC && DIR == 1) {
        if (twisted_x(g, x)) {
            V4<R> om = { tw.c1, R(0), tw.s1, R(0) };
            m = su2_mul(om, m);
        }
    }
    return m;
}
template <typename R, bool TBC>
__device__ __forceinline__ M3r2<R> load3_rt(const R* table, const Geom& g, uint32_t gindex, uint32_t x, int dir, const Twist<R>& tw) {
    V4<R> a = ld_plane(table, g, gindex, dir + 0);
    V4<R> b = ld_plane(table, g, gindex, dir + 4);
    V4<R> c = ld_plane(table, g, gindex, dir + 8);
    M3r2<R> m = unpack3(a, b, c);
    if (TBC) {
        if (dir == 1 && twisted_x(g, x)) {
            const Cx<R> p1 = { tw.c1, tw.s1 }, p2 = { tw.c2, tw.s2 };
#pragma unroll
            for (int j = 0; j < 3; j++) { m.e[0][j] = cmul(m.e[0][j], p1); m.e[1][j] = cmul(m.e[1][j], p2); }
        }
    }
    return m;
}
template <typename R, bool TBC>
__device__ __forceinline__ V4<R> load2_rt(const R* table, const Geom& g, uint32_t gindex, uint32_t x, int dir, const Twist<R>& tw) {
    V4<R> m = ld_plane(table, g, gindex, dir);
    if (TBC) {
        if (dir == 1 && twisted_x(g, x)) {
            V4<R> om = { tw.c1, R(0), tw.s1, R(0) };
            m = su2_mul(om, m);
        }
    }
    return m;
}
template <typename R, int DIR>
__device__ __forceinline__ void store3(R* table, const Geom& g, uint32_t gindex, const Cx<R> u[3], const Cx<R> v[3]) {
    V4<R> a, b, c;
    pack3(u, v, a, b, c);
    st_plane(table, g, gindex, DIR + 0, a);
    st_plane(table, g, gindex, DIR + 4, b);
    st_plane(table, g, gindex, DIR + 8, c);
    if (QCD_PEER_ON(g)) {
        st_plane_peer(g, gindex, DIR + 0, a);
        st_plane_peer(g, gindex, DIR + 4, b);
        st_plane_peer(g, gindex, DIR + 8, c);
    }
}

// ---------------------------------------------------------------- Gram-Schmidt (su3_update_cl.cl:228-291)
template <typename R> __device__ __forceinline__ R rsqrt_r(R x);
// 1 / sqrt(x) and 1 / x for positive normal x (norms of rows of SU(3) matrices, |det| of a 2x2 block): hardware seed (2^-22) + two
// Newton steps, about a third of the instructions of sqrt + IEEE division and no slow path; relative error ~1e-16
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
}
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0); y = fma(y, e, y);
    e = fma(-x, y, 1.0); y = fma(y, e, y);
    return y;
}
template <> __device__ __forceinline__ double rsqrt_r<double>(double x) { return fast_rsqrt(x); }
template <> __device__ __forceinline__ float  rsqrt_r<float>(float x)   { return 1.0f / sqrtf(x); }

template <typename R>
__device__ __forceinline__ void gram_schmidt3(Cx<R> u[3], Cx<R> v[3]) {
    R nu = R(0);
#pragma unroll
    for (int j = 0; j < 3; j++) nu += u[j].re * u[j].re + u[j].im * u[j].im;
    R inu = rsqrt_r(nu);
#pragma unroll
    for (int j = 0; j < 3; j++) { u[j].re *= inu; u[j].im *= inu; }
    Cx<R> sc = { R(0), R(0) };   // v . conj(u)
#pragma unroll
    for (int j = 0; j < 3; j++) cfmac(sc, v[j], u[j]);
    R nv = R(0);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Cx<R> t = cmul(u[j], sc);
        v[j].re -= t.re; v[j].im -= t.im;
        nv += v[j].re * v[j].re + v[j].im * v[j].im;
    }
    R inv = rsqrt_r(nv);
#pragma unroll
    for (int j = 0; j < 3; j++) { v[j].re *= inv; v[j].im *= inv; }
}

}  // namespace qcd
