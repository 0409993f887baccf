// update.cuh -- one (parity, direction) pass of the Cabibbo-Marinari pseudo-heat-bath
// with Kennedy-Pendleton SU(2) subgroup sampling, plus overrelaxation, for sm_100a.
//
// Replaces the reference kernels update_{even,odd}_{X,Y,Z,T} (SU(N)/suncl/suncl.cl:398-977)
// and everything they inline: lattice_staple_{2,3} (su2_update_cl.cl:206-431,
// su3_update_cl.cl:341-565), lattice_heatbath2/3 (su3_update_cl.cl:568-779,
// su2_update_cl.cl:434-550), lattice_GramSchmidt3 (su3_update_cl.cl:228-291).
//
// Same inputs -> same links within fp tolerance; PRN consumption order is identical
// (one vec4 per KP trial, one for the direction, SU(3) burns one slot on NHIT
// rejections, SU(2) keeps the old link instead).  What is different from the reference:
//   * staples are built from the two stored rows (third rows only where a product needs
//     them: 2x(2-row x full) products + cross products instead of 2 full 3x3 products);
//   * W = U*S is formed once and updated together with U by each SU(2) subgroup rotation
//     (the reference recomputes the full 3x3 product three times);
//   * the PRN source is a template parameter: the reference-layout buffer, the closed-form
//     INTS=2 pseudo-randoms, or a generator whose state lives in registers (fused mode).
#pragma once
#include <math_constants.h>
#include "lattice.cuh"
#include "prng.cuh"

#ifndef QCD_BUF_MINBLOCKS
#define QCD_BUF_MINBLOCKS 3
#endif

namespace qcd {

// ---------------------------------------------------------------- PRN sources
// Reference-layout buffer: thread GID reads prns[GID + k*PRNGSTEP]  (su3_update_cl.cl:613-617, 679)
struct PrnBuffered {
    struct Args { const float4* prns; };
    // (keeping the next two slots in flight in registers was measured slower: 3.10 vs 2.97 ms per 64^4 fp64 pass)
    const float4* p; uint32_t ind, step;
    __device__ __forceinline__ PrnBuffered(const Args& a, const Geom& g, uint32_t GID, uint32_t) : p(a.prns), ind(GID), step(g.prngstep) {}
    __device__ __forceinline__ F4 trial() { float4 r = __ldg(p + ind); ind += step; return { r.x, r.y, r.z, r.w }; }
    __device__ __forceinline__ F4 direction() { return trial(); }
    __device__ __forceinline__ F4 peek() { float4 r = __ldg(p + ind); return { r.x, r.y, r.z, r.w }; }     // next slot, not consumed
    __device__ __forceinline__ void burn() { ind += step; }
    __device__ __forceinline__ void finish() {}
    static constexpr bool kGid = false;
    static constexpr bool kCanPeek = true;
    static constexpr int kMinBlocks = QCD_BUF_MINBLOCKS;
};
// INTS=2 / -D GID_UPD closed-form pseudo-randoms (su3_update_cl.cl:598-637); every trial is accepted
struct PrnGid {
    struct Args { int dummy; };
    double gid, sites, a;
    __device__ __forceinline__ PrnGid(const Args&, const Geom& g, uint32_t, uint32_t gindex) {
        // suncl.cl:424-430: prns[GID].x = (float) gindex  (BIGLAT: global index)
        uint32_t gi = g.biglat ? (gindex - g.n234 + g.left_sites) : gindex;
        gid = double(float(gi));
        sites = g.biglat ? double(g.full_sites) : double(g.sites);
        a = 0.005 * double(1 + int(g.nhit));
    }
    __device__ __forceinline__ F4 trial() {   // values returned as float are only used by the fp32 path
        return { 0.f, 0.f, 0.f, 0.f };
    }
    __device__ __forceinline__ void trial_d(double r[4]) {
        r[0] = fabs(sin((a + 270.0 / sites) * gid)); r[1] = fabs(cos((a + 60.0 / sites) * gid));
        r[2] = fabs(sin((a - 150.0 / sites) * gid)); r[3] = fabs(cos((a - 380.0 / sites) * gid));
    }
    __device__ __forceinline__ void direction_d(double r[4]) {
        r[0] = fabs(cos((0.08 - 270.0 / sites) * gid)); r[1] = fabs(sin((0.08 - 60.0 / sites) * gid));
        r[2] = fabs(cos((0.08 + 150.0 / sites) * gid)); r[3] = fabs(sin((0.08 + 380.0 / sites) * gid));
    }
    __device__ __forceinline__ F4 direction() { return { 0.f, 0.f, 0.f, 0.f }; }
    __device__ __forceinline__ void burn() {}
    __device__ __forceinline__ void finish() {}
    static constexpr bool kGid = true;
    static constexpr bool kCanPeek = false;
    static constexpr int kMinBlocks = 2;
};
// Fused generators: instance GID lives in the registers of update thread GID.  Bit-exact with the
// buffered path when the run is configured with PRNG_instances == PRNGSTEP (hard part 1):
// every pass advances every instance by exactly `samples` vec4, consumed or not.
template <int NSKIP>
struct PrnRanlux {
    struct Args { uint32_t* table; uint32_t instances; uint32_t samples; };
    Ranlux r; uint32_t used; uint32_t* table; uint32_t instances, samples, id;
    __device__ __forceinline__ PrnRanlux(const Args& a, const Geom&, uint32_t GID, uint32_t)
        : used(0), table(a.table), instances(a.instances), samples(a.samples), id(GID) { ranlux_load(r, table, instances, id); }
    __device__ __forceinline__ F4 trial() { used++; return r.template next4<NSKIP>(); }
    __device__ __forceinline__ F4 direction() { used++; return r.template next4<NSKIP>(); }
    __device__ __forceinline__ void burn() { used++; r.template skip4<NSKIP>(); }
    __device__ __forceinline__ void finish() { r.template skip_samples<NSKIP>(samples - used); ranlux_store(r, table, instances, id); }
    static constexpr bool kGid = false;
    static constexpr bool kCanPeek = false;
    static constexpr int kMinBlocks = 2;
};
// RANLUX, two-phase: phase 1 of the kernel runs the whole generator work of the pass (all `samples` vec4,
// luxury skips included) and parks the first `keep` vec4 in shared memory; phase 2 (the fp64/fp32 physics) reads them
// from there.  The 26 state registers are dead during phase 2, which is what lets three blocks share an SM.
// A link consumes 6-8 vec4 on average and at most 3 (NHIT+1) NHITPar = 33: `keep` defaults to 12 (update_inst.cu).  The
// advanced generator state is written to a SECOND table (the backend swaps the two after the pass), so the state in front
// of the pass stays intact and a thread that needs more than `keep` samples regenerates them from it (ranlux_overflow;
// probability ~1e-5 per link at beta = 6).  Shared memory per block: 24.6 KB instead of 67.6 KB.
template <int NSKIP>
struct PrnSmemT {
    struct Args { int dummy; };
    const uint4* p; uint32_t k, stride, keep; unsigned long long converted; const uint32_t* state; uint32_t instances;
    __device__ __forceinline__ PrnSmemT(const uint4* base, uint32_t stride_, unsigned long long conv, uint32_t keep_,
                                        const uint32_t* state_, uint32_t instances_)
        : p(base), k(0), stride(stride_), keep(keep_), converted(conv), state(state_), instances(instances_) {}
    // parked samples are raw seeds << 8 unless flagged as already converted (RanluxPk::pass, prng.cuh)
    __device__ __forceinline__ F4 get(uint32_t i) const {
        if (i >= keep) {
            float4 v = ranlux_overflow<NSKIP>(state, instances, i);
            return { v.x, v.y, v.z, v.w };
        }
        uint4 r = p[i * stride];
        if ((converted >> i) & 1ull) return { __uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w) };
        const float c = 2.3283064365386963e-10f;      // (seed << 8) * 2^-32 = seed * 2^-24, exact
        return { __uint2float_rn(r.x) * c, __uint2float_rn(r.y) * c, __uint2float_rn(r.z) * c, __uint2float_rn(r.w) * c };
    }
    __device__ __forceinline__ F4 trial() { return get(k++); }
    __device__ __forceinline__ F4 direction() { return get(k++); }
    __device__ __forceinline__ F4 peek() const { return get(k); }
    __device__ __forceinline__ void burn() { k++; }
    __device__ __forceinline__ void finish() {}
    // parked sample i WITHOUT a branch (selects); a caller that is past the parked samples discards the value (cm_round_pipelined)
    __device__ __forceinline__ F4 get_parked(uint32_t i) const {
        i = min(i, keep - 1u);
        const uint4 r = p[i * stride];
        const bool cv = ((converted >> i) & 1ull) != 0ull;
        const float c = 2.3283064365386963e-10f;
        F4 o;
        o.x = cv ? __uint_as_float(r.x) : __uint2float_rn(r.x) * c;
        o.y = cv ? __uint_as_float(r.y) : __uint2float_rn(r.y) * c;
        o.z = cv ? __uint_as_float(r.z) : __uint2float_rn(r.z) * c;
        o.w = cv ? __uint_as_float(r.w) : __uint2float_rn(r.w) * c;
        return o;
    }
    static constexpr bool kGid = false;
    static constexpr bool kCanPeek = true;
    static constexpr bool kParked = true;
    static constexpr int kMinBlocks = 3;
};
template <class P, class = void> struct HasParked { static constexpr bool value = false; };
template <class P> struct HasParked<P, decltype((void)P::kParked)> { static constexpr bool value = P::kParked; };
template <class Gen, int WORDS>
struct PrnSmallState {   // XOR128 (4 words), PM (4 words), MRG32k3a / XOR7 / RANECU (2 x uint4 rows)
    struct Args { uint32_t* table; uint32_t instances; uint32_t samples; };
    Gen gen; uint32_t used; uint32_t* table; uint32_t instances, samples, id;
    F4 ahead; bool have_ahead;     // peek(): the next sample, generated but not yet consumed
    __device__ __forceinline__ PrnSmallState(const Args& a, const Geom&, uint32_t GID, uint32_t)
        : used(0), table(a.table), instances(a.instances), samples(a.samples), id(GID), have_ahead(false) { load(); }
    __device__ __forceinline__ void load();
    __device__ __forceinline__ void store();
    __device__ __forceinline__ F4 trial() {
        used++;
        if (have_ahead) { have_ahead = false; return ahead; }
        return gen.next4();
    }
    __device__ __forceinline__ F4 direction() { return trial(); }
    __device__ __forceinline__ F4 peek() { if (!have_ahead) { ahead = gen.next4(); have_ahead = true; } return ahead; }
    __device__ __forceinline__ void burn() { used++; if (have_ahead) have_ahead = false; else gen.skip4(); }
    __device__ __forceinline__ void finish() {
#pragma unroll 1
        for (uint32_t k = used + (have_ahead ? 1u : 0u); k < samples; k++) gen.skip4();
        store();
    }
    static constexpr bool kGid = false;
    static constexpr bool kCanPeek = true;
    static constexpr int kMinBlocks = 3;
};
template <> __device__ __forceinline__ void PrnSmallState<Xor128, 4>::load() {
    uint4 s = reinterpret_cast<const uint4*>(table)[id]; gen.x = s.x; gen.y = s.y; gen.z = s.z; gen.w = s.w;
}
template <> __device__ __forceinline__ void PrnSmallState<Xor128, 4>::store() {
    reinterpret_cast<uint4*>(table)[id] = make_uint4(gen.x, gen.y, gen.z, gen.w);
}
template <> __device__ __forceinline__ void PrnSmallState<ParkMiller, 4>::load() {
    uint4 s = reinterpret_cast<const uint4*>(table)[id]; gen.s[0] = s.x; gen.s[1] = s.y; gen.s[2] = s.z; gen.s[3] = s.w;
}
template <> __device__ __forceinline__ void PrnSmallState<ParkMiller, 4>::store() {
    reinterpret_cast<uint4*>(table)[id] = make_uint4(gen.s[0], gen.s[1], gen.s[2], gen.s[3]);
}
template <> __device__ __forceinline__ void PrnSmallState<Mrg32k3a, 8>::load() {
    uint4 s1 = reinterpret_cast<const uint4*>(table)[id];
    uint4 s2 = reinterpret_cast<const uint4*>(table)[size_t(instances) + id];
    gen.a[0] = s1.x; gen.a[1] = s1.y; gen.a[2] = s1.z; gen.b[0] = s2.x; gen.b[1] = s2.y; gen.b[2] = s2.z;
}
template <> __device__ __forceinline__ void PrnSmallState<Mrg32k3a, 8>::store() {
    uint4* t = reinterpret_cast<uint4*>(table);
    uint4 s1 = t[id], s2 = t[size_t(instances) + id];
    s1.x = gen.a[0]; s1.y = gen.a[1]; s1.z = gen.a[2]; s2.x = gen.b[0]; s2.y = gen.b[1]; s2.z = gen.b[2];
    t[id] = s1; t[size_t(instances) + id] = s2;
}
// XOR7 / RANECU: seed_table[GID] and seed_table[GID + instances] (prngcl_xor7.cl:111-112,137-138; prngcl_ranecu.cl:139-140,160-161)
template <> __device__ __forceinline__ void PrnSmallState<Xor7, 8>::load() {
    uint4 s1 = reinterpret_cast<const uint4*>(table)[id];
    uint4 s2 = reinterpret_cast<const uint4*>(table)[size_t(instances) + id];
    gen.s[0] = s1.x; gen.s[1] = s1.y; gen.s[2] = s1.z; gen.s[3] = s1.w; gen.s[4] = s2.x; gen.s[5] = s2.y; gen.s[6] = s2.z; gen.s[7] = s2.w;
}
template <> __device__ __forceinline__ void PrnSmallState<Xor7, 8>::store() {
    uint4* t = reinterpret_cast<uint4*>(table);
    t[id] = make_uint4(gen.s[0], gen.s[1], gen.s[2], gen.s[3]);
    t[size_t(instances) + id] = make_uint4(gen.s[4], gen.s[5], gen.s[6], gen.s[7]);
}
template <> __device__ __forceinline__ void PrnSmallState<Ranecu, 8>::load() {
    uint4 s1 = reinterpret_cast<const uint4*>(table)[id];
    uint4 s2 = reinterpret_cast<const uint4*>(table)[size_t(instances) + id];
    gen.a[0] = s1.x; gen.a[1] = s1.y; gen.a[2] = s1.z; gen.a[3] = s1.w; gen.b[0] = s2.x; gen.b[1] = s2.y; gen.b[2] = s2.z; gen.b[3] = s2.w;
}
template <> __device__ __forceinline__ void PrnSmallState<Ranecu, 8>::store() {
    uint4* t = reinterpret_cast<uint4*>(table);
    t[id] = make_uint4(gen.a[0], gen.a[1], gen.a[2], gen.a[3]);
    t[size_t(instances) + id] = make_uint4(gen.b[0], gen.b[1], gen.b[2], gen.b[3]);
}
using PrnXor128 = PrnSmallState<Xor128, 4>;
using PrnXor7 = PrnSmallState<Xor7, 8>;
using PrnRanecu = PrnSmallState<Ranecu, 8>;
using PrnPM = PrnSmallState<ParkMiller, 4>;
using PrnMrg = PrnSmallState<Mrg32k3a, 8>;

// ---------------------------------------------------------------- math helpers
template <typename R> struct Mth;
template <> struct Mth<double> {
#ifdef QCD_LIBDEVICE_SQRT
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
#else
    // sqrt of 0 <= x <= 1 (1 - cos^2 of the Kennedy-Pendleton angles: zero or a normal number) WITHOUT a branch: reciprocal-square-root
    // seed + two Newton steps + one residual step; zero through a select.  The library call has a slow-path branch, and a branch ends the
    // basic block: the five independent chains of a Kennedy-Pendleton trial (two logarithms, cos, sqrt, sincos) then run one after the
    // other instead of interleaved, which is what made the Cabibbo-Marinari section 60 % of the kernel's time (profiles/r02_*).
    // (one Newton step leaves 1/sqrt(x) with relative error e ~ 2^-44; the residual step squares it: root error 1.5 e^2)
    static __device__ __forceinline__ double sqrt_(double x) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        const double inv = fma(y, fma(-0.5 * x * y, y, 0.5), y);
        const double r = x * inv;
        const double root = fma(fma(-r, r, x), 0.5 * inv, r);
        return (x == 0.0) ? 0.0 : root;
    }
#endif
    static __device__ __forceinline__ double rcp_(double x) { return fast_rcp(x); }      // positive normal arguments only
    // sqrt(x) and 1 / sqrt(x) of a positive normal x from one reciprocal-square-root (one residual step for the root)
    static __device__ __forceinline__ void sqrt_and_rcp(double x, double& root, double& inv) {
        inv = fast_rsqrt(x);
        const double r = x * inv;
        root = fma(fma(-r, r, x), 0.5 * inv, r);
    }
#ifdef QCD_LIBDEVICE_LOG
    static __device__ __forceinline__ double log_(double x) { return log(x); }
#else
    // log of a positive normal number (the Kennedy-Pendleton arguments are 1 - r, r a float in (0, 1): 2^-24 <= x < 1), error below
    // 1 ulp: the classic argument reduction x = 2^k (1 + f), sqrt(2)/2 < 1 + f < sqrt(2), s = f / (2 + f), log(1 + f) = f - f^2/2 +
    // s (f^2/2 + R(s^2)) with the degree-7 minimax polynomial R of Sun's fdlibm (e_log.c).  libdevice's
    // log costs about three times the instructions; with it the 64^4 pass takes 3.85 instead of 3.67 ms.
    static __device__ __forceinline__ double log_(double x) {
        int hx = __double2hiint(x);
        const int lx = __double2loint(x);
        // zero (XOR128 / PM / the INTS=2 closed forms can deliver r = 1): -inf as the library; an argument 1 - r is never negative or
        // subnormal (those, inf and nan give nan).  No branch (selects at the end): the logarithms of a trial must stay in one basic
        // block with the other chains so that ptxas interleaves them.
        const bool special = hx < 0x00100000 || hx >= 0x7ff00000;
        const double special_value = (x == 0.0) ? -CUDART_INF : CUDART_NAN;
        int k = (hx >> 20) - 1023;
        hx &= 0x000fffff;
        const int i = (hx + 0x95f64) & 0x100000;
        hx |= i ^ 0x3ff00000;                    // normalise x or x / 2
        k += i >> 20;
        const double f = __hiloint2double(hx, lx) - 1.0;
        // s = f / (2 + f), 1.7 < 2 + f < 2.42: reciprocal seed + two Newton steps + one residual correction (no slow path needed)
        const double d = 2.0 + f;
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
        double e = fma(-d, y, 1.0); y = fma(y, e, y);
        e = fma(-d, y, 1.0); y = fma(y, e, y);
        double s = f * y;
        s = fma(fma(-d, s, f), y, s);
        const double dk = double(k);
        const double z = s * s, w = z * z;
        const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
        const double t2 = z * (6.666666666666735130e-01 + w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
        const double R = t2 + t1, hfsq = 0.5 * f * f;
        const double res = dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
        return special ? special_value : res;
    }
#endif
#ifdef QCD_LIBDEVICE_TRIG
    static __device__ __forceinline__ double cos2pi(double x) { return cospi(2.0 * x); }
    static __device__ __forceinline__ void sincos2pi(double x, double* s, double* c) { sincospi(2.0 * x, s, c); }
#else
    // sin and cos of 2 pi x for 0 <= x <= 1 (a float PRN): 4 x = q + r with q the nearest integer, |r| <= 1/2, so the angle is
    // q pi/2 + z with |z| = |r| pi/2 <= pi/4: Taylor series of sin z (to z^17) and cos z (to z^18) in Horner form, truncation below
    // 5e-17, then the quadrant swap.  No argument checks and none of sincospi's generality: cospi + sincospi cost the pass up to
    // 0.18 ms (measured with float stand-ins), this about half of it.
    static __device__ __forceinline__ void sincos_quadrant(double x, double& sn, double& cs, int& q) {
        const double t = 4.0 * x;
        const double qd = rint(t);
        q = int(qd);
        const double z = (t - qd) * 1.5707963267948966192;
        // Both polynomials in Estrin form on w = z^2 (pairs, then w^2 and w^4): five dependent operations deep instead of nine -- the
        // Kennedy-Pendleton step is a chain of dependent FP64 operations run by three warps per scheduler, its depth is its cost.
        const double w = z * z, w2 = w * w, w4 = w2 * w2;
        // sin z = z + z w (s0 + s1 w + ... + s7 w^7),  s_k = (-1)^(k+1) / (2k+3)!
        const double s01 = fma(w, 8.3333333333333333333e-03, -1.6666666666666666667e-01);     //  1/5! w - 1/3!
        const double s23 = fma(w, 2.7557319223985890653e-06, -1.9841269841269841270e-04);     //  1/9! w - 1/7!
        const double s45 = fma(w, 1.6059043836821614599e-10, -2.5052108385441718775e-08);     //  1/13! w - 1/11!
        const double s67 = fma(w, 2.8114572543455207632e-15, -7.6471637318198164759e-13);     //  1/17! w - 1/15!
        const double ps = fma(w4, fma(w2, s67, s45), fma(w2, s23, s01));
        sn = fma(ps * w, z, z);
        // cos z = 1 + w (c0 + c1 w + ... + c8 w^8),  c_k = (-1)^(k+1) / (2k+2)!
        const double c01 = fma(w, 4.1666666666666666667e-02, -0.5);                            //  1/4! w - 1/2!
        const double c23 = fma(w, 2.4801587301587301587e-05, -1.3888888888888888889e-03);     //  1/8! w - 1/6!
        const double c45 = fma(w, 2.0876756987868098979e-09, -2.7557319223985890653e-07);     //  1/12! w - 1/10!
        const double c67 = fma(w, 4.7794773323873852974e-14, -1.1470745597729724714e-11);     //  1/16! w - 1/14!
        const double c8 = -1.5619206968586226462e-16;                                          // -1/18!
        const double pc = fma(w4, fma(w4, c8, fma(w2, c67, c45)), fma(w2, c23, c01));
        cs = fma(pc, w, 1.0);
    }
    static __device__ __forceinline__ double cos2pi(double x) {
        double sn, cs; int q;
        sincos_quadrant(x, sn, cs, q);
        const double v = (q & 1) ? sn : cs;                 // q = 0..4: cos, -sin, -cos, sin, cos
        return ((q + 1) & 2) ? -v : v;
    }
    static __device__ __forceinline__ void sincos2pi(double x, double* s, double* c) {
        double sn, cs; int q;
        sincos_quadrant(x, sn, cs, q);
        const double sv = (q & 1) ? cs : sn, cv = (q & 1) ? sn : cs;   // sin: sin, cos, -sin, -cos; cos: cos, -sin, -cos, sin
        *s = (q & 2) ? -sv : sv;
        *c = ((q + 1) & 2) ? -cv : cv;
    }
#endif
};
template <> struct Mth<float> {
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rcp_(float x) { return 1.0f / x; }
    static __device__ __forceinline__ void sqrt_and_rcp(float x, float& root, float& inv) { root = sqrtf(x); inv = 1.0f / root; }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float cos2pi(float x) { return cospif(2.0f * x); }
    static __device__ __forceinline__ void sincos2pi(float x, float* s, float* c) { sincospif(2.0f * x, s, c); }
};

// Kennedy-Pendleton sampling of the new SU(2) element (reference: lattice_heatbath2).
// aH = (r/sqrt(det r))^+ as a 4-vector, ibdet = 1 / (beta*sqrt(det)).  Returns false after NHIT rejections.
// direction part of the new element: uniform point on S^2 from (r.x, r.y)
template <typename R> struct KpDir { R costh, sinth, sinphi, cosphi; };
template <typename R> __device__ __forceinline__ KpDir<R> kp_direction(R rx, R ry) {
    KpDir<R> d;
    d.costh = R(2) * rx - R(1);
    d.sinth = Mth<R>::sqrt_(R(1) - d.costh * d.costh);
    Mth<R>::sincos2pi(ry, &d.sinphi, &d.cosphi);
    return d;
}
template <typename R> __device__ __forceinline__ V4<R> kp_compose(R delta, const KpDir<R>& d, const V4<R>& aH) {
    R cosal = R(1) - delta;
    R sinal = Mth<R>::sqrt_(R(1) - cosal * cosal);
    V4<R> c;
    c.x = cosal;
    c.z = sinal * d.costh;
    c.y = sinal * d.sinth * d.sinphi;
    c.w = sinal * d.sinth * d.cosphi;
    return su2_mul(c, aH);
}
// The lattice-independent part of a Kennedy-Pendleton trial with sample r, and of the direction with sample r2:
// E = -(log(1 - r.x) + cos^2(2 pi r.y) log(1 - r.z)) (delta = E / (beta sqrt(det))), r.w^2, and the point on the sphere.
template <typename R> struct KpTrial { R E, w2; KpDir<R> D; };
template <typename R> __device__ __forceinline__ KpTrial<R> kp_trial_eval(const F4& r, const F4& r2) {
    KpTrial<R> t;
    R c = Mth<R>::cos2pi(R(r.y));
    R l1 = Mth<R>::log_(R(1) - R(r.x)), l2 = Mth<R>::log_(R(1) - R(r.z));
    t.D = kp_direction<R>(R(r2.x), R(r2.y));
    t.E = -(l1 + c * c * l2);
    t.w2 = R(r.w) * R(r.w);
    return t;
}
// nhit: trials allowed (default: the run's NHIT)
template <typename R, class Prn>
__device__ __forceinline__ bool kp_heatbath(Prn& prn, const Geom& g, const V4<R>& aH, R ibdet, V4<R>& out, uint32_t nhit = 0xffffffffu) {
    if (nhit == 0xffffffffu) nhit = g.nhit;
    if constexpr (Prn::kGid) {
        double r[4];
        prn.trial_d(r);
        R c = Mth<R>::cos2pi(R(r[1]));
        R delta = c * c * ibdet;
        prn.direction_d(r);
        out = kp_compose(delta, kp_direction<R>(R(r[0]), R(r[1])), aH);
        return true;
    } else if constexpr (Prn::kCanPeek) {
        // The slot after a trial is the direction sample if the trial is accepted and the next trial if not:
        // evaluate both roles together (five independent transcendental chains instead of one after the other).
        F4 r = prn.trial();
        for (uint32_t i = 0; i < nhit; i++) {
            F4 r2 = prn.peek();
            const KpTrial<R> t = kp_trial_eval<R>(r, r2);
            R delta = t.E * ibdet;
            if (t.w2 <= R(1) - R(0.5) * delta) {
                prn.burn();                        // the peeked slot is consumed as the direction sample
                out = kp_compose(delta, t.D, aH);
                return true;
            }
            if (i + 1 < nhit) r = prn.trial();   // the peeked slot becomes the next trial
        }
        return false;
    } else {
        bool flag = false;
        R delta = R(0);
        for (uint32_t i = 0; i < nhit && !flag; i++) {
            F4 r = prn.trial();
            R c = Mth<R>::cos2pi(R(r.y));
            delta = -(Mth<R>::log_(R(1) - R(r.x)) + c * c * Mth<R>::log_(R(1) - R(r.z))) * ibdet;
            if (R(r.w) * R(r.w) <= R(1) - R(0.5) * delta) flag = true;
        }
        if (!flag) return false;
        F4 r = prn.direction();
        out = kp_compose(delta, kp_direction<R>(R(r.x), R(r.y)), aH);
        return true;
    }
}

// ---------------------------------------------------------------- SU(3) products on 2-row operands
// D(rows 0,1) = C(rows 0,1) * B(full)
template <typename R> __device__ __forceinline__ void mul_r2_full(const M3r2<R>& C, const M3<R>& B, M3<R>& D) {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<R> acc = cmul(C.e[i][0], B.e[0][j]);
            cfma(acc, C.e[i][1], B.e[1][j]);
            cfma(acc, C.e[i][2], B.e[2][j]);
            D.e[i][j] = acc;
        }
    third_row(D.e[0], D.e[1], D.e[2]);
}
// F(rows 0,1) = A(rows 0,1) * D^+ ;  third row by cross product
template <typename R> __device__ __forceinline__ void mul_r2_dag(const M3r2<R>& A, const M3<R>& D, M3<R>& F) {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<R> acc = cmulc(A.e[i][0], D.e[j][0]);
            cfmac(acc, A.e[i][1], D.e[j][1]);
            cfmac(acc, A.e[i][2], D.e[j][2]);
            F.e[i][j] = acc;
        }
    third_row(F.e[0], F.e[1], F.e[2]);
}
// G(rows 0,1) = E^+ * C ; third row by cross product
template <typename R> __device__ __forceinline__ void mul_dag_full(const M3<R>& E, const M3<R>& C, M3<R>& G) {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<R> acc = { E.e[0][i].re * C.e[0][j].re + E.e[0][i].im * C.e[0][j].im,
                          E.e[0][i].re * C.e[0][j].im - E.e[0][i].im * C.e[0][j].re };
            cfmacc(acc, E.e[1][i], C.e[1][j]);
            cfmacc(acc, E.e[2][i], C.e[2][j]);
            G.e[i][j] = acc;
        }
    third_row(G.e[0], G.e[1], G.e[2]);
}
// the last product of a staple, added to the staple sum S: rows 0, 1 of the product are formed (the cross product needs them) and
// added, the third row goes into S by fused operations (six additions less per staple than forming it first)
template <typename R> __device__ __forceinline__ void mul_r2_dag_add(const M3r2<R>& A, const M3<R>& D, M3<R>& S) {
    Cx<R> F[2][3];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<R> acc = cmulc(A.e[i][0], D.e[j][0]);
            cfmac(acc, A.e[i][1], D.e[j][1]);
            cfmac(acc, A.e[i][2], D.e[j][2]);
            F[i][j] = acc;
            S.e[i][j].re += acc.re; S.e[i][j].im += acc.im;
        }
    third_row_add(F[0], F[1], S.e[2]);
}
template <typename R> __device__ __forceinline__ void mul_dag_full_add(const M3<R>& E, const M3<R>& C, M3<R>& S) {
    Cx<R> G[2][3];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Cx<R> acc = { E.e[0][i].re * C.e[0][j].re + E.e[0][i].im * C.e[0][j].im,
                          E.e[0][i].re * C.e[0][j].im - E.e[0][i].im * C.e[0][j].re };
            cfmacc(acc, E.e[1][i], C.e[1][j]);
            cfmacc(acc, E.e[2][i], C.e[2][j]);
            G[i][j] = acc;
            S.e[i][j].re += acc.re; S.e[i][j].im += acc.im;
        }
    third_row_add(G[0], G[1], S.e[2]);
}
#ifdef QCD_S_SMEM
// staple accumulator parked in shared memory: element e of thread t at sm[e * kThreads + t] (conflict-free)
template <typename R> struct SmemAcc {
    R* p;
    __device__ __forceinline__ void set(const M3<R>& F) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) { p[(6 * i + 2 * j) * 128] = F.e[i][j].re; p[(6 * i + 2 * j + 1) * 128] = F.e[i][j].im; }
    }
    __device__ __forceinline__ void add(const M3<R>& F) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) { p[(6 * i + 2 * j) * 128] += F.e[i][j].re; p[(6 * i + 2 * j + 1) * 128] += F.e[i][j].im; }
    }
    __device__ __forceinline__ M3<R> get() const {
        M3<R> S;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) { S.e[i][j].re = p[(6 * i + 2 * j) * 128]; S.e[i][j].im = p[(6 * i + 2 * j + 1) * 128]; }
        return S;
    }
    // (experiment only: forms the product first, so its last bits differ from the register accumulator's fused third row)
    __device__ __forceinline__ void add_r2_dag(const M3r2<R>& A, const M3<R>& D) { M3<R> F; mul_r2_dag(A, D, F); add(F); }
    __device__ __forceinline__ void add_dag_full(const M3<R>& E, const M3<R>& C) { M3<R> G; mul_dag_full(E, C, G); add(G); }
};
#endif
template <typename R> __device__ __forceinline__ void acc3(M3<R>& S, const M3<R>& F) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { S.e[i][j].re += F.e[i][j].re; S.e[i][j].im += F.e[i][j].im; }
}

// The six links of the (MU,NU) plane that enter the staple of U_mu(p): site indices and x coordinates (twist test)
struct PlaneSites { uint32_t g_pmu, g_pnu, g_pmn, g_pmnmu; uint32_t x_pmu, x_pnu, x_pmn, x_pmnmu; };
template <int MU, int NU>
__device__ __forceinline__ PlaneSites plane_sites(const Geom& g, const Coord& p, const Coord& pmu) {
    Coord pnu = nb_plus<NU>(g, p);
    Coord pmn = nb_minus<NU>(g, p);
    Coord pmnmu = nb_plus<MU>(g, pmn);
    PlaneSites s;
    s.g_pmu = coords_to_gid(g, pmu); s.g_pnu = coords_to_gid(g, pnu);
    s.g_pmn = coords_to_gid(g, pmn); s.g_pmnmu = coords_to_gid(g, pmnmu);
    s.x_pmu = pmu.x; s.x_pnu = pnu.x; s.x_pmn = pmn.x; s.x_pmnmu = pmnmu.x;
    return s;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
template <typename R, int DIR, int NPL>
__device__ __forceinline__ void prefetch_link(const R* table, const Geom& g, uint32_t gindex) {
#pragma unroll
    for (int k = 0; k < NPL; k++) prefetch_l2(table + 4 * (size_t(gindex) + size_t(DIR + 4 * k) * g.rowsize));
}
// pull the six links of a plane towards L2 while the previous plane is being multiplied (no registers held)
template <typename R, int MU, int NU, int NPL>
__device__ __forceinline__ void plane_prefetch(const R* table, const Geom& g, uint32_t gindex, const PlaneSites& s) {
    prefetch_link<R, NU, NPL>(table, g, s.g_pmu);
    prefetch_link<R, MU, NPL>(table, g, s.g_pnu);
    prefetch_link<R, NU, NPL>(table, g, gindex);
    prefetch_link<R, NU, NPL>(table, g, s.g_pmnmu);
    prefetch_link<R, MU, NPL>(table, g, s.g_pmn);
    prefetch_link<R, NU, NPL>(table, g, s.g_pmn);
}
// staple contribution of the (MU,NU) plane at site p: forward + backward
template <typename R, int MU, int NU, bool TBC, class Acc>
__device__ __forceinline__ void staple3_plane(const R* table, const Geom& g, uint32_t px, uint32_t gindex,
                                              const PlaneSites& ps, const Twist<R>& tw, Acc& S, bool first) {
    // issue all six link loads of the plane before the arithmetic
    M3r2<R> A  = load3<R, NU, TBC>(table, g, ps.g_pmu, ps.x_pmu, tw);      // U_nu(p+mu)
    M3r2<R> B  = load3<R, MU, TBC>(table, g, ps.g_pnu, ps.x_pnu, tw);      // U_mu(p+nu)
    M3r2<R> C  = load3<R, NU, TBC>(table, g, gindex, px, tw);              // U_nu(p)
    M3r2<R> A2 = load3<R, NU, TBC>(table, g, ps.g_pmnmu, ps.x_pmnmu, tw);  // U_nu(p-nu+mu)
    M3r2<R> B2 = load3<R, MU, TBC>(table, g, ps.g_pmn, ps.x_pmn, tw);      // U_mu(p-nu)
    M3r2<R> C2 = load3<R, NU, TBC>(table, g, ps.g_pmn, ps.x_pmn, tw);      // U_nu(p-nu)
    {   // forward: A B^+ C^+ = A (C B)^+
        M3<R> Bf = complete(B), D;
        mul_r2_full(C, Bf, D);
        if (first) { M3<R> F; mul_r2_dag(A, D, F); S.set(F); } else S.add_r2_dag(A, D);
    }
    {   // backward: A2^+ B2^+ C2 = (B2 A2)^+ C2
        M3<R> Af = complete(A2), Cf = complete(C2), E;
        mul_r2_full(B2, Af, E);
        S.add_dag_full(E, Cf);
    }
}
template <typename R> struct RegAcc {
    M3<R> S;
    __device__ __forceinline__ void set(const M3<R>& F) { S = F; }
    __device__ __forceinline__ void add(const M3<R>& F) { acc3(S, F); }
    __device__ __forceinline__ void add_r2_dag(const M3r2<R>& A, const M3<R>& D) { mul_r2_dag_add(A, D, S); }
    __device__ __forceinline__ void add_dag_full(const M3<R>& E, const M3<R>& C) { mul_dag_full_add(E, C, S); }
    __device__ __forceinline__ M3<R> get() const { return S; }
};
// plane order of the reference: X:{Y,Z,T} Y:{X,Z,T} Z:{Y,X,T} T:{Y,Z,X}  (su3_update_cl.cl:367-559)
template <int MU> struct PlaneOrder {
    static constexpr int n0 = (MU == 1) ? 0 : 1;
    static constexpr int n1 = (MU == 0 || MU == 1 || MU == 3) ? 2 : 0;
    static constexpr int n2 = (MU == 3) ? 0 : 3;
};
// Sum of the six staples, planes unrolled (a run-time loop over the planes was measured 4 % slower: spills
// and run-time plane indexing cost more than the smaller code saves in instruction fetch).
template <typename R, int MU, bool TBC>
__device__ __forceinline__ M3<R> staple3(const R* table, const Geom& g, const Coord& p, uint32_t gindex, const Twist<R>& tw) {
#ifdef QCD_S_SMEM
    __shared__ R s_acc[18 * 128];
    SmemAcc<R> S = { s_acc + threadIdx.x };
#else
    RegAcc<R> S;
#endif
    Coord pmu = nb_plus<MU>(g, p);
    constexpr int n0 = PlaneOrder<MU>::n0, n1 = PlaneOrder<MU>::n1, n2 = PlaneOrder<MU>::n2;
    PlaneSites s0 = plane_sites<MU, n0>(g, p, pmu);
    PlaneSites s1 = plane_sites<MU, n1>(g, p, pmu);
    PlaneSites s2 = plane_sites<MU, n2>(g, p, pmu);
    plane_prefetch<R, MU, n1, 3>(table, g, gindex, s1);
    staple3_plane<R, MU, n0, TBC>(table, g, p.x, gindex, s0, tw, S, true);
    plane_prefetch<R, MU, n2, 3>(table, g, gindex, s2);
    staple3_plane<R, MU, n1, TBC>(table, g, p.x, gindex, s1, tw, S, false);
    staple3_plane<R, MU, n2, TBC>(table, g, p.x, gindex, s2, tw, S, false);
    return S.get();
}
// ---------------------------------------------------------------- SU(3) link staging through shared memory
// Each thread owns a ring of QCD_STAGE_K link slots (3 vec4 each) in shared memory.  The 19 links a site update reads
// (18 staple links + its own) are copied global -> shared by cp.async (LDGSTS) QCD_STAGE_K links ahead of their use, so
// their memory latency is covered by the products of the preceding links without holding any register.
// Layout: 16-byte unit u of thread t at (u * kUpdateThreads + t) * 16: conflict-free LDS.128 / LDGSTS.128.
// Measured (64^4 fp64, PRNs from memory): K=3 3.12 ms, K=6 3.12 ms per pass against 2.97 ms for plain LDG.256 loads --
// each LDGSTS.128 of a warp touches 32 half-used sectors, L1 throughput goes 36 % -> 75 % and eats the latency gain
// (cp.async.cg: 3.73 ms).  Kept as a build-time experiment (QCD_NVCC_EXTRA=-DQCD_STAGE_K=3), off by default.
#ifndef QCD_STAGE_K
#define QCD_STAGE_K 0
#endif
#ifndef QCD_UPDATE_THREADS
#define QCD_UPDATE_THREADS 128
#endif
constexpr int kUpdateThreads = QCD_UPDATE_THREADS;
constexpr int kStageK = QCD_STAGE_K;
template <typename R> constexpr size_t stage_bytes() { return size_t(kStageK) * 3 * (sizeof(V4<R>) / 16) * 16 * kUpdateThreads; }

template <int MU> __host__ __device__ constexpr int plane_nu(int pl) {
    return pl == 0 ? PlaneOrder<MU>::n0 : pl == 1 ? PlaneOrder<MU>::n1 : PlaneOrder<MU>::n2;
}
template <typename R, int MU, bool TBC> struct Su3Stage {
    static constexpr int H = sizeof(V4<R>) / 16;
    static constexpr int NLINKS = 19;
    const char* sm;         // this thread's first unit (generic pointer into shared memory)
    const R* table;
    uint32_t rowsize, gindex, px;
    PlaneSites s0, s1, s2;

    template <int SLOT> __device__ __forceinline__ void copy_link(uint32_t gidx, int dir) const {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const char* src = reinterpret_cast<const char*>(table + 4 * (size_t(gidx) + size_t(dir + 4 * k) * rowsize));
#pragma unroll
            for (int h = 0; h < H; h++)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::
                             "r"(sa + uint32_t(((SLOT * 3 + k) * H + h) * 16 * kUpdateThreads)), "l"(src + 16 * h) : "memory");
        }
    }
    // link L of the site's read sequence: plane L/6, {U_mu(p+nu), U_nu(p), U_nu(p+mu), U_nu(p-nu+mu), U_mu(p-nu), U_nu(p-nu)}; 18 = U_mu(p)
    template <int L> __device__ __forceinline__ void issue() const {
        if constexpr (L < 18) {
            constexpr int pl = L / 6, i = L % 6, NU = plane_nu<MU>(pl);
            const PlaneSites& s = (pl == 0) ? s0 : (pl == 1) ? s1 : s2;
            if constexpr (i == 0) copy_link<L % kStageK>(s.g_pnu, MU);
            if constexpr (i == 1) copy_link<L % kStageK>(gindex, NU);
            if constexpr (i == 2) copy_link<L % kStageK>(s.g_pmu, NU);
            if constexpr (i == 3) copy_link<L % kStageK>(s.g_pmnmu, NU);
            if constexpr (i == 4) copy_link<L % kStageK>(s.g_pmn, MU);
            if constexpr (i == 5) copy_link<L % kStageK>(s.g_pmn, NU);
        } else if constexpr (L == 18) {
            copy_link<L % kStageK>(gindex, MU);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // also for L > 18: one (empty) group per link keeps the wait count fixed
    }
    template <int L> __device__ __forceinline__ M3r2<R> take(const Geom& g, const Twist<R>& tw) const {
        asm volatile("cp.async.wait_group %0;" :: "n"(kStageK - 1) : "memory");
        constexpr int SLOT = L % kStageK;
        V4<R> v[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if constexpr (H == 2) {
                const double2 lo = *reinterpret_cast<const double2*>(sm + ((SLOT * 3 + k) * 2 + 0) * 16 * kUpdateThreads);
                const double2 hi = *reinterpret_cast<const double2*>(sm + ((SLOT * 3 + k) * 2 + 1) * 16 * kUpdateThreads);
                v[k].x = lo.x; v[k].y = lo.y; v[k].z = hi.x; v[k].w = hi.y;
            } else {
                const float4 q = *reinterpret_cast<const float4*>(sm + (SLOT * 3 + k) * 16 * kUpdateThreads);
                v[k].x = q.x; v[k].y = q.y; v[k].z = q.z; v[k].w = q.w;
            }
        }
        M3r2<R> m = unpack3(v[0], v[1], v[2]);
        // which links are Y-links (the only twisted ones) and at which x they sit
        constexpr int pl = (L < 18) ? L / 6 : 0, i = L % 6;
        constexpr int DIR = (L == 18) ? MU : ((i == 0 || i == 4) ? MU : plane_nu<MU>(pl));
        if constexpr (TBC && DIR == 1) {
            const PlaneSites& s = (pl == 0) ? s0 : (pl == 1) ? s1 : s2;
            const uint32_t x = (L == 18 || i == 1) ? px : (i == 0) ? s.x_pnu : (i == 2) ? s.x_pmu : (i == 3) ? s.x_pmnmu : s.x_pmn;
            twist3(m, g, x, tw);
        }
        asm volatile("" ::: "memory");
        issue<L + kStageK>();
        return m;
    }
    __device__ __forceinline__ void begin(const void* smem_thread, const R* tab, const Geom& g, uint32_t gi) {
        sm = reinterpret_cast<const char*>(smem_thread); table = tab; rowsize = g.rowsize; gindex = gi;
        Coord p = gid_to_coords(g, gi);
        px = p.x;
        Coord pmu = nb_plus<MU>(g, p);
        s0 = plane_sites<MU, PlaneOrder<MU>::n0>(g, p, pmu);
        s1 = plane_sites<MU, PlaneOrder<MU>::n1>(g, p, pmu);
        s2 = plane_sites<MU, PlaneOrder<MU>::n2>(g, p, pmu);
        prologue<0>();
    }
    template <int L> __device__ __forceinline__ void prologue() const {
        if constexpr (L < kStageK) { issue<L>(); prologue<L + 1>(); }
    }
    template <int PL> __device__ __forceinline__ void plane(const Geom& g, const Twist<R>& tw, M3<R>& S) const {
        {   // forward: A B^+ C^+ = A (C B)^+
            M3r2<R> B = take<6 * PL + 0>(g, tw);
            M3<R> Bf = complete(B), D, F;
            M3r2<R> C = take<6 * PL + 1>(g, tw);
            mul_r2_full(C, Bf, D);
            M3r2<R> A = take<6 * PL + 2>(g, tw);
            if (PL == 0) { mul_r2_dag(A, D, F); S = F; } else mul_r2_dag_add(A, D, S);
        }
        {   // backward: A2^+ B2^+ C2 = (B2 A2)^+ C2
            M3r2<R> A2 = take<6 * PL + 3>(g, tw);
            M3<R> Af = complete(A2), E;
            M3r2<R> B2 = take<6 * PL + 4>(g, tw);
            mul_r2_full(B2, Af, E);
            M3r2<R> C2 = take<6 * PL + 5>(g, tw);
            M3<R> Cf = complete(C2);
            mul_dag_full_add(E, Cf, S);
        }
    }
    // staple sum, then the site's own link
    __device__ __forceinline__ void staple_and_link(const Geom& g, const Twist<R>& tw, M3<R>& S, M3r2<R>& m0) const {
        plane<0>(g, tw, S);
        plane<1>(g, tw, S);
        plane<2>(g, tw, S);
        m0 = take<18>(g, tw);
    }
};

// first things a site update will read: its own link and the first plane (issued before the generator phase)
template <int SUN, typename R, int MU>
__device__ __forceinline__ void prefetch_site_head(const R* table, const Geom& g, uint32_t gindex, const Coord& p) {
    constexpr int NPL = (SUN == 3) ? 3 : 1;
    Coord pmu = nb_plus<MU>(g, p);
    PlaneSites s0 = plane_sites<MU, PlaneOrder<MU>::n0>(g, p, pmu);
    prefetch_link<R, MU, NPL>(table, g, gindex);
    plane_prefetch<R, MU, PlaneOrder<MU>::n0, NPL>(table, g, gindex, s0);
}

// left-multiply rows (a,b) of M by the SU(2) element d (4-vector): rows' = [[u1,u2],[-conj u2, conj u1]] rows
template <typename R, int A, int B>
__device__ __forceinline__ void rotate_rows(M3<R>& M, const V4<R>& d) {
    const Cx<R> u1 = { d.x, d.z }, u2 = { d.y, d.w };
#pragma unroll
    for (int j = 0; j < 3; j++) {
        Cx<R> ra = M.e[A][j], rb = M.e[B][j];
        Cx<R> na = cmul(u1, ra); cfma(na, u2, rb);
        // v1 = -conj(u2), v2 = conj(u1)
        Cx<R> nb = { -(u2.re * ra.re + u2.im * ra.im), -(u2.re * ra.im - u2.im * ra.re) };
        cfmacc(nb, u1, rb);
        M.e[A][j] = na; M.e[B][j] = nb;
    }
}
// the same, computing only the columns of the new row A in mask MA and of the new row B in mask MB (bit j = column j): inside the
// Cabibbo-Marinari loop the compiler cannot see which elements of W and U the later subgroups still read
template <typename R, int A, int B, int MA, int MB>
__device__ __forceinline__ void rotate_rows_sel(M3<R>& M, const V4<R>& d) {
    const Cx<R> u1 = { d.x, d.z }, u2 = { d.y, d.w };
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const Cx<R> ra = M.e[A][j], rb = M.e[B][j];
        if ((MA >> j) & 1) {
            Cx<R> na = cmul(u1, ra); cfma(na, u2, rb);
            M.e[A][j] = na;
        }
        if ((MB >> j) & 1) {
            Cx<R> nb = { -(u2.re * ra.re + u2.im * ra.im), -(u2.re * ra.im - u2.im * ra.re) };
            cfmacc(nb, u1, rb);
            M.e[B][j] = nb;
        }
    }
}
// 1 / beta: formed once per thread at kernel entry and parked in shared memory (every thread of the block writes the same value and
// reads it back inside the Cabibbo-Marinari loop).  Left to the compiler, beta was re-loaded from lattice_parameters and divided by
// in every subgroup step, at the head of the step's dependency chain.
template <typename R> struct InvBeta {
    static __device__ __forceinline__ R* slot() { __shared__ R s; return &s; }
    static __device__ __forceinline__ void park(const R* params) { *reinterpret_cast<volatile R*>(slot()) = R(1) / params[0]; }
    static __device__ __forceinline__ R get() { return *reinterpret_cast<volatile R*>(slot()); }
};
// aH (normalised dagger of the 2x2 block (a,b) of W), its norm and the reciprocal of the norm
template <typename R, int A, int B>
__device__ __forceinline__ V4<R> block_dagger_inv(const M3<R>& W, R& idet) {
    V4<R> aH;
    aH.x =  (W.e[A][A].re + W.e[B][B].re);
    aH.z = -(W.e[A][A].im - W.e[B][B].im);
    aH.y = -(W.e[A][B].re - W.e[B][A].re);
    aH.w = -(W.e[A][B].im + W.e[B][A].im);
    idet = rsqrt_r(aH.x * aH.x + aH.y * aH.y + aH.z * aH.z + aH.w * aH.w);
    aH.x *= idet; aH.y *= idet; aH.z *= idet; aH.w *= idet;
    return aH;
}
// aH (normalised dagger of the 2x2 block (a,b) of W) and its norm
template <typename R, int A, int B>
__device__ __forceinline__ V4<R> block_dagger(const M3<R>& W, R& det) {
    V4<R> aH;
    aH.x =  (W.e[A][A].re + W.e[B][B].re);
    aH.z = -(W.e[A][A].im - W.e[B][B].im);
    aH.y = -(W.e[A][B].re - W.e[B][A].re);
    aH.w = -(W.e[A][B].im + W.e[B][A].im);
    R id;
    Mth<R>::sqrt_and_rcp(aH.x * aH.x + aH.y * aH.y + aH.z * aH.z + aH.w * aH.w, det, id);
    aH.x *= id; aH.y *= id; aH.z *= id; aH.w *= id;
    return aH;
}
// One Cabibbo-Marinari round over the subgroups (0,1), (0,2), (1,2) in the reference's order.  The loop is NOT unrolled:
// the Kennedy-Pendleton code (log, cospi, sincospi, sqrt) is instantiated once; only the small block-extraction and
// row-rotation pieces are per-subgroup (uniform switch, no divergence).
#ifndef QCD_CM_UNROLL      // experiment knob: 3 = the three subgroup steps as straight-line code (Kennedy-Pendleton instantiated three times)
#define QCD_CM_UNROLL 1
#endif
constexpr int kCmUnroll = QCD_CM_UNROLL;
template <typename R, class Prn>
__device__ __forceinline__ void cm_round(M3<R>& U, M3<R>& W, const Geom& g, R beta, Prn& prn) {
#pragma unroll kCmUnroll
    for (int sub = 0; sub < 3; sub++) {
        R idet;
        V4<R> aH;
        switch (sub) {
            case 0: aH = block_dagger_inv<R, 0, 1>(W, idet); break;
            case 1: aH = block_dagger_inv<R, 0, 2>(W, idet); break;
            default: aH = block_dagger_inv<R, 1, 2>(W, idet); break;
        }
        V4<R> d;
        if (kp_heatbath<R, Prn>(prn, g, aH, idet * InvBeta<R>::get(), d)) {       // 1 / (beta sqrt(det))
            // Only what is read later is formed.  After the round the caller uses rows 0 and 1 of U (Gram-Schmidt, store; the
            // third row and W are rebuilt for another round): the last subgroup needs no W and no new third row of U.  Of W the
            // subgroup (0,2) reads the corners of rows 0 and 2 and rotates row 0 into row 2, (1,2) reads columns 1, 2 of rows 1, 2.
            switch (sub) {
                case 0: rotate_rows<R, 0, 1>(U, d); rotate_rows_sel<R, 0, 1, 7, 6>(W, d); break;
                case 1: rotate_rows<R, 0, 2>(U, d); rotate_rows_sel<R, 0, 2, 0, 6>(W, d); break;
                default: rotate_rows_sel<R, 1, 2, 7, 0>(U, d); break;
            }
        } else {
            prn.burn();    // su3_update_cl.cl:665-668
        }
    }
}
// One Cabibbo-Marinari round, software-pipelined over the three subgroups (samples parked in shared memory, PrnSmemT).  The
// transcendental part of a Kennedy-Pendleton trial depends on the random numbers only, not on the lattice: the first trial of subgroup
// s + 1 is evaluated in the same basic block as the composition of subgroup s's update, so two dependency chains run interleaved where
// cm_round runs them one after the other (the section is a chain of dependent FP64 operations at three warps per scheduler: its depth
// is its cost).  A first trial that is rejected (5 % of the lanes at beta = 6), or samples beyond the parked ones, continue in the
// general code.  Same functions on the same values as cm_round: the lattices stay bit-identical to the other kernels'.
// MEASURED, NOT SHIPPED (build with QCD_NVCC_EXTRA=-DQCD_CM_PIPELINE=1): the Cabibbo-Marinari section's share of the warps' life time
// falls from 49 % to 17 %, but the pass takes 3.97 instead of 3.32 ms -- `no_instruction` stalls rise from 13 % to 30 % of all samples,
// in every phase, the generator's small loop included.  The kernel executes ~7.6 k mostly straight-line instructions per warp (~34 B of
// instruction fetch per clock and SM): warps that get through their straight-line code faster starve the others of instructions.
#ifndef QCD_CM_PIPELINE
#define QCD_CM_PIPELINE 0
#endif
template <typename R, int NSKIP>
__device__ __forceinline__ void cm_round_pipelined(M3<R>& U, M3<R>& W, const Geom& g, R beta, PrnSmemT<NSKIP>& prn) {
    typedef PrnSmemT<NSKIP> Prn;
    if (g.nhit == 0u) { cm_round<R, Prn>(U, W, g, beta, prn); return; }
    const R ibeta = InvBeta<R>::get();
    uint32_t k = prn.k;
    const uint32_t last = (prn.keep >= 2u) ? prn.keep - 2u : 0u;          // k <= last: samples k and k + 1 are parked
    bool parked = prn.keep >= 2u && k <= last;
    KpTrial<R> T = kp_trial_eval<R>(prn.get_parked(k), prn.get_parked(k + 1u));
#pragma unroll 1
    for (int sub = 0; sub < 3; sub++) {
        R idet;
        V4<R> aH;
        switch (sub) {
            case 0: aH = block_dagger_inv<R, 0, 1>(W, idet); break;
            case 1: aH = block_dagger_inv<R, 0, 2>(W, idet); break;
            default: aH = block_dagger_inv<R, 1, 2>(W, idet); break;
        }
        const R ibdet = idet * ibeta;
        const R delta = T.E * ibdet;
        const bool fast = parked && (T.w2 <= R(1) - R(0.5) * delta);
        V4<R> d = { R(1), R(0), R(0), R(0) };      // the identity leaves U and W as they are (NHIT rejections, su3_update_cl.cl:665-668)
        if (!fast) {
            // parked: sample k was the rejected first trial, trials 2 .. NHIT follow from sample k + 1 on; otherwise (samples beyond
            // the parked ones) the whole step by the general code
            const uint32_t tries = parked ? g.nhit - 1u : g.nhit;
            prn.k = parked ? k + 1u : k;
            const bool ok = tries > 0u && kp_heatbath<R, Prn>(prn, g, aH, ibdet, d, tries);
            if (!ok) { prn.burn(); d = { R(1), R(0), R(0), R(0) }; }
            k = prn.k;
        } else {
            k += 2u;
        }
        // first trial of the next subgroup together with the composition of this one: independent chains in one basic block
        KpTrial<R> Tn = T;
        if (sub < 2) {
            parked = prn.keep >= 2u && k <= last;
            Tn = kp_trial_eval<R>(prn.get_parked(k), prn.get_parked(k + 1u));
        }
        if (fast) d = kp_compose(delta, T.D, aH);
        switch (sub) {
            case 0: rotate_rows<R, 0, 1>(U, d); rotate_rows_sel<R, 0, 1, 7, 6>(W, d); break;
            case 1: rotate_rows<R, 0, 2>(U, d); rotate_rows_sel<R, 0, 2, 0, 6>(W, d); break;
            default: rotate_rows_sel<R, 1, 2, 7, 0>(U, d); break;
        }
        T = Tn;
    }
    prn.k = k;
}
template <typename R, int A, int B>
__device__ __forceinline__ void or_subgroup(M3<R>& U, M3<R>& W) {
    R det;
    V4<R> aH = block_dagger<R, A, B>(W, det);
    V4<R> d = su2_mul(aH, aH);
    rotate_rows<R, A, B>(U, d);
    rotate_rows<R, A, B>(W, d);
}
// ---------------------------------------------------------------- SU(2) staple (4-vector algebra)
template <typename R, int MU, int NU, bool TBC>
__device__ __forceinline__ void staple2_plane(const R* table, const Geom& g, const Coord& p, uint32_t gindex,
                                              const Coord& pmu, const Twist<R>& tw, V4<R>& S, bool first) {
    Coord pnu = nb_plus<NU>(g, p);
    Coord pmn = nb_minus<NU>(g, p);
    Coord pmnmu = nb_plus<MU>(g, pmn);
    uint32_t g_pmn = coords_to_gid(g, pmn);
    V4<R> A  = load2<R, NU, TBC>(table, g, coords_to_gid(g, pmu), pmu.x, tw);
    V4<R> B  = load2<R, MU, TBC>(table, g, coords_to_gid(g, pnu), pnu.x, tw);
    V4<R> C  = load2<R, NU, TBC>(table, g, gindex, p.x, tw);
    V4<R> A2 = load2<R, NU, TBC>(table, g, coords_to_gid(g, pmnmu), pmnmu.x, tw);
    V4<R> B2 = load2<R, MU, TBC>(table, g, g_pmn, pmn.x, tw);
    V4<R> C2 = load2<R, NU, TBC>(table, g, g_pmn, pmn.x, tw);
    V4<R> f = su2_mul(su2_mul(A, su2_dag(B)), su2_dag(C));
    V4<R> b = su2_mul(su2_mul(su2_dag(A2), su2_dag(B2)), C2);
    if (first) { S = f; } else { S.x += f.x; S.y += f.y; S.z += f.z; S.w += f.w; }
    S.x += b.x; S.y += b.y; S.z += b.z; S.w += b.w;
}
template <typename R, int MU, bool TBC>
__device__ __forceinline__ V4<R> staple2(const R* table, const Geom& g, const Coord& p, uint32_t gindex, const Twist<R>& tw) {
    V4<R> S;
    Coord pmu = nb_plus<MU>(g, p);
    constexpr int n0 = (MU == 1) ? 0 : 1;
    constexpr int n1 = (MU == 0 || MU == 1 || MU == 3) ? 2 : 0;
    constexpr int n2 = (MU == 3) ? 0 : 3;
    staple2_plane<R, MU, n0, TBC>(table, g, p, gindex, pmu, tw, S, true);
    staple2_plane<R, MU, n1, TBC>(table, g, p, gindex, pmu, tw, S, false);
    staple2_plane<R, MU, n2, TBC>(table, g, p, gindex, pmu, tw, S, false);
    return S;
}

// ---------------------------------------------------------------- site updates
template <typename R, int MU, class Prn>
__device__ __forceinline__ void heatbath_su3(R* table, const Geom& g, R beta, uint32_t gindex, const M3r2<R>& m0, const M3<R>& S, Prn& prn);
template <typename R, int MU, bool TBC, class Prn>
__device__ __forceinline__ void update_site_su3(R* table, const Geom& g, R beta, const Twist<R>& tw, uint32_t gindex, const Coord& p, Prn& prn) {
    M3<R> S = staple3<R, MU, TBC>(table, g, p, gindex, tw);
    M3r2<R> m0 = load3<R, MU, TBC>(table, g, gindex, p.x, tw);     // after the staples: 24 registers less while they are formed
    heatbath_su3<R, MU, Prn>(table, g, beta, gindex, m0, S, prn);
}
// ---------------------------------------------------------------- SU(3) link staging by TMA (warp = one y-row of one parity)
// With N2 a multiple of 64 the 32 lanes of a warp update 32 consecutive sites of one parity of one y-row, so every (link, plane)
// they read is 32 vec4 at a stride of two elements: ONE tensor-map box per link (cp.async.bulk.tensor.3d, traversal strides (1, 2, 4),
// box 4 x 64 x 9 -> 3 planes of 32 x sizeof(vec4) bytes) issued by lane 0, landing in a per-warp ring of kTmaK link slots (3 planes each) and
// announced through one mbarrier per slot.  The links of the next staple so arrive while the current one is multiplied, without
// registers and without the half-used 16-byte LDGSTS sectors that made the cp.async ring above lose; the physics phase otherwise
// waits for the first use of a loaded link 19 times per update (43 % of its stall samples, profiles/r01_ncu_update_v6_packed.json).
// The box starts at lane 0's site; a lane whose own site is not at (lane 0's site + 2 lane) -- the one y +- 1 neighbour per row that
// wraps around -- loads its three planes itself.  tools/microbench/tma_rows.cu measures the access pattern in isolation.
#ifndef QCD_TMA_PF     // experiment knob: lane 0 also prefetches the boxes of the next plane into the L2 (measured slower)
#define QCD_TMA_PF 0
#endif
#ifndef QCD_TMA_SYNC_AFTER_ISSUE     // experiment knob: warp barrier behind lane 0's TMA issue (not needed for correctness)
#define QCD_TMA_SYNC_AFTER_ISSUE 0
#endif
#ifndef QCD_TMA_K     // links in flight per warp (ring slots of 3 planes each)
#define QCD_TMA_K 3
#endif
constexpr int kTmaK = QCD_TMA_K;
template <typename R> __host__ __device__ constexpr uint32_t tma_plane_bytes() { return 32u * uint32_t(sizeof(V4<R>)); }
template <typename R> __host__ __device__ constexpr uint32_t tma_slot_bytes() { return 3u * tma_plane_bytes<R>(); }
// per warp: kTmaK link slots, kTmaK 128-byte slots for the one wrapped-around site of a row, kTmaK mbarriers
template <typename R> __host__ __device__ constexpr size_t tma_stage_bytes() { return size_t(kUpdateThreads / 32) * kTmaK * (tma_slot_bytes<R>() + 128 + 8); }

// WS: the ring lives across the items of a persistent warp (update_kernel_rlx_ws): the mbarriers are initialised once and the
// phase a slot waits for depends on how often it was used before (ring_par, one bit per slot).
template <typename R, int MU, bool TBC, bool WS = false> struct Su3StageTma {
    uint32_t buf, xbuf, bar;    // shared-memory addresses of this warp's slot 0, wrapped-site slot 0 and mbarrier 0
    uint32_t ring_par;          // WS: parity of the uses of each slot by the items before this one
    const void* tmap;           // CUtensorMap over the link table: (4 reals, ROWSIZE sites, 12 planes), box = 32 sites x 3 planes
    const void* tmap1;          // same tensor, box = 1 site x 3 planes
    const R* table;
    uint32_t gindex, px, lane;
    Coord pc;                   // the site (plane loop: the sites of a plane's links are formed when the plane comes up)
    PlaneSites s0, s1, s2;

    // link L of the site's read sequence (same order as Su3Stage): site and direction
    template <int L> __device__ __forceinline__ uint32_t link_gid() const {
        if constexpr (L >= 18) return gindex;
        constexpr int pl = L / 6, i = L % 6;
        const PlaneSites& s = (pl == 0) ? s0 : (pl == 1) ? s1 : s2;
        return (i == 0) ? s.g_pnu : (i == 1) ? gindex : (i == 2) ? s.g_pmu : (i == 3) ? s.g_pmnmu : s.g_pmn;
    }
    template <int L> __host__ __device__ static constexpr int link_dir() {
        if (L >= 18) return MU;
        return (L % 6 == 0 || L % 6 == 4) ? MU : plane_nu<MU>(L / 6);
    }
    // can the site of link L be a y +- 1 neighbour (the only move that breaks "lane 0's site + 2 lane")?
    template <int L> __host__ __device__ static constexpr bool may_wrap() {
        if (L >= 18) return false;
        const int i = L % 6, NU = plane_nu<MU>(L / 6);
        return (i == 1) ? false : (i == 2) ? (MU == 1) : (i == 3) ? (MU == 1 || NU == 1) : (NU == 1);
    }
    // One link into ring slot SLOT: the three planes of direction `dir` of the 32 sites of this warp's row, box start = lane 0's site.
    // MW: 0 = the link's site is never a y +- 1 neighbour that wraps around, 1 = it may be, 2 = it may be if mw_rt (plane loop: NU == 1).
    // dep: a word (always 0) whose computation needed the loaded data of every lane that read the slot being refilled
    template <int SLOT, int MW> __device__ __forceinline__ void issue_link(uint32_t gid_, int dir, bool mw_rt, uint32_t dep) const {
        uint32_t gidb = 0, wrapped = 0;
        int start = int(gid_);              // box start = "site of lane 0" of the unwrapped row
        if (MW == 1 || (MW == 2 && mw_rt)) {    // the first or the last lane of a row may wrap: lane 1 never does, it is the reference;
            const int gid = start;          // the wrapped lane's site comes through a box of its own
            start = __shfl_sync(0xffffffffu, gid, 1) - 2;
            const uint32_t bad = __ballot_sync(0xffffffffu, gid != start + 2 * int(lane));
            wrapped = bad;
            gidb = uint32_t(__shfl_sync(0xffffffffu, gid, bad ? (__ffs(int(bad)) - 1) : 0));
        }
        if (lane == 0) {
            const int gid0 = start + int(dep);
            const uint32_t b = bar + 8u * SLOT, d = buf + SLOT * tma_slot_bytes<R>();
            const uint32_t bytes = tma_slot_bytes<R>() + (wrapped ? 3u * uint32_t(sizeof(V4<R>)) : 0u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
            // one box per link: planes dir, dir + 4, dir + 8 through the traversal stride of the plane dimension
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(d), "l"(tmap), "r"(0), "r"(gid0), "r"(dir), "r"(b) : "memory");
            if (wrapped)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             :: "r"(xbuf + 128u * SLOT), "l"(tmap1), "r"(0), "r"(int(gidb)), "r"(dir), "r"(b) : "memory");
        }
#if QCD_TMA_SYNC_AFTER_ISSUE
        __syncwarp();
#endif
    }
    template <int L> __device__ __forceinline__ void issue(uint32_t dep = 0u) const {
        if constexpr (L <= 18) issue_link<L % kTmaK, may_wrap<L>() ? 1 : 0>(link_gid<L>(), link_dir<L>(), false, dep);
    }
    // lane 0: pull the three boxes of link L towards the L2 (one instruction per KB instead of one prefetch per thread and sector)
    template <int L> __device__ __forceinline__ void prefetch_l2_link() const {
        constexpr int DIR = link_dir<L>();
        const uint32_t gid0 = link_gid<L>();
        asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" :: "l"(tmap), "r"(0), "r"(int(gid0)), "r"(DIR) : "memory");
    }
    template <int L0, int L1> __device__ __forceinline__ void prefetch_l2_links() const {
        if constexpr (L0 < L1) { prefetch_l2_link<L0>(); prefetch_l2_links<L0 + 1, L1>(); }
    }
    // The link in ring slot SLOT (use number 2 n + PHASE of the slot): wait for its box, read this lane's three vec4, unpack, twist.
    // `dep` comes back as the (always zero) word that orders the slot's refill behind the reads (see below).
    template <int SLOT, int PHASE, int MW>
    __device__ __forceinline__ M3r2<R> take_link(const Geom& g, const Twist<R>& tw, uint32_t gid, int dir, uint32_t x, bool mw_rt, uint32_t& dep) const {
        const bool mw = MW == 1 || (MW == 2 && mw_rt);
        const int start = mw ? __shfl_sync(0xffffffffu, int(gid), 1) - 2 : 0;
        {
            const uint32_t b = bar + 8u * SLOT, phase = uint32_t(PHASE) ^ (WS ? ((ring_par >> SLOT) & 1u) : 0u);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(b), "r"(phase) : "memory");
        }
        V4<R> v[3];
        // this lane's vec4 of plane k: in the link slot, or (the row's wrapped-around y neighbour) in the one-site slot
        uint32_t a = buf + SLOT * tma_slot_bytes<R>() + lane * uint32_t(sizeof(V4<R>)), step = tma_plane_bytes<R>();
        if (mw) {
            if (int(gid) != start + 2 * int(lane)) { a = xbuf + 128u * SLOT; step = uint32_t(sizeof(V4<R>)); }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if constexpr (sizeof(R) == 8) {
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[k].x), "=d"(v[k].y) : "r"(a + k * step) : "memory");
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[k].z), "=d"(v[k].w) : "r"(a + k * step + 16u) : "memory");
            } else {
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "r"(a + k * step) : "memory");
            }
        }
        M3r2<R> m = unpack3(v[0], v[1], v[2]);
        if (TBC) {
            if (dir == 1) twist3(m, g, x, tw);
        }
        // The slot can be refilled once every lane's reads have RETURNED (the refill is a write through the async proxy: issuing it
        // right behind the reads let a whole warp see a half-refilled slot about once in 10^6 updates).  A vote over a word taken
        // from each of the lane's loads waits for exactly that; its (always zero) result enters the box coordinate, so the TMA
        // cannot be issued before.  (fence.proxy.async + __syncwarp does the same for 7 % of the kernel's stall samples.)
        uint32_t probe;
        if constexpr (sizeof(R) == 8)
            probe = uint32_t(__double2hiint(v[0].x) | __double2hiint(v[0].z) | __double2hiint(v[1].x) | __double2hiint(v[1].z) |
                             __double2hiint(v[2].x) | __double2hiint(v[2].z));
        else
            probe = __float_as_uint(float(v[0].x)) | __float_as_uint(float(v[1].x)) | __float_as_uint(float(v[2].x));
        dep = __ballot_sync(0xffffffffu, probe == 0x7ff5a5a5u);
        asm volatile("and.b32 %0, %0, 0;" : "+r"(dep));
        return m;
    }
    template <int L> __device__ __forceinline__ M3r2<R> take(const Geom& g, const Twist<R>& tw) const {
        constexpr int pl = (L < 18) ? L / 6 : 0, i = L % 6;
        const PlaneSites& s = (pl == 0) ? s0 : (pl == 1) ? s1 : s2;
        const uint32_t x = (L == 18 || i == 1) ? px : (i == 0) ? s.x_pnu : (i == 2) ? s.x_pmu : (i == 3) ? s.x_pmnmu : s.x_pmn;
        uint32_t dep;
        M3r2<R> m = take_link<L % kTmaK, (L / kTmaK) & 1, may_wrap<L>() ? 1 : 0>(g, tw, link_gid<L>(), link_dir<L>(), x, false, dep);
        issue<L + kTmaK>(dep);
        return m;
    }
    // called by all 32 lanes of a full warp; smem_warp = this warp's ring, bars = this warp's kTmaK mbarriers
    __device__ __forceinline__ void begin(uint32_t smem_warp, uint32_t xslots, uint32_t bars, const void* tm, const void* tm1, const R* tab, const Geom& g, uint32_t gi, const Coord& p) {
        buf = smem_warp; xbuf = xslots; bar = bars; tmap = tm; tmap1 = tm1; table = tab; gindex = gi; px = p.x; pc = p; lane = threadIdx.x & 31u;
        Coord pmu = nb_plus<MU>(g, p);
        s0 = plane_sites<MU, PlaneOrder<MU>::n0>(g, p, pmu);
        s1 = plane_sites<MU, PlaneOrder<MU>::n1>(g, p, pmu);
        s2 = plane_sites<MU, PlaneOrder<MU>::n2>(g, p, pmu);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kTmaK; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar + 8u * k) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        prologue<0>();
        if (QCD_TMA_PF && lane == 0) prefetch_l2_links<kTmaK, kTmaK + 6>();      // the next six links towards the L2; the rest follows one plane ahead of the ring
        __syncwarp();
    }
    // WS: once per warp -- ring addresses and mbarriers; then begin_item() / staple_and_link() / end_item() per item
    __device__ __forceinline__ void init_ws(uint32_t smem_warp, uint32_t xslots, uint32_t bars, const void* tm, const void* tm1, const R* tab) {
        buf = smem_warp; xbuf = xslots; bar = bars; tmap = tm; tmap1 = tm1; table = tab; lane = threadIdx.x & 31u; ring_par = 0u;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kTmaK; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar + 8u * k) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    __device__ __forceinline__ void begin_item(const Geom& g, uint32_t gi, const Coord& p) {
        gindex = gi; px = p.x; pc = p;
        Coord pmu = nb_plus<MU>(g, p);
        s0 = plane_sites<MU, PlaneOrder<MU>::n0>(g, p, pmu);
        s1 = plane_sites<MU, PlaneOrder<MU>::n1>(g, p, pmu);
        s2 = plane_sites<MU, PlaneOrder<MU>::n2>(g, p, pmu);
        prologue<0>();
        __syncwarp();
    }
    // slot s was used by the links L = s, s + K, ... <= 18 of the item
    __device__ __forceinline__ void end_item() {
#pragma unroll
        for (int s = 0; s < kTmaK; s++) if (((18 - s) / kTmaK + 1) & 1) ring_par ^= 1u << s;
    }
    template <int L> __device__ __forceinline__ void prologue() const {
        if constexpr (L < kTmaK) { issue<L>(); prologue<L + 1>(); }
    }
    template <int PL> __device__ __forceinline__ void plane(const Geom& g, const Twist<R>& tw, M3<R>& S) const {
        if (QCD_TMA_PF && lane == 0) prefetch_l2_links<(6 * PL + 6 + kTmaK < 19) ? 6 * PL + 6 + kTmaK : 19, (6 * PL + 12 + kTmaK < 19) ? 6 * PL + 12 + kTmaK : 19>();   // L2 prefetch runs six links ahead of the ring
        __syncwarp();
        {   // forward: A B^+ C^+ = A (C B)^+
            M3r2<R> B = take<6 * PL + 0>(g, tw);
            M3<R> Bf = complete(B), D, F;
            M3r2<R> C = take<6 * PL + 1>(g, tw);
            mul_r2_full(C, Bf, D);
            M3r2<R> A = take<6 * PL + 2>(g, tw);
            if (PL == 0) { mul_r2_dag(A, D, F); S = F; } else mul_r2_dag_add(A, D, S);
        }
        {   // backward: A2^+ B2^+ C2 = (B2 A2)^+ C2
            M3r2<R> A2 = take<6 * PL + 3>(g, tw);
            M3<R> Af = complete(A2), E;
            M3r2<R> B2 = take<6 * PL + 4>(g, tw);
            mul_r2_full(B2, Af, E);
            M3r2<R> C2 = take<6 * PL + 5>(g, tw);
            M3<R> Cf = complete(C2);
            mul_dag_full_add(E, Cf, S);
        }
    }
#ifndef QCD_PLANE_LOOP     // 1: the three planes as a run-time loop (a third of the staple code: 36 KB -> 13 KB of SASS).  Measured slower
#define QCD_PLANE_LOOP 0   // (4.20 against 3.97 ms per pass with the pipelined Cabibbo-Marinari round): the run-time plane costs 870
#endif                     // instructions per update (sites, wrap tests, copies, spill traffic) and the fetch stalls do not go away
    __device__ __forceinline__ PlaneSites sites_rt(const Geom& g, int nu) const {
        const Coord pmu = nb_plus<MU>(g, pc), pnu = nb_plus_rt(g, pc, nu), pmn = nb_minus_rt(g, pc, nu), pmnmu = nb_plus<MU>(g, pmn);
        PlaneSites s;
        s.g_pmu = coords_to_gid(g, pmu); s.g_pnu = coords_to_gid(g, pnu);
        s.g_pmn = coords_to_gid(g, pmn); s.g_pmnmu = coords_to_gid(g, pmnmu);
        s.x_pmu = pmu.x; s.x_pnu = pnu.x; s.x_pmn = pmn.x; s.x_pmnmu = pmnmu.x;
        return s;
    }
    // Which links of a plane can sit on a wrapped y neighbour (may_wrap above with a run-time NU): 0 never, 1 yes, 2 if NU == 1
    template <int I> __host__ __device__ static constexpr int mw_of() {
        return (I == 1) ? 0 : (I == 2) ? (MU == 1 ? 1 : 0) : (I == 3) ? (MU == 1 ? 1 : 2) : (MU == 1 ? 0 : 2);
    }
    // The same 19 reads as plane<0..2> + take<18>, the planes as a run-time loop: link 6 pl + i uses ring slot i % 3 in its use number
    // 2 pl + i / 3, so slot and mbarrier phase do not depend on the plane; only the sites and the direction NU do.  Bit-identical
    // to the unrolled planes (experiment, see QCD_PLANE_LOOP).
    __device__ __forceinline__ void staple_and_link_loop(const Geom& g, const Twist<R>& tw, M3<R>& S, M3r2<R>& m0) const {
        PlaneSites cur = s0;
#pragma unroll 1
        for (int pl = 0; pl < 3; pl++) {
            const int nu = (pl == 0) ? PlaneOrder<MU>::n0 : (pl == 1) ? PlaneOrder<MU>::n1 : PlaneOrder<MU>::n2;
            const int nun = (pl == 0) ? PlaneOrder<MU>::n1 : PlaneOrder<MU>::n2;
            const bool mw = nu == 1, mwn = nun == 1, lastp = pl == 2;
            PlaneSites nxt = cur;
            if (!lastp) nxt = sites_rt(g, nun);
            uint32_t dep;
            __syncwarp();
            {   // forward: A B^+ C^+ = A (C B)^+
                M3r2<R> B = take_link<0, 0, mw_of<0>()>(g, tw, cur.g_pnu, MU, cur.x_pnu, mw, dep);
                issue_link<0, mw_of<3>()>(cur.g_pmnmu, nu, mw, dep);
                M3<R> Bf = complete(B), D;
                M3r2<R> C = take_link<1, 0, 0>(g, tw, gindex, nu, px, false, dep);
                issue_link<1, mw_of<4>()>(cur.g_pmn, MU, mw, dep);
                mul_r2_full(C, Bf, D);
                M3r2<R> A = take_link<2, 0, mw_of<2>()>(g, tw, cur.g_pmu, nu, cur.x_pmu, mw, dep);
                issue_link<2, mw_of<5>()>(cur.g_pmn, nu, mw, dep);
                if (pl == 0) { M3<R> F; mul_r2_dag(A, D, F); S = F; } else mul_r2_dag_add(A, D, S);
            }
            {   // backward: A2^+ B2^+ C2 = (B2 A2)^+ C2
                M3r2<R> A2 = take_link<0, 1, mw_of<3>()>(g, tw, cur.g_pmnmu, nu, cur.x_pmnmu, mw, dep);
                if (!lastp) issue_link<0, mw_of<0>()>(nxt.g_pnu, MU, mwn, dep);
                else        issue_link<0, 0>(gindex, MU, false, dep);                  // the site's own link (read 18)
                M3<R> Af = complete(A2), E;
                M3r2<R> B2 = take_link<1, 1, mw_of<4>()>(g, tw, cur.g_pmn, MU, cur.x_pmn, mw, dep);
                if (!lastp) issue_link<1, 0>(gindex, nun, false, dep);
                mul_r2_full(B2, Af, E);
                M3r2<R> C2 = take_link<2, 1, mw_of<5>()>(g, tw, cur.g_pmn, nu, cur.x_pmn, mw, dep);
                if (!lastp) issue_link<2, mw_of<2>()>(nxt.g_pmu, nun, mwn, dep);
                M3<R> Cf = complete(C2);
                mul_dag_full_add(E, Cf, S);
            }
            cur = nxt;
        }
        uint32_t dep;
        m0 = take_link<0, 0, 0>(g, tw, gindex, MU, px, false, dep);
    }
    __device__ __forceinline__ void staple_and_link(const Geom& g, const Twist<R>& tw, M3<R>& S, M3r2<R>& m0) const {
#if QCD_PLANE_LOOP
        static_assert(kTmaK == 3, "the plane loop is written for a ring of three slots");
        staple_and_link_loop(g, tw, S, m0);
#else
        plane<0>(g, tw, S);
        plane<1>(g, tw, S);
        plane<2>(g, tw, S);
        m0 = take<18>(g, tw);
#endif
    }
};
// same with the links arriving through the shared-memory ring (begin() was called earlier)
template <typename R, int MU, bool TBC, class Prn, class Stage>
__device__ __forceinline__ void update_site_su3_staged(R* table, const Geom& g, R beta, const Twist<R>& tw,
                                                       const Stage& stage, Prn& prn) {
    M3<R> S;
    M3r2<R> m0;
    stage.staple_and_link(g, tw, S, m0);
    heatbath_su3<R, MU, Prn>(table, g, beta, stage.gindex, m0, S, prn);
}
template <typename R, int MU, class Prn>
__device__ __forceinline__ void heatbath_su3(R* table, const Geom& g, R beta, uint32_t gindex, const M3r2<R>& m0, const M3<R>& S, Prn& prn) {
    M3<R> U;
#pragma unroll
    for (int j = 0; j < 3; j++) { U.e[0][j] = m0.e[0][j]; U.e[1][j] = m0.e[1][j]; }
#pragma unroll 1
    for (uint32_t h = 0; h < g.nhitpar; h++) {
        third_row(U.e[0], U.e[1], U.e[2]);
        M3<R> W = mul_full(U, S);
#if QCD_CM_PIPELINE
        if constexpr (HasParked<Prn>::value) cm_round_pipelined(U, W, g, beta, prn);
        else
#endif
        cm_round<R, Prn>(U, W, g, beta, prn);
        gram_schmidt3(U.e[0], U.e[1]);
    }
    if (g.fold_gs) gram_schmidt3(U.e[0], U.e[1]);      // the sweep's reunitarisation pass, applied here (same function, same values)
    store3<R, MU>(table, g, gindex, U.e[0], U.e[1]);
}
template <typename R, int MU, bool TBC, class Prn>
__device__ __forceinline__ void update_site_su2(R* table, const Geom& g, R beta, const Twist<R>& tw, uint32_t gindex, const Coord& p, Prn& prn) {
    V4<R> S = staple2<R, MU, TBC>(table, g, p, gindex, tw);
    // su2_update_cl.cl:448-461: aH = S^+/|S|, det = 2|S|
    R in = rsqrt_r(S.x * S.x + S.y * S.y + S.z * S.z + S.w * S.w);
    V4<R> aH = { S.x * in, -S.y * in, -S.z * in, -S.w * in };
    V4<R> d;
    if (kp_heatbath<R, Prn>(prn, g, aH, R(0.5) * in * InvBeta<R>::get(), d)) {     // 1 / (beta 2 |S|)
        st_plane(table, g, gindex, MU, d);
        if (QCD_PEER_ON(g)) st_plane_peer(g, gindex, MU, d);
    }
    // on NHIT rejections the old (untwisted, stored) link stays -- su2_update_cl.cl:541-548.  With TBC the
    // reference would store the twisted m0 back; that only differs on the x = N1-1 Y-links.
    else if (TBC && MU == 1) {
        V4<R> m0 = load2<R, MU, TBC>(table, g, gindex, p.x, tw);
        st_plane(table, g, gindex, MU, m0);
        if (QCD_PEER_ON(g)) st_plane_peer(g, gindex, MU, m0);
    }
}
template <typename R, int MU>
__device__ __forceinline__ void overrelax_site_su3(R* table, const Geom& g, uint32_t gindex) {
    Twist<R> tw = { R(1), R(0), R(1), R(0) };
    Coord p = gid_to_coords(g, gindex);
    M3r2<R> m0 = load3<R, MU, false>(table, g, gindex, p.x, tw);
    M3<R> S = staple3<R, MU, false>(table, g, p, gindex, tw);
    M3<R> U = complete(m0);
    M3<R> W = mul_full(U, S);
    or_subgroup<R, 0, 1>(U, W);
    or_subgroup<R, 0, 2>(U, W);
    or_subgroup<R, 1, 2>(U, W);
    gram_schmidt3(U.e[0], U.e[1]);
    if (g.fold_gs) gram_schmidt3(U.e[0], U.e[1]);
    store3<R, MU>(table, g, gindex, U.e[0], U.e[1]);
}
template <typename R, int MU>
__device__ __forceinline__ void overrelax_site_su2(R* table, const Geom& g, uint32_t gindex) {
    Twist<R> tw = { R(1), R(0), R(1), R(0) };
    Coord p = gid_to_coords(g, gindex);
    V4<R> m0 = load2<R, MU, false>(table, g, gindex, p.x, tw);
    V4<R> S = staple2<R, MU, false>(table, g, p, gindex, tw);
    R in = R(1) / Mth<R>::sqrt_(S.x * S.x + S.y * S.y + S.z * S.z + S.w * S.w);
    V4<R> aH = { S.x * in, -S.y * in, -S.z * in, -S.w * in };
    V4<R> r = su2_mul(su2_mul(aH, su2_dag(m0)), aH);
    st_plane(table, g, gindex, MU, r);
    if (QCD_PEER_ON(g)) st_plane_peer(g, gindex, MU, r);
}


// twist phases from lattice_parameters[1..2] (su3_matrix_memory.cl:64-66, su2_matrix_memory.cl:73-75)
template <int SUN, typename R, bool TBC>
__device__ __forceinline__ Twist<R> make_twist(const R* params) {
    Twist<R> tw = { R(1), R(0), R(1), R(0) };
    if (TBC) {
        R phi = params[1], omega = (SUN == 3) ? params[2] : R(0);
        if (SUN == 3) {
            R a = (phi + omega) / 2, b = (omega - phi) / 2;
            if (sizeof(R) == 8) { tw.c1 = cos(a); tw.s1 = sin(a); tw.c2 = cos(b); tw.s2 = sin(b); }
            else { tw.c1 = cosf(a); tw.s1 = sinf(a); tw.c2 = cosf(b); tw.s2 = sinf(b); }
        } else {
            if (sizeof(R) == 8) { tw.c1 = cos(phi / 2); tw.s1 = sin(phi / 2); }
            else { tw.c1 = cosf(phi / 2); tw.s1 = sinf(phi / 2); }
        }
    }
    return tw;
}

template <int SUN, typename R, int MU, bool TBC, class Prn>
__global__ void __launch_bounds__(kUpdateThreads, (Prn::kMinBlocks * 128) / kUpdateThreads > 0 ? (Prn::kMinBlocks * 128) / kUpdateThreads : 1)
update_kernel(R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, int odd, typename Prn::Args pargs) {
    uint32_t GID = blockIdx.x * blockDim.x + threadIdx.x;
    if (GID >= g.nthreads) return;
    const SiteId site = thread_site(g, GID, odd != 0);
    GID = site.GID;
    const R beta = params[0];
    InvBeta<R>::park(params);
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    const uint32_t gindex = site.gindex;
    Prn prn(pargs, g, GID, gindex);
    if constexpr (SUN == 3 && kStageK > 0) {
        extern __shared__ __align__(128) uint4 dyn_smem[];
        Su3Stage<R, MU, TBC> stage;
        stage.begin(dyn_smem + threadIdx.x, table, g, gindex);
        update_site_su3_staged<R, MU, TBC, Prn, Su3Stage<R, MU, TBC>>(table, g, beta, tw, stage, prn);
    } else if constexpr (SUN == 3) update_site_su3<R, MU, TBC, Prn>(table, g, beta, tw, gindex, site.p, prn);
    else          update_site_su2<R, MU, TBC, Prn>(table, g, beta, tw, gindex, site.p, prn);
    prn.finish();
}
#ifdef QCD_TRACE
// diagnostic build (-DQCD_TRACE): clock64 of thread 0 at kernel entry, lock acquired, generator done, site done; and %smid
static __device__ unsigned long long g_trace[5 * 65536];
#define QCD_TRACE_MARK(i) if (threadIdx.x == 0 && blockIdx.x < 65536u) g_trace[5 * blockIdx.x + (i)] = clock64()
#else
#define QCD_TRACE_MARK(i)
#endif

template <int SUN, typename R, int MU, bool TBC, int NSKIP>
#ifndef QCD_RLX_MINBLOCKS     // experiment knob: resident blocks per SM the fp64 SU(3) two-phase kernel is compiled for
#define QCD_RLX_MINBLOCKS ((3 * 128) / kUpdateThreads)
#endif
__global__ void __launch_bounds__(kUpdateThreads, (sizeof(R) == 8 && SUN == 3) ? QCD_RLX_MINBLOCKS : (4 * 128) / kUpdateThreads)
update_kernel_rlx(R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, int odd,
                  const uint32_t* rtable, uint32_t* rtable_out, uint32_t instances, uint32_t samples, uint32_t keep,
                  const uint32_t* __restrict__ jump) {
    extern __shared__ __align__(128) uint4 prn_smem[];
    QCD_TRACE_MARK(0);
    uint32_t GID = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = GID < g.nthreads;
    const SiteId site = thread_site(g, active ? GID : 0u, odd != 0);
    GID = site.GID;
    unsigned long long converted[1] = { 0ull };
    RanluxPk<1> r;
    uint32_t left = 0;
    if (active && g.debug_skip != 1u) left = r.load(0, rtable, instances, GID);     // 20 DRAM loads in flight while the site's addresses are formed
    // (prefetching the state of the block that will take this block's place into the L2 was measured: the 26 extra instructions
    //  per warp cost more than the shorter wait saves, 34.4 -> 35.0 ms per sweep)
    const uint32_t gindex = site.gindex;
    const Coord p = site.p;                                               // kept across the generator phase (5 registers)
    if (active) prefetch_site_head<SUN, R, MU>(table, g, gindex, p);
    if (g.debug_skip != 1u) {   // phase 1: the pass's whole RANLUX work
        // (Serialising the generator phase per SM with a lock was measured: the resident blocks of an SM are already out of
        //  step -- 1.17 blocks in this phase at a time, tools/trace_phases.py -- and the lock costs 4.04 -> 4.11 ms per pass.)
        QCD_TRACE_MARK(1);
        if (active) {
            uint4* const out[1] = { prn_smem + threadIdx.x };
            r.pass(left, samples, keep, out, kUpdateThreads, converted, [] {}, jump);
            r.store(0, rtable_out, instances, GID, left);
        }
        QCD_TRACE_MARK(2);
    }
    if (!active) return;
    if (g.debug_skip == 2u) return;
    // phase 2: the link update
    const R beta = params[0];
    InvBeta<R>::park(params);
    const Twist<R> tw = make_twist<SUN, R, TBC>(params);
    typedef PrnSmemT<NSKIP> PrnSmem;
    // in place (rtable_out == rtable) the launcher parks every sample a link can consume, the rare path is unreachable
    PrnSmem prn(prn_smem + threadIdx.x, kUpdateThreads, converted[0], (rtable_out != rtable) ? keep : 0xffffffffu, rtable + GID, instances);
    if (SUN == 3) update_site_su3<R, MU, TBC, PrnSmem>(table, g, beta, tw, gindex, p, prn);
    else          update_site_su2<R, MU, TBC, PrnSmem>(table, g, beta, tw, gindex, p, prn);
    QCD_TRACE_MARK(3);
#ifdef QCD_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 65536u) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_trace[5 * blockIdx.x + 4] = sm; }
#endif
}
// update_kernel_rlx with the links of the staples staged by TMA (SU(3), full warps, N2 a multiple of 64; chosen by the launcher)
#ifdef QCD_RLX_MAXNREG        // experiment knob: register cap of the TMA-staged kernel below what three blocks per SM allow (168)
#define QCD_RLX_MAXNREG_ATTR __maxnreg__(QCD_RLX_MAXNREG)
#else
#define QCD_RLX_MAXNREG_ATTR __launch_bounds__(kUpdateThreads, (sizeof(R) == 8) ? QCD_RLX_MINBLOCKS : (4 * 128) / kUpdateThreads)
#endif
template <typename R, int MU, bool TBC, int NSKIP>
__global__ void QCD_RLX_MAXNREG_ATTR
update_kernel_rlx_tma(R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, int odd,
                      const uint32_t* rtable, uint32_t* rtable_out, uint32_t instances, uint32_t samples, uint32_t keep,
                      const __grid_constant__ TensorMapBlob tmap, const __grid_constant__ TensorMapBlob tmap1, const uint32_t* __restrict__ jump) {
    extern __shared__ __align__(128) uint4 prn_smem[];
    if (blockIdx.x * blockDim.x + threadIdx.x >= g.nthreads) return;                          // whole warps (the launcher guarantees full warps)
    const SiteId site = thread_site(g, blockIdx.x * blockDim.x + threadIdx.x, odd != 0);
    const uint32_t GID = site.GID;
    unsigned long long converted[1] = { 0ull };
    RanluxPk<1> r;
    uint32_t left = 0;
    if (g.debug_skip != 1u) left = r.load(0, rtable, instances, GID);
    const uint32_t gindex = site.gindex;
    const Coord p = site.p;
    // ring of this warp behind the parked samples, mbarriers behind the rings
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(prn_smem) + keep * uint32_t(kUpdateThreads * sizeof(uint4));
    const uint32_t warp = threadIdx.x >> 5;
    Su3StageTma<R, MU, TBC> stage;
    constexpr uint32_t NW = kUpdateThreads / 32, RING = kTmaK * tma_slot_bytes<R>();
    stage.begin(base + warp * RING, base + NW * RING + warp * (kTmaK * 128u), base + NW * (RING + kTmaK * 128u) + warp * (kTmaK * 8u),
                &tmap, &tmap1, table, g, gindex, p);
    if (g.debug_skip != 1u) {   // phase 1: the pass's whole RANLUX work (the first three links land meanwhile)
        uint4* const out[1] = { prn_smem + threadIdx.x };
        r.pass(left, samples, keep, out, kUpdateThreads, converted, [] {}, jump);
        r.store(0, rtable_out, instances, GID, left);
    }
    if (g.debug_skip == 2u) return;
    const R beta = params[0];
    InvBeta<R>::park(params);
    const Twist<R> tw = make_twist<3, R, TBC>(params);
    typedef PrnSmemT<NSKIP> PrnSmem;
    PrnSmem prn(prn_smem + threadIdx.x, kUpdateThreads, converted[0], (rtable_out != rtable) ? keep : 0xffffffffu, rtable + GID, instances);
    update_site_su3_staged<R, MU, TBC, PrnSmem, Su3StageTma<R, MU, TBC>>(table, g, beta, tw, stage, prn);
}
template <int SUN, typename R, int MU>
__global__ void __launch_bounds__(kUpdateThreads)
overrelax_kernel(R* __restrict__ table, const __grid_constant__ Geom g, int odd) {
    uint32_t GID = blockIdx.x * blockDim.x + threadIdx.x;
    if (GID >= g.nthreads) return;
    const uint32_t gindex = thread_site(g, GID, odd != 0).gindex;
    if (SUN == 3) overrelax_site_su3<R, MU>(table, g, gindex);
    else          overrelax_site_su2<R, MU>(table, g, gindex);
}

// overrelaxation pass with the links staged by TMA (SU(3), full warps, N2 a multiple of 64; chosen by the launcher)
#ifndef QCD_OR_TMA_MINBLOCKS
#define QCD_OR_TMA_MINBLOCKS 3
#endif
template <typename R, int MU>
__global__ void __launch_bounds__(kUpdateThreads, (sizeof(R) == 8) ? (QCD_OR_TMA_MINBLOCKS * 128) / kUpdateThreads : (4 * 128) / kUpdateThreads)
overrelax_kernel_tma(R* __restrict__ table, const __grid_constant__ Geom g, int odd,
                     const __grid_constant__ TensorMapBlob tmap, const __grid_constant__ TensorMapBlob tmap1) {
    extern __shared__ __align__(128) uint4 prn_smem[];
    const SiteId site = thread_site(g, blockIdx.x * blockDim.x + threadIdx.x, odd != 0);      // the launcher guarantees full blocks
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(prn_smem), warp = threadIdx.x >> 5;
    constexpr uint32_t NW = kUpdateThreads / 32, RING = kTmaK * tma_slot_bytes<R>();
    Su3StageTma<R, MU, false> stage;
    stage.begin(base + warp * RING, base + NW * RING + warp * (kTmaK * 128u), base + NW * (RING + kTmaK * 128u) + warp * (kTmaK * 8u),
                &tmap, &tmap1, table, g, site.gindex, site.p);
    const Twist<R> tw = { R(1), R(0), R(1), R(0) };
    M3<R> S;
    M3r2<R> m0;
    stage.staple_and_link(g, tw, S, m0);
    M3<R> U = complete(m0);
    M3<R> W = mul_full(U, S);
    or_subgroup<R, 0, 1>(U, W);
    or_subgroup<R, 0, 2>(U, W);
    or_subgroup<R, 1, 2>(U, W);
    gram_schmidt3(U.e[0], U.e[1]);
    if (g.fold_gs) gram_schmidt3(U.e[0], U.e[1]);
    store3<R, MU>(table, g, site.gindex, U.e[0], U.e[1]);
}

// ---------------------------------------------------------------- EXPERIMENT: the pass as two kernels (build with -DQCD_SPLIT_EXPERIMENT, run with QCDGPU_SPLIT=1)
// A warp of the fused kernel is bound by its own dependent chains at three warps per scheduler, and 168 registers allow no more.  Cut
// between staples and heat bath, each half needs fewer: staple_kernel_tma (TMA-staged links -> the staple sum S of every updated link,
// 9 complex numbers, to a scratch buffer: plane e of launch thread k at scratch[e * nthreads + k]) and heatbath_kernel_rlx (generator
// phase as in update_kernel_rlx, then S from the scratch buffer and the link itself from the table, Cabibbo-Marinari, store).  Same functions on the same values:
// bit-identical lattices.  Costs 288 B of extra traffic per update.
// MEASURED, NOT SHIPPED (64^4 fp64): 32.7-33.3 against 28.1 ms per sweep.  staple_kernel_tma 1.87 ms at 4 blocks per SM with the issue
// slots 31 % busy and 8.8 GB of DRAM traffic = 4.7 TB/s, 73 % of the measured copy bandwidth: the staple phase on its own is bound by
// HBM; heatbath_kernel_rlx 2.10 ms (128 registers, 184 B of spills; 168 registers / 3 blocks per SM: the same).  The fused kernel's
// 3.43 ms is less than the sum because its generator and Cabibbo-Marinari phases run under the staple phase's memory traffic.
#ifdef QCD_SPLIT_EXPERIMENT
#ifndef QCD_SPLIT_MINBLOCKS
#define QCD_SPLIT_MINBLOCKS 4
#endif
#ifndef QCD_SPLIT_B_MINBLOCKS
#define QCD_SPLIT_B_MINBLOCKS 4
#endif
template <typename R> struct Pair { R a, b; };
template <typename R> __device__ __forceinline__ void st_pair(R* p, R a, R b);
template <> __device__ __forceinline__ void st_pair<double>(double* p, double a, double b) { asm volatile("st.global.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(a), "d"(b) : "memory"); }
template <> __device__ __forceinline__ void st_pair<float>(float* p, float a, float b) { asm volatile("st.global.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory"); }
template <typename R> __device__ __forceinline__ Cx<R> ld_pair(const R* p);
template <> __device__ __forceinline__ Cx<double> ld_pair<double>(const double* p) { Cx<double> c; asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(c.re), "=d"(c.im) : "l"(p)); return c; }
template <> __device__ __forceinline__ Cx<float> ld_pair<float>(const float* p) { Cx<float> c; asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(c.re), "=f"(c.im) : "l"(p)); return c; }

template <typename R, int MU, bool TBC>
__global__ void __launch_bounds__(kUpdateThreads, (QCD_SPLIT_MINBLOCKS * 128) / kUpdateThreads)
staple_kernel_tma(const R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, int odd,
                  const __grid_constant__ TensorMapBlob tmap, const __grid_constant__ TensorMapBlob tmap1, R* __restrict__ scratch) {
    extern __shared__ __align__(128) uint4 prn_smem[];
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.nthreads) return;
    const SiteId site = thread_site(g, k, odd != 0);
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(prn_smem), warp = threadIdx.x >> 5;
    constexpr uint32_t NW = kUpdateThreads / 32, RING = kTmaK * tma_slot_bytes<R>();
    Su3StageTma<R, MU, TBC> stage;
    stage.begin(base + warp * RING, base + NW * RING + warp * (kTmaK * 128u), base + NW * (RING + kTmaK * 128u) + warp * (kTmaK * 8u),
                &tmap, &tmap1, table, g, site.gindex, site.p);
    const Twist<R> tw = make_twist<3, R, TBC>(params);
    M3<R> S;
    M3r2<R> m0;
    stage.staple_and_link(g, tw, S, m0);
    R* out = scratch + 2 * size_t(k);
    const size_t plane = 2 * size_t(g.nthreads);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) st_pair<R>(out + (3 * i + j) * plane, S.e[i][j].re, S.e[i][j].im);
}
template <typename R, int MU, bool TBC, int NSKIP>
__global__ void __launch_bounds__(kUpdateThreads, (QCD_SPLIT_B_MINBLOCKS * 128) / kUpdateThreads)
heatbath_kernel_rlx(R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, int odd,
                    const uint32_t* rtable, uint32_t* rtable_out, uint32_t instances, uint32_t samples, uint32_t keep,
                    const uint32_t* __restrict__ jump, const R* __restrict__ scratch) {
    extern __shared__ __align__(128) uint4 prn_smem[];
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.nthreads) return;
    const SiteId site = thread_site(g, k, odd != 0);
    const uint32_t GID = site.GID;
    unsigned long long converted[1] = { 0ull };
    RanluxPk<1> r;
    uint32_t left = r.load(0, rtable, instances, GID);
    {
        uint4* const out[1] = { prn_smem + threadIdx.x };
        r.pass(left, samples, keep, out, kUpdateThreads, converted, [] {}, jump);
        r.store(0, rtable_out, instances, GID, left);
    }
    const R beta = params[0];
    InvBeta<R>::park(params);
    typedef PrnSmemT<NSKIP> PrnSmem;
    PrnSmem prn(prn_smem + threadIdx.x, kUpdateThreads, converted[0], (rtable_out != rtable) ? keep : 0xffffffffu, rtable + GID, instances);
    const R* in = scratch + 2 * size_t(k);
    const size_t plane = 2 * size_t(g.nthreads);
    M3<R> S;
    M3r2<R> m0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S.e[i][j] = ld_pair<R>(in + (3 * i + j) * plane);
    const Twist<R> tw = make_twist<3, R, TBC>(params);
    m0 = load3<R, MU, TBC>(table, g, site.gindex, site.p.x, tw);
    heatbath_su3<R, MU, PrnSmem>(table, g, beta, site.gindex, m0, S, prn);
}
#endif  // QCD_SPLIT_EXPERIMENT

// ---------------------------------------------------------------- warp-specialised persistent update kernel (RANLUX3, SU(3), TMA rows)
// One block of 512 threads per SM for the whole pass: warp groups 0-2 are 12 "physics" warps (staples, Cabibbo-Marinari, store), warp
// group 3 is four "generator" warps.  Generator warp q and the physics warps q, q + 4, q + 8 share a scheduler and form an independent
// unit: the generator warp draws a ticket (96 consecutive update threads = three y-rows of one parity, in launch order, so the L2 window
// of remap_gid stays), hands one row to each of its physics warps and runs the RANLUX work of the three rows as THREE instances per lane,
// stepped together -- the three borrow chains interleave in the instruction stream and keep the integer pipe busy on their own, while the
// physics warps keep the FP64 pipe and the memory system busy.  In the two-phase kernels a warp alternates between the two kinds of work
// and the scheduler, three warps deep, rarely sees both at the same time (fused time = generator time + physics time).
// setmaxnreg moves registers from the generator warp group (80) to the physics warp groups (144).  Samples travel through the same
// shared-memory columns as in update_kernel_rlx_tma (first `keep` float4 per link, raw; rare overflow path from the state in front of the
// pass), one buffer per physics warp: the generator parks early in its round and the physics warp consumes late in its item, so a single
// buffer with a full / empty mbarrier pair pipelines the two.  Results are bit-identical to the other kernels.
#ifndef QCD_WS_GEN_REGS       // 4 x gen + 12 x phys = 16 x 128 (the block's registers at launch); multiples of 8
#define QCD_WS_GEN_REGS 104
#define QCD_WS_PHYS_REGS 136
#endif
#ifdef QCD_WS_TRACE
// diagnostic build (-DQCD_WS_TRACE): clock64 stamps of the first 64 rounds of every unit (generator warp + its three physics warps):
// slots 0-5 generator (round start, ticket, states loaded, buffers free, samples parked, states stored), 6 + 4 c .. physics warp c
// (ticket, staples done, samples there, link stored)
static __device__ unsigned long long g_ws_trace[592 * 64 * 24];
#define QCD_WS_MARK(unit, round, slot) if ((threadIdx.x & 31u) == 0 && (round) < 64u && (unit) < 592u) g_ws_trace[((unit) * 64u + (round)) * 24u + (slot)] = clock64()
#else
#define QCD_WS_MARK(unit, round, slot)
#endif
constexpr int kWsPhysWarps = 12, kWsGenWarps = 4, kWsThreads = 32 * (kWsPhysWarps + kWsGenWarps), kWsChains = kWsPhysWarps / kWsGenWarps;
constexpr uint32_t kWsDone = 0xffffffffu;
template <typename R> __host__ __device__ constexpr size_t ws_smem_bytes(uint32_t keep) {
    return size_t(kWsPhysWarps) * (size_t(keep) * 32 * 16            // parked samples
                                   + 32 * 8                          // "already converted" masks
                                   + kTmaK * (tma_slot_bytes<R>() + 128 + 8)   // link ring, wrapped-site slots, ring mbarriers
                                   + 2 * 8 + 2 * 8                   // item (x2), full, empty mbarriers
                                   + 2 * 4 + 8);                     // tickets (x2), pad
}
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(a) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
template <typename R, int MU, bool TBC, int NSKIP>
__global__ void __launch_bounds__(kWsThreads, 1)
update_kernel_rlx_ws(R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, int odd,
                     const uint32_t* rtable, uint32_t* rtable_out, uint32_t instances, uint32_t samples, uint32_t keep,
                     const __grid_constant__ TensorMapBlob tmap, const __grid_constant__ TensorMapBlob tmap1, uint32_t* counters) {
    static_assert(NSKIP == 223, "written for RANLUX3");
    extern __shared__ __align__(128) uint4 ws_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    // carve-up: [samples 12 x keep x 32 uint4][link rings 12 x K slots][wrapped-site slots][conv masks][ring mbarriers][item/full/empty][tickets]
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(ws_smem);
    const uint32_t SAMP = keep * 32u * 16u, RING = kTmaK * tma_slot_bytes<R>();
    const uint32_t o_ring = kWsPhysWarps * SAMP, o_x = o_ring + kWsPhysWarps * RING, o_conv = o_x + kWsPhysWarps * kTmaK * 128u,
                   o_rbar = o_conv + kWsPhysWarps * 256u, o_bar = o_rbar + kWsPhysWarps * kTmaK * 8u, o_tick = o_bar + kWsPhysWarps * 32u;
    // per physics warp w: item barriers at o_bar + 32 w + {0, 8}, full at + 16, empty at + 24; tickets at o_tick + 8 w + {0, 4}
    if (threadIdx.x < kWsPhysWarps) {
        const uint32_t b = sm0 + o_bar + 32u * threadIdx.x;
        mbar_init(b, 1); mbar_init(b + 8, 1); mbar_init(b + 16, 32); mbar_init(b + 24, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t nthreads = g.nthreads;                 // a multiple of 32 (launcher)
    if (warp >= kWsPhysWarps) {
        // ------------------------------------------------------------ generator warp
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(QCD_WS_GEN_REGS));
        const uint32_t q = warp - kWsPhysWarps;
        uint4* out[kWsChains];
#pragma unroll
        for (int c = 0; c < kWsChains; c++) out[c] = ws_smem + (q + 4 * c) * (SAMP / 16u) + lane;
        for (uint32_t n = 0;; n++) {
            QCD_WS_MARK(blockIdx.x * 4u + q, n, 0);
            uint32_t T = 0;
            if (lane == 0) T = atomicAdd(counters, 1u);
            T = __shfl_sync(0xffffffffu, T, 0);
            const uint32_t kb0 = T * (32u * kWsChains);
            const bool any = kb0 < nthreads;
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < kWsChains; c++) {
                    const uint32_t w = q + 4 * c, kb = kb0 + 32u * c;
                    asm volatile("st.shared.u32 [%0], %1;" :: "r"(sm0 + o_tick + 8u * w + 4u * (n & 1u)), "r"(kb < nthreads ? kb : kWsDone) : "memory");
                    mbar_arrive(sm0 + o_bar + 32u * w + 8u * (n & 1u));
                }
            }
            if (!any) break;
            QCD_WS_MARK(blockIdx.x * 4u + q, n, 1);
            RanluxPk<kWsChains> G;
            uint32_t left = 0, GID[kWsChains];
            bool uniform = true;
            if (g.debug_skip != 1u) {
#pragma unroll
                for (int c = 0; c < kWsChains; c++) {
                    const bool valid = kb0 + 32u * c < nthreads;
                    GID[c] = thread_site(g, valid ? kb0 + 32u * c + lane : 0u, odd != 0).GID;
                    const uint32_t l = G.load(c, rtable, instances, GID[c]);
                    if (c == 0) left = __shfl_sync(0xffffffffu, l, 0);
                    uniform = uniform && (!valid || l == left);
                }
                uniform = __all_sync(0xffffffffu, uniform) && g.debug_skip != 3u;      // QCDGPU_DEBUG_SKIP=3: test the general path
            }
            QCD_WS_MARK(blockIdx.x * 4u + q, n, 2);
            // the buffers are free once the physics warps are through with the previous item
            if (n > 0) {
#pragma unroll
                for (int c = 0; c < kWsChains; c++) if (kb0 + 32u * c < nthreads) mbar_wait(sm0 + o_bar + 32u * (q + 4 * c) + 24u, (n - 1u) & 1u);
            }
            QCD_WS_MARK(blockIdx.x * 4u + q, n, 3);
            unsigned long long conv[kWsChains];
#pragma unroll
            for (int c = 0; c < kWsChains; c++) conv[c] = 0ull;
            auto signal_full = [&]() {
#pragma unroll
                for (int c = 0; c < kWsChains; c++) {
                    const uint32_t w = q + 4 * c;
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(sm0 + o_conv + 256u * w + 8u * lane), "r"((uint32_t)conv[c]), "r"((uint32_t)(conv[c] >> 32)) : "memory");
                    if (kb0 + 32u * c < nthreads) mbar_arrive(sm0 + o_bar + 32u * w + 16u);
                }
                QCD_WS_MARK(blockIdx.x * 4u + q, n, 4);
            };
            if (g.debug_skip == 1u) { signal_full(); continue; }
            if (uniform) {
                G.pass(left, samples, keep, out, 32u, conv, signal_full);
#pragma unroll
                for (int c = 0; c < kWsChains; c++)
                    if (kb0 + 32u * c < nthreads) G.store(c, rtable_out, instances, GID[c], left);
                QCD_WS_MARK(blockIdx.x * 4u + q, n, 5);
            } else {
                // instances out of step with each other (never after ranlux_init + whole passes): one instance at a time
#pragma unroll 1
                for (int c = 0; c < kWsChains; c++) {
                    if (kb0 + 32u * c >= nthreads) continue;
                    const uint32_t gid = thread_site(g, kb0 + 32u * c + lane, odd != 0).GID;
                    RanluxPk<1> r;
                    uint32_t l = r.load(0, rtable, instances, gid);
                    unsigned long long cv[1] = { 0ull };
                    uint4* const o1[1] = { ws_smem + (q + 4 * c) * (SAMP / 16u) + lane };
                    r.pass(l, samples, keep, o1, 32u, cv, [] {});
                    r.store(0, rtable_out, instances, gid, l);
                    conv[c] = cv[0];
                }
                signal_full();
            }
        }
        // the last generator warp to run out of tickets re-arms the counters for the next launch
        if (lane == 0) {
            const uint32_t d = atomicAdd(counters + 1, 1u);
            if (d == gridDim.x * kWsGenWarps - 1u) { counters[0] = 0u; counters[1] = 0u; __threadfence(); }
        }
        return;
    }
    // ---------------------------------------------------------------- physics warp
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(QCD_WS_PHYS_REGS));
    const uint32_t w = warp;
    Su3StageTma<R, MU, TBC, true> stage;
    stage.init_ws(sm0 + o_ring + w * RING, sm0 + o_x + w * (kTmaK * 128u), sm0 + o_rbar + w * (kTmaK * 8u), &tmap, &tmap1, table);
    const R beta = params[0];
    InvBeta<R>::park(params);
    const Twist<R> tw = make_twist<3, R, TBC>(params);
    const uint32_t bar = sm0 + o_bar + 32u * w;
    for (uint32_t n = 0;; n++) {
        mbar_wait(bar + 8u * (n & 1u), (n >> 1) & 1u);
        uint32_t kb;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(kb) : "r"(sm0 + o_tick + 8u * w + 4u * (n & 1u)) : "memory");
        if (kb == kWsDone) break;
        QCD_WS_MARK(blockIdx.x * 4u + (w & 3u), n, 6u + 4u * (w >> 2));
        const SiteId site = thread_site(g, kb + lane, odd != 0);
        if (g.debug_skip == 2u) { mbar_wait(bar + 16u, n & 1u); mbar_arrive(bar + 24u); continue; }
        stage.begin_item(g, site.gindex, site.p);
        M3<R> S;
        M3r2<R> m0;
        stage.staple_and_link(g, tw, S, m0);
        stage.end_item();
        QCD_WS_MARK(blockIdx.x * 4u + (w & 3u), n, 7u + 4u * (w >> 2));
        mbar_wait(bar + 16u, n & 1u);
        QCD_WS_MARK(blockIdx.x * 4u + (w & 3u), n, 8u + 4u * (w >> 2));
        uint32_t cv0, cv1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cv0), "=r"(cv1) : "r"(sm0 + o_conv + 256u * w + 8u * lane) : "memory");
        typedef PrnSmemT<NSKIP> PrnSmem;
        PrnSmem prn(ws_smem + w * (SAMP / 16u) + lane, 32u, ((unsigned long long)cv1 << 32) | cv0, keep, rtable + site.GID, instances);
        heatbath_su3<R, MU, PrnSmem>(table, g, beta, site.gindex, m0, S, prn);
        QCD_WS_MARK(blockIdx.x * 4u + (w & 3u), n, 9u + 4u * (w >> 2));
        mbar_arrive(bar + 24u);
    }
}


// ---------------------------------------------------------------- EXPERIMENT: persistent kernel, symmetric warp groups, registers on loan
// (build with -DQCD_SYM_EXPERIMENT, run with QCDGPU_SYM=1).  The fused kernel's time is generator phase + physics phase because a block keeps its 168 registers per thread while
// it runs the generator, which needs 50.  Here one block of 512 threads per SM holds the whole register file and its four warp groups
// (4 warps = 128 update threads = four y-rows) each loop over tickets on their own: generator phase at QCD_SYM_GEN_REGS registers
// (setmaxnreg.dec), physics phase at QCD_SYM_PHYS_REGS (setmaxnreg.inc, which waits until another warp group has given registers back):
// 3 x 136 + 104 = 512 = the launch's 128 per thread, so three warp groups can be in their physics phase while the fourth generates --
// sixteen warps per SM instead of twelve, the generator phase under the physics phases.  Unlike update_kernel_rlx_ws no warp works for
// another one: samples stay in the warp's own shared-memory columns, links in its own TMA ring.  Bit-identical to the other kernels.
// MEASURED, NOT SHIPPED (64^4 fp64): 36.1-36.4 against 27.9 ms per sweep (with the jump in the generator phase 36.8).  The plan needs a
// generator phase that takes a third of a physics phase (5 us); a warp's generator phase is one dependent chain (state load, two
// luxury skips, the tail) of ~10 k cycles even undisturbed, and sharing a scheduler with three physics warps it takes ~10 us: the one
// generating warp group is the bottleneck (443 tickets x 10 us = 4.4 ms per pass), as in update_kernel_rlx_ws.  Also: the four warps
// of a group wait for each other at every setmaxnreg, physics at 136 registers is 4 % slower, 8 parked samples 2 %.
#ifdef QCD_SYM_EXPERIMENT
#ifndef QCD_SYM_GEN_REGS
#define QCD_SYM_GEN_REGS 104
#define QCD_SYM_PHYS_REGS 136
#endif
constexpr int kSymWarps = 16, kSymThreads = 32 * kSymWarps;
template <typename R> __host__ __device__ constexpr size_t sym_smem_bytes(uint32_t keep) {
    return size_t(kSymWarps) * (size_t(keep) * 32 * 16 + kTmaK * (tma_slot_bytes<R>() + 128 + 8)) + 64;
}
template <typename R, int MU, bool TBC, int NSKIP>
__global__ void __launch_bounds__(kSymThreads, 1)
update_kernel_rlx_sym(R* __restrict__ table, const R* __restrict__ params, const __grid_constant__ Geom g, int odd,
                      const uint32_t* rtable, uint32_t* rtable_out, uint32_t instances, uint32_t samples, uint32_t keep,
                      const __grid_constant__ TensorMapBlob tmap, const __grid_constant__ TensorMapBlob tmap1,
                      const uint32_t* __restrict__ jump, uint32_t* counters) {
    static_assert(NSKIP == 223, "written for RANLUX3");
    extern __shared__ __align__(128) uint4 sym_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, wg = warp >> 2;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(sym_smem);
    const uint32_t SAMP = keep * 32u * 16u, RING = kTmaK * tma_slot_bytes<R>();
    const uint32_t o_ring = kSymWarps * SAMP, o_x = o_ring + kSymWarps * RING, o_rbar = o_x + kSymWarps * kTmaK * 128u,
                   o_tick = o_rbar + kSymWarps * kTmaK * 8u;
    Su3StageTma<R, MU, TBC, true> stage;
    stage.init_ws(sm0 + o_ring + warp * RING, sm0 + o_x + warp * (kTmaK * 128u), sm0 + o_rbar + warp * (kTmaK * 8u), &tmap, &tmap1, table);
    const R beta = params[0];
    InvBeta<R>::park(params);
    const Twist<R> tw = make_twist<3, R, TBC>(params);
    const uint32_t nthreads = g.nthreads;                 // a multiple of 32 (launcher)
    uint4* const column = sym_smem + warp * (SAMP / 16u) + lane;
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(QCD_SYM_GEN_REGS));
    for (uint32_t n = 0;; n++) {
        // the warp group's ticket: 128 consecutive update threads in launch order (so the L2 window of remap_gid stays)
        const uint32_t slot = sm0 + o_tick + 16u * wg + 4u * (n & 1u);
        if ((threadIdx.x & 127u) == 0u) {
            const uint32_t t = atomicAdd(counters, 1u);
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(slot), "r"(t) : "memory");
        }
        asm volatile("bar.sync %0, 128;" :: "r"(1u + wg) : "memory");
        uint32_t T;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(T) : "r"(slot) : "memory");
        if (T * 128u >= nthreads || T >= 0x01000000u) break;
        const uint32_t kb = T * 128u + (warp & 3u) * 32u;
        const bool active = kb < nthreads;
        SiteId site = {};
        unsigned long long conv[1] = { 0ull };
        if (active) {
            site = thread_site(g, kb + lane, odd != 0);
            stage.begin_item(g, site.gindex, site.p);          // the first links travel during the generator phase
            if (g.debug_skip != 1u) {                          // QCDGPU_DEBUG_SKIP=1: physics only (diagnostic: stale samples)
                RanluxPk<1> r;
                uint32_t left = r.load(0, rtable, instances, site.GID);
                uint4* const out[1] = { column };
                r.pass(left, samples, keep, out, 32u, conv, [] {}, jump);
                r.store(0, rtable_out, instances, site.GID, left);
            }
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(QCD_SYM_PHYS_REGS));
        if (active && g.debug_skip != 2u) {                    // QCDGPU_DEBUG_SKIP=2: generator only
            M3<R> S;
            M3r2<R> m0;
            stage.staple_and_link(g, tw, S, m0);
            stage.end_item();
            typedef PrnSmemT<NSKIP> PrnSmem;
            PrnSmem prn(column, 32u, conv[0], g.debug_skip == 1u ? 0xffffffffu : keep, rtable + site.GID, instances);
            heatbath_su3<R, MU, PrnSmem>(table, g, beta, site.gindex, m0, S, prn);
        }
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(QCD_SYM_GEN_REGS));
    }
    // the last warp group to run out of tickets re-arms the counters for the next launch
    if ((threadIdx.x & 127u) == 0u) {
        const uint32_t d = atomicAdd(counters + 1, 1u);
        if (d == gridDim.x * 4u - 1u) { counters[0] = 0u; counters[1] = 0u; __threadfence(); }
    }
}
#endif  // QCD_SYM_EXPERIMENT

}  // namespace qcd
