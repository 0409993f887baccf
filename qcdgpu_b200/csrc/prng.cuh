// prng.cuh -- device-side pseudo-random generators for the gauge-update path.
//
// Streams are bit-exact with the reference's generators:
//   RANLUX  (Luescher/James subtract-with-borrow, luxury skip 24/48/97/223/389)
//           reference: SU(N)/random/prngcl_ranlux.cl:97-563, host replica random.cpp:504-574
//   XOR128  (Marsaglia xorshift128; emitted word = state[0] after the shift, the
//           host/device-consistent O(N) flavour: O(N)/random/random.cl:137-173, random.cpp:636-652)
//   PM      (Park-Miller minimal standard, 4 independent lanes: prngcl_pm.cl:86-123)
//   MRG32k3a(L'Ecuyer; prngcl_mrg32k3a.cl:73-150 -- no host code in the reference)
//
// New design: the reference keeps the 24 RANLUX seeds as fp32 multiples of 2^-24
// and steps them with float compares; here they are 24-bit integers, one
// subtract-with-borrow is an IADD3 + a mask, and numbers are converted to float
// only when a consumer actually asks for them.  The float results are identical:
// every reference operation on those values is exact in fp32 except the
// "uni < 2^-12" fix-up, which is reproduced with round-to-nearest float adds.
//
// Each generator is a small struct living in registers; `Buffered` reads the
// reference-layout PRN buffer instead (prns[GID + k*PRNGSTEP]).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qcd {

struct F4 { float x, y, z, w; };

// ---------------------------------------------------------------- RANLUX
// Register state: s[0..23] hold the ring buffer ROTATED so that the next number to be
// produced always uses i24 = 23, j24 = 9: all indices are compile-time constants and the
// 24 words stay in registers.  `left` = numbers left before the next luxury skip (0..24).
struct Ranlux {
    uint32_t s[24];
    uint32_t carry;   // 0 / 1
    uint32_t left;    // in24 of the host replica

    // one block of 24 subtract-with-borrow steps: afterwards the ring is back in canonical rotation
    __device__ __forceinline__ void block24() {
#pragma unroll
        for (int k = 0; k < 24; k++) {
            const int i = 23 - k, j = (9 - k + 24) % 24;
            int32_t d = int32_t(s[j]) - int32_t(s[i]) - int32_t(carry);
            carry = uint32_t(d) >> 31;
            s[i] = uint32_t(d) & 0xFFFFFFu;
        }
    }
    // n (< 24) single steps followed by a rotation that restores the canonical indices
    template <int N> __device__ __forceinline__ void steps_rot() {
        if (N == 0) return;
#pragma unroll
        for (int k = 0; k < N; k++) {
            const int i = 23 - k, j = (9 - k + 24) % 24;
            int32_t d = int32_t(s[j]) - int32_t(s[i]) - int32_t(carry);
            carry = uint32_t(d) >> 31;
            s[i] = uint32_t(d) & 0xFFFFFFu;
        }
        // after N steps the live indices are i24 = 23-N, j24 = 9-N: rotate up by N
        uint32_t t[24];
#pragma unroll
        for (int k = 0; k < 24; k++) t[(k + N) % 24] = s[k];
#pragma unroll
        for (int k = 0; k < 24; k++) s[k] = t[k];
    }
    template <int NSKIP> __device__ __forceinline__ void luxury_skip() {
#pragma unroll 1
        for (int b = 0; b < NSKIP / 24; b++) block24();
        steps_rot<NSKIP % 24>();
    }
    // four consecutive numbers (one float4 sample).  24 % 4 == 0, so a sample never straddles a skip.
    template <int NSKIP> __device__ __forceinline__ F4 next4() {
        if (left == 0) { luxury_skip<NSKIP>(); left = 24; }
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = 23 - k, j = (9 - k + 24) % 24, jn = (8 - k + 24) % 24;
            int32_t d = int32_t(s[j]) - int32_t(s[i]) - int32_t(carry);
            carry = uint32_t(d) >> 31;
            uint32_t u = uint32_t(d) & 0xFFFFFFu;
            s[i] = u;
            float f = __uint2float_rn(u) * 5.9604644775390625e-8f;          // exact
            if (u < 4096u) {                                               // uni < 2^-12
                f = __fadd_rn(f, __fmul_rn(5.9604644775390625e-8f, __uint2float_rn(s[jn]) * 5.9604644775390625e-8f));
                if (f == 0.0f) f = 5.9604644775390625e-8f * 5.9604644775390625e-8f;
            }
            o[k] = f;
        }
        // rotate by 4
        uint32_t t[24];
#pragma unroll
        for (int k = 0; k < 24; k++) t[(k + 4) % 24] = s[k];
#pragma unroll
        for (int k = 0; k < 24; k++) s[k] = t[k];
        left -= 4;
        return { o[0], o[1], o[2], o[3] };
    }
    // advance by one float4 sample without converting
    template <int NSKIP> __device__ __forceinline__ void skip4() {
        if (left == 0) { luxury_skip<NSKIP>(); left = 24; }
        steps_rot<4>();
        left -= 4;
    }
    // advance by n float4 samples
    template <int NSKIP> __device__ __forceinline__ void skip_samples(uint32_t n) {
        // bring to a 24-boundary, then whole periods of (24 outputs + NSKIP) in block form
        while (n > 0 && left != 0) { steps_rot<4>(); left -= 4; n--; }
        while (n >= 6) {
            luxury_skip<NSKIP>();
            block24();
            n -= 6;
        }
        while (n > 0) { skip4<NSKIP>(); n--; }
    }
};

// state table layout in global memory (this backend's own; the reference's float4 x 7 table is
// never read by the host except through PRNG::check_seeds):
//   word w (0..25) of instance i at  table[w * instances + i]; w 0..23 seeds, 24 carry, 25 left
__device__ __forceinline__ void ranlux_load(Ranlux& r, const uint32_t* table, uint32_t instances, uint32_t i) {
#pragma unroll
    for (int w = 0; w < 24; w++) r.s[w] = table[size_t(w) * instances + i];
    r.carry = table[size_t(24) * instances + i];
    r.left = table[size_t(25) * instances + i];
}
__device__ __forceinline__ void ranlux_store(const Ranlux& r, uint32_t* table, uint32_t instances, uint32_t i) {
#pragma unroll
    for (int w = 0; w < 24; w++) table[size_t(w) * instances + i] = r.s[w];
    table[size_t(24) * instances + i] = r.carry;
    table[size_t(25) * instances + i] = r.left;
}
// ranlux_init (prngcl_ranlux.cl:97-157): 24 seeds from one 31-bit seed through L'Ecuyer's LCG
__device__ __forceinline__ void ranlux_seed(Ranlux& r, int32_t jseed) {
#pragma unroll 1
    for (int i = 0; i < 24; i++) {
        int32_t k = jseed / 53668;
        jseed = 40014 * (jseed - k * 53668) - k * 12211;
        if (jseed < 0) jseed += 2147483563;
        r.s[i] = uint32_t(jseed % 16777216);
    }
    r.carry = (r.s[23] == 0u) ? 1u : 0u;
    r.left = 0u;        // the first request performs the luxury skip first (RL_in24 = 0 / in24 = 6.0f)
}

// ---------------------------------------------------------------- XOR128
struct Xor128 {
    uint32_t x, y, z, w;
    __device__ __forceinline__ uint32_t step() {
        uint32_t t = x ^ (x << 11);
        x = y; y = z; z = w;
        w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
        return x;
    }
    __device__ __forceinline__ F4 next4() {
        F4 r;
        r.x = __uint2float_rn(step()) / 4294967296.0f;
        r.y = __uint2float_rn(step()) / 4294967296.0f;
        r.z = __uint2float_rn(step()) / 4294967296.0f;
        r.w = __uint2float_rn(step()) / 4294967296.0f;
        return r;
    }
    __device__ __forceinline__ void skip4() { step(); step(); step(); step(); }
};

// ---------------------------------------------------------------- Park-Miller (4 lanes)
struct ParkMiller {
    uint32_t s[4];
    __device__ __forceinline__ static uint32_t step1(uint32_t seed) {
        uint32_t t1 = 16807u * (seed % 127773u);
        uint32_t t2 = 2836u * (seed / 127773u);
        return (t1 > t2) ? (t1 - t2) : (t1 - t2 + 2147483647u);
    }
    __device__ __forceinline__ F4 next4() {
#pragma unroll
        for (int k = 0; k < 4; k++) s[k] = step1(s[k]);
        return { __uint2float_rn(s[0]) / 2147483648.0f, __uint2float_rn(s[1]) / 2147483648.0f,
                 __uint2float_rn(s[2]) / 2147483648.0f, __uint2float_rn(s[3]) / 2147483648.0f };
    }
    __device__ __forceinline__ void skip4() {
#pragma unroll
        for (int k = 0; k < 4; k++) s[k] = step1(s[k]);
    }
};

// ---------------------------------------------------------------- MRG32k3a
struct Mrg32k3a {
    uint32_t a[3], b[3];
    __device__ __forceinline__ float step() {
        long long x1n = 1403580LL * (long long)a[1] - 810728LL * (long long)a[0];
        long long k = x1n / 4294967087LL;
        x1n -= k * 4294967087LL;
        if (x1n < 0) x1n += 4294967087LL;
        a[0] = a[1]; a[1] = a[2]; a[2] = uint32_t(x1n);
        long long x2n = 527612LL * (long long)b[2] - 1370589LL * (long long)b[0];
        k = x2n / 4294944443LL;
        x2n -= k * 4294944443LL;
        if (x2n < 0) x2n += 4294944443LL;
        b[0] = b[1]; b[1] = b[2]; b[2] = uint32_t(x2n);
        const double m = 2.3283065492957276882397381481401e-10;
        long long d = x1n - x2n;
        if (!(x1n > x2n)) d += 4294967087LL;
        return __double2float_rn(__dmul_rn(double(d), m));
    }
    __device__ __forceinline__ F4 next4() { F4 r; r.x = step(); r.y = step(); r.z = step(); r.w = step(); return r; }
    __device__ __forceinline__ void skip4() { step(); step(); step(); step(); }
};

}  // namespace qcd
