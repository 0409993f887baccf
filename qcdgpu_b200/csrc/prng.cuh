// prng.cuh -- device-side pseudo-random generators for the gauge-update path.
//
// Streams are bit-exact with the reference's generators:
//   RANLUX  (Luescher/James subtract-with-borrow, luxury skip 24/48/97/223/389)
//           reference: SU(N)/random/prngcl_ranlux.cl:97-563, host replica random.cpp:504-574
//   XOR128  (Marsaglia xorshift128; emitted word = state[0] after the shift, the
//           host/device-consistent O(N) flavour: O(N)/random/random.cl:137-173, random.cpp:636-652)
//   PM      (Park-Miller minimal standard, 4 independent lanes: prngcl_pm.cl:86-123)
//   MRG32k3a(L'Ecuyer; prngcl_mrg32k3a.cl:73-150 -- no host code in the reference)
//   XOR7, RANECU, RANMAR (prngcl_xor7.cl, prngcl_ranecu.cl, prngcl_ranmar.cl; host replicas random.cpp:773-816, 965-986, 1044-1068)
//
// New design: the reference keeps the 24 RANLUX seeds as fp32 multiples of 2^-24
// and steps them with float compares; here they are 24-bit integers, one
// subtract-with-borrow is a carry-flag subtract + a mask, and numbers are converted
// to float only when a consumer can actually use them.  The float results are identical:
// every reference operation on those values is exact in fp32 except the
// "uni < 2^-12" fix-up, which is reproduced with round-to-nearest float adds.
//
// Each generator is a small struct living in registers; `Buffered` reads the
// reference-layout PRN buffer instead (prns[GID + k*PRNGSTEP]).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ranlux_jump.cuh"

namespace qcd {

struct F4 { float x, y, z, w; };

// ---------------------------------------------------------------- RANLUX
// Register state: s[0..23] hold the ring buffer ROTATED so that the next number to be produced
// always uses i24 = 23, j24 = 9: all indices are compile-time constants and the 24 words stay in
// registers.  Each word holds the 24-bit seed in its TOP 24 bits (x << 8, low byte zero): a
// subtract-with-borrow step is then one `subc.cc.u32` whose borrow travels in the hardware carry flag
// (the only serial dependency: ~1 ALU latency per step) plus one `and` that clears the low byte off the
// critical path.  `left` = numbers left before the next luxury skip (0..24), as RL_in24 of the host replica.
#define QCD_SWB(I, J) "subc.cc.u32 %" #I ", %" #J ", %" #I ";\n\tand.b32 %" #I ", %" #I ", 0xFFFFFF00;\n\t"
#define QCD_SWB_OPERANDS(s, c)                                                                              \
    "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]), "+r"(s[6]), "+r"(s[7]),         \
    "+r"(s[8]), "+r"(s[9]), "+r"(s[10]), "+r"(s[11]), "+r"(s[12]), "+r"(s[13]), "+r"(s[14]), "+r"(s[15]),   \
    "+r"(s[16]), "+r"(s[17]), "+r"(s[18]), "+r"(s[19]), "+r"(s[20]), "+r"(s[21]), "+r"(s[22]), "+r"(s[23]), \
    "+r"(c)
// carry (0/1) -> CF, N steps, CF -> carry
#define QCD_SWB_PRE  "{\n\t.reg .u32 t;\n\tsub.cc.u32 t, 0, %24;\n\t"
#define QCD_SWB_POST "subc.u32 t, 0, 0;\n\tand.b32 %24, t, 1;\n\t}"
#define QCD_SWB_1  QCD_SWB(23, 9)
#define QCD_SWB_4  QCD_SWB_1 QCD_SWB(22, 8) QCD_SWB(21, 7) QCD_SWB(20, 6)
#define QCD_SWB_5  QCD_SWB_4 QCD_SWB(19, 5)
#define QCD_SWB_7  QCD_SWB_5 QCD_SWB(18, 4) QCD_SWB(17, 3)
#define QCD_SWB_24 QCD_SWB_7 QCD_SWB(16, 2) QCD_SWB(15, 1) QCD_SWB(14, 0) QCD_SWB(13, 23) QCD_SWB(12, 22)  \
    QCD_SWB(11, 21) QCD_SWB(10, 20) QCD_SWB(9, 19) QCD_SWB(8, 18) QCD_SWB(7, 17) QCD_SWB(6, 16)             \
    QCD_SWB(5, 15) QCD_SWB(4, 14) QCD_SWB(3, 13) QCD_SWB(2, 12) QCD_SWB(1, 11) QCD_SWB(0, 10)

// Packed ring (whole luxury periods of the two-phase kernel): the 24 seeds as one 576-bit number in 18 words, limb m
// (m = 0 the oldest, x[n-24]) at bit 24 m.  x[n] = x[n-10] - x[n-24] - borrow for four consecutive n is then one
// 96-bit subtraction: three subc.cc on words, the minuend being the same number 240 bits = 7.5 words further on (one
// funnel shift per word), and no low byte to clear: 1.5 ALU operations per step instead of 2, with the borrow chain
// 0.75 instructions per step long (measured 3.5 against 4.5 clocks per step with three warps per scheduler,
// tools/microbench/swb_pack.cu).  24 steps bring the ring back to the same registers.
#define QCD_PK_OPERANDS(w, c)                                                                                        \
    "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]),      \
    "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(c)
#define QCD_PK_PRE  "{\n\t.reg .u32 t, a;\n\tsub.cc.u32 t, 0, %18;\n\t"
#define QCD_PK_POST "subc.u32 t, 0, 0;\n\tand.b32 %18, t, 1;\n\t}"
#define QCD_PK(D, LO, HI) "shf.r.wrap.b32 a, %" #LO ", %" #HI ", 16;\n\tsubc.cc.u32 %" #D ", a, %" #D ";\n\t"
#define QCD_PK_CHUNK0 QCD_PK(0, 10, 11) QCD_PK(1, 11, 12) QCD_PK(2, 12, 13)
#define QCD_PK_24 QCD_PK_CHUNK0 QCD_PK(3, 13, 14) QCD_PK(4, 14, 15) QCD_PK(5, 15, 16) QCD_PK(6, 16, 17) QCD_PK(7, 17, 0)   \
    QCD_PK(8, 0, 1) QCD_PK(9, 1, 2) QCD_PK(10, 2, 3) QCD_PK(11, 3, 4) QCD_PK(12, 4, 5) QCD_PK(13, 5, 6) QCD_PK(14, 6, 7)   \
    QCD_PK(15, 7, 8) QCD_PK(16, 8, 9) QCD_PK(17, 9, 10)
#define QCD_PK_CHUNK1 QCD_PK(3, 13, 14) QCD_PK(4, 14, 15) QCD_PK(5, 15, 16)
#define QCD_PK_CHUNK2 QCD_PK(6, 16, 17) QCD_PK(7, 17, 0) QCD_PK(8, 0, 1)
#define QCD_PK_CHUNK3 QCD_PK(9, 1, 2) QCD_PK(10, 2, 3) QCD_PK(11, 3, 4)
#define QCD_PK_CHUNK4 QCD_PK(12, 4, 5) QCD_PK(13, 5, 6) QCD_PK(14, 6, 7)
#define QCD_PK_CHUNK5 QCD_PK(15, 7, 8) QCD_PK(16, 8, 9) QCD_PK(17, 9, 10)
// seven steps = one chunk of four + three limbs (72 bits: two words and the low byte of a third, %19 in / out)
#define QCD_PK_TAIL7 QCD_PK_CHUNK0 QCD_PK(3, 13, 14) QCD_PK(4, 14, 15)                                                 \
    "shf.r.wrap.b32 a, %15, %16, 16;\n\tand.b32 a, a, 0xFF;\n\tand.b32 %19, %5, 0xFF;\n\tsubc.cc.u32 %19, a, %19;\n\t"

struct RanluxPacked {
    uint32_t w[18];
    // s[23 - m] = limb m << 8
    __device__ __forceinline__ void pack(const uint32_t (&s)[24]) {
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const uint32_t y0 = s[23 - 4 * q], y1 = s[22 - 4 * q], y2 = s[21 - 4 * q], y3 = s[20 - 4 * q];
            w[3 * q + 0] = __byte_perm(y0, y1, 0x5321);
            w[3 * q + 1] = __byte_perm(y1, y2, 0x6532);
            w[3 * q + 2] = __byte_perm(y2, y3, 0x7653);
        }
    }
    // limb m << 8 (m compile-time after unrolling)
    __device__ __forceinline__ uint32_t limb8(int m) const {
        const int j = 3 * (m >> 2);
        switch (m & 3) {
            case 0: return w[j] << 8;
            case 1: return __byte_perm(w[j], w[j + 1], 0x5430) & 0xFFFFFF00u;
            case 2: return __byte_perm(w[j + 1], w[j + 2], 0x4320) & 0xFFFFFF00u;
            default: return w[j + 2] & 0xFFFFFF00u;
        }
    }
    __device__ __forceinline__ void unpack(uint32_t (&s)[24]) const {
#pragma unroll
        for (int m = 0; m < 24; m++) s[23 - m] = limb8(m);
    }
    __device__ __forceinline__ void rotation24(uint32_t& carry) {
        asm volatile(QCD_PK_PRE QCD_PK_24 QCD_PK_POST : QCD_PK_OPERANDS(w, carry));
    }
    // seven steps, then the window is moved back onto the word grid (it starts at bit 168 = word 5, bit 8)
    __device__ __forceinline__ void tail7(uint32_t& carry) {
        uint32_t n5;
        asm volatile("{\n\t.reg .u32 t, a;\n\tsub.cc.u32 t, 0, %18;\n\t" QCD_PK_TAIL7
                     "subc.u32 t, 0, 0;\n\tand.b32 %18, t, 1;\n\tand.b32 %19, %19, 0xFF;\n\t}"
                     : QCD_PK_OPERANDS(w, carry), "=r"(n5));
        uint32_t t[18];
#pragma unroll
        for (int k = 0; k < 18; k++) {
            const uint32_t lo = (5 + k < 18) ? w[5 + k] : w[5 + k - 18];
            const uint32_t hi = (6 + k < 18) ? w[6 + k] : (6 + k - 18 < 5 ? w[6 + k - 18] : n5);
            t[k] = __funnelshift_r(lo, hi, 8);
        }
#pragma unroll
        for (int k = 0; k < 18; k++) w[k] = t[k];
    }
    // four steps = one float4 sample: chunk C (0..5, a compile-time constant after unrolling) of a rotation that started at word 0
    __device__ __forceinline__ void chunk(int C, uint32_t& carry) {
        switch (C) {
            case 0: asm volatile(QCD_PK_PRE QCD_PK_CHUNK0 QCD_PK_POST : QCD_PK_OPERANDS(w, carry)); break;
            case 1: asm volatile(QCD_PK_PRE QCD_PK_CHUNK1 QCD_PK_POST : QCD_PK_OPERANDS(w, carry)); break;
            case 2: asm volatile(QCD_PK_PRE QCD_PK_CHUNK2 QCD_PK_POST : QCD_PK_OPERANDS(w, carry)); break;
            case 3: asm volatile(QCD_PK_PRE QCD_PK_CHUNK3 QCD_PK_POST : QCD_PK_OPERANDS(w, carry)); break;
            case 4: asm volatile(QCD_PK_PRE QCD_PK_CHUNK4 QCD_PK_POST : QCD_PK_OPERANDS(w, carry)); break;
            default: asm volatile(QCD_PK_PRE QCD_PK_CHUNK5 QCD_PK_POST : QCD_PK_OPERANDS(w, carry)); break;
        }
    }
    // after J chunks of a rotation: renumber the words so that the next chunk is chunk 0 again (w'[i] = w[(i + 3 J) % 18])
    template <int J> __device__ __forceinline__ void renumber() {
        uint32_t t[18];
#pragma unroll
        for (int i = 0; i < 18; i++) t[i] = w[(i + 3 * J) % 18];
#pragma unroll
        for (int i = 0; i < 18; i++) w[i] = t[i];
    }
    // the float4 sample made by chunk C (numbers 4 C .. 4 C + 3 of the rotation), raw (seed << 8) unless it holds a number below 2^-12,
    // whose fix-up reads x[n-9]: limb n - 9 of this rotation, or limb n + 15 (still the previous rotation's: chunks 3.. come later).
    // Call right after chunk(C).  Returns true if the sample was converted to floats.
    __device__ __forceinline__ bool sample(int C, uint4& out) const;
    // three rotations in one statement: the borrow stays in the carry flag across them (each statement boundary costs three
    // dependent instructions to park and restore it)
    __device__ __forceinline__ void rotation72(uint32_t& carry) {
        asm volatile(QCD_PK_PRE QCD_PK_24 QCD_PK_24 QCD_PK_24 QCD_PK_POST : QCD_PK_OPERANDS(w, carry));
    }
    // luxury skip of 223 = 9 x 24 + 7 steps
    __device__ __forceinline__ void skip223(uint32_t& carry) {
#pragma unroll 1
        for (int b = 0; b < 3; b++) rotation72(carry);
        tail7(carry);
    }
};

struct Ranlux {
    uint32_t s[24];   // seed << 8
    uint32_t carry;   // 0 / 1
    uint32_t left;    // in24 of the host replica

    // one block of 24 subtract-with-borrow steps: afterwards the ring is back in canonical rotation
    __device__ __forceinline__ void block24() {
        asm volatile(QCD_SWB_PRE QCD_SWB_24 QCD_SWB_POST : QCD_SWB_OPERANDS(s, carry));
    }
    template <int N> __device__ __forceinline__ void rotate() {
        uint32_t t[24];
#pragma unroll
        for (int k = 0; k < 24; k++) t[(k + N) % 24] = s[k];
#pragma unroll
        for (int k = 0; k < 24; k++) s[k] = t[k];
    }
    // N (< 24) single steps followed by the rotation that restores the canonical indices
    template <int N> __device__ __forceinline__ void steps_rot() {
        static_assert(N == 0 || N == 1 || N == 4 || N == 5 || N == 7, "RL_skip % 24 must be 0, 1, 5 or 7");
        if (N == 0) return;
        if (N == 1) asm volatile(QCD_SWB_PRE QCD_SWB_1 QCD_SWB_POST : QCD_SWB_OPERANDS(s, carry));
        if (N == 4) asm volatile(QCD_SWB_PRE QCD_SWB_4 QCD_SWB_POST : QCD_SWB_OPERANDS(s, carry));
        if (N == 5) asm volatile(QCD_SWB_PRE QCD_SWB_5 QCD_SWB_POST : QCD_SWB_OPERANDS(s, carry));
        if (N == 7) asm volatile(QCD_SWB_PRE QCD_SWB_7 QCD_SWB_POST : QCD_SWB_OPERANDS(s, carry));
        rotate<N>();
    }
    template <int NSKIP> __device__ __forceinline__ void luxury_skip() {
#pragma unroll 1
        for (int b = 0; b < NSKIP / 24; b++) block24();
        steps_rot<NSKIP % 24>();
    }
    // four consecutive numbers (one float4 sample).  24 % 4 == 0, so a sample never straddles a skip.
    template <int NSKIP> __device__ __forceinline__ F4 next4() {
        if (left == 0) { luxury_skip<NSKIP>(); left = 24; }
        asm volatile(QCD_SWB_PRE QCD_SWB_4 QCD_SWB_POST : QCD_SWB_OPERANDS(s, carry));
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t x = s[23 - k];                       // new seed << 8
            float f = __uint2float_rn(x) * 2.3283064365386963e-10f;        // (x >> 8) * 2^-24, exact
            if (x < (4096u << 8)) {                                        // uni < 2^-12: add 2^-24 * seeds[j24 after the step]
                f = __fadd_rn(f, __fmul_rn(5.9604644775390625e-8f, __uint2float_rn(s[8 - k]) * 2.3283064365386963e-10f));
                if (f == 0.0f) f = 5.9604644775390625e-8f * 5.9604644775390625e-8f;
            }
            o[k] = f;
        }
        rotate<4>();
        left -= 4;
        return { o[0], o[1], o[2], o[3] };
    }
    // float of the k-th number just produced (x = new seed << 8, nx = seeds[j24 after the step] << 8), reference
    // semantics incl. the "uni < 2^-12" fix-up (random.cpp:565-569)
    static __device__ __forceinline__ float to_float(uint32_t x, uint32_t nx) {
        float f = __uint2float_rn(x) * 2.3283064365386963e-10f;            // (x >> 8) * 2^-24, exact
        if (x < (4096u << 8)) {
            f = __fadd_rn(f, __fmul_rn(5.9604644775390625e-8f, __uint2float_rn(nx) * 2.3283064365386963e-10f));
            if (f == 0.0f) f = 5.9604644775390625e-8f * 5.9604644775390625e-8f;
        }
        return f;
    }
    // advance by one float4 sample without converting
    template <int NSKIP> __device__ __forceinline__ void skip4() {
        if (left == 0) { luxury_skip<NSKIP>(); left = 24; }
        steps_rot<4>();
        left -= 4;
    }
    // advance by n float4 samples
    template <int NSKIP> __device__ __forceinline__ void skip_samples(uint32_t n) {
        // bring to a 24-boundary, then whole periods of (NSKIP + 24 outputs) in block form
        while (n > 0 && left != 0) { steps_rot<4>(); left -= 4; n--; }
#pragma unroll 1
        while (n >= 6) {
            luxury_skip<NSKIP>();
            block24();
            n -= 6;
        }
        while (n > 0) { skip4<NSKIP>(); n--; }
    }
};

__device__ __forceinline__ bool RanluxPacked::sample(int C, uint4& out) const {
    uint32_t v[4], nx[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int n = 4 * C + e;
        v[e] = limb8(n);
        nx[e] = limb8(n >= 9 ? n - 9 : n + 15);
    }
    bool conv = false;
    if (min(min(v[0], v[1]), min(v[2], v[3])) < (4096u << 8)) {
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = __float_as_uint(Ranlux::to_float(v[e], nx[e]));
        conv = true;
    }
    out = make_uint4(v[0], v[1], v[2], v[3]);
    return conv;
}

// State table in global memory (this backend's own; the reference's float4 x 7 table is never read by the host except through
// PRNG::check_seeds): the generator in PACKED form, at a sample boundary (the next subtract-with-borrow step replaces
// limb 0):  word w (0..19) of instance i at table[w * instances + i];  w 0..17 the 576-bit ring (RanluxPacked), 18 carry, 19 left.
// 20 words instead of the 26 of the unpacked form: 0.4 GB less DRAM traffic per 64^4 pass, and the two-phase kernels need no
// pack / unpack at all.
constexpr int kRanluxStateWords = 20;
__device__ __forceinline__ void ranlux_load(Ranlux& r, const uint32_t* table, uint32_t instances, uint32_t i) {
    RanluxPacked pk;
#pragma unroll
    for (int w = 0; w < 18; w++) pk.w[w] = table[size_t(w) * instances + i];
    pk.unpack(r.s);
    r.carry = table[size_t(18) * instances + i];
    r.left = table[size_t(19) * instances + i];
}
__device__ __forceinline__ void ranlux_store(const Ranlux& r, uint32_t* table, uint32_t instances, uint32_t i) {
    RanluxPacked pk;
    pk.pack(r.s);
#pragma unroll
    for (int w = 0; w < 18; w++) table[size_t(w) * instances + i] = pk.w[w];
    table[size_t(18) * instances + i] = r.carry;
    table[size_t(19) * instances + i] = r.left;
}
// NC RANLUX3 instances per thread in packed form, stepped together (the borrow chains of different instances interleave in the
// instruction stream).  NC = 1: the generator phase of the two-phase update kernels; NC = 3: the generator warps of the
// warp-specialised kernel, which require the same `left` for all instances.
template <int NC> struct RanluxPk {
    RanluxPacked pk[NC];
    uint32_t carry[NC];
    __device__ __forceinline__ uint32_t load(int c, const uint32_t* table, uint32_t instances, uint32_t i) {
#pragma unroll
        for (int w = 0; w < 18; w++) pk[c].w[w] = table[size_t(w) * instances + i];
        carry[c] = table[size_t(18) * instances + i];
        return table[size_t(19) * instances + i];       // left
    }
    __device__ __forceinline__ void store(int c, uint32_t* table, uint32_t instances, uint32_t i, uint32_t left) const {
#pragma unroll
        for (int w = 0; w < 18; w++) table[size_t(w) * instances + i] = pk[c].w[w];
        table[size_t(18) * instances + i] = carry[c];
        table[size_t(19) * instances + i] = left;
    }
    __device__ __forceinline__ void skip223() {
#pragma unroll 1
        for (int b = 0; b < 3; b++) {
#pragma unroll
            for (int c = 0; c < NC; c++) pk[c].rotation72(carry[c]);
        }
#pragma unroll
        for (int c = 0; c < NC; c++) pk[c].tail7(carry[c]);
    }
    // j (1..6) samples from the start of a rotation; sample number k0 + C of the pass is parked at out[c][(k0 + C) * stride] if below keep
    __device__ __forceinline__ void deliver(uint32_t j, uint32_t k0, uint32_t keep, uint4* const (&out)[NC], uint32_t stride,
                                            unsigned long long (&conv)[NC]) {
        if (j == 6 && k0 >= keep) {
#pragma unroll
            for (int c = 0; c < NC; c++) pk[c].rotation24(carry[c]);
            return;
        }
        if (NC == 1 && j == 6) {
            // a whole rotation in one statement (one borrow chain, no parking of the carry between the samples), then the six samples:
            // the "uni < 2^-12" fix-up of number n reads x[n-9], for n <= 8 a limb that leaves the window in this rotation
            uint32_t old[9];
#pragma unroll
            for (int i = 0; i < 9; i++) old[i] = pk[0].limb8(15 + i);
            pk[0].rotation24(carry[0]);
#pragma unroll
            for (int q = 0; q < 6; q++) {
                uint32_t v[4], nx[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int n = 4 * q + e;
                    v[e] = pk[0].limb8(n);
                    nx[e] = (n <= 8) ? old[n] : pk[0].limb8(n - 9);
                }
                if (min(min(v[0], v[1]), min(v[2], v[3])) < (4096u << 8)) {
#pragma unroll
                    for (int e = 0; e < 4; e++) v[e] = __float_as_uint(Ranlux::to_float(v[e], nx[e]));
                    conv[0] |= 1ull << (k0 + q);
                }
                if (k0 + q < keep) out[0][(k0 + q) * stride] = make_uint4(v[0], v[1], v[2], v[3]);
            }
            return;
        }
#pragma unroll
        for (int C = 0; C < 6; C++) {
            if (uint32_t(C) < j) {
#pragma unroll
                for (int c = 0; c < NC; c++) pk[c].chunk(C, carry[c]);
                if (k0 + C < keep) {
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        uint4 v;
                        if (pk[c].sample(C, v)) conv[c] |= 1ull << (k0 + C);
                        out[c][(k0 + C) * stride] = v;
                    }
                }
            }
        }
        // a started rotation: renumber the words so that the next chunk is chunk 0 again (a small loop: once per pass)
        if (j < 6) {
#pragma unroll 1
            for (uint32_t i = 0; i < j; i++) {
#pragma unroll
                for (int c = 0; c < NC; c++) pk[c].template renumber<1>();
            }
        }
    }
    // One pass of generator work: advance by `samples` float4 and park the first `keep` of them, RAW (four seeds << 8): only the ~7 of
    // 33 that a link really consumes get converted, by the consumer.  The exception is a sample containing a number below 2^-12, whose
    // fix-up needs a seed that is gone later: it is converted here and flagged in conv (bit k).  parked() is called once all `keep`
    // samples are in place.  `left` (numbers left before the next luxury skip) is read and updated.
    // jump (NC = 1; may be null): the skip-ahead table of this (samples, keep) -- ranlux_jump_host.h, entry = `left` on entry / 4 --
    // with which the rest of the pass behind the parked samples is ONE multiplication mod 2^576 - 2^240 + 1 (ranlux_jump.cuh)
    // instead of ~1000 subtract-with-borrow steps; same state afterwards, bit for bit.
    template <class F>
    __device__ __forceinline__ void pass(uint32_t& left, uint32_t samples, uint32_t keep, uint4* const (&out)[NC], uint32_t stride,
                                         unsigned long long (&conv)[NC], F parked, const uint32_t* __restrict__ jump = nullptr) {
        uint32_t k = 0;
        bool told = false;
        const uint32_t* entry = (NC == 1 && jump != nullptr && left <= 24u && (left & 3u) == 0u) ? jump + (left >> 2) * 20u : nullptr;
#pragma unroll 1
        while (k < samples) {
            if (left == 0u) { skip223(); left = 24u; }
            const uint32_t j = min(left >> 2, samples - k);
            deliver(j, k, keep, out, stride, conv);
            k += j; left -= 4u * j;
            if (!told && k >= keep) {
                parked(); told = true;
                if (NC == 1 && entry != nullptr && k < samples) {
                    const uint32_t flags = __ldg(entry + 19);           // bit 0: entry usable; the other bits are zero
                    if ((flags & 1u) && !ranlux_is_fixed_point(pk[0].w, carry[0])) {
                        uint32_t A[18];
#pragma unroll
                        for (int w = 0; w < 18; w++) A[w] = __ldg(entry + w);
                        ranlux_jump(pk[0].w, carry[0], A, flags >> 1);
                        left = __ldg(entry + 18);
                        return;
                    }
                }
            }
        }
        if (!told) parked();
    }
};
// Rare path of the two-phase kernel: sample number `index` of the pass, regenerated from the generator state in front of
// the pass (word w at state[w * instances]; the kernel writes the advanced state to the other table).  Not inlined: its
// 26 state registers must not count against the physics phase.
template <int NSKIP>
__device__ __noinline__ float4 ranlux_overflow(const uint32_t* state, uint32_t instances, uint32_t index) {
    Ranlux r;
    ranlux_load(r, state, instances, 0u);
    F4 v = r.template next4<NSKIP>();
#pragma unroll 1
    for (uint32_t i = 0; i < index; i++) v = r.template next4<NSKIP>();
    return make_float4(v.x, v.y, v.z, v.w);
}
// ranlux_init (prngcl_ranlux.cl:97-157): 24 seeds from one 31-bit seed through L'Ecuyer's LCG
__device__ __forceinline__ void ranlux_seed(Ranlux& r, int32_t jseed) {
#pragma unroll 1
    for (int i = 0; i < 24; i++) {
        int32_t k = jseed / 53668;
        jseed = 40014 * (jseed - k * 53668) - k * 12211;
        if (jseed < 0) jseed += 2147483563;
        r.s[i] = uint32_t(jseed % 16777216) << 8;
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

// ---------------------------------------------------------------- XOR7 (Panneton-L'Ecuyer seven-xorshift generator on eight words)
// prngcl_xor7.cl:59-75 (device: the state as two uint4 that shift by one word per step) == random.cpp:975-986 (host: a ring of
// eight words with a moving index).  s[0] is the oldest word, s[7] the one just produced.
struct Xor7 {
    uint32_t s[8];
    __device__ __forceinline__ uint32_t step() {
        uint32_t t, y;
        t = s[7]; t ^= t << 13; y = t ^ (t << 9);
        t = s[4]; y ^= t ^ (t << 7);
        t = s[3]; y ^= t ^ (t >> 3);
        t = s[1]; y ^= t ^ (t >> 10);
        t = s[0]; t ^= t >> 7; y ^= t ^ (t << 24);
#pragma unroll
        for (int k = 0; k < 7; k++) s[k] = s[k + 1];
        s[7] = y;
        return y;
    }
    __device__ __forceinline__ F4 next4() {      // prngcl_xor7.cl:125-134: (float) y / 2^32 (an exact scaling; 1.0f is reachable)
        F4 r;
        r.x = __uint2float_rn(step()) / 4294967296.0f;
        r.y = __uint2float_rn(step()) / 4294967296.0f;
        r.z = __uint2float_rn(step()) / 4294967296.0f;
        r.w = __uint2float_rn(step()) / 4294967296.0f;
        return r;
    }
    __device__ __forceinline__ void skip4() { step(); step(); step(); step(); }
};

// ---------------------------------------------------------------- RANECU (L'Ecuyer's combined LCG, four independent lanes)
// prngcl_ranecu.cl:77-103 (device, unsigned compares) / random.cpp:1049-1068 (host replica of lane x of instance 0, signed
// arithmetic: the two agree unless an intermediate difference is exactly zero, probability 2^-31 per step; the device form is
// what is followed, as for Park-Miller).
struct Ranecu {
    uint32_t a[4], b[4];
    __device__ __forceinline__ F4 next4() {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t q = a[k] / 53668u;
            uint32_t t1 = 40014u * (a[k] - q * 53668u), t2 = q * 12211u;
            a[k] = (t1 > t2) ? (t1 - t2) : (t1 - t2 + 2147483563u);
            q = b[k] / 52774u;
            t1 = 40692u * (b[k] - q * 52774u); t2 = q * 3791u;
            b[k] = (t1 > t2) ? (t1 - t2) : (t1 - t2 + 2147483399u);
            const uint32_t z = (a[k] > b[k]) ? (a[k] - b[k]) : (a[k] - b[k] + 2147483562u);
            o[k] = __uint2float_rn(z) / 2147483648.0f;
        }
        return { o[0], o[1], o[2], o[3] };
    }
    __device__ __forceinline__ void skip4() { (void)next4(); }
};

// ---------------------------------------------------------------- RANMAR (Marsaglia-Zaman-Tsang lagged Fibonacci, four lanes)
// prngcl_ranmar.cl:65-191: 97 float4 seeds per instance in global memory (seedtable[GID + n * instances]), a 98th element
// (I97, J97 as integers, the arithmetic-sequence value c shared by the four lanes, 0).  Every value is a multiple of 2^-24
// in [0, 1), so the float subtractions are exact and the stream equals the host replica's (random.cpp:801-816) bit for bit.
struct RanmarLane4 {
    // the 97 x 24 bits of one lane from its two seeds (ranmar_init / RANMAR_initialize_CPU, random.cpp:773-799)
    static __device__ __forceinline__ void seed(uint32_t seed1, uint32_t seed2, float* out, size_t stride) {
        uint32_t i = ((seed1 / 177u) % 177u) + 2u, j = (seed1 % 177u) + 2u, k = ((seed2 / 169u) % 178u) + 1u, l = seed2 % 169u;
#pragma unroll 1
        for (int n = 0; n < 97; n++) {
            uint32_t s = 0u, t = 8388608u;
#pragma unroll 1
            for (int m = 0; m < 24; m++) {
                uint32_t u = (i * j) % 179u;
                u = (u * k) % 179u;
                i = j; j = k; k = u;
                l = (53u * l + 1u) % 169u;
                if (((l * u) % 64u) >= 32u) s += t;
                t >>= 1;
            }
            out[size_t(n) * stride] = __uint2float_rn(s) / 16777216.0f;
        }
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
