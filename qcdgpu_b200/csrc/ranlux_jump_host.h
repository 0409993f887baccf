// ranlux_jump_host.h -- host side of the RANLUX skip-ahead (ranlux_jump.cuh): the multipliers b^-(steps - 48) mod m,
// b = 2^24, m = b^24 - b^10 + 1, and the plan of a generator pass (which part is stepped, which part is jumped over).
//
// The subtract-with-borrow recurrence x[n] = x[n-10] - x[n-24] - borrow (random.cpp:504-574) multiplies
// X = N - (N >> 336) + borrow (N = the 24 seeds as one number, oldest in the low limb) by b^-1 mod m per step.
#pragma once
#include <cstdint>
#include <cstring>

namespace qcd {

// numbers mod m as 24 limbs of 24 bits
struct Mod576 {
    uint64_t l[24];

    // t: 48 signed limbs (weights b^k) -> canonical residue
    static Mod576 reduce(int64_t (&t)[48]) {
        for (int k = 47; k >= 24; k--) {        // b^k = b^(k-24) (b^10 - 1)
            const int64_t v = t[k];
            t[k] = 0; t[k - 14] += v; t[k - 24] -= v;
        }
        for (;;) {
            int64_t carry = 0;
            for (int k = 0; k < 24; k++) {
                const int64_t v = t[k] + carry;
                carry = v >> 24;                // floor division (arithmetic shift)
                t[k] = v & 0xFFFFFF;
            }
            if (carry == 0) break;
            t[10] += carry; t[0] -= carry;      // carry b^24 = carry (b^10 - 1)
        }
        // 0 <= value < b^24; subtract m once if value >= m = b^24 - b^10 + 1: limbs 10..23 all ones and the low ten limbs >= 1
        bool top_ones = true, low_zero = true;
        for (int k = 10; k < 24; k++) top_ones = top_ones && t[k] == 0xFFFFFF;
        for (int k = 0; k < 10; k++) low_zero = low_zero && t[k] == 0;
        Mod576 r;
        if (top_ones && !low_zero) {            // value - m = low ten limbs - 1
            int64_t borrow = 1;
            for (int k = 0; k < 10; k++) { int64_t v = t[k] - borrow; borrow = v < 0; t[k] = v & 0xFFFFFF; }
            for (int k = 10; k < 24; k++) t[k] = 0;
        }
        for (int k = 0; k < 24; k++) r.l[k] = (uint64_t)t[k];
        return r;
    }
    static Mod576 mul(const Mod576& a, const Mod576& b) {
        int64_t t[48];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < 24; i++)
            for (int j = 0; j < 24; j++) t[i + j] += (int64_t)(a.l[i] * b.l[j]);     // 24 products below 2^48 per column
        return reduce(t);
    }
    static Mod576 one() { Mod576 r; memset(&r, 0, sizeof(r)); r.l[0] = 1; return r; }
    // b^-1 = m - (m - 1) / b = b^24 - b^23 - b^10 + b^9 + 1
    static Mod576 b_inverse() {
        int64_t t[48];
        memset(t, 0, sizeof(t));
        t[24] = 1; t[23] = -1; t[10] = -1; t[9] = 1; t[0] = 1;
        return reduce(t);
    }
    static Mod576 b_inverse_pow(uint64_t e) {
        Mod576 r = one(), base = b_inverse();
        while (e) {
            if (e & 1u) r = mul(r, base);
            base = mul(base, base);
            e >>= 1;
        }
        return r;
    }
    void to_words(uint32_t (&out)[18]) const {
        memset(out, 0, sizeof(out));
        for (int k = 0; k < 24; k++) {
            const int bit = 24 * k, wd = bit >> 5, sh = bit & 31;
            out[wd] |= (uint32_t)l[k] << sh;
            if (sh > 8) out[wd + 1] |= (uint32_t)l[k] >> (32 - sh);
        }
    }
};

// 18 words of b^-(steps - 48) mod m; false if steps < 48 (too short to jump)
inline bool ranlux_jump_multiplier(uint64_t steps, uint32_t (&out)[18]) {
    if (steps < 48u) return false;
    Mod576::b_inverse_pow(steps - 48u).to_words(out);
    return true;
}

// One generator pass of `samples` float4 (RanluxPk::pass, prng.cuh), replayed on the host for a generator that enters it with `left`
// numbers to go before its next luxury skip: the kernel steps until `keep` samples are parked and jumps over the rest.
// jump_steps = subtract-with-borrow steps of that rest, left_after = `left` at the end of the pass.
// from_start: nothing is parked or stepped, the whole pass is jumped over (ranlux_skip_kernel: fast-forward of a resumed run).
struct RanluxPassPlan { uint64_t jump_steps; uint32_t left_after; };
inline RanluxPassPlan ranlux_pass_plan(uint32_t left, uint64_t samples, uint64_t keep, uint32_t skip, bool from_start = false) {
    uint64_t k = 0, steps = 0;
    bool told = from_start;
    while (k < samples) {
        if (left == 0u) { if (told) steps += skip; left = 24u; }
        uint64_t j = left >> 2;
        if (j > samples - k) j = samples - k;
        if (told && k + 6 <= samples && left == 24u) {      // whole periods in one go (fast-forward over many passes)
            const uint64_t periods = (samples - k) / 6;
            steps += periods * 24u + (periods - 1) * skip;
            k += periods * 6; left = 0u;
            continue;
        }
        if (told) steps += 4u * j;
        k += j; left -= 4u * (uint32_t)j;
        if (!told && k >= keep) told = true;
    }
    return { steps, left };
}

// Table the kernels read: entry e = left / 4 (0..6) of 20 words: 18 multiplier words, `left` after the pass, 1 if the entry is usable
constexpr int kRanluxJumpEntryWords = 20;
inline void ranlux_jump_table(uint64_t samples, uint64_t keep, uint32_t skip, uint32_t (&table)[7 * kRanluxJumpEntryWords],
                              bool from_start = false) {
    memset(table, 0, sizeof(table));
    for (uint32_t e = 0; e < 7u; e++) {
        const RanluxPassPlan plan = ranlux_pass_plan(4u * e, samples, keep, skip, from_start);
        uint32_t a[18];
        if (!ranlux_jump_multiplier(plan.jump_steps, a)) continue;
        for (int w = 0; w < 18; w++) table[e * kRanluxJumpEntryWords + w] = a[w];
        table[e * kRanluxJumpEntryWords + 18] = plan.left_after;
        table[e * kRanluxJumpEntryWords + 19] = 1u;
    }
}

}  // namespace qcd
