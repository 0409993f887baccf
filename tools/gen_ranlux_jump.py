#!/usr/bin/env python3
"""Generator (and CPU model) of qcdgpu_b200/csrc/ranlux_jump.cuh: RANLUX skip-ahead by ONE modular multiplication.

RANLUX's subtract-with-borrow recurrence x[n] = x[n-10] - x[n-24] - borrow (mod b = 2^24; prngcl_ranlux.cl:97-563, host replica
random.cpp:504-574) is the linear congruential generator X' = X / b (mod m), m = b^24 - b^10 + 1, in disguise (Luescher 1994;
Tezuka, L'Ecuyer, Couture 1993).  With the 24 seeds as one 576-bit number N (oldest seed in the low limb: exactly the packed ring of
prng.cuh) and the borrow c:

    X = N - (N >> 336) + c  (mod m)        one step:  X' * b = X (mod m)

so R steps are a multiplication by b^-R mod m, and the 24 numbers that FOLLOW a state are the limbs of
D = -X * (1 + b^10 + b^20) mod b^24  (1 + b^10 + b^20 = m^-1 mod b^24).  The part of a pass whose samples nobody can consume
(21 of 33 float4 = 976 steps with RANLUX3's luxury skips in the SU(3) 64^4 run) so costs ~650 instructions instead of ~1500:

    X  = N - (N >> 336) + c                                        (kept as X - 1: the +1 is the addend A of the product)
    P  = X * A,  A = b^-(R-48) mod m from the host                 18 x 18 words, IMAD.WIDE.U32 with carry chains
    X1 = P / b^24 mod m  (Montgomery step, canonical)              = X at step R - 24
    D  = -X1 (1 + b^10 + b^20) mod b^24                            = the ring at step R;  borrow from the top words of X1 + D - (D << 240)

The two fixed points (all zero / borrow 0, all ones / borrow 1) both have X = 0 and are left alone.

This script holds the algorithm as a list of 32-bit operations with PTX carry-flag semantics, RUNS it on random and edge states
against the plain recurrence (python integers), and prints the same list as CUDA.  tests/test_abi_and_host.py re-runs the model
and checks that the committed header is what the script prints.
"""
from __future__ import annotations

import random
import sys

B = 1 << 24
LAG = 3            # multiplication carry chains in flight (see build)
M = B**24 - B**10 + 1
MASK32 = 0xFFFFFFFF
W576 = (1 << 576) - 1


# ----------------------------------------------------------------------------- the plain recurrence (the definition)
def swb_step(x, c):
    new = x[14] - x[0] - c
    if new < 0:
        return x[1:] + [new + B], 1
    return x[1:] + [new], 0


def ring_number(x):
    return sum(v << (24 * k) for k, v in enumerate(x))


def state_to_x(x, c):
    n = ring_number(x)
    return n - (n >> 336) + c


def jump_multiplier(steps):
    """A = b^-(steps-48) mod m, the table entry of a jump over `steps` subtract-with-borrow steps (steps >= 48)."""
    assert steps >= 48
    return pow(B, -(steps - 48), M)


# ----------------------------------------------------------------------------- program = list of operations
class Prog:
    def __init__(self):
        self.ops = []
        self.vars = []
        self.n = 0

    def v(self, stem):
        self.n += 1
        name = f"{stem}_{self.n}"
        self.vars.append(name)
        return name

    def emit(self, op, d, *src):
        self.ops.append((op, d, src))
        return d

    # plain operations (C++ expressions)
    def shf(self, lo, hi, n):           # funnel shift right of hi:lo by n
        return self.emit("shf", self.v("t"), lo, hi, n)

    def shl(self, a, n):
        return self.emit("shl", self.v("t"), a, n)

    def shr(self, a, n):
        return self.emit("shr", self.v("t"), a, n)

    def top(self, a, b, s):             # ((a + b + 1 - s) >> 32) & 1 in 64 bits
        return self.emit("top", self.v("t"), a, b, s)

    # carry-flag operations (one PTX instruction each)
    def cc(self, op, a, b, c=None, stem="r"):
        d = self.v(stem)
        if c is None:
            return self.emit(op, d, a, b)
        return self.emit(op, d, a, b, c)


def words(stem, n):
    return [f"{stem}[{i}]" for i in range(n)]


def chain_add(p, a, b, last_cc=False):
    """multi-word a + b (lists of operands, ints = immediates); returns result words; the carry out stays in CF if last_cc"""
    out = []
    for i, (x, y) in enumerate(zip(a, b)):
        last = i == len(a) - 1
        if i == 0:
            op = "add.cc"
        else:
            op = "addc.cc" if (not last or last_cc) else "addc"
        out.append(p.cc(op, x, y))
    return out


def chain_sub(p, a, b, first="sub.cc", last_cc=False):
    out = []
    for i, (x, y) in enumerate(zip(a, b)):
        last = i == len(a) - 1
        if i == 0:
            op = first
        else:
            op = "subc.cc" if (not last or last_cc) else "subc"
        out.append(p.cc(op, x, y))
    return out


def shifted_240(p, x, nwords):
    """words 7 .. 7+nwords-1 of (x << 240): word 7 = x[0] << 16, word 7+k = x[k-1]:x[k] >> 16"""
    out = [p.shl(x[0], 16)]
    for k in range(1, nwords):
        out.append(p.shf(x[k - 1], x[k], 16))
    return out


def top_240(p, x):
    """the eight words of x >> 336 (x: 18 words)"""
    return [p.shf(x[10 + k], x[11 + k], 16) for k in range(7)] + [p.shr(x[17], 16)]


def neg_times_minv(p, x):
    """-x (1 + 2^240 + 2^480) mod 2^576"""
    s = shifted_240(p, x, 11)
    y = list(x[:7]) + chain_add(p, x[7:18], s)
    y = y[:15] + chain_add(p, y[15:18], x[0:3])
    return chain_sub(p, [0] * 18, y)


def build() -> Prog:
    p = Prog()
    w = words("w", 18)
    A = words("A", 18)
    # X - 1 = N - (N >> 336) - (1 - c)
    g = top_240(p, w)
    nb = p.emit("xor1", p.v("t"), "carry")
    p.cc("sub.cc", 0, nb, stem="junk")                               # CF = 1 - c
    x = chain_sub(p, w, g + [0] * 10, first="subc.cc")
    # P = (X - 1) * A + A: products a[j] * b_i go to E when i + j is even and to O when it is odd, so that every 64-bit product lands
    # on a word pair that is aligned in its array and the carry chain of a row runs through adjacent pairs
    #
    # ptxas would interleave all 36 chains (one live carry predicate each, seven predicate registers: measured 268 LOP3 / P2R of
    # predicate spilling); chain n therefore takes its multiplier word through `A[i] | (top word of chain n - LAG) & zero`, with
    # `zero` a zero the compilers cannot see through (the caller loads it), which keeps LAG chains in flight
    E = [None] * 37
    O = [None] * 37
    tops = []
    for i in range(18):
        for acc, parity in ((E, 0), (O, 1)):
            js = [j for j in range(18) if (i + j) % 2 == parity]
            first = True
            bi = A[i]
            if len(tops) >= LAG:
                bi = p.emit("tie", p.v("b"), A[i], tops[-LAG], "zero")
            for j in js:
                pos = i + j
                lo = acc[pos] if acc[pos] is not None else (A[pos] if (acc is E and i == 0) else 0)
                hi = acc[pos + 1] if acc[pos + 1] is not None else (A[pos + 1] if (acc is E and i == 0) else 0)
                acc[pos] = p.cc("mad.lo.cc" if first else "madc.lo.cc", x[j], bi, lo, stem="e" if acc is E else "o")
                acc[pos + 1] = p.cc("madc.hi.cc", x[j], bi, hi, stem="e" if acc is E else "o")
                first = False
            topw = i + js[-1] + 2
            acc[topw] = p.cc("addc", acc[topw] if acc[topw] is not None else 0, 0, stem="e" if acc is E else "o")
            tops.append(acc[topw])
    assert E[36] is None or True
    P = [E[0]] + chain_add(p, [E[k] if E[k] is not None else 0 for k in range(1, 36)],
                           [O[k] if O[k] is not None else 0 for k in range(1, 36)])
    Plo, Phi = P[:18], P[18:]
    # Montgomery step: Q = -Plo / m mod 2^576, X1 = (P + Q m) / 2^576 = Phi + Q - (Q >> 336) + carry of Plo + Q - (Q << 240)
    Q = neg_times_minv(p, Plo)
    cy = p.top(Plo[17], Q[17], p.shf(Q[9], Q[10], 16))
    p.cc("add.cc", cy, MASK32, stem="junk")                         # CF = cy
    r = []
    for k in range(18):
        r.append(p.cc("addc.cc", Phi[k], Q[k]))
    ov = p.cc("addc", 0, 0)
    g2 = top_240(p, Q)
    r = chain_sub(p, r, g2 + [0] * 10, last_cc=True)
    ov = p.cc("subc", ov, 0)
    # canonical: X1 >= m  <=>  X1 + 2^240 - 1 >= 2^576
    z = chain_add(p, r, [MASK32] * 7 + [0xFFFF] + [0] * 10, last_cc=True)
    ov2 = p.cc("addc", ov, 0)
    x1 = [p.emit("sel", p.v("x"), ov2, z[k], r[k]) for k in range(18)]
    # the ring at the target and its borrow
    D = neg_times_minv(p, x1)
    c = p.top(x1[17], D[17], p.shf(D[9], D[10], 16))
    for k in range(18):
        p.emit("mov", f"w[{k}]", D[k])
    p.emit("mov", "carry", c)
    return p


# ----------------------------------------------------------------------------- CPU model of the operation list
def run(p: Prog, w, carry, A):
    env = {f"w[{i}]": w[i] for i in range(18)}
    env.update({f"A[{i}]": A[i] for i in range(18)})
    env["carry"] = carry
    env["zero"] = 0
    cf = 0

    def val(s):
        return s if isinstance(s, int) else env[s]

    for op, d, src in p.ops:
        a = [val(s) for s in src]
        if op == "shf":
            r = (((a[1] << 32) | a[0]) >> a[2]) & MASK32
        elif op == "shl":
            r = (a[0] << a[1]) & MASK32
        elif op == "shr":
            r = a[0] >> a[1]
        elif op == "xor1":
            r = a[0] ^ 1
        elif op == "tie":
            r = a[0] | (a[1] & a[2])
        elif op == "top":
            r = ((a[0] + a[1] + 1 - a[2]) >> 32) & 1
        elif op == "sel":
            r = a[1] if a[0] else a[2]
        elif op == "mov":
            r = a[0]
        elif op in ("add.cc", "addc.cc", "addc"):
            t = a[0] + a[1] + (cf if op != "add.cc" else 0)
            r = t & MASK32
            if op != "addc":
                cf = t >> 32
        elif op in ("sub.cc", "subc.cc", "subc"):
            t = a[0] - a[1] - (cf if op != "sub.cc" else 0)
            r = t & MASK32
            if op != "subc":
                cf = 1 if t < 0 else 0
        elif op in ("mad.lo.cc", "madc.lo.cc", "madc.hi.cc"):
            prod = a[0] * a[1]
            part = (prod >> 32) if "hi" in op else (prod & MASK32)
            t = part + a[2] + (cf if op != "mad.lo.cc" else 0)
            r = t & MASK32
            cf = t >> 32
        else:
            raise ValueError(op)
        assert 0 <= r <= MASK32
        env[d] = r
    return [env[f"w[{i}]"] for i in range(18)], env["carry"]


def is_fixed_point(x, c):
    return (all(v == 0 for v in x) and c == 0) or (all(v == B - 1 for v in x) and c == 1)


def to_words(n, k=18):
    return [(n >> (32 * i)) & MASK32 for i in range(k)]


def self_test(p: Prog, trials=400, seed=7):
    rnd = random.Random(seed)
    for t in range(trials):
        kind = t % 6
        if kind == 0 or kind == 5:
            x = [rnd.randrange(B) for _ in range(24)]
        elif kind == 1:
            x = [rnd.choice([0, B - 1, 1, B - 2]) for _ in range(24)]
        elif kind == 2:
            x = [0] * 24
            x[rnd.randrange(24)] = rnd.choice([1, B - 1])
        elif kind == 3:
            x = [B - 1] * 24
            x[rnd.randrange(24)] = rnd.choice([0, B - 2, B - 1])
        else:
            x = [rnd.randrange(B) if rnd.random() < 0.5 else 0 for _ in range(24)]
        c = rnd.randrange(2)
        if is_fixed_point(x, c):
            continue
        steps = rnd.choice([48, 49, 84, 307, 976, 1223, 1358])
        xs, cs = x, c
        for _ in range(steps):
            xs, cs = swb_step(xs, cs)
        got_w, got_c = run(p, to_words(ring_number(x)), c, to_words(jump_multiplier(steps)))
        assert got_w == to_words(ring_number(xs)) and got_c == cs, (t, kind, steps)
    return trials


# ----------------------------------------------------------------------------- CUDA
HEADER = '''// ranlux_jump.cuh -- GENERATED by tools/gen_ranlux_jump.py (which also runs this operation list on the CPU against the plain
// recurrence); do not edit.  RANLUX skip-ahead: the subtract-with-borrow recurrence (prngcl_ranlux.cl:97-563, random.cpp:504-574) is
// X' = X / 2^24 mod m, m = 2^576 - 2^240 + 1, on X = N - (N >> 336) + borrow (N = the packed ring of prng.cuh), so a fixed number of
// steps is one multiplication by a constant mod m; the ring after the jump is -X1 (1 + 2^240 + 2^480) mod 2^576 with X1 the state
// 24 steps before the target.  w: the packed ring (RanluxPacked::w), carry: its borrow, A: 2^-24(steps - 48) mod m (qgpu backend:
// ranlux_jump_table), zero: a zero
// loaded from memory (it only ties instructions together for the scheduler).  The caller must not pass one of the two fixed points
// (ranlux_is_fixed_point).
#pragma once
#include <cstdint>
namespace qcd {
#define QCD_J2(OP, D, X, Y) asm volatile(OP ".u32 %0, %1, %2;" : "=r"(D) : "r"(X), "r"(Y))
#define QCD_J3(OP, D, X, Y, Z) asm volatile(OP ".u32 %0, %1, %2, %3;" : "=r"(D) : "r"(X), "r"(Y), "r"(Z))
// all 18 words zero with borrow 0, or all ones with borrow 1
__device__ __forceinline__ bool ranlux_is_fixed_point(const uint32_t (&w)[18], uint32_t carry) {
    const uint32_t m = 0u - carry;
    uint32_t acc = 0u;
#pragma unroll
    for (int k = 0; k < 18; k += 2) acc |= (w[k] ^ m) | (w[k + 1] ^ m);
    return acc == 0u;
}
__device__ __forceinline__ void ranlux_jump(uint32_t (&w)[18], uint32_t& carry, const uint32_t (&A)[18], const uint32_t zero) {
'''


def operand(s):
    if isinstance(s, int):
        return f"0x{s:X}u" if s > 9 else f"{s}u"
    return s


def to_cuda(p: Prog) -> str:
    out = [HEADER]
    names = [v for v in p.vars]
    for i in range(0, len(names), 12):
        out.append("    uint32_t " + ", ".join(names[i:i + 12]) + ";\n")
    for op, d, src in p.ops:
        s = [operand(x) for x in src]
        if op == "shf":
            out.append(f"    {d} = __funnelshift_r({s[0]}, {s[1]}, {src[2]});\n")
        elif op == "shl":
            out.append(f"    {d} = {s[0]} << {src[1]};\n")
        elif op == "shr":
            out.append(f"    {d} = {s[0]} >> {src[1]};\n")
        elif op == "xor1":
            out.append(f"    {d} = {s[0]} ^ 1u;\n")
        elif op == "tie":
            out.append(f"    {d} = {s[0]} | ({s[1]} & {s[2]});\n")
        elif op == "top":
            out.append(f"    {d} = uint32_t(((unsigned long long){s[0]} + {s[1]} + 1ull - {s[2]}) >> 32) & 1u;\n")
        elif op == "sel":
            out.append(f"    {d} = {s[0]} ? {s[1]} : {s[2]};\n")
        elif op == "mov":
            out.append(f"    {d} = {s[0]};\n")
        elif len(src) == 2:
            out.append(f'    QCD_J2("{op}", {d}, {s[0]}, {s[1]});\n')
        else:
            out.append(f'    QCD_J3("{op}", {d}, {s[0]}, {s[1]}, {s[2]});\n')
    out.append("}\n#undef QCD_J2\n#undef QCD_J3\n}  // namespace qcd\n")
    return "".join(out)


def main():
    p = build()
    n = self_test(p)
    carry_ops = sum(1 for op, _, _ in p.ops if "." in op or op in ("addc", "subc"))
    sys.stderr.write(f"ranlux_jump: {len(p.ops)} operations ({carry_ops} with the carry flag), model checked on {n} states\n")
    text = to_cuda(p)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
