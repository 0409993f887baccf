// Microbenchmark + check: RANLUX3 skip-ahead by one modular multiplication (csrc/ranlux_jump.cuh) against the stepped generator.
//   stepped: RanluxPk<1>::pass over the tail of an update pass (21 float4 samples nobody consumes = 976 steps with the luxury skips)
//   jumped:  X = N - (N >> 336) + borrow, X * A mod m, ring = -X1 (1 + 2^240 + 2^480)
// Both start from the same random states (incl. degenerate ones); the final states must agree bit for bit for every entry
// `left` = 0, 4, .., 24 of the jump table.  Then both are timed at 3 blocks of 128 threads per SM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../qcdgpu_b200/csrc/prng.cuh"
#include "../../qcdgpu_b200/csrc/ranlux_jump.cuh"
#include "../../qcdgpu_b200/csrc/ranlux_jump_host.h"

using namespace qcd;

__global__ void __launch_bounds__(128, 3) k_stepped(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t samples) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RanluxPk<1> r;
    uint32_t left = r.load(0, in, n, i);
    unsigned long long conv[1] = { 0ull };
    uint4* const o[1] = { nullptr };
    r.pass(left, samples, 0u, o, 0u, conv, [] {});
    r.store(0, out, n, i, left);
}
__global__ void __launch_bounds__(128, 3) k_jumped(const uint32_t* in, uint32_t* out, uint32_t n, const uint32_t* __restrict__ table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RanluxPk<1> r;
    uint32_t left = r.load(0, in, n, i);
    const uint32_t* e = table + (left >> 2) * kRanluxJumpEntryWords;
    if (!ranlux_is_fixed_point(r.pk[0].w, r.carry[0])) {
        uint32_t A[18];
#pragma unroll
        for (int w = 0; w < 18; w++) A[w] = __ldg(e + w);
        ranlux_jump(r.pk[0].w, r.carry[0], A, __ldg(e + 19) >> 1);
    }
    left = __ldg(e + 18);
    r.store(0, out, n, i, left);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const uint32_t samples = argc > 1 ? atoi(argv[1]) : 21u;
    const uint32_t n = 148u * 3u * 128u * 16u;
    std::vector<uint32_t> h(size_t(20) * n);
    uint32_t *din, *dout1, *dout2, *dtab;
    CK(cudaMalloc(&din, h.size() * 4)); CK(cudaMalloc(&dout1, h.size() * 4)); CK(cudaMalloc(&dout2, h.size() * 4));
    CK(cudaMalloc(&dtab, 7 * kRanluxJumpEntryWords * 4));
    std::vector<uint32_t> o1(h.size()), o2(h.size());
    srand(12345);
    auto rnd = [] { return (uint32_t(rand()) << 16) ^ uint32_t(rand()) ^ (uint32_t(rand()) << 31); };
    int bad_total = 0;
    for (int32_t sl = 24; sl >= 0; sl -= 4) { const uint32_t left = (uint32_t)sl;
        // the pass as the kernel sees it: keep = 0 parks nothing, but the first batch of samples is still stepped (RanluxPk::pass), so
        // the stepped reference runs `samples` samples with keep = 0 from a state whose first batch has been delivered already:
        // use the plan of (left, samples, keep = 0) minus its stepped head by starting the jump table at the same point
        for (uint32_t i = 0; i < n; i++) {
            const int kind = i % 8;
            for (int w = 0; w < 18; w++) {
                uint32_t v = rnd();
                if (kind == 1) v = (rnd() & 1) ? 0xFFFFFFFFu : 0u;
                if (kind == 2) v = 0u;
                if (kind == 3) v = 0xFFFFFFFFu;
                if (kind == 4) v = (w == int(i / 8 % 18)) ? (1u << (i / 144 % 32)) : 0u;
                if (kind == 5) v = (w == int(i / 8 % 18)) ? ~(1u << (i / 144 % 32)) : 0xFFFFFFFFu;
                h[size_t(w) * n + i] = v;
            }
            h[size_t(18) * n + i] = rnd() & 1u;
            h[size_t(19) * n + i] = left;
        }
        // table entry for THIS comparison: all of the pass is jumped (the stepped kernel steps all of it)
        uint32_t table[7 * kRanluxJumpEntryWords] = {};
        {
            uint32_t k = 0, l = left, steps = 0;
            while (k < samples) { if (l == 0u) { steps += 223u; l = 24u; } uint32_t j = (l >> 2) < (samples - k) ? (l >> 2) : (samples - k); steps += 4u * j; k += j; l -= 4u * j; }
            uint32_t a[18];
            if (!ranlux_jump_multiplier(steps, a)) { printf("left %u: %u steps, too short to jump\n", left, steps); continue; }
            uint32_t* e = table + (left >> 2) * kRanluxJumpEntryWords;
            for (int w = 0; w < 18; w++) e[w] = a[w];
            e[18] = l; e[19] = 1u;
            printf("left %2u: %u steps, left after %u: ", left, steps, l);
        }
        CK(cudaMemcpy(din, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dtab, table, sizeof(table), cudaMemcpyHostToDevice));
        k_stepped<<<n / 128, 128>>>(din, dout1, n, samples);
        k_jumped<<<n / 128, 128>>>(din, dout2, n, dtab);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(o1.data(), dout1, h.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(o2.data(), dout2, h.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (uint32_t i = 0; i < n; i++) {
            bool same = true;
            for (int w = 0; w < 20; w++) same = same && o1[size_t(w) * n + i] == o2[size_t(w) * n + i];
            if (!same && bad++ < 3) printf("\n  instance %u (kind %u) differs", i, i % 8);
        }
        printf("%s (%d of %u differ)\n", bad ? "MISMATCH" : "identical", bad, n);
        bad_total += bad;
    }
    // timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms1 = 0, ms2 = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); for (int r = 0; r < 10; r++) k_stepped<<<n / 128, 128>>>(din, dout1, n, samples); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms1, e0, e1);
        cudaEventRecord(e0); for (int r = 0; r < 10; r++) k_jumped<<<n / 128, 128>>>(din, dout2, n, dtab); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms2, e0, e1);
    }
    printf("%u instances, %u samples: stepped %.3f ms, jumped %.3f ms per launch (%.2fx)\n", n, samples, ms1 / 10, ms2 / 10, ms1 / ms2);
    return bad_total ? 1 : 0;
}
