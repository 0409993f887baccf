#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define OPS18(w, c)                                                                              \
    "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]),     \
    "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(c)
#define PRE18  "{\n\t.reg .u32 t, a, h;\n\tsub.cc.u32 t, 0, %18;\n\t"
#define POST18 "subc.u32 t, 0, 0;\n\tand.b32 %18, t, 1;\n\t}"
#define WS(D, LO, HI) "shf.r.wrap.b32 a, %" #LO ", %" #HI ", 16;\n\tsubc.cc.u32 %" #D ", a, %" #D ";\n\t"
#define ROT(A, B, C) \
    A(0, 10, 11) B(1, 11, 12) C(2, 12, 13) A(3, 13, 14) B(4, 14, 15) C(5, 15, 16) A(6, 16, 17) B(7, 17, 0) C(8, 0, 1) \
    A(9, 1, 2) B(10, 2, 3) C(11, 3, 4) A(12, 4, 5) B(13, 5, 6) C(14, 6, 7) A(15, 7, 8) B(16, 8, 9) C(17, 9, 10)
__device__ __forceinline__ void rot(uint32_t (&w)[18], uint32_t& c) { asm volatile(PRE18 ROT(WS, WS, WS) POST18 : OPS18(w, c)); }

template <int NC> __global__ void __launch_bounds__(128, 1) k_chains(uint32_t* st, int R) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    uint32_t w[NC][18], c[NC];
#pragma unroll
    for (int q = 0; q < NC; q++) {
#pragma unroll
        for (int k = 0; k < 18; k++) w[q][k] = st[(q * 19 + k) * n + i];
        c[q] = st[(q * 19 + 18) * n + i];
    }
#pragma unroll 1
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int q = 0; q < NC; q++) rot(w[q], c[q]);
    }
#pragma unroll
    for (int q = 0; q < NC; q++) {
#pragma unroll
        for (int k = 0; k < 18; k++) st[(q * 19 + k) * n + i] = w[q][k];
        st[(q * 19 + 18) * n + i] = c[q];
    }
}
int main(int argc, char** argv) {
    int R = argc > 1 ? atoi(argv[1]) : 4000;
    const int blocks = 148, threads = 128, n = blocks * threads;
    uint32_t* d; cudaMalloc(&d, 19 * 4 * (size_t)n * 4); cudaMemset(d, 0x5a, 19 * 4 * (size_t)n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int nc = 1; nc <= 4; nc++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (nc == 1) k_chains<1><<<blocks, threads>>>(d, R);
            if (nc == 2) k_chains<2><<<blocks, threads>>>(d, R);
            if (nc == 3) k_chains<3><<<blocks, threads>>>(d, R);
            if (nc == 4) k_chains<4><<<blocks, threads>>>(d, R);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double steps = 24.0 * R * nc, cyc = best * 1e-3 * 1.965e9;
        printf("chains/thread=%d, 1 warp per scheduler: %8.3f ms  %6.2f clk per chain-step  err=%s\n", nc, best, cyc / steps, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
