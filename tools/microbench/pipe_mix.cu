// Microbenchmark: do the FP64 pipe and the integer (ALU) pipe of an sm_100a scheduler overlap when DIFFERENT warps feed them?
// Block of 512 threads = 16 warps, one block per SM: warps 12-15 ("generator", one per scheduler) run three interleaved RANLUX
// subtract-with-borrow chains each (ALU pipe only), warps 0-11 ("physics", three per scheduler) run chains of DFMA with ILP I.
// Times: generator warps alone, physics warps alone, both together.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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

template <int ILP>
__global__ void __launch_bounds__(512, 1) k_mix(uint32_t* st, double* dd, int Rgen, int Rphys, int mode, long long* clk) {
    const uint32_t warp = threadIdx.x >> 5, i = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    long long t0 = clock64();
    if (warp >= 12) {
        if (mode == 1) return;      // physics only
        uint32_t w[3][18], c[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
#pragma unroll
            for (int k = 0; k < 18; k++) w[q][k] = st[(q * 19 + k) * n + i];
            c[q] = st[(q * 19 + 18) * n + i] & 1;
        }
#pragma unroll 1
        for (int r = 0; r < Rgen; r++) {
#pragma unroll
            for (int q = 0; q < 3; q++) rot(w[q], c[q]);
        }
#pragma unroll
        for (int q = 0; q < 3; q++) {
#pragma unroll
            for (int k = 0; k < 18; k++) st[(q * 19 + k) * n + i] = w[q][k];
            st[(q * 19 + 18) * n + i] = c[q];
        }
        if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 16 + warp] = clock64() - t0;
    } else {
        if (mode == 2) return;      // generator only
        double a[ILP], x = dd[i], y = dd[n + i];
#pragma unroll
        for (int k = 0; k < ILP; k++) a[k] = dd[(2 + k) * n + i];
#pragma unroll 1
        for (int r = 0; r < Rphys; r++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int k = 0; k < ILP; k++) a[k] = fma(a[k], x, y);
            }
        }
#pragma unroll
        for (int k = 0; k < ILP; k++) dd[(2 + k) * n + i] = a[k];
        if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 16 + warp] = clock64() - t0;
    }
}
template <int ILP> void run(uint32_t* d, double* dd, long long* clk, int Rgen, int Rphys) {
    long long h[148 * 16];
    for (int mode = 0; mode < 3; mode++) {
        cudaMemset(clk, 0, sizeof(h));
        k_mix<ILP><<<148, 512>>>(d, dd, Rgen, Rphys, mode, clk);
        cudaDeviceSynchronize();
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double g = 0, p = 0;
        for (int b = 0; b < 148; b++) { for (int w = 12; w < 16; w++) g += h[b * 16 + w]; for (int w = 0; w < 12; w++) p += h[b * 16 + w]; }
        g /= 148 * 4; p /= 148 * 12;
        // generator: ALU instructions per warp = Rgen * 3 chains * 36 (+ 9 bookkeeping); physics: DFMA per warp = Rphys * 8 * ILP
        printf("ILP %d %-14s gen warp %9.0f clk (%.2f clk per ALU instr)   physics warp %9.0f clk (%.2f clk per DFMA, 3 warps per scheduler: pipe busy %.0f %%)  %s\n", ILP,
               mode == 0 ? "both" : mode == 1 ? "physics only" : "generator only", g, g / (Rgen * 3.0 * 39), p, p / (Rphys * 8.0 * ILP),
               p > 0 ? 100.0 * 3 * 2 * Rphys * 8.0 * ILP / p : 0.0, cudaGetErrorString(cudaGetLastError()));
    }
}
int main(int argc, char** argv) {
    int Rgen = argc > 1 ? atoi(argv[1]) : 2000, Rphys = argc > 2 ? atoi(argv[2]) : 4000;
    const int n = 148 * 512;
    uint32_t* d; cudaMalloc(&d, 19 * 3 * (size_t)n * 4); cudaMemset(d, 0x5a, 19 * 3 * (size_t)n * 4);
    double* dd; cudaMalloc(&dd, 12 * (size_t)n * 8); cudaMemset(dd, 0, 12 * (size_t)n * 8);
    long long* clk; cudaMalloc(&clk, 148 * 16 * 8);
    run<1>(d, dd, clk, Rgen, Rphys);
    run<2>(d, dd, clk, Rgen, Rphys / 2);
    run<4>(d, dd, clk, Rgen, Rphys / 4);
    run<8>(d, dd, clk, Rgen, Rphys / 8);
    return 0;
}
