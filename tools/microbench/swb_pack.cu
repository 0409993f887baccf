// Microbenchmark: subtract-with-borrow (RANLUX) step cost on sm_100a for three state layouts.
//   V0  unpacked: 24 words (x << 8), per step subc.cc + and            (2 alu ops / step)   -- prng.cuh today
//   V1  packed: 18 words = 24 x 24 bit, per word funnel shift (shf) + subc.cc   (1.5 alu ops / step)
//   V2  packed, alignment through mul.hi + mad.lo (fma pipe)            (0.75 alu + 1.5 fma ops / step)
//   V3  packed, one word of three aligned by shf, two by mul.hi/mad.lo
// Every variant advances the same generator by R x 24 steps; the final states must agree.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define SWB(I, J) "subc.cc.u32 %" #I ", %" #J ", %" #I ";\n\tand.b32 %" #I ", %" #I ", 0xFFFFFF00;\n\t"
#define OPS24(s, c)                                                                              \
    "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]), "+r"(s[6]), "+r"(s[7]),         \
    "+r"(s[8]), "+r"(s[9]), "+r"(s[10]), "+r"(s[11]), "+r"(s[12]), "+r"(s[13]), "+r"(s[14]), "+r"(s[15]),   \
    "+r"(s[16]), "+r"(s[17]), "+r"(s[18]), "+r"(s[19]), "+r"(s[20]), "+r"(s[21]), "+r"(s[22]), "+r"(s[23]), \
    "+r"(c)
#define PRE24  "{\n\t.reg .u32 t;\n\tsub.cc.u32 t, 0, %24;\n\t"
#define POST24 "subc.u32 t, 0, 0;\n\tand.b32 %24, t, 1;\n\t}"
#define SWB_24 SWB(23, 9) SWB(22, 8) SWB(21, 7) SWB(20, 6) SWB(19, 5) SWB(18, 4) SWB(17, 3) SWB(16, 2) SWB(15, 1) SWB(14, 0) \
    SWB(13, 23) SWB(12, 22) SWB(11, 21) SWB(10, 20) SWB(9, 19) SWB(8, 18) SWB(7, 17) SWB(6, 16) SWB(5, 15) SWB(4, 14) SWB(3, 13) SWB(2, 12) SWB(1, 11) SWB(0, 10)

__global__ void __launch_bounds__(128, 3) k_unpacked(uint32_t* st, int R) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    uint32_t s[24], c;
    for (int w = 0; w < 24; w++) s[w] = st[w * n + i];
    c = st[24 * n + i];
#pragma unroll 1
    for (int r = 0; r < R; r++) asm volatile(PRE24 SWB_24 POST24 : OPS24(s, c));
    for (int w = 0; w < 24; w++) st[w * n + i] = s[w];
    st[24 * n + i] = c;
}

// packed ring: W[0..17], oldest limb at bit 0 of W[0].  chunk c: W[3c+k] = funnel16(W[(10+3c+k)%18], W[(11+3c+k)%18]) - W[3c+k] - borrow
#define OPS18(w, c)                                                                              \
    "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]),     \
    "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(c)
#define PRE18  "{\n\t.reg .u32 t, a, h;\n\tsub.cc.u32 t, 0, %18;\n\t"
#define POST18 "subc.u32 t, 0, 0;\n\tand.b32 %18, t, 1;\n\t}"
// alignment by funnel shift (alu)
#define WS(D, LO, HI) "shf.r.wrap.b32 a, %" #LO ", %" #HI ", 16;\n\tsubc.cc.u32 %" #D ", a, %" #D ";\n\t"
// alignment by mul.hi / mad.lo (fma pipe)
#define WM(D, LO, HI) "mul.hi.u32 h, %" #LO ", 65536;\n\tmad.lo.u32 a, %" #HI ", 65536, h;\n\tsubc.cc.u32 %" #D ", a, %" #D ";\n\t"
#define ROT(A, B, C) \
    A(0, 10, 11) B(1, 11, 12) C(2, 12, 13) A(3, 13, 14) B(4, 14, 15) C(5, 15, 16) A(6, 16, 17) B(7, 17, 0) C(8, 0, 1) \
    A(9, 1, 2) B(10, 2, 3) C(11, 3, 4) A(12, 4, 5) B(13, 5, 6) C(14, 6, 7) A(15, 7, 8) B(16, 8, 9) C(17, 9, 10)

template <int V> __device__ __forceinline__ void rot(uint32_t (&w)[18], uint32_t& c) {
    if (V == 1) asm volatile(PRE18 ROT(WS, WS, WS) POST18 : OPS18(w, c));
    if (V == 2) asm volatile(PRE18 ROT(WM, WM, WM) POST18 : OPS18(w, c));
    if (V == 3) asm volatile(PRE18 ROT(WS, WM, WM) POST18 : OPS18(w, c));
}
__device__ __forceinline__ void pack(const uint32_t (&s)[24], uint32_t (&w)[18]) {
    // limb m (oldest first) = s[23 - m] >> 8
    for (int q = 0; q < 6; q++) {
        uint32_t y0 = s[23 - 4 * q], y1 = s[22 - 4 * q], y2 = s[21 - 4 * q], y3 = s[20 - 4 * q];
        w[3 * q + 0] = (y0 >> 8) | (y1 << 16);
        w[3 * q + 1] = (y1 >> 16) | (y2 << 8);
        w[3 * q + 2] = (y2 >> 24) | y3;
    }
}
__device__ __forceinline__ void unpack(const uint32_t (&w)[18], uint32_t (&s)[24]) {
    for (int q = 0; q < 6; q++) {
        uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
        s[23 - 4 * q] = w0 << 8;
        s[22 - 4 * q] = ((w0 >> 24) | (w1 << 8)) << 8;
        s[21 - 4 * q] = ((w1 >> 16) | (w2 << 16)) << 8;
        s[20 - 4 * q] = w2 & 0xFFFFFF00u;
    }
}
template <int V> __global__ void __launch_bounds__(128, 3) k_packed(uint32_t* st, int R) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    uint32_t s[24], w[18], c;
    for (int k = 0; k < 24; k++) s[k] = st[k * n + i];
    c = st[24 * n + i];
    pack(s, w);
#pragma unroll 1
    for (int r = 0; r < R; r++) rot<V>(w, c);
    unpack(w, s);
    for (int k = 0; k < 24; k++) st[k * n + i] = s[k];
    st[24 * n + i] = c;
}

int main(int argc, char** argv) {
    int R = argc > 1 ? atoi(argv[1]) : 4000;
    const int blocks = 148 * 3, threads = 128, n = blocks * threads;
    std::vector<uint32_t> h(25 * (size_t)n);
    srand(1);
    for (size_t k = 0; k < 24 * (size_t)n; k++) h[k] = ((uint32_t)rand() & 0xFFFFFFu) << 8;
    for (size_t k = 24 * (size_t)n; k < 25 * (size_t)n; k++) h[k] = rand() & 1;
    uint32_t* d[4];
    std::vector<uint32_t> out[4];
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = { "V0 unpacked subc+and", "V1 packed shf+subc", "V2 packed mul.hi+mad.lo+subc", "V3 packed mixed 1 shf : 2 imad" };
    for (int v = 0; v < 4; v++) {
        cudaMalloc(&d[v], h.size() * 4);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaMemcpy(d[v], h.data(), h.size() * 4, cudaMemcpyHostToDevice);
            cudaEventRecord(e0);
            if (v == 0) k_unpacked<<<blocks, threads>>>(d[v], R);
            if (v == 1) k_packed<1><<<blocks, threads>>>(d[v], R);
            if (v == 2) k_packed<2><<<blocks, threads>>>(d[v], R);
            if (v == 3) k_packed<3><<<blocks, threads>>>(d[v], R);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        out[v].resize(h.size());
        cudaMemcpy(out[v].data(), d[v], h.size() * 4, cudaMemcpyDeviceToHost);
        // cycles per step per SMSP-warp: 12 warps/SM -> 3 warps per scheduler
        double steps = 24.0 * R, cyc = best * 1e-3 * 1.965e9;
        printf("%-34s %8.3f ms  %6.2f clk/step/warp-slot (3 warps per scheduler)  same as V0: %s  err=%s\n", names[v], best, cyc / steps / 3.0,
               v == 0 ? "-" : (out[v] == out[0] ? "yes" : "NO"), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
