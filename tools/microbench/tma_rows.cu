// Microbenchmark for the round-2 plan: staging the links of a staple through shared memory.
// At 64^4 a warp of an update pass is one y-row of one parity: every (link, plane) it reads is 32 elements of 32 B at a
// stride of 64 B.  Three ways to get 6 groups ("staples") of 9 such rows to the 32 lanes, each group followed by a block
// of dependent FP64 arithmetic that stands for the matrix products:
//   A   ld.global.v4.f64 per lane at the start of the group (what the update kernel does), L2 prefetch of the next group
//   B   one TMA box per (link, plane): cp.async.bulk.tensor, element strides (1, 2), box 4 x 64 -> 1 KB in shared memory,
//       issued by one lane for the NEXT group while the current one is multiplied, mbarrier wait, 32-byte LDS per lane
// Same bytes, same arithmetic, same occupancy (3 blocks of 128 threads per SM through shared-memory padding).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int kGroups = 6, kBoxes = 9;
__constant__ int kFma;                                // kFma x 8 dependent DFMA per lane and group (argv[2])
struct Pattern { int row_off[kGroups][kBoxes]; int plane[kGroups][kBoxes]; int par[kGroups][kBoxes]; };

__device__ __forceinline__ void work(double (&acc)[8], const double (&v)[kBoxes][4]) {
    double m[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
    for (int b = 0; b < kBoxes; b++)
#pragma unroll
        for (int e = 0; e < 4; e++) m[(4 * b + e) & 7] += v[b][e];
#pragma unroll 1
    for (int it = 0; it < kFma; it++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = fma(acc[k], m[k], m[(k + 3) & 7]);
}

__global__ void __launch_bounds__(128, 3)
k_ldg(const double* __restrict__ table, size_t rowsize, uint32_t nrows, Pattern pt, double* out) {
    extern __shared__ uint8_t pad[];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    double acc[8] = { 1, 1, 1, 1, 1, 1, 1, 1 };
    auto addr = [&](int g, int b) {
        uint32_t row = (warp + nrows + pt.row_off[g][b]) % nrows;
        return table + 4 * (size_t(row) * 64 + pt.par[g][b] + 2 * lane + size_t(pt.plane[g][b]) * rowsize);
    };
#pragma unroll 1
    for (int g = 0; g < kGroups; g++) {
        double v[kBoxes][4];
#pragma unroll
        for (int b = 0; b < kBoxes; b++)
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[b][0]), "=d"(v[b][1]), "=d"(v[b][2]), "=d"(v[b][3]) : "l"(addr(g, b)));
        if (g + 1 < kGroups) {
#pragma unroll
            for (int b = 0; b < kBoxes; b++) asm volatile("prefetch.global.L2 [%0];" :: "l"(addr(g + 1, b)));
        }
        work(acc, v);
    }
    double s = 0; for (int k = 0; k < 8; k++) s += acc[k];
    out[size_t(warp) * 32 + lane] = s;
}

__global__ void __launch_bounds__(128, 3)
k_tma(const __grid_constant__ CUtensorMap tmap, uint32_t nrows, Pattern pt, double* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t* buf = smem + wib * (kBoxes * 1024);                       // 9 KB per warp
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * kBoxes * 1024) + wib;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), buf_a = (uint32_t)__cvta_generic_to_shared(buf);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (warp >= nrows) return;
    auto issue = [&](int g) {
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(kBoxes * 1024) : "memory");
#pragma unroll
            for (int b = 0; b < kBoxes; b++) {
                uint32_t row = (warp + nrows + pt.row_off[g][b]) % nrows;
                int c1 = int(row * 64 + pt.par[g][b]), c2 = pt.plane[g][b];
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             :: "r"(buf_a + b * 1024), "l"(&tmap), "r"(0), "r"(c1), "r"(c2), "r"(bar_a) : "memory");
            }
        }
    };
    double acc[8] = { 1, 1, 1, 1, 1, 1, 1, 1 };
    issue(0);
#pragma unroll 1
    for (int g = 0; g < kGroups; g++) {
        uint32_t done = 0, phase = g & 1;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        double v[kBoxes][4];
#pragma unroll
        for (int b = 0; b < kBoxes; b++) {
            const double2* p = reinterpret_cast<const double2*>(buf + b * 1024 + lane * 32);
            double2 lo = p[0], hi = p[1];
            v[b][0] = lo.x; v[b][1] = lo.y; v[b][2] = hi.x; v[b][3] = hi.y;
        }
        __syncwarp();
        if (g + 1 < kGroups) issue(g + 1);
        work(acc, v);
    }
    double s = 0; for (int k = 0; k < 8; k++) s += acc[k];
    out[size_t(warp) * 32 + lane] = s;
}

int main(int argc, char** argv) {
    const int L = argc > 1 ? atoi(argv[1]) : 64;
    const int nfma = argc > 2 ? atoi(argv[2]) : 40;
    cudaMemcpyToSymbol(kFma, &nfma, sizeof(int));
    const uint32_t nrows = (uint32_t)L * L * L;                       // (x, t, z) triples; a row is 64 sites
    const size_t rowsize = size_t(nrows) * 64;
    const int planes = 12;
    double* table; double* out[2];
    cudaMalloc(&table, rowsize * planes * 32);
    cudaMemset(table, 0, rowsize * planes * 32);
    {   // a few non-zero values so that the two paths can be compared
        std::vector<double> h(1 << 20);
        for (size_t i = 0; i < h.size(); i++) h[i] = 1e-3 * double((i * 2654435761u) % 1000);
        for (int p = 0; p < planes; p++) cudaMemcpy(table + 4 * size_t(p) * rowsize, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    }
    for (int i = 0; i < 2; i++) { cudaMalloc(&out[i], size_t(nrows) * 32 * 8); cudaMemset(out[i], 0, size_t(nrows) * 32 * 8); }
    Pattern pt;
    const int offs[6] = { 1, -1, L, -L, L * L, -L * L };
    for (int g = 0; g < kGroups; g++)
        for (int b = 0; b < kBoxes; b++) {
            pt.row_off[g][b] = (b / 3 == 2) ? 0 : offs[(g + b / 3) % 6];
            pt.plane[g][b] = (g % 4) + 4 * (b % 3);
            pt.par[g][b] = (g + b / 3) & 1;
        }
    // tensor map: dims (4 doubles, sites, planes), traversal strides (1, 2, 1), box 4 x 64 x 1 -> 32 elements of 32 B
    CUtensorMap tmap;
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    cuuint64_t gdim[3] = { 4, rowsize, (cuuint64_t)planes }, gstr[2] = { 32, rowsize * 32 };
    cuuint32_t box[3] = { 4, 64, 1 }, estr[3] = { 1, 2, 1 };
    CUresult r = ((Enc)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, table, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    const int threads = 128, blocks = (nrows * 32 + threads - 1) / threads;
    const size_t smem_tma = 4 * kBoxes * 1024 + 64, smem_pad = smem_tma;        // same occupancy for both
    cudaFuncSetAttribute(k_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pad);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = double(nrows) * kGroups * kBoxes * 1024;
    for (int v = 0; v < 2; v++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (v == 0) k_ldg<<<blocks, threads, smem_pad>>>(table, rowsize, nrows, pt, out[0]);
            else        k_tma<<<blocks, threads, smem_tma>>>(tmap, nrows, pt, out[1]);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("[%d x 8 DFMA per group] %-44s %8.3f ms   %7.1f GB/s of row data   (%s)\n",
               nfma, v == 0 ? "A  ld.global.v4.f64 + L2 prefetch" : "B  TMA strided boxes -> smem, one group ahead", best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<double> a(size_t(nrows) * 32), b(a.size());
    cudaMemcpy(a.data(), out[0], a.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), out[1], b.size() * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0; for (size_t i = 0; i < a.size(); i++) if (a[i] != b[i]) bad++;
    printf("results identical: %s (%zu of %zu differ)\n", bad ? "NO" : "yes", bad, a.size());
    return 0;
}
