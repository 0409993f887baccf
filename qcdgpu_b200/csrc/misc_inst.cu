// misc_inst.cu -- precision-independent kernels: PRN producers and double2 reducers.
#include "launch.h"
#include "measure.cuh"
#include "init_prng.cuh"

namespace qcd {

cudaError_t launch_reduce_double2(const double2* table, uint32_t off, uint32_t size, double2* dst, uint32_t index, cudaStream_t st) {
    reduce_double2_kernel<<<1, 1024, 0, st>>>(table, off, size, dst, index);
    return cudaGetLastError();
}
cudaError_t launch_reduce_slices(const double2* table, uint32_t size, uint32_t nslices, double2* dst, uint32_t index, uint32_t stride,
                                 uint32_t extra_off, uint32_t extra_size, bool with_extra, cudaStream_t st) {
    reduce_slices_kernel<<<nslices + (with_extra ? 1u : 0u), 256, 0, st>>>(table, size, nslices, dst, index, stride, extra_off, extra_size);
    return cudaGetLastError();
}
cudaError_t launch_reduce_double_x(const double2* table, uint32_t size, double* dst, uint32_t index, cudaStream_t st) {
    reduce_double_x_kernel<<<1, 1024, 0, st>>>(table, size, dst, index);
    return cudaGetLastError();
}
cudaError_t launch_clear_double2(double2* p, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    clear_double2_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n);
    return cudaGetLastError();
}
cudaError_t launch_ranlux_init(const uint32_t* seeds, uint32_t* table, uint32_t instances, cudaStream_t st) {
    ranlux_init_kernel<<<(instances + 127) / 128, 128, 0, st>>>(seeds, table, instances);
    return cudaGetLastError();
}
cudaError_t launch_ranmar_init(const uint32_t* seeds, float* table, uint32_t instances, cudaStream_t st) {
    ranmar_init_kernel<<<(4u * instances + 127) / 128, 128, 0, st>>>(seeds, table, instances);
    return cudaGetLastError();
}
cudaError_t launch_prn_produce(int kind, uint32_t* table, void* prns, uint32_t instances, uint32_t samples, cudaStream_t st) {
    dim3 grid((instances + 127) / 128), block(128);
    float4* out = static_cast<float4*>(prns);
    switch (kind) {
        case PRN_RANLUX0: ranlux_produce_kernel<24><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_RANLUX1: ranlux_produce_kernel<48><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_RANLUX2: ranlux_produce_kernel<97><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_RANLUX3: ranlux_produce_kernel<223><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_RANLUX4: ranlux_produce_kernel<389><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_XOR128:  small_produce_kernel<PrnXor128><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_PM:      small_produce_kernel<PrnPM><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_MRG32K3A: small_produce_kernel<PrnMrg><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_XOR7:    small_produce_kernel<PrnXor7><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_RANECU:  small_produce_kernel<PrnRanecu><<<grid, block, 0, st>>>(table, out, instances, samples); break;
        case PRN_RANMAR:  ranmar_produce_kernel<<<grid, block, 0, st>>>(reinterpret_cast<float4*>(table), out, instances, samples); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_ranlux_skip(uint32_t* table, uint32_t instances, uint64_t samples, cudaStream_t st) {
    if (samples == 0 || instances == 0) return cudaSuccess;
    const uint32_t* jump = ranlux_jump_table_device(samples, 0, 223u, true);      // null: every instance is stepped
    ranlux_skip_kernel<<<(instances + 127) / 128, 128, 0, st>>>(table, instances, jump, (unsigned long long)samples);
    return cudaGetLastError();
}

// Hand-shake of the halo exchange that is fused into the update kernels (Geom::peer_*): after its boundary launch of pass n a part
// writes n into a flag word in each neighbour's memory; before its next boundary launch it waits until both neighbours have
// announced the pass before.  The boundary kernel has completed when the signal kernel starts (same stream), so its stores into
// peer memory are performed; the fence + system-scope release / acquire order them against the flag.
static __global__ void halo_signal_kernel(unsigned long long* a, unsigned long long* b, unsigned long long value) {
    __threadfence_system();
    if (threadIdx.x == 0 && a) asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(a), "l"(value) : "memory");
    if (threadIdx.x == 1 && b) asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(b), "l"(value) : "memory");
}
// flags[0] = from the left neighbour, flags[1] = from the right one; flags[2] counts waits that gave up after `timeout` clocks
// (a neighbour that died must not hang this GPU: the host finds the count at its next synchronisation)
static __global__ void halo_wait_kernel(unsigned long long* flags, unsigned long long value, unsigned long long timeout) {
    if (threadIdx.x >= 2) return;
    if (*reinterpret_cast<volatile unsigned long long*>(flags + 2) != 0ull) return;     // a neighbour was lost before: every later wait would time out too
    const unsigned long long* f = flags + threadIdx.x;
    const long long t0 = clock64();
    unsigned long long v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= value) break;
        if ((unsigned long long)(clock64() - t0) > timeout) { atomicAdd(flags + 2, 1ull); break; }
        __nanosleep(200);
    }
}
cudaError_t launch_halo_signal(unsigned long long* a, unsigned long long* b, unsigned long long value, cudaStream_t st) {
    halo_signal_kernel<<<1, 32, 0, st>>>(a, b, value);
    return cudaGetLastError();
}
cudaError_t launch_halo_wait(unsigned long long* flags, unsigned long long value, unsigned long long timeout_clocks, cudaStream_t st) {
    halo_wait_kernel<<<1, 32, 0, st>>>(flags, value, timeout_clocks);
    return cudaGetLastError();
}

// Halo exchange building block: copy `nplanes` vec4 planes of one x-slice (n234 elements each)
// between two plane-strided arrays.  Plane k lives at (p0 + k*ps)*row_elems + x*n234.
// Used table->table (peer pointers over NVLink), table->packed and packed->table.
template <typename R>
__global__ void copy_planes_kernel(R* __restrict__ dst, size_t dst_row, uint32_t xd, int dp0, int dps,
                                   const R* __restrict__ src, size_t src_row, uint32_t xs, int sp0, int sps, uint32_t n234) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n234) return;
    int k = blockIdx.y;
    const R* s = src + 4 * (size_t(xs) * n234 + i + size_t(sp0 + k * sps) * src_row);
    R* d = dst + 4 * (size_t(xd) * n234 + i + size_t(dp0 + k * dps) * dst_row);
    stv4(d, ldv4(s));
}
cudaError_t launch_copy_planes(int prec, void* dst, size_t dst_row, uint32_t xd, int dp0, int dps,
                               const void* src, size_t src_row, uint32_t xs, int sp0, int sps,
                               uint32_t n234, int nplanes, cudaStream_t st) {
    dim3 grid((n234 + 127) / 128, nplanes), block(128);
    if (prec == 1) copy_planes_kernel<double><<<grid, block, 0, st>>>((double*)dst, dst_row, xd, dp0, dps, (const double*)src, src_row, xs, sp0, sps, n234);
    else           copy_planes_kernel<float><<<grid, block, 0, st>>>((float*)dst, dst_row, xd, dp0, dps, (const float*)src, src_row, xs, sp0, sps, n234);
    return cudaGetLastError();
}

}  // namespace qcd
