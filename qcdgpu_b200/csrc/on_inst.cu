// on_inst.cu -- instantiation unit of the O(N) scalar branch (on_model.cuh), both precisions.
#include "on_model.cuh"

namespace qcd {

template <typename R>
static cudaError_t on_update_t(const OnArgs& a, cudaStream_t st) {
    dim3 block(kOnThreads), grid((a.geom.nthreads + kOnThreads - 1) / kOnThreads);
    if (grid.x == 0) return cudaSuccess;
    R* t = static_cast<R*>(a.table);
    const R* p = static_cast<const R*>(a.params);
    switch (a.prn_kind) {
        case PRN_BUFFERED: {
            PrnBuffered::Args pa = { static_cast<const float4*>(a.prns) };
            on_update_kernel<R, PrnBuffered><<<grid, block, 0, st>>>(t, p, a.geom, a.skgroup, a.odd, pa, a.acc);
            break;
        }
        case PRN_XOR128: {
            PrnXor128::Args pa = { a.prn_table, a.prn_instances, a.prn_samples };
            on_update_kernel<R, PrnXor128><<<grid, block, 0, st>>>(t, p, a.geom, a.skgroup, a.odd, pa, a.acc);
            break;
        }
        case PRN_PM: {
            PrnPM::Args pa = { a.prn_table, a.prn_instances, a.prn_samples };
            on_update_kernel<R, PrnPM><<<grid, block, 0, st>>>(t, p, a.geom, a.skgroup, a.odd, pa, a.acc);
            break;
        }
        case PRN_MRG32K3A: {
            PrnMrg::Args pa = { a.prn_table, a.prn_instances, a.prn_samples };
            on_update_kernel<R, PrnMrg><<<grid, block, 0, st>>>(t, p, a.geom, a.skgroup, a.odd, pa, a.acc);
            break;
        }
        case PRN_XOR7: {
            PrnXor7::Args pa = { a.prn_table, a.prn_instances, a.prn_samples };
            on_update_kernel<R, PrnXor7><<<grid, block, 0, st>>>(t, p, a.geom, a.skgroup, a.odd, pa, a.acc);
            break;
        }
        case PRN_RANECU: {
            PrnRanecu::Args pa = { a.prn_table, a.prn_instances, a.prn_samples };
            on_update_kernel<R, PrnRanecu><<<grid, block, 0, st>>>(t, p, a.geom, a.skgroup, a.odd, pa, a.acc);
            break;
        }
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_on_update(int prec, const OnArgs& a, cudaStream_t st) {
    return prec == 1 ? on_update_t<double>(a, st) : on_update_t<float>(a, st);
}
cudaError_t launch_on_init(int prec, void* table, const void* params, const void* prns, uint32_t sites, cudaStream_t st) {
    dim3 block(kOnThreads), grid((sites + kOnThreads - 1) / kOnThreads);
    if (prec == 1) on_init_kernel<double><<<grid, block, 0, st>>>((double*)table, (const double*)params, (const float4*)prns, sites);
    else           on_init_kernel<float><<<grid, block, 0, st>>>((float*)table, (const float*)params, (const float4*)prns, sites);
    return cudaGetLastError();
}
cudaError_t launch_on_measure(int prec, const OnArgs& a, int what, const uint32_t step[4], double2* out, cudaStream_t st) {
    dim3 block(kOnThreads), grid((a.geom.sites + kOnThreads - 1) / kOnThreads);
    uint4 s = make_uint4(step[0], step[1], step[2], step[3]);
    if (prec == 1) on_measure_kernel<double><<<grid, block, 0, st>>>((const double*)a.table, (const double*)a.params, a.geom, a.skgroup, what, s, out);
    else           on_measure_kernel<float><<<grid, block, 0, st>>>((const float*)a.table, (const float*)a.params, a.geom, a.skgroup, what, s, out);
    return cudaGetLastError();
}
cudaError_t launch_on_reduce_add(const double2* table, uint32_t size, double2* dst, uint32_t index, cudaStream_t st) {
    on_reduce_add_kernel<<<1, 1024, 0, st>>>(table, size, dst, index);
    return cudaGetLastError();
}

}  // namespace qcd
