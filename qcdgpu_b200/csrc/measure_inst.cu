// measure_inst.cu -- instantiation unit for measurement, start and reunitarisation kernels.  Built once per (INST_SUN, INST_PREC).
#include "launch.h"
#include "measure.cuh"
#include "init_prng.cuh"

#ifndef INST_SUN
#error "compile with -DINST_SUN=2|3 -DINST_PREC=0|1"
#endif

namespace qcd {

#if INST_PREC == 1
typedef double RealT;
#else
typedef float RealT;
#endif

#define CAT_(a, b, c) a##_##b##_##c
#define CAT(a, b, c) CAT_(a, b, c)
#define NAME(f) CAT(f, INST_SUN, INST_PREC)

cudaError_t NAME(launch_measure)(const MeasureArgs& a, cudaStream_t st) {
    uint32_t block = a.block ? a.block : kMeasThreads;
    if (block > kMeasThreads) block = kMeasThreads;
    dim3 grid((a.nsites + block - 1) / block);
    if (grid.x < a.min_blocks) grid.x = a.min_blocks;      // blocks without a site write a zero partial
    if (grid.x == 0) return cudaSuccess;
    const RealT* t = static_cast<const RealT*>(a.table);
    const RealT* p = static_cast<const RealT*>(a.params);
    if (a.wilson) {
        if (a.tbc) wilson_kernel<INST_SUN, RealT, true><<<grid, block, 0, st>>>(t, p, a.geom, a.out_plq, a.first_site, a.nsites);
        else       wilson_kernel<INST_SUN, RealT, false><<<grid, block, 0, st>>>(t, p, a.geom, a.out_plq, a.first_site, a.nsites);
        return cudaGetLastError();
    }
    if (a.fmode) {
        if (a.tbc) measure_f_kernel<INST_SUN, RealT, true><<<grid, block, 0, st>>>(t, p, a.geom, a.out_plq, a.offset, a.fmode, a.gen1, a.gen2, a.first_site, a.nsites);
        else       measure_f_kernel<INST_SUN, RealT, false><<<grid, block, 0, st>>>(t, p, a.geom, a.out_plq, a.offset, a.fmode, a.gen1, a.gen2, a.first_site, a.nsites);
        return cudaGetLastError();
    }
    if (a.diff_axis >= 0) {    // slice-resolved action: the host's grid and work-group size define the partial layout
        if (a.min_blocks == 0 || a.block == 0 || a.block > (uint32_t)kMeasThreads || a.plk == 0) return cudaErrorInvalidValue;
        if (a.tbc) measure_kernel<INST_SUN, RealT, true, true><<<a.min_blocks, a.block, 0, st>>>(t, p, a.geom, a.out_action, nullptr, 0u, 0u, a.diff_axis, a.plk);
        else       measure_kernel<INST_SUN, RealT, false, true><<<a.min_blocks, a.block, 0, st>>>(t, p, a.geom, a.out_action, nullptr, 0u, 0u, a.diff_axis, a.plk);
        return cudaGetLastError();
    }
    if (a.tbc) measure_kernel<INST_SUN, RealT, true><<<grid, block, 0, st>>>(t, p, a.geom, a.out_action, a.out_plq, a.first_site, a.nsites);
    else       measure_kernel<INST_SUN, RealT, false><<<grid, block, 0, st>>>(t, p, a.geom, a.out_action, a.out_plq, a.first_site, a.nsites);
    return cudaGetLastError();
}
cudaError_t NAME(launch_polyakov)(const PolyakovArgs& a, cudaStream_t st) {
    uint32_t block = a.block ? a.block : kMeasThreads;
    if (block > kMeasThreads) block = kMeasThreads;
    const RealT* t = static_cast<const RealT*>(a.table);
    const RealT* p = static_cast<const RealT*>(a.params);
    if (a.axis >= 0) {     // slice-resolved: the host's grid and work-group size define the partial layout
        if (a.min_blocks == 0 || a.block == 0 || a.block > 256) return cudaErrorInvalidValue;
        if (a.tbc) polyakov_diff_kernel<INST_SUN, RealT, true><<<a.min_blocks, a.block, 0, st>>>(t, p, a.geom, a.out, a.axis, a.plk);
        else       polyakov_diff_kernel<INST_SUN, RealT, false><<<a.min_blocks, a.block, 0, st>>>(t, p, a.geom, a.out, a.axis, a.plk);
        return cudaGetLastError();
    }
    uint32_t n = a.x_count * a.geom.n23;
    dim3 grid((n + block - 1) / block);
    if (grid.x < a.min_blocks) grid.x = a.min_blocks;
    if (grid.x == 0) return cudaSuccess;
    if (a.tbc) polyakov_kernel<INST_SUN, RealT, true><<<grid, block, 0, st>>>(t, p, a.geom, a.out, a.offset, a.pl_level, a.x_first, a.x_count);
    else       polyakov_kernel<INST_SUN, RealT, false><<<grid, block, 0, st>>>(t, p, a.geom, a.out, a.offset, a.pl_level, a.x_first, a.x_count);
    return cudaGetLastError();
}
cudaError_t NAME(launch_init_cold)(void* table, const Geom& g, cudaStream_t st) {
    init_cold_kernel<INST_SUN, RealT><<<(g.sites + 127) / 128, 128, 0, st>>>(static_cast<RealT*>(table), g);
    return cudaGetLastError();
}
cudaError_t NAME(launch_init_gid)(void* table, const Geom& g, cudaStream_t st) {
    init_gid_kernel<INST_SUN, RealT><<<(g.sites + 127) / 128, 128, 0, st>>>(static_cast<RealT*>(table), g);
    return cudaGetLastError();
}
cudaError_t NAME(launch_init_hot)(void* table, const Geom& g, const void* prns, int dir, cudaStream_t st) {
    init_hot_kernel<INST_SUN, RealT><<<(g.sites + 127) / 128, 128, 0, st>>>(static_cast<RealT*>(table), g, static_cast<const float4*>(prns), dir);
    return cudaGetLastError();
}
cudaError_t NAME(launch_gramschmidt)(void* table, const Geom& g, cudaStream_t st) {
    dim3 grid((g.sites + 127) / 128, 4);
    gramschmidt_kernel<INST_SUN, RealT><<<grid, 128, 0, st>>>(static_cast<RealT*>(table), g);
    return cudaGetLastError();
}

}  // namespace qcd
