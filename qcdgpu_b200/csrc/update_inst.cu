// update_inst.cu -- instantiation unit for the update kernels.  Built once per
// (INST_SUN, INST_PREC, INST_MU) by build.py so that the 16 units compile in parallel.
#include "launch.h"
#include "update.cuh"
#include <cstdlib>

#ifndef INST_SUN
#error "compile with -DINST_SUN=2|3 -DINST_PREC=0|1 -DINST_MU=0..3"
#endif

namespace qcd {

#if INST_PREC == 1
typedef double RealT;
#else
typedef float RealT;
#endif

template <bool TBC, class Prn>
static cudaError_t launch_one(const UpdateArgs& a, typename Prn::Args pa, cudaStream_t st) {
    dim3 block(kUpdateThreads), grid((a.geom.nthreads + kUpdateThreads - 1) / kUpdateThreads);
    if (grid.x == 0) return cudaSuccess;
    const size_t smem = (INST_SUN == 3 && kStageK > 0) ? stage_bytes<RealT>() : 0;   // link ring (update.cuh, Su3Stage)
    static bool configured = false;
    if (smem > 48 * 1024 && !configured) {
        cudaError_t e = cudaFuncSetAttribute(update_kernel<INST_SUN, RealT, INST_MU, TBC, Prn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    update_kernel<INST_SUN, RealT, INST_MU, TBC, Prn><<<grid, block, smem, st>>>(
        static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd, pa);
    return cudaGetLastError();
}
// two-phase RANLUX kernel (generator phase -> shared memory -> physics phase)
template <bool TBC>
static cudaError_t launch_rlx(const UpdateArgs& a, cudaStream_t st) {
    // the most a link can consume in a pass (SU(2): one Kennedy-Pendleton instance, su2_update_cl.cl:534-550)
    uint32_t need = (INST_SUN == 3 ? 3u : 1u) * (a.geom.nhit + 1u) * a.geom.nhitpar;
    if (need > a.prn_samples) need = a.prn_samples;
    // With a separate output table the state in front of the pass survives the launch: park `keep` samples per thread, the rest
    // is regenerated on demand (update.cuh, PrnSmemT).  In place: park everything a link can consume.
    static int keep_env = -1;
    if (keep_env < 0) { const char* e = getenv("QCDGPU_RLX_KEEP"); keep_env = e ? atoi(e) : 12; if (keep_env < 1) keep_env = 1; if (keep_env > 64) keep_env = 64; }
    uint32_t* out = a.prn_table_out ? a.prn_table_out : a.prn_table;
    uint32_t keep = need;
    if (out != a.prn_table && (uint32_t)keep_env < need) keep = (uint32_t)keep_env;
    size_t smem = size_t(keep) * kUpdateThreads * sizeof(float4);
    // diagnostic: unused shared memory per block, to measure the kernel at 2 or 1 resident blocks per SM
    static int pad_env = -1;
    if (pad_env < 0) { const char* e = getenv("QCDGPU_SMEM_PAD_KB"); pad_env = e ? atoi(e) : 0; }
    smem += size_t(pad_env) * 1024;
    if (smem > 200 * 1024 || keep > 64) {   // does not fit (smem, or the 64-bit "already converted" mask): state-in-registers variant
        if (out != a.prn_table) return cudaErrorInvalidValue;
        PrnRanlux<223>::Args pa = { a.prn_table, a.prn_instances, a.prn_samples };
        return launch_one<TBC, PrnRanlux<223>>(a, pa, st);
    }
#if INST_SUN == 3
    // links of the staples staged by TMA: full warps that are one y-row of one parity each (update.cuh, Su3StageTma)
    static int tma_env = -1;
    if (tma_env < 0) { const char* e = getenv("QCDGPU_TMA"); tma_env = e ? atoi(e) : 1; }
    if (tma_env && a.tmap && a.tmap1 && kUpdateThreads == 128 && a.geom.n2 % 64u == 0u && a.geom.nthreads % kUpdateThreads == 0u) {
        auto kt = update_kernel_rlx_tma<RealT, INST_MU, TBC, 223>;
        const size_t smem_t = smem + tma_stage_bytes<RealT>();
        static size_t configured_t = 0;
        if (smem_t > configured_t) {
            cudaError_t e = cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
            if (e != cudaSuccess) return e;
            configured_t = smem_t;
        }
        kt<<<a.geom.nthreads / kUpdateThreads, kUpdateThreads, smem_t, st>>>(static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd,
                                                                           a.prn_table, out, a.prn_instances, a.prn_samples, keep, *a.tmap, *a.tmap1);
        return cudaGetLastError();
    }
#endif
    auto kern = update_kernel_rlx<INST_SUN, RealT, INST_MU, TBC, 223>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 block(kUpdateThreads), grid((a.geom.nthreads + kUpdateThreads - 1) / kUpdateThreads);
    if (grid.x == 0) return cudaSuccess;
    kern<<<grid, block, smem, st>>>(static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd,
                                    a.prn_table, out, a.prn_instances, a.prn_samples, keep);
    return cudaGetLastError();
}
template <bool TBC>
static cudaError_t launch_tbc(const UpdateArgs& a, cudaStream_t st) {
    switch (a.prn_kind) {
        case PRN_BUFFERED: { PrnBuffered::Args pa = { static_cast<const float4*>(a.prns) }; return launch_one<TBC, PrnBuffered>(a, pa, st); }
        case PRN_GID:      { PrnGid::Args pa = { 0 }; return launch_one<TBC, PrnGid>(a, pa, st); }
        case PRN_RANLUX3:  return launch_rlx<TBC>(a, st);
        case PRN_XOR128:   { PrnXor128::Args pa = { a.prn_table, a.prn_instances, a.prn_samples }; return launch_one<TBC, PrnXor128>(a, pa, st); }
        case PRN_PM:       { PrnPM::Args pa = { a.prn_table, a.prn_instances, a.prn_samples }; return launch_one<TBC, PrnPM>(a, pa, st); }
        case PRN_MRG32K3A: { PrnMrg::Args pa = { a.prn_table, a.prn_instances, a.prn_samples }; return launch_one<TBC, PrnMrg>(a, pa, st); }
        default: return cudaErrorInvalidValue;
    }
}

#define CAT_(a, b, c, d) a##_##b##_##c##_##d
#define CAT(a, b, c, d) CAT_(a, b, c, d)
#define LAUNCH_NAME CAT(launch_update, INST_SUN, INST_PREC, INST_MU)
#define REGS_NAME CAT(regs_update, INST_SUN, INST_PREC, INST_MU)

cudaError_t LAUNCH_NAME(const UpdateArgs& a, cudaStream_t st) {
    if (a.overrelax) {
#if INST_SUN == 3
        static int or_tma_env = -1;
        if (or_tma_env < 0) { const char* e = getenv("QCDGPU_TMA"); or_tma_env = e ? atoi(e) : 1; }
        if (or_tma_env && a.tmap && a.tmap1 && kUpdateThreads == 128 && a.geom.n2 % 64u == 0u && a.geom.nthreads > 0 && a.geom.nthreads % kUpdateThreads == 0u) {
            auto kt = overrelax_kernel_tma<RealT, INST_MU>;
            const size_t smem_t = tma_stage_bytes<RealT>();
            static bool configured_or = false;
            if (!configured_or) {
                cudaError_t e = cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
                if (e != cudaSuccess) return e;
                configured_or = true;
            }
            kt<<<a.geom.nthreads / kUpdateThreads, kUpdateThreads, smem_t, st>>>(static_cast<RealT*>(a.table), a.geom, a.odd, *a.tmap, *a.tmap1);
            return cudaGetLastError();
        }
#endif
        dim3 block(kUpdateThreads), grid((a.geom.nthreads + kUpdateThreads - 1) / kUpdateThreads);
        if (grid.x == 0) return cudaSuccess;
        overrelax_kernel<INST_SUN, RealT, INST_MU><<<grid, block, 0, st>>>(static_cast<RealT*>(a.table), a.geom, a.odd);
        return cudaGetLastError();
    }
    return a.tbc ? launch_tbc<true>(a, st) : launch_tbc<false>(a, st);
}

template <bool TBC, class Prn> static int regs_of() {
    cudaFuncAttributes at;
    if (cudaFuncGetAttributes(&at, update_kernel<INST_SUN, RealT, INST_MU, TBC, Prn>) != cudaSuccess) return -1;
    return at.numRegs;
}
template <bool TBC> static int regs_of_rlx() {
    cudaFuncAttributes at;
    if (cudaFuncGetAttributes(&at, update_kernel_rlx<INST_SUN, RealT, INST_MU, TBC, 223>) != cudaSuccess) return -1;
    return at.numRegs;
}
int REGS_NAME(int tbc, int prn_kind) {
    switch (prn_kind) {
        case PRN_BUFFERED: return tbc ? regs_of<true, PrnBuffered>() : regs_of<false, PrnBuffered>();
        case PRN_GID:      return tbc ? regs_of<true, PrnGid>() : regs_of<false, PrnGid>();
        case PRN_RANLUX3:  return tbc ? regs_of_rlx<true>() : regs_of_rlx<false>();
        case PRN_XOR128:   return tbc ? regs_of<true, PrnXor128>() : regs_of<false, PrnXor128>();
        case PRN_PM:       return tbc ? regs_of<true, PrnPM>() : regs_of<false, PrnPM>();
        case PRN_MRG32K3A: return tbc ? regs_of<true, PrnMrg>() : regs_of<false, PrnMrg>();
        default: return -1;
    }
}

}  // namespace qcd

#if defined(QCD_TRACE) && INST_SUN == 3 && INST_PREC == 1 && INST_MU == 0
// diagnostic build only: copy out the phase time stamps of the last SU(3) fp64 X-direction launch
extern "C" int qgpu_debug_trace(unsigned long long* out, size_t n) {
    return (int)cudaMemcpyFromSymbol(out, qcd::g_trace, n * sizeof(unsigned long long));
}
#endif
