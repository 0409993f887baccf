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

// Opt-in to more than 48 KB of dynamic shared memory.  cudaFuncSetAttribute applies to the CURRENT device only, and one process may
// drive several (BigLattice parts in one process): remember the configured size per device.
constexpr int kMaxDevices = 64;
struct SmemConfigured { size_t bytes[kMaxDevices] = {}; };
template <class K>
static cudaError_t ensure_smem(K kern, size_t bytes, SmemConfigured& done) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices && bytes <= done.bytes[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) done.bytes[dev] = bytes;
    return cudaSuccess;
}
static int sm_count() {
    static int n[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 1;
    if (n[dev] == 0 && cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 1;
    return n[dev];
}
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

template <bool TBC, class Prn>
static cudaError_t launch_one(const UpdateArgs& a, typename Prn::Args pa, cudaStream_t st) {
    dim3 block(kUpdateThreads), grid((a.geom.nthreads + kUpdateThreads - 1) / kUpdateThreads);
    if (grid.x == 0) return cudaSuccess;
    const size_t smem = (INST_SUN == 3 && kStageK > 0) ? stage_bytes<RealT>() : 0;   // link ring (update.cuh, Su3Stage)
    static SmemConfigured configured;
    if (cudaError_t e = ensure_smem(update_kernel<INST_SUN, RealT, INST_MU, TBC, Prn>, smem, configured); e != cudaSuccess) return e;
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
    if (keep_env < 0) { keep_env = env_int("QCDGPU_RLX_KEEP", 12); if (keep_env < 1) keep_env = 1; if (keep_env > 64) keep_env = 64; }
    uint32_t* out = a.prn_table_out ? a.prn_table_out : a.prn_table;
    uint32_t keep = need;
    if (out != a.prn_table && (uint32_t)keep_env < need) keep = (uint32_t)keep_env;
#if INST_SUN == 3 && defined(QCD_SYM_EXPERIMENT)
    // experiment: persistent kernel with symmetric warp groups and registers on loan (update.cuh, update_kernel_rlx_sym); 16 warps per SM
    // leave room for 8 parked samples per thread
    static int sym_env = -1;
    if (sym_env < 0) sym_env = env_int("QCDGPU_SYM", 0);
    const bool sym = sym_env && a.tmap && a.tmap1 && a.ws_counters && out != a.prn_table && a.geom.n2 % 64u == 0u && a.geom.nthreads % 32u == 0u &&
                     a.geom.nthreads > 0u;
    if (sym && keep > 8u) keep = 8u;
#else
    const bool sym = false; (void)sym;
#endif
    size_t smem = size_t(keep) * kUpdateThreads * sizeof(float4);
    // the rest of the pass behind the parked samples as one modular multiplication (ranlux_jump.cuh); QCDGPU_RLX_JUMP=0: stepped
    static int jump_env = -1;
    if (jump_env < 0) jump_env = env_int("QCDGPU_RLX_JUMP", 1);
    const uint32_t* jump = (jump_env && keep < a.prn_samples) ? ranlux_jump_table_device(a.prn_samples, keep, 223u) : nullptr;
    // diagnostic: unused shared memory per block, to measure the kernel at 2 or 1 resident blocks per SM
    static int pad_env = -1;
    if (pad_env < 0) pad_env = env_int("QCDGPU_SMEM_PAD_KB", 0);
    smem += size_t(pad_env) * 1024;
    if (smem > 200 * 1024 || keep > 64) {   // does not fit (smem, or the 64-bit "already converted" mask): state-in-registers variant
        if (out != a.prn_table) return cudaErrorInvalidValue;
        PrnRanlux<223>::Args pa = { a.prn_table, a.prn_instances, a.prn_samples };
        return launch_one<TBC, PrnRanlux<223>>(a, pa, st);
    }
#if INST_SUN == 3
    // links of the staples staged by TMA: full warps that are one y-row of one parity each (update.cuh, Su3StageTma)
    static int tma_env = -1;
    if (tma_env < 0) tma_env = env_int("QCDGPU_TMA", 1);
    // warp-specialised persistent kernel (update.cuh, update_kernel_rlx_ws): needs the second state table (rare path of the parked
    // samples), rows of 32 update threads and room for 12 sample buffers + 12 link rings in shared memory.
    // Measured slower than the two-phase kernel below (4.46 against 3.62 ms per 64^4 pass, DESIGN.md decision 14): opt-in, QCDGPU_WS=1.
    // QCDGPU_WS_MIN: smallest launch (update threads) that takes it; QCDGPU_WS_BLOCKS: blocks instead of one per SM.
#ifdef QCD_SYM_EXPERIMENT
    if (sym && sym_smem_bytes<RealT>(keep) <= 227u * 1024u) {
        auto ks = update_kernel_rlx_sym<RealT, INST_MU, TBC, 223>;
        const size_t smem_s = sym_smem_bytes<RealT>(keep);
        static SmemConfigured configured_s;
        if (cudaError_t e = ensure_smem(ks, smem_s, configured_s); e != cudaSuccess) return e;
        static int sym_jump = -1;
        if (sym_jump < 0) sym_jump = env_int("QCDGPU_SYM_JUMP", 0);
        const uint32_t tickets = (a.geom.nthreads + 127u) / 128u;
        uint32_t blocks = (tickets + 3u) / 4u;
        if (blocks > (uint32_t)sm_count()) blocks = (uint32_t)sm_count();
        ks<<<blocks, kSymThreads, smem_s, st>>>(static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd,
                                               a.prn_table, out, a.prn_instances, a.prn_samples, keep, *a.tmap, *a.tmap1, sym_jump ? jump : nullptr, a.ws_counters);
        return cudaGetLastError();
    }
#endif
    static int ws_env = -1, ws_min = 0, ws_blocks = 0;
    if (ws_env < 0) { ws_env = env_int("QCDGPU_WS", 0); ws_min = env_int("QCDGPU_WS_MIN", 0); ws_blocks = env_int("QCDGPU_WS_BLOCKS", 0); }
    if (ws_env && tma_env && a.tmap && a.tmap1 && a.ws_counters && out != a.prn_table && keep <= 16u && pad_env == 0 &&
        a.geom.n2 % 64u == 0u && a.geom.nthreads % 32u == 0u && a.geom.nthreads >= (uint32_t)ws_min &&
        ws_smem_bytes<RealT>(keep) <= 227u * 1024u) {
        auto kw = update_kernel_rlx_ws<RealT, INST_MU, TBC, 223>;
        const size_t smem_w = ws_smem_bytes<RealT>(keep);
        static SmemConfigured configured_w;
        if (cudaError_t e = ensure_smem(kw, smem_w, configured_w); e != cudaSuccess) return e;
        const uint32_t tickets = (a.geom.nthreads + 32u * kWsChains - 1u) / (32u * kWsChains);
        uint32_t blocks = (tickets + kWsGenWarps - 1u) / kWsGenWarps;
        const uint32_t cap = ws_blocks > 0 ? (uint32_t)ws_blocks : (uint32_t)sm_count();
        if (blocks > cap) blocks = cap;
        kw<<<blocks, kWsThreads, smem_w, st>>>(static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd,
                                              a.prn_table, out, a.prn_instances, a.prn_samples, keep, *a.tmap, *a.tmap1, a.ws_counters);
        return cudaGetLastError();
    }
#ifdef QCD_SPLIT_EXPERIMENT
    // experiment (update.cuh): the pass as a staple kernel + a heat-bath kernel with the staple sums in a scratch buffer
    static int split_env = -1;
    if (split_env < 0) split_env = env_int("QCDGPU_SPLIT", 0);
    if (split_env && tma_env && a.tmap && a.tmap1 && a.geom.n2 % 64u == 0u && a.geom.nthreads % 32u == 0u && a.geom.nthreads > 0u &&
        a.split_scratch && a.split_scratch_bytes >= size_t(a.geom.nthreads) * 9 * 2 * sizeof(RealT) && out != a.prn_table) {
        auto ka = staple_kernel_tma<RealT, INST_MU, TBC>;
        auto kb = heatbath_kernel_rlx<RealT, INST_MU, TBC, 223>;
        const size_t smem_a = tma_stage_bytes<RealT>();
        static SmemConfigured configured_a, configured_b;
        if (cudaError_t e = ensure_smem(ka, smem_a, configured_a); e != cudaSuccess) return e;
        if (cudaError_t e = ensure_smem(kb, smem, configured_b); e != cudaSuccess) return e;
        const uint32_t blocks = (a.geom.nthreads + kUpdateThreads - 1) / kUpdateThreads;
        ka<<<blocks, kUpdateThreads, smem_a, st>>>(static_cast<const RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd, *a.tmap, *a.tmap1,
                                                  static_cast<RealT*>(a.split_scratch));
        if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
        kb<<<blocks, kUpdateThreads, smem, st>>>(static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd, a.prn_table, out,
                                                 a.prn_instances, a.prn_samples, keep, jump, static_cast<const RealT*>(a.split_scratch));
        return cudaGetLastError();
    }
#endif
    if (tma_env && a.tmap && a.tmap1 && kUpdateThreads % 32 == 0 && a.geom.n2 % 64u == 0u && a.geom.nthreads % 32u == 0u && a.geom.nthreads > 0u) {
        auto kt = update_kernel_rlx_tma<RealT, INST_MU, TBC, 223>;
        const size_t smem_t = smem + tma_stage_bytes<RealT>();
        static SmemConfigured configured_t;
        if (cudaError_t e = ensure_smem(kt, smem_t, configured_t); e != cudaSuccess) return e;
        kt<<<(a.geom.nthreads + kUpdateThreads - 1) / kUpdateThreads, kUpdateThreads, smem_t, st>>>(static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd,
                                                                           a.prn_table, out, a.prn_instances, a.prn_samples, keep, *a.tmap, *a.tmap1, jump);
        return cudaGetLastError();
    }
#endif
    auto kern = update_kernel_rlx<INST_SUN, RealT, INST_MU, TBC, 223>;
    static SmemConfigured configured;
    if (cudaError_t e = ensure_smem(kern, smem, configured); e != cudaSuccess) return e;
    dim3 block(kUpdateThreads), grid((a.geom.nthreads + kUpdateThreads - 1) / kUpdateThreads);
    if (grid.x == 0) return cudaSuccess;
    kern<<<grid, block, smem, st>>>(static_cast<RealT*>(a.table), static_cast<const RealT*>(a.params), a.geom, a.odd,
                                    a.prn_table, out, a.prn_instances, a.prn_samples, keep, jump);
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
        case PRN_XOR7:     { PrnXor7::Args pa = { a.prn_table, a.prn_instances, a.prn_samples }; return launch_one<TBC, PrnXor7>(a, pa, st); }
        case PRN_RANECU:   { PrnRanecu::Args pa = { a.prn_table, a.prn_instances, a.prn_samples }; return launch_one<TBC, PrnRanecu>(a, pa, st); }
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
        if (or_tma_env < 0) or_tma_env = env_int("QCDGPU_TMA", 1);
        if (or_tma_env && a.tmap && a.tmap1 && kUpdateThreads == 128 && a.geom.n2 % 64u == 0u && a.geom.nthreads > 0 && a.geom.nthreads % kUpdateThreads == 0u) {
            auto kt = overrelax_kernel_tma<RealT, INST_MU>;
            const size_t smem_t = tma_stage_bytes<RealT>();
            static SmemConfigured configured_or;
            if (cudaError_t e = ensure_smem(kt, smem_t, configured_or); e != cudaSuccess) return e;
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
        case PRN_XOR7:     return tbc ? regs_of<true, PrnXor7>() : regs_of<false, PrnXor7>();
        case PRN_RANECU:   return tbc ? regs_of<true, PrnRanecu>() : regs_of<false, PrnRanecu>();
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
#if defined(QCD_WS_TRACE) && INST_SUN == 3 && INST_PREC == 1 && INST_MU == 0
// diagnostic build only: time stamps of the last SU(3) fp64 X-direction launch of the warp-specialised kernel
extern "C" int qgpu_debug_ws_trace(unsigned long long* out, size_t n) {
    return (int)cudaMemcpyFromSymbol(out, qcd::g_ws_trace, n * sizeof(unsigned long long));
}
#endif
