// dispatch.cu -- tables that route (SUN, precision, direction) to the separately compiled
// instantiation units.
#include "launch.h"
#include "ranlux_jump_host.h"
#include <mutex>
#include <vector>

namespace qcd {

// Device copy of the RANLUX skip-ahead table (ranlux_jump_host.h) for a pass of `samples` float4 of which `keep` are parked, on the
// current device (from_start: the whole pass is jumped over, for ranlux_skip_kernel); cached for the life of the process.  Null on any failure: the kernels then step.
const uint32_t* ranlux_jump_table_device(uint64_t samples, uint64_t keep, uint32_t skip, bool from_start) {
    struct Entry { int dev; uint64_t samples, keep; uint32_t skip; bool from_start; uint32_t* dptr; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
        if (e.dev == dev && e.samples == samples && e.keep == keep && e.skip == skip && e.from_start == from_start) return e.dptr;
    uint32_t table[7 * kRanluxJumpEntryWords];
    ranlux_jump_table(samples, keep, skip, table, from_start);
    bool any = false;
    for (int e = 0; e < 7; e++) any = any || table[e * kRanluxJumpEntryWords + 19] != 0u;
    uint32_t* d = nullptr;
    if (any && (cudaMalloc(&d, sizeof(table)) != cudaSuccess || cudaMemcpy(d, table, sizeof(table), cudaMemcpyHostToDevice) != cudaSuccess)) {
        (void)cudaGetLastError();
        if (d) cudaFree(d);
        d = nullptr;
    }
    cache.push_back({ dev, samples, keep, skip, from_start, d });
    return d;
}

#define DECL_UPD(S, P, M) cudaError_t launch_update_##S##_##P##_##M(const UpdateArgs&, cudaStream_t); int regs_update_##S##_##P##_##M(int, int);
#define DECL_UPD4(S, P) DECL_UPD(S, P, 0) DECL_UPD(S, P, 1) DECL_UPD(S, P, 2) DECL_UPD(S, P, 3)
DECL_UPD4(2, 0) DECL_UPD4(2, 1) DECL_UPD4(3, 0) DECL_UPD4(3, 1)

#define ROW(S, P) { launch_update_##S##_##P##_0, launch_update_##S##_##P##_1, launch_update_##S##_##P##_2, launch_update_##S##_##P##_3 }
static const UpdateLaunchFn kUpdate[2][2][4] = { { ROW(2, 0), ROW(2, 1) }, { ROW(3, 0), ROW(3, 1) } };
typedef int (*RegsFn)(int, int);
#define RROW(S, P) { regs_update_##S##_##P##_0, regs_update_##S##_##P##_1, regs_update_##S##_##P##_2, regs_update_##S##_##P##_3 }
static const RegsFn kRegs[2][2][4] = { { RROW(2, 0), RROW(2, 1) }, { RROW(3, 0), RROW(3, 1) } };

UpdateLaunchFn get_update_launcher(int sun, int prec, int mu) {
    if (sun < 2 || sun > 3 || prec < 0 || prec > 1 || mu < 0 || mu > 3) return nullptr;
    return kUpdate[sun - 2][prec][mu];
}
int update_kernel_regs(int sun, int prec, int mu, int tbc, int prn_kind) {
    if (sun < 2 || sun > 3 || prec < 0 || prec > 1 || mu < 0 || mu > 3) return -1;
    return kRegs[sun - 2][prec][mu](tbc, prn_kind);
}

#define DECL_M(S, P) \
    cudaError_t launch_measure_##S##_##P(const MeasureArgs&, cudaStream_t); \
    cudaError_t launch_polyakov_##S##_##P(const PolyakovArgs&, cudaStream_t); \
    cudaError_t launch_init_cold_##S##_##P(void*, const Geom&, cudaStream_t); \
    cudaError_t launch_init_gid_##S##_##P(void*, const Geom&, cudaStream_t); \
    cudaError_t launch_init_hot_##S##_##P(void*, const Geom&, const void*, int, cudaStream_t); \
    cudaError_t launch_gramschmidt_##S##_##P(void*, const Geom&, cudaStream_t);
DECL_M(2, 0) DECL_M(2, 1) DECL_M(3, 0) DECL_M(3, 1)

#define SWITCH4(call) \
    if (sun == 2 && prec == 0) return call(2, 0); \
    if (sun == 2 && prec == 1) return call(2, 1); \
    if (sun == 3 && prec == 0) return call(3, 0); \
    if (sun == 3 && prec == 1) return call(3, 1); \
    return cudaErrorInvalidValue;

cudaError_t launch_measure(int sun, int prec, const MeasureArgs& a, cudaStream_t st) {
#define C(S, P) launch_measure_##S##_##P(a, st)
    SWITCH4(C)
#undef C
}
cudaError_t launch_polyakov(int sun, int prec, const PolyakovArgs& a, cudaStream_t st) {
#define C(S, P) launch_polyakov_##S##_##P(a, st)
    SWITCH4(C)
#undef C
}
cudaError_t launch_init_cold(int sun, int prec, void* t, const Geom& g, cudaStream_t st) {
#define C(S, P) launch_init_cold_##S##_##P(t, g, st)
    SWITCH4(C)
#undef C
}
cudaError_t launch_init_gid(int sun, int prec, void* t, const Geom& g, cudaStream_t st) {
#define C(S, P) launch_init_gid_##S##_##P(t, g, st)
    SWITCH4(C)
#undef C
}
cudaError_t launch_init_hot(int sun, int prec, void* t, const Geom& g, const void* prns, int dir, cudaStream_t st) {
#define C(S, P) launch_init_hot_##S##_##P(t, g, prns, dir, st)
    SWITCH4(C)
#undef C
}
cudaError_t launch_gramschmidt(int sun, int prec, void* t, const Geom& g, cudaStream_t st) {
#define C(S, P) launch_gramschmidt_##S##_##P(t, g, st)
    SWITCH4(C)
#undef C
}

}  // namespace qcd
