// dispatch.cu -- tables that route (SUN, precision, direction) to the separately compiled
// instantiation units.
#include "launch.h"

namespace qcd {

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
