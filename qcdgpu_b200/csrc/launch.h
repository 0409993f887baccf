// launch.h -- host-side launch interface between the C-ABI backend (backend.cu) and the
// separately compiled kernel instantiation units (update_inst.cu is built once per
// (SUN, precision, direction); measure_inst.cu / misc_inst.cu once per (SUN, precision)).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "lattice.cuh"

namespace qcd {

enum PrnKind : int {
    PRN_BUFFERED = 0,   // reference-layout float4 buffer
    PRN_GID      = 1,   // -D GID_UPD closed form
    PRN_RANLUX3  = 2,   // fused, state in registers (RL_skip = 223)
    PRN_XOR128   = 3,
    PRN_PM       = 4,
    PRN_MRG32K3A = 5,
    PRN_RANLUX0  = 6, PRN_RANLUX1 = 7, PRN_RANLUX2 = 8, PRN_RANLUX4 = 9,   // buffered-only luxury levels
    PRN_XOR7     = 10,  // fused or buffered (8 words of state)
    PRN_RANECU   = 11,  // fused or buffered (8 words of state)
    PRN_RANMAR   = 12,  // buffered only (97 float4 of state per instance, in global memory)
};

// CUtensorMap (128 bytes, 64-byte aligned) without pulling cuda.h into this header
struct alignas(64) TensorMapBlob { unsigned char bytes[128]; };

struct UpdateArgs {
    void* table;            // R*
    const void* params;     // R*: [0]=beta/N [1]=phi [2]=omega   (suncl.cpp:3939-3953)
    Geom geom;
    int odd;                // 0 even, 1 odd
    int tbc;
    int prn_kind;
    const void* prns;       // float4* (buffered)
    uint32_t* prn_table;    // generator state (fused)
    uint32_t* prn_table_out;  // RANLUX two-phase kernel: where the advanced state goes (!= prn_table: the state in front of the
                              // pass stays readable for the whole launch; == prn_table or null: in place)
    uint32_t prn_instances;
    uint32_t prn_samples;
    int overrelax;          // 1: overrelaxation pass instead of heat bath (no PRNs)
    const TensorMapBlob* tmap;   // tensor maps over the link table (4 reals, ROWSIZE sites, planes), or null: TMA staging of the staple links;
    const TensorMapBlob* tmap1;  // box = 32 sites of one parity x 3 planes, and 1 site x 3 planes
    uint32_t* ws_counters;       // two zeroed words owned by the handle: ticket and exit counters of the persistent kernel (it re-arms them itself)
    void* split_scratch; size_t split_scratch_bytes;   // QCDGPU_SPLIT=1: staple sums + links between staple_kernel_tma and heatbath_kernel_rlx
};
typedef cudaError_t (*UpdateLaunchFn)(const UpdateArgs&, cudaStream_t);
// [sun-2][prec][mu]
UpdateLaunchFn get_update_launcher(int sun, int prec, int mu);
// RANLUX skip-ahead table on the current device (dispatch.cu, ranlux_jump_host.h); null: none, the kernels step through the pass
const uint32_t* ranlux_jump_table_device(uint64_t samples, uint64_t keep, uint32_t skip, bool from_start = false);

struct MeasureArgs {
    const void* table; const void* params; Geom geom; int tbc;
    double2* out_action; double2* out_plq;      // either may be null
    uint32_t first_site, nsites; uint32_t block;
    // -D FMUNU (fmode 1) / -D F0MU (fmode 2): out_plq also receives six projected partial arrays, `offset` apart
    int fmode, gen1, gen2; uint32_t offset;
    int wilson;                                 // 1: Wilson loop W(R,T) into out_plq[BID].x instead of the plaquettes
    uint32_t min_blocks;                        // the host's grid: every partial slot its reducer will sum must be written (zeros beyond the sites)
    int diff_axis; uint32_t plk;                // diff_axis >= 0: lattice_action_diff_{x,y,z} on the index space padded to PLK (memset gives 0: set -1!)
};
cudaError_t launch_measure(int sun, int prec, const MeasureArgs&, cudaStream_t);   // dispatch.cu

struct PolyakovArgs {
    const void* table; const void* params; Geom geom; int tbc;
    double2* out; uint32_t offset; int pl_level; uint32_t x_first, x_count; uint32_t block;
    uint32_t min_blocks;                        // the host's grid: every partial slot it will sum must be written
    int axis; uint32_t plk;                     // axis >= 0: slice-resolved variant lattice_polyakov_diff_{x,y,z}
};
cudaError_t launch_polyakov(int sun, int prec, const PolyakovArgs&, cudaStream_t);
cudaError_t launch_reduce_double2(const double2* table, uint32_t off, uint32_t size, double2* dst, uint32_t index, cudaStream_t);
cudaError_t launch_clear_double2(double2* p, uint32_t n, cudaStream_t);
cudaError_t launch_reduce_slices(const double2* table, uint32_t size, uint32_t nslices, double2* dst, uint32_t index, uint32_t stride,
                                 uint32_t extra_off, uint32_t extra_size, bool with_extra, cudaStream_t);
cudaError_t launch_reduce_double_x(const double2* table, uint32_t size, double* dst, uint32_t index, cudaStream_t);

cudaError_t launch_init_cold(int sun, int prec, void* table, const Geom&, cudaStream_t);
cudaError_t launch_init_gid(int sun, int prec, void* table, const Geom&, cudaStream_t);
cudaError_t launch_init_hot(int sun, int prec, void* table, const Geom&, const void* prns, int dir, cudaStream_t);
cudaError_t launch_gramschmidt(int sun, int prec, void* table, const Geom&, cudaStream_t);

cudaError_t launch_ranlux_init(const uint32_t* seeds, uint32_t* table, uint32_t instances, cudaStream_t);
cudaError_t launch_halo_signal(unsigned long long* flag_a, unsigned long long* flag_b, unsigned long long value, cudaStream_t);
cudaError_t launch_halo_wait(unsigned long long* flags, unsigned long long value, unsigned long long timeout_clocks, cudaStream_t);
cudaError_t launch_ranmar_init(const uint32_t* seeds, float* table, uint32_t instances, cudaStream_t);    // seeds: uint4[2 x instances], table: float4[98 x instances]
cudaError_t launch_prn_produce(int prn_kind, uint32_t* table, void* prns, uint32_t instances, uint32_t samples, cudaStream_t);
// advance every RANLUX3 instance of a state table by `samples` float4 without producing them (one modular multiplication each)
cudaError_t launch_ranlux_skip(uint32_t* table, uint32_t instances, uint64_t samples, cudaStream_t);

// halo exchange building block (BigLattice): copy nplanes vec4 planes of one x-slice between plane-strided arrays
cudaError_t launch_copy_planes(int prec, void* dst, size_t dst_row, uint32_t xd, int dp0, int dps,
                               const void* src, size_t src_row, uint32_t xs, int sp0, int sps,
                               uint32_t n234, int nplanes, cudaStream_t);

// ---- O(N) scalar branch (on_model.cuh)
struct OnArgs {
    void* table; const void* params; Geom geom; uint32_t skgroup; int odd;
    int prn_kind; const void* prns; uint32_t* prn_table; uint32_t prn_instances, prn_samples;
    double2* acc;            // -D ACC_RATE: per-block acceptance sums (.x even pass, .y odd pass), or null
};
constexpr uint32_t kOnBlock = 128;   // threads per block of every O(N) kernel (= what kernel_get_worksize reports)
cudaError_t launch_on_update(int prec, const OnArgs&, cudaStream_t);
cudaError_t launch_on_init(int prec, void* table, const void* params, const void* prns /* null: cold */, uint32_t sites, cudaStream_t);
cudaError_t launch_on_measure(int prec, const OnArgs&, int what, const uint32_t step[4], double2* out, cudaStream_t);
cudaError_t launch_on_reduce_add(const double2* table, uint32_t size, double2* dst, uint32_t index, cudaStream_t);

int update_kernel_regs(int sun, int prec, int mu, int tbc, int prn_kind);   // registers/thread (cudaFuncGetAttributes)

}  // namespace qcd
