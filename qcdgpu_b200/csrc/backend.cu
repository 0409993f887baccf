// backend.cu -- the C ABI of include/qcdgpu_b200.h: a CUDA replacement for the reference's
// GPU_CL::GPU class (SU(N)/clinterface/clinterface.{h,cpp}).
//
// Model kept from the reference: per-handle tables of programs / kernels / buffers addressed
// by small integer ids starting at 1 (clinterface.cpp:84-87, 1134, 1298), positional kernel
// arguments set once by kernel_init_buffer / kernel_init_constant*, kernel_run(id).
// New: programs are option strings parsed into a KernelConfig (no OpenCL compilation);
// kernel names are bound to hand-written sm_100a kernels; execution is stream-ordered;
// PRN production can be fused into the update kernel (see qgpu_set_fused_prng).
#include "../../include/qcdgpu_b200.h"
#include "launch.h"
#include "ranlux_jump_host.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <utility>
#include <cuda.h>

using namespace qcd;

namespace {

struct ProgramCfg {
    std::string options;
    int sun = 3, nd = 4;
    uint32_t n1 = 0, n2 = 0, n3 = 0, n4 = 0;
    uint32_t rowsize = 0, prngstep = 0;
    int precision = 0;              // 0 single, 1 double, 2 mixed (suncl.h:73)
    int nhit = 10, nhitpar = 1;
    bool tbc = false, gid_upd = false, biglat = false;
    uint32_t full_sites = 0, left_sites = 0;
    int pl = 0;
    int rl_skip = 223;
    int prng_precision = 0;
    bool fmunu = false, f0mu = false;
    int fmunu_gen1 = 0, fmunu_gen2 = 0;      // -D FMUNU1/2/4 and FMUNU5/6/7 (0: lambda_3 resp. lambda_8)
    uint32_t plk = 0, plk_axis[3] = { 0, 0, 0 };   // -D PLK, PLKx, PLKy, PLKz (slice-resolved kernels)
    bool has_lattice = false;
    // O(N) scalar branch (reference tree O(N)/, options assembled in O(N)/suncl/suncl.cpp:1454-1482)
    bool on_model = false, acc_rate = false, on_sub_scheme = false;
    uint32_t skgroup = 4, n1exact = 0;
};

enum KernelKind {
    K_NONE = 0, K_INIT_COLD, K_INIT_GID, K_INIT_HOT, K_GRAMSCHMIDT, K_UPDATE, K_OVERRELAX, K_CLEAR,
    K_MEASURE, K_REDUCE_MEAS, K_MEASURE_PLQ, K_REDUCE_PLQ, K_POLYAKOV, K_REDUCE_POLYAKOV,
    K_POLYAKOV_DIFF, K_REDUCE_POLYAKOV_DIFF, K_WILSON, K_REDUCE_WILSON, K_ACTION_DIFF, K_REDUCE_ACTION_DIFF,
    K_RANLUX_INIT, K_RANMAR_INIT, K_PRODUCE,
    K_ON_INIT_COLD, K_ON_INIT_HOT, K_ON_UPDATE, K_ON_MEASURE, K_ON_REDUCE
};

struct Arg {
    enum Type { BUF, I32, U32X4, F32, F64 } type;
    int buf;
    int32_t i;
    uint32_t u4[4];
    float f;
    double d;
};

struct Kernel {
    std::string name;
    KernelKind kind = K_NONE;
    int program = 0;
    int dir = 0, odd = 0, prn_kind = 0;
    size_t global[3] = { 1, 1, 1 };
    size_t local0 = 0;
    bool local_given = false;
    int x_first = 0, x_count = -1;   // update kernels: x-range of the launch inside the updated region (-1 = all)
    int x_stride = 1;                // distance between the slices of the range
    struct PeerHalo { void* table = nullptr; uint32_t src_x = 0, dst_x = 0, rowsize = 0; } peer[2];   // qgpu_kernel_set_peer_halo
    std::vector<Arg> args;
    // profiling (GPU_debug.profiling)
    double t_sum = 0, t_sq = 0; int starts = 0;
};

struct Buffer {
    int type = 0;
    size_t size = 0, size_of = 0, bytes = 0;
    size_t capacity = 0;             // allocated bytes (>= bytes): see qgpu_buffer_init
    void* dptr = nullptr;
    const void* host_ptr = nullptr;
    void* shadow = nullptr;          // host memory returned by buffer_map
    size_t map_off = 0, map_len = 0;
    int pending_produce = 0;         // kernel id of a deferred PRN producer writing this buffer
    uint32_t pending_skips = 0;      // produce() calls in front of it whose numbers nobody looked at (RANLUX3: skipped in one jump)
    bool pending_partly_fused = false;   // x-range launches of the current pass already ran that generator in registers
    // RANLUX state table of the two-phase update kernel: a second table of the same size receives the advanced state and the
    // two are swapped when the pass is complete, so the state in front of a pass stays readable during it (update.cuh, PrnSmemT)
    void* twin = nullptr;
    int swap_seed_buf = 0;           // (on the PRN buffer) state table whose swap waits for the pass's last x-range launch
    // link table: tensor map (4 reals, ROWSIZE sites, planes) for the TMA staging of the update kernel, made on first use
    TensorMapBlob tmap, tmap1; uint32_t tmap_rowsize = 0; int tmap_prec = -1; bool tmap_failed = false;
    // link table, folded reunitarisation (qgpu_set_fold_reunitarisation): gs_cover[(odd * 4 + dir) * nx + x] = 1 once slice x was
    // rewritten by a (parity, dir) pass that reunitarised its links twice.  Fully covered: every link already is what
    // lattice_GramSchmidt would make of it, and that launch is skipped.  Any other writer of the table clears it.
    std::vector<unsigned char> gs_cover; uint32_t gs_nx = 0;
    bool alive = false;
};

}  // namespace

struct qgpu {
    int device = 0;
    bool initialized = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr, ev_halo = nullptr;
    unsigned limit_wg = 0;
    size_t max_memory_width = 0;
    qgpu_device_info info;
    bool profiling = false, brief = false, show_stage = false;
    // profiling == 2 ("timeline"): an event pair per launch, recorded without waiting, read back by qgpu_profile_collect -- per-kernel
    // times INSIDE a run of back-to-back launches (the one-by-one mode above waits after every launch and lets the board cool down)
    int profiling_mode = 0;
    struct TimedLaunch { int kid; cudaEvent_t e0, e1; };
    std::vector<TimedLaunch> timeline;
    std::vector<cudaEvent_t> event_pool;
    bool return_errors = false, async = false, fused = false;
    bool fold_gs = true;
    int tile_x = -1;
    std::string last_error;
    std::vector<ProgramCfg> programs;
    std::vector<Kernel> kernels;
    std::vector<Buffer> buffers;
    int active_program = 0;
    uint64_t launches = 0;
    void* split_scratch = nullptr; size_t split_scratch_bytes = 0;   // QCDGPU_SPLIT=1 (experiment): staple sums between the two kernels of a pass
    uint32_t* ws_counters = nullptr;     // ticket / exit counters of the persistent update kernel (update.cuh), zeroed once: the kernel re-arms them
    qgpu() {
        programs.resize(1); kernels.resize(1); buffers.resize(1); memset(&info, 0, sizeof(info));
        if (const char* e = getenv("QCDGPU_FOLD_GS")) fold_gs = atoi(e) != 0;
    }
};

namespace {

int fail(qgpu* h, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (h) h->last_error = buf;
    if (!h || !h->return_errors) {
        // reference behaviour: OpenCL_Check_Error prints, finalizes, exit(-1)  (clinterface.cpp:225-284)
        fprintf(stderr, "qcdgpu_b200 error: %s\n", buf);
        exit(-1);
    }
    return -1;
}
#define CUDA_OK(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(h, "%s failed: %s", #call, cudaGetErrorString(e__)); } while (0)
#define CUDA_OKP(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fail(h, "%s failed: %s", #call, cudaGetErrorString(e__)); return nullptr; } } while (0)

bool valid_buf(qgpu* h, int id) { return id > 0 && id < (int)h->buffers.size() && h->buffers[id].alive; }
bool valid_kernel(qgpu* h, int id) { return id > 0 && id < (int)h->kernels.size(); }

// "-D NAME=VALUE", "-D NAME", "-DNAME=VALUE", "-I path", "-Werror"
void parse_options(const char* opt, ProgramCfg& c) {
    c.options = opt ? opt : "";
    bool nhit_given = false;
    std::vector<std::string> tok;
    { std::string cur; for (const char* p = c.options.c_str();; p++) { if (*p == ' ' || *p == '\t' || *p == 0) { if (!cur.empty()) tok.push_back(cur); cur.clear(); if (!*p) break; } else cur += *p; } }
    for (size_t i = 0; i < tok.size(); i++) {
        std::string d;
        if (tok[i] == "-D" && i + 1 < tok.size()) d = tok[++i];
        else if (tok[i].rfind("-D", 0) == 0 && tok[i].size() > 2) d = tok[i].substr(2);
        else if (tok[i] == "-I" && i + 1 < tok.size()) { i++; continue; }
        else continue;
        std::string name = d, val = "1";
        size_t eq = d.find('=');
        if (eq != std::string::npos) { name = d.substr(0, eq); val = d.substr(eq + 1); }
        unsigned long v = strtoul(val.c_str(), nullptr, 10);
        if (name == "SUN") { c.sun = (int)v; c.has_lattice = true; }
        else if (name == "ND") c.nd = (int)v;
        else if (name == "N1") c.n1 = (uint32_t)v;
        else if (name == "N2") c.n2 = (uint32_t)v;
        else if (name == "N3") c.n3 = (uint32_t)v;
        else if (name == "N4") c.n4 = (uint32_t)v;
        else if (name == "ROWSIZE") c.rowsize = (uint32_t)v;
        else if (name == "PRNGSTEP") c.prngstep = (uint32_t)v;
        else if (name == "PRECISION") c.precision = (int)v;
        else if (name == "NHIT") { c.nhit = (int)v; nhit_given = true; }
        else if (name == "NHITPar") c.nhitpar = (int)v;
        else if (name == "TBC") c.tbc = true;
        else if (name == "GID_UPD") c.gid_upd = true;
        else if (name == "BIGLAT") c.biglat = true;
        else if (name == "FULL_SITES") c.full_sites = (uint32_t)v;
        else if (name == "LEFT_SITES") c.left_sites = (uint32_t)v;
        else if (name == "PL") c.pl = (int)v;
        else if (name == "RL_skip") c.rl_skip = (int)v;
        else if (name == "PRNG_PRECISION") c.prng_precision = (int)v;
        else if (name == "FMUNU") c.fmunu = true;
        else if (name == "F0MU") c.f0mu = true;
        else if (name == "FMUNU1") c.fmunu_gen1 = 1;
        else if (name == "FMUNU2") c.fmunu_gen1 = 2;
        else if (name == "FMUNU4") c.fmunu_gen1 = 4;
        else if (name == "FMUNU5") c.fmunu_gen2 = 5;
        else if (name == "FMUNU6") c.fmunu_gen2 = 6;
        else if (name == "FMUNU7") c.fmunu_gen2 = 7;
        else if (name == "PLK") c.plk = (uint32_t)v;
        else if (name == "PLKx") c.plk_axis[0] = (uint32_t)v;
        else if (name == "PLKy") c.plk_axis[1] = (uint32_t)v;
        else if (name == "PLKz") c.plk_axis[2] = (uint32_t)v;
        else if (name == "ON_MODEL") c.on_model = true;
        else if (name == "ACC_RATE") c.acc_rate = true;
        else if (name == "ON_SUB_SCHEME") c.on_sub_scheme = true;
        else if (name == "SKGROUP") c.skgroup = (uint32_t)v;
        else if (name == "N1EXACT") c.n1exact = (uint32_t)v;
    }
    // The reference's BIGLAT host builds its update program with N1 = Nx + 2 (halo slices included, suncl.cpp:2422) and its measurement,
    // Polyakov and Wilson programs with N1 = Nx (suncl.cpp:2597, 2889; their kernels test gindex < (N1 + 1) N2N3N4,
    // sun_measurements_cl.cl:211), both with -D BIGLAT and the ROWSIZE of the stored part.  The update program is the one that carries
    // -D NHIT; everything below works on the stored extent.
    if (c.biglat && !nhit_given) c.n1 += 2;
}

Geom make_geom(const ProgramCfg& c, int tile_override = -1) {
    Geom g;
    memset(&g, 0, sizeof(g));
    g.n1 = c.n1; g.n2 = c.n2; g.n3 = c.n3; g.n4 = c.n4;
    g.n23 = c.n2 * c.n3; g.n234 = g.n23 * c.n4; g.sites = c.n1 * g.n234;
    g.rowsize = c.rowsize ? c.rowsize : g.sites;                            // model.cl:101-103
    g.biglat = c.biglat ? 1u : 0u;
    g.nthreads = c.biglat ? (c.n1 - 2) * g.n234 / 2 : g.sites / 2;          // suncl.cl:404-418
    g.prngstep = c.prngstep ? c.prngstep : g.sites / 2;                     // model.cl:126-128
    g.left_sites = c.left_sites; g.full_sites = c.full_sites;
    g.nhit = (uint32_t)c.nhit; g.nhitpar = (uint32_t)c.nhitpar;
    g.upd_x0 = 0; g.upd_nx = c.biglat ? c.n1 - 2 : c.n1; g.upd_xstride = 1;
    if (c.biglat && g.n234) { g.tw_lo = c.left_sites / g.n234; g.tw_hi = c.full_sites / g.n234; }
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("QCDGPU_DEBUG_SKIP"); dbg = e ? atoi(e) : 0; } g.debug_skip = (uint32_t)dbg; }
    // L2 tiling (lattice.cuh remap_gid): worthwhile once three x-slices of all links no longer fit the L2
    {
        static int env_tile = -2;
        if (env_tile == -2) { const char* e = getenv("QCDGPU_TILE_X"); env_tile = e ? atoi(e) : -1; }
        const uint32_t nx = c.biglat ? c.n1 - 2 : c.n1;
        const size_t slice_bytes = (size_t)g.n234 * 4 * (c.sun == 3 ? 3 : 1) * (c.precision == 1 ? 32 : 16);
        uint32_t tile = 0;
        if (tile_override >= 0) tile = (uint32_t)tile_override;
        else if (env_tile >= 0) tile = (uint32_t)env_tile;
        else if (3 * slice_bytes > (size_t)48 << 20) tile = 8;
        if (tile >= nx || (c.n2 & 1)) tile = 0;
        (void)nx;
        g.tile_x = tile;
    }
    return g;
}
int prec_index(const ProgramCfg& c) { return c.precision == 1 ? 1 : 0; }   // mixed == single storage/arithmetic (complex.h:61-91)

struct NameMap { const char* name; KernelKind kind; int dir; int odd; int prn; };
const NameMap kNames[] = {
    { "lattice_init_cold", K_INIT_COLD, 0, 0, 0 }, { "lattice_init_gid", K_INIT_GID, 0, 0, 0 },
    { "lattice_init_hot_X", K_INIT_HOT, 0, 0, 0 }, { "lattice_init_hot_Y", K_INIT_HOT, 1, 0, 0 },
    { "lattice_init_hot_Z", K_INIT_HOT, 2, 0, 0 }, { "lattice_init_hot_T", K_INIT_HOT, 3, 0, 0 },
    { "lattice_GramSchmidt", K_GRAMSCHMIDT, 0, 0, 0 },
    { "update_even_X", K_UPDATE, 0, 0, 0 }, { "update_even_Y", K_UPDATE, 1, 0, 0 }, { "update_even_Z", K_UPDATE, 2, 0, 0 }, { "update_even_T", K_UPDATE, 3, 0, 0 },
    { "update_odd_X", K_UPDATE, 0, 1, 0 }, { "update_odd_Y", K_UPDATE, 1, 1, 0 }, { "update_odd_Z", K_UPDATE, 2, 1, 0 }, { "update_odd_T", K_UPDATE, 3, 1, 0 },
    // extension (north_star: overrelaxation; not in the reference)
    { "overrelax_even_X", K_OVERRELAX, 0, 0, 0 }, { "overrelax_even_Y", K_OVERRELAX, 1, 0, 0 }, { "overrelax_even_Z", K_OVERRELAX, 2, 0, 0 }, { "overrelax_even_T", K_OVERRELAX, 3, 0, 0 },
    { "overrelax_odd_X", K_OVERRELAX, 0, 1, 0 }, { "overrelax_odd_Y", K_OVERRELAX, 1, 1, 0 }, { "overrelax_odd_Z", K_OVERRELAX, 2, 1, 0 }, { "overrelax_odd_T", K_OVERRELAX, 3, 1, 0 },
    { "clear_measurement", K_CLEAR, 0, 0, 0 },
    { "lattice_measurement", K_MEASURE, 0, 0, 0 }, { "reduce_measurement_double2", K_REDUCE_MEAS, 0, 0, 0 },
    { "lattice_measurement_plq", K_MEASURE_PLQ, 0, 0, 0 }, { "reduce_measurement_plq_double2", K_REDUCE_PLQ, 0, 0, 0 },
    { "lattice_polyakov", K_POLYAKOV, 0, 0, 0 }, { "reduce_polyakov_double2", K_REDUCE_POLYAKOV, 0, 0, 0 },
    { "lattice_polyakov_diff_x", K_POLYAKOV_DIFF, 0, 0, 0 }, { "lattice_polyakov_diff_y", K_POLYAKOV_DIFF, 1, 0, 0 }, { "lattice_polyakov_diff_z", K_POLYAKOV_DIFF, 2, 0, 0 },
    { "reduce_polyakov_diff_x_double2", K_REDUCE_POLYAKOV_DIFF, 0, 0, 0 }, { "reduce_polyakov_diff_y_double2", K_REDUCE_POLYAKOV_DIFF, 1, 0, 0 },
    { "reduce_polyakov_diff_z_double2", K_REDUCE_POLYAKOV_DIFF, 2, 0, 0 },
    { "lattice_measurement_wilson", K_WILSON, 0, 0, 0 }, { "reduce_wilson_double2", K_REDUCE_WILSON, 0, 0, 0 },
    { "lattice_action_diff_x", K_ACTION_DIFF, 0, 0, 0 }, { "lattice_action_diff_y", K_ACTION_DIFF, 1, 0, 0 }, { "lattice_action_diff_z", K_ACTION_DIFF, 2, 0, 0 },
    { "reduce_action_diff_x_double2", K_REDUCE_ACTION_DIFF, 0, 0, 0 }, { "reduce_action_diff_y_double2", K_REDUCE_ACTION_DIFF, 1, 0, 0 },
    { "reduce_action_diff_z_double2", K_REDUCE_ACTION_DIFF, 2, 0, 0 },
    { "ranlux_init", K_RANLUX_INIT, 0, 0, 0 },
    { "ranlux", K_PRODUCE, 0, 0, -1 }, { "xor128", K_PRODUCE, 0, 0, PRN_XOR128 }, { "pm", K_PRODUCE, 0, 0, PRN_PM }, { "mrg32k3a", K_PRODUCE, 0, 0, PRN_MRG32K3A },
    { "xor7", K_PRODUCE, 0, 0, PRN_XOR7 }, { "ranecu", K_PRODUCE, 0, 0, PRN_RANECU },
    // RANMAR: the kernels are ranmar_init / ranmar in prngcl_ranmar.cl:65,159, the reference's host asks for "rmseed" / "rmproduce" (random.cpp:757,763)
    { "ranmar_init", K_RANMAR_INIT, 0, 0, 0 }, { "rmseed", K_RANMAR_INIT, 0, 0, 0 },
    { "ranmar", K_PRODUCE, 0, 0, PRN_RANMAR }, { "rmproduce", K_PRODUCE, 0, 0, PRN_RANMAR },
};

// kernel names of a program created with -D ON_MODEL (O(N)/oncl/oncl.cl, on_measurements_cl.cl); dir = what is measured
const NameMap kNamesOn[] = {
    { "lattice_init_cold", K_ON_INIT_COLD, 0, 0, 0 }, { "lattice_init_hot", K_ON_INIT_HOT, 0, 0, 0 },
    { "update_even", K_ON_UPDATE, 0, 0, 0 }, { "update_odd", K_ON_UPDATE, 0, 1, 0 },
    { "clear_measurement", K_CLEAR, 0, 0, 0 },
    { "lattice_measurement", K_ON_MEASURE, 0, 0, 0 }, { "lattice_measurement_plq", K_ON_MEASURE, 1, 0, 0 },
    { "lattice_measurement_correlator", K_ON_MEASURE, 2, 0, 0 },
    { "reduce_measurement_double2", K_ON_REDUCE, 0, 0, 0 }, { "reduce_measurement_plq_double2", K_ON_REDUCE, 0, 0, 0 },
    { "reduce_acceptance_rate", K_ON_REDUCE, 0, 0, 0 },
};

int ranlux_kind(int skip) {
    switch (skip) { case 24: return PRN_RANLUX0; case 48: return PRN_RANLUX1; case 97: return PRN_RANLUX2; case 223: return PRN_RANLUX3; case 389: return PRN_RANLUX4; }
    return -1;
}

void* buf_ptr(qgpu* h, const Kernel& k, size_t i) {
    if (i >= k.args.size() || k.args[i].type != Arg::BUF) return nullptr;
    int b = k.args[i].buf;
    if (!valid_buf(h, b)) return nullptr;
    return h->buffers[b].dptr;
}
int buf_id(const Kernel& k, size_t i) { return (i < k.args.size() && k.args[i].type == Arg::BUF) ? k.args[i].buf : 0; }
bool arg_int(const Kernel& k, size_t i, int* out) {
    if (i >= k.args.size() || k.args[i].type != Arg::I32) return false;
    *out = k.args[i].i; return true;
}

int launch_kernel(qgpu* h, int kid);

// the x-range launches of a pass have all been issued: drop the deferred producer and publish the advanced RANLUX states
void finish_partial_pass(qgpu* h, int pb) {
    Buffer& b = h->buffers[pb];
    b.pending_produce = 0; b.pending_partly_fused = false;
    if (b.swap_seed_buf && valid_buf(h, b.swap_seed_buf) && h->buffers[b.swap_seed_buf].twin)
        std::swap(h->buffers[b.swap_seed_buf].dptr, h->buffers[b.swap_seed_buf].twin);
    b.swap_seed_buf = 0;
}

// the host is about to look at buffer `bid`: if it is a RANLUX state table whose swap waits for the end of a pass, publish it now
void settle_state_table(qgpu* h, int bid) {
    for (size_t i = 1; i < h->buffers.size(); i++)
        if (h->buffers[i].alive && h->buffers[i].swap_seed_buf == bid) finish_partial_pass(h, (int)i);
}

// A RANLUX state table with a twin (Buffer::twin) was written outside the two-phase kernel (host write, unmap, the stand-alone
// producer advancing it in place): bring the twin up to date, because the two-phase kernel writes the twin only for the instances its
// update threads cover (padding instances and BIGLAT halo slices keep what the twin held) and the tables are swapped after each pass.
int resync_twin(qgpu* h, int bid) {
    if (!valid_buf(h, bid)) return 0;
    Buffer& b = h->buffers[bid];
    if (!b.twin || !b.dptr) return 0;
    CUDA_OK(h, cudaMemcpyAsync(b.twin, b.dptr, b.capacity, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}
// produce() calls that were overtaken by the next one without anybody reading their numbers (a resumed run replays PRNG_counter of
// them back to back, suncl.cpp:5350): advance the generators past them in one jump per instance (ranlux_skip_kernel)
int apply_pending_skips(qgpu* h, int bid) {
    Buffer& b = h->buffers[bid];
    if (!b.pending_skips || !b.pending_produce) { b.pending_skips = 0; return 0; }
    Kernel& k = h->kernels[b.pending_produce];
    int samples = 0;
    if (!arg_int(k, 2, &samples) || !buf_ptr(h, k, 0)) return fail(h, "kernel %s: arguments (seed table, samples) not set", k.name.c_str());
    const uint64_t total = (uint64_t)b.pending_skips * (uint64_t)samples;
    b.pending_skips = 0;
    CUDA_OK(h, launch_ranlux_skip((uint32_t*)buf_ptr(h, k, 0), (uint32_t)k.global[0], total, h->stream));
    h->launches++;
    return resync_twin(h, buf_id(k, 0));
}
// run a deferred PRN producer now (writes the reference-layout buffer)
int flush_pending(qgpu* h, int bid) {
    if (!valid_buf(h, bid)) return 0;
    int kid = h->buffers[bid].pending_produce;
    if (!kid) return 0;
    if (apply_pending_skips(h, bid) < 0) return -1;
    if (h->buffers[bid].pending_partly_fused)
        return fail(h, "PRN buffer %d is needed in memory, but x-range launches of this pass already consumed its generator in registers", bid);
    h->buffers[bid].pending_produce = 0;
    Kernel& k = h->kernels[kid];
    int samples = 0;
    if (!arg_int(k, 2, &samples)) return fail(h, "kernel %s: argument 2 (samples) not set", k.name.c_str());
    CUDA_OK(h, launch_prn_produce(k.prn_kind, (uint32_t*)buf_ptr(h, k, 0), buf_ptr(h, k, 1), (uint32_t)k.global[0], (uint32_t)samples, h->stream));
    h->launches++;
    return resync_twin(h, buf_id(k, 0));
}

// Tensor map over a link table for cp.async.bulk.tensor: dims (4 reals, ROWSIZE, planes), traversal strides (1, 2, 4), box 4 x 64 x 9,
// i.e. the three vec4 planes of one link direction for 32 sites of one parity of a y-row per box.  Null when the driver entry point is missing or the table does not qualify.
const TensorMapBlob* link_tensor_map(qgpu* h, int bid, const ProgramCfg& c, const Geom& g, const TensorMapBlob** one_site) {
    *one_site = nullptr;
    if (!valid_buf(h, bid)) return nullptr;
    Buffer& b = h->buffers[bid];
    const int prec = prec_index(c);
    if (b.tmap_failed) return nullptr;
    if (b.tmap_rowsize == g.rowsize && b.tmap_prec == prec) { *one_site = &b.tmap1; return &b.tmap; }
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Enc enc = nullptr; static bool looked = false;
    if (!looked) {
        looked = true;
        void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) enc = (Enc)fn;
        (void)cudaGetLastError();
    }
    const size_t rb = prec ? 8 : 4;
    const uint64_t planes = (c.sun == 3) ? 12 : 4;
    if (!enc || ((uintptr_t)b.dptr & 15u) || b.bytes < (size_t)g.rowsize * planes * 4 * rb) { b.tmap_failed = true; return nullptr; }
    static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "CUtensorMap is 128 bytes");
    cuuint64_t gdim[3] = { 4, g.rowsize, planes }, gstr[2] = { 4 * rb, (cuuint64_t)g.rowsize * 4 * rb };
    cuuint32_t box[3] = { 4, 64, 9 }, estr[3] = { 1, 2, 4 };      // 32 sites of one parity x planes DIR, DIR + 4, DIR + 8
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(&b.tmap), prec ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, b.dptr,
                     gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b.tmap_failed = true; return nullptr; }
    cuuint32_t box1[3] = { 4, 2, 9 };                              // the same for a single site
    r = enc(reinterpret_cast<CUtensorMap*>(&b.tmap1), prec ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, b.dptr,
            gdim, gstr, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b.tmap_failed = true; return nullptr; }
    b.tmap_rowsize = g.rowsize; b.tmap_prec = prec;
    *one_site = &b.tmap1;
    return &b.tmap;
}

void gs_cover_clear(qgpu* h, int bid) { if (valid_buf(h, bid)) h->buffers[bid].gs_cover.clear(); }
void gs_cover_mark(qgpu* h, int bid, int odd, int dir, uint32_t nx, uint32_t x0, uint32_t count) {
    if (!valid_buf(h, bid)) return;
    Buffer& b = h->buffers[bid];
    if (b.gs_nx != nx || b.gs_cover.size() != (size_t)8 * nx) { b.gs_cover.assign((size_t)8 * nx, 0); b.gs_nx = nx; }
    for (uint32_t x = x0; x < x0 + count && x < nx; x++) b.gs_cover[(size_t)(odd * 4 + dir) * nx + x] = 1;
}
bool gs_cover_complete(qgpu* h, int bid, uint32_t nx) {
    if (!valid_buf(h, bid)) return false;
    const Buffer& b = h->buffers[bid];
    if (nx == 0 || b.gs_nx != nx || b.gs_cover.size() != (size_t)8 * nx) return false;
    for (unsigned char c : b.gs_cover) if (!c) return false;
    return true;
}

// reduce_*_double2: partials [off, off + size) of buffer `src` summed into element `index` of buffer `dst`, both ranges checked
// against what is allocated (a wrong host-side size or index must be a loud error, not an out-of-bounds access)
int checked_reduce(qgpu* h, const Kernel& k, int src, int dst, size_t off, size_t size, size_t index, cudaStream_t st) {
    if (!valid_buf(h, src) || !valid_buf(h, dst) || !h->buffers[src].dptr || !h->buffers[dst].dptr)
        return fail(h, "kernel %s: buffers not set", k.name.c_str());
    if ((off + size) * 16 > h->buffers[src].capacity || (index + 1) * 16 > h->buffers[dst].capacity)
        return fail(h, "kernel %s: partials [%zu, %zu) or result slot %zu outside the buffers", k.name.c_str(), off, off + size, index);
    CUDA_OK(h, launch_reduce_double2((const double2*)h->buffers[src].dptr, (uint32_t)off, (uint32_t)size, (double2*)h->buffers[dst].dptr, (uint32_t)index, st));
    return 0;
}

int launch_kernel(qgpu* h, int kid) {
    Kernel& k = h->kernels[kid];
    const ProgramCfg& c = h->programs[k.program];
    cudaStream_t st = h->stream;
    const int prec = prec_index(c);
    switch (k.kind) {
        case K_INIT_COLD: case K_INIT_GID: {
            if (!buf_ptr(h, k, 0)) return fail(h, "kernel %s: argument 0 (lattice_table) not set", k.name.c_str());
            Geom g = make_geom(c);
            gs_cover_clear(h, buf_id(k, 0));
            CUDA_OK(h, k.kind == K_INIT_COLD ? launch_init_cold(c.sun, prec, buf_ptr(h, k, 0), g, st) : launch_init_gid(c.sun, prec, buf_ptr(h, k, 0), g, st));
            break;
        }
        case K_INIT_HOT: {
            if (!buf_ptr(h, k, 0) || !buf_ptr(h, k, 1)) return fail(h, "kernel %s: arguments (lattice_table, prns) not set", k.name.c_str());
            if (flush_pending(h, buf_id(k, 1)) < 0) return -1;
            Geom g = make_geom(c);
            size_t need = (c.sun == 3 ? 2 * (size_t)g.rowsize : 0) + g.sites;
            if (h->buffers[buf_id(k, 1)].bytes < need * 16) return fail(h, "kernel %s: PRN buffer too small for a hot start (%zu < %zu vec4)", k.name.c_str(), h->buffers[buf_id(k, 1)].bytes / 16, need);
            gs_cover_clear(h, buf_id(k, 0));
            CUDA_OK(h, launch_init_hot(c.sun, prec, buf_ptr(h, k, 0), g, buf_ptr(h, k, 1), k.dir, st));
            break;
        }
        case K_GRAMSCHMIDT: {
            if (!buf_ptr(h, k, 0)) return fail(h, "kernel %s: argument 0 (lattice_table) not set", k.name.c_str());
            {   // every link rewritten since the last call by passes that already reunitarised twice: nothing left to do
                const uint32_t nx = c.biglat ? c.n1 - 2 : c.n1;
                const bool done = h->fold_gs && c.sun == 3 && gs_cover_complete(h, buf_id(k, 0), nx);
                gs_cover_clear(h, buf_id(k, 0));
                if (done) return 0;      // not a launch
            }
            CUDA_OK(h, launch_gramschmidt(c.sun, prec, buf_ptr(h, k, 0), make_geom(c), st));
            break;
        }
        case K_UPDATE: case K_OVERRELAX: {
            UpdateArgs a;
            memset(&a, 0, sizeof(a));
            a.table = buf_ptr(h, k, 0); a.params = buf_ptr(h, k, 1);
            if (!a.table || !a.params) return fail(h, "kernel %s: arguments (lattice_table, lattice_parameters) not set", k.name.c_str());
            a.geom = make_geom(c, h->tile_x); a.odd = k.odd; a.tbc = c.tbc ? 1 : 0;
            a.geom.fold_gs = (h->fold_gs && c.sun == 3) ? 1u : 0u;
            for (int sd = 0; sd < 2; sd++) {     // halo exchange fused into the update: boundary-slice links also go to the neighbours' halo slices
                if (!k.peer[sd].table) continue;
                a.geom.peer_on = 1u;
                a.geom.peer_table[sd] = (unsigned long long)(uintptr_t)k.peer[sd].table;
                a.geom.peer_lo[sd] = k.peer[sd].src_x * a.geom.n234;
                a.geom.peer_shift[sd] = (k.peer[sd].dst_x - k.peer[sd].src_x) * a.geom.n234;
                a.geom.peer_rowsize[sd] = k.peer[sd].rowsize;
            }
            const uint32_t nx_all = a.geom.upd_nx;
            if (k.x_count >= 0) {
                if (k.x_stride < 1 || (k.x_count > 0 && (uint64_t)k.x_first + (uint64_t)(k.x_count - 1) * k.x_stride >= a.geom.upd_nx) || k.x_first < 0)
                    return fail(h, "kernel %s: x-slices %d + i * %d, i < %d, outside the %u updated slices", k.name.c_str(), k.x_first, k.x_stride, k.x_count, a.geom.upd_nx);
                if (k.x_stride == 1 && (uint32_t)(k.x_first + k.x_count) > a.geom.upd_nx) return fail(h, "kernel %s: x-range [%d,+%d) outside the %u updated slices", k.name.c_str(), k.x_first, k.x_count, a.geom.upd_nx);
                a.geom.upd_x0 = (uint32_t)k.x_first; a.geom.upd_nx = (uint32_t)k.x_count;
                a.geom.nthreads = a.geom.upd_nx * (a.geom.n234 / 2);
                a.geom.upd_xstride = (uint32_t)k.x_stride;
                if (a.geom.tile_x >= a.geom.upd_nx || k.x_stride != 1) a.geom.tile_x = 0;
                if (a.geom.nthreads == 0) break;
            }
            if (a.geom.fold_gs) {
                if (a.geom.upd_xstride == 1u) gs_cover_mark(h, buf_id(k, 0), k.odd, k.dir, nx_all, a.geom.upd_x0, a.geom.upd_nx);
                else for (uint32_t i = 0; i < a.geom.upd_nx; i++) gs_cover_mark(h, buf_id(k, 0), k.odd, k.dir, nx_all, a.geom.upd_x0 + i * a.geom.upd_xstride, 1u);
            }
            const uint32_t full_threads = (c.biglat ? c.n1 - 2 : c.n1) * (a.geom.n234 / 2);
            UpdateLaunchFn fn = get_update_launcher(c.sun, prec, k.dir);
            if (!fn) return fail(h, "kernel %s: no CUDA instantiation for SUN=%d PRECISION=%d", k.name.c_str(), c.sun, c.precision);
            if (k.kind == K_OVERRELAX) {
                a.overrelax = 1;
                if (c.sun == 3) a.tmap = link_tensor_map(h, buf_id(k, 0), c, a.geom, &a.tmap1);
                CUDA_OK(h, fn(a, st));
                break;
            }
            if (c.gid_upd) { a.prn_kind = PRN_GID; CUDA_OK(h, fn(a, st)); break; }
            int pb = buf_id(k, 2);
            if (!valid_buf(h, pb)) return fail(h, "kernel %s: argument 2 (prns) not set", k.name.c_str());
            if (c.prng_precision == 2) return fail(h, "kernel %s: PRNG_PRECISION=2 (double PRNs) is not implemented", k.name.c_str());
            Buffer& pbuf = h->buffers[pb];
            bool fused = false;
            int swap_seed = 0;       // full-range launch of the two-phase RANLUX kernel: swap the state tables right after it
            if (pbuf.pending_produce) {
                Kernel& pk = h->kernels[pbuf.pending_produce];
                int samples = 0; arg_int(pk, 2, &samples);
                bool fusable = h->fused && (uint32_t)pk.global[0] == a.geom.prngstep && full_threads <= a.geom.prngstep &&
                               (pk.prn_kind == PRN_RANLUX3 || pk.prn_kind == PRN_XOR128 || pk.prn_kind == PRN_PM || pk.prn_kind == PRN_MRG32K3A ||
                                pk.prn_kind == PRN_XOR7 || pk.prn_kind == PRN_RANECU) &&
                               (uint32_t)samples >= 3u * (uint32_t)(c.nhit + 1) * (uint32_t)c.nhitpar;
                if (fusable) {
                    if (apply_pending_skips(h, pb) < 0) return -1;
                    a.prn_kind = pk.prn_kind; a.prn_table = (uint32_t*)buf_ptr(h, pk, 0);
                    a.prn_instances = (uint32_t)pk.global[0]; a.prn_samples = (uint32_t)samples;
                    if (pk.prn_kind == PRN_RANLUX3) {      // second state table (see Buffer::twin); without it the kernel works in place
                        if (c.sun == 3) a.tmap = link_tensor_map(h, buf_id(k, 0), c, a.geom, &a.tmap1);
                        Buffer& sb = h->buffers[buf_id(pk, 0)];
                        if (!sb.twin && cudaMalloc(&sb.twin, sb.capacity) == cudaSuccess)
                            CUDA_OK(h, cudaMemcpyAsync(sb.twin, sb.dptr, sb.capacity, cudaMemcpyDeviceToDevice, st));
                        else if (!sb.twin) { sb.twin = nullptr; (void)cudaGetLastError(); }
                        a.prn_table_out = (uint32_t*)sb.twin;
                        swap_seed = sb.twin ? buf_id(pk, 0) : 0;
                        if (!h->ws_counters) {
                            if (cudaMalloc(&h->ws_counters, 2 * sizeof(uint32_t)) == cudaSuccess)
                                CUDA_OK(h, cudaMemsetAsync(h->ws_counters, 0, 2 * sizeof(uint32_t), st));
                            else { h->ws_counters = nullptr; (void)cudaGetLastError(); }
                        }
                        a.ws_counters = h->ws_counters;
                        {   // QCDGPU_SPLIT=1 (experiment): scratch for the staple sums and links of this launch
                            static int split_env = -1;
                            if (split_env < 0) { const char* e = getenv("QCDGPU_SPLIT"); split_env = e ? atoi(e) : 0; }
                            if (split_env) {
                                const size_t need = (size_t)a.geom.nthreads * 9 * 2 * (prec ? 8 : 4);
                                if (h->split_scratch_bytes < need) {
                                    if (h->split_scratch) cudaFree(h->split_scratch);
                                    h->split_scratch = nullptr; h->split_scratch_bytes = 0;
                                    if (cudaMalloc(&h->split_scratch, need) == cudaSuccess) h->split_scratch_bytes = need; else (void)cudaGetLastError();
                                }
                                a.split_scratch = h->split_scratch; a.split_scratch_bytes = h->split_scratch_bytes;
                            }
                        }
                    }
                    if (k.x_count < 0) { pbuf.pending_produce = 0; pbuf.pending_partly_fused = false; }
                    else { pbuf.pending_partly_fused = true; pbuf.swap_seed_buf = swap_seed; swap_seed = 0; }   // the other x-ranges of this pass follow
                    fused = true;
                } else if (flush_pending(h, pb) < 0) return -1;
            }
            if (!fused) {
                a.prn_kind = PRN_BUFFERED; a.prns = pbuf.dptr;
                size_t need = (size_t)a.geom.prngstep * (3u * (c.nhit + 1) * c.nhitpar - 1) + full_threads;
                if (pbuf.bytes < need * 16) return fail(h, "kernel %s: PRN buffer holds %zu vec4, the pass may read %zu", k.name.c_str(), pbuf.bytes / 16, need);
            }
            CUDA_OK(h, fn(a, st));
            if (swap_seed) std::swap(h->buffers[swap_seed].dptr, h->buffers[swap_seed].twin);
            break;
        }
        case K_CLEAR: {
            int b = buf_id(k, 0);
            if (!valid_buf(h, b)) return fail(h, "kernel %s: argument 0 not set", k.name.c_str());
            size_t n = k.global[0];
            if (n * 16 > h->buffers[b].bytes) n = h->buffers[b].bytes / 16;
            CUDA_OK(h, launch_clear_double2((double2*)h->buffers[b].dptr, (uint32_t)n, st));
            break;
        }
        case K_MEASURE: case K_MEASURE_PLQ: {
            MeasureArgs a;
            memset(&a, 0, sizeof(a));
            a.diff_axis = -1;
            a.table = buf_ptr(h, k, 0); a.params = buf_ptr(h, k, 2);
            double2* out = (double2*)buf_ptr(h, k, 1);
            if (!a.table || !a.params || !out) return fail(h, "kernel %s: arguments (table, measurement, parameters) not set", k.name.c_str());
            a.geom = make_geom(c); a.tbc = c.tbc ? 1 : 0;
            if (k.kind == K_MEASURE) a.out_action = out; else a.out_plq = out;
            a.first_site = c.biglat ? a.geom.n234 : 0;
            a.nsites = c.biglat ? (c.n1 - 2) * a.geom.n234 : a.geom.sites;
            a.block = (uint32_t)k.local0;
            size_t blocks = (a.nsites + a.block - 1) / a.block;
            // The host sizes its reducer from its own (aligned) grid, like the reference: write every slot it will sum.
            a.min_blocks = (uint32_t)((k.global[0] + a.block - 1) / a.block);
            if (blocks < a.min_blocks) blocks = a.min_blocks;
            size_t slots = blocks;
            if (k.kind == K_MEASURE_PLQ && (c.fmunu || c.f0mu)) {      // only the plaquette kernel has Fmunu variants (sun_measurements_cl.cl:710-1175)
                int offset = 0;
                if (c.biglat) return fail(h, "kernel %s: FMUNU/F0MU with BIGLAT is not implemented in qcdgpu_b200", k.name.c_str());
                if (!arg_int(k, 4, &offset) || offset < (int)blocks) return fail(h, "kernel %s: argument 4 (offset) not set or smaller than the partial count", k.name.c_str());
                a.fmode = c.fmunu ? 1 : 2; a.offset = (uint32_t)offset;
                if (c.sun == 3) { a.gen1 = c.fmunu_gen1 ? c.fmunu_gen1 : 3; a.gen2 = c.fmunu_gen2 ? c.fmunu_gen2 : 8; }
                else            { a.gen1 = (c.fmunu_gen1 == 1) ? 1 : 3; a.gen2 = 2; }
                slots = 6 * (size_t)offset + blocks;
            }
            if (h->buffers[buf_id(k, 1)].capacity < slots * 16) return fail(h, "kernel %s: measurement buffer too small (%zu partials needed)", k.name.c_str(), slots);
            CUDA_OK(h, launch_measure(c.sun, prec, a, st));
            break;
        }
        case K_REDUCE_MEAS: {
            int size = 0, index = 0;
            if (!arg_int(k, 3, &size) || !arg_int(k, 4, &index)) return fail(h, "kernel %s: arguments (size, index) not set", k.name.c_str());
            if (size < 0 || index < 0) return fail(h, "kernel %s: negative size or index", k.name.c_str());
            if (checked_reduce(h, k, buf_id(k, 0), buf_id(k, 1), 0, (size_t)size, (size_t)index, st) < 0) return -1;
            break;
        }
        case K_REDUCE_PLQ: {
            int index = 0;
            if (k.args.size() < 5 || k.args[3].type != Arg::U32X4 || !arg_int(k, 4, &index)) return fail(h, "kernel %s: arguments (param, index) not set", k.name.c_str());
            if (index < 0) return fail(h, "kernel %s: negative index", k.name.c_str());
            if (checked_reduce(h, k, buf_id(k, 0), buf_id(k, 1), 0, k.args[3].u4[0], (size_t)index, st) < 0) return -1;
            if (c.fmunu || c.f0mu) {   // six projected sums: partials at q*offset, results at index + q*offset2  (sun_measurements_cl.cl:1322-1357)
                const uint32_t size = k.args[3].u4[0], offset = k.args[3].u4[1], offset2 = k.args[3].u4[2];
                if (h->buffers[buf_id(k, 1)].capacity < ((size_t)index + 6 * (size_t)offset2 + 1) * 16) return fail(h, "kernel %s: result buffer too small for the Fmunu sums", k.name.c_str());
                for (uint32_t q = 1; q <= 6; q++) {
                    if (checked_reduce(h, k, buf_id(k, 0), buf_id(k, 1), (size_t)q * offset, size, (size_t)index + (size_t)q * offset2, st) < 0) return -1;
                    h->launches++;
                }
            }
            break;
        }
        case K_WILSON: {
            MeasureArgs a;
            memset(&a, 0, sizeof(a));
            a.diff_axis = -1;
            a.table = buf_ptr(h, k, 0); a.params = buf_ptr(h, k, 2);
            a.out_plq = (double2*)buf_ptr(h, k, 1);
            if (!a.table || !a.params || !a.out_plq) return fail(h, "kernel %s: arguments (table, measurement, parameters) not set", k.name.c_str());
            if (c.biglat) return fail(h, "kernel %s: the BIGLAT Wilson-loop kernels are not implemented in qcdgpu_b200", k.name.c_str());
            if (h->buffers[buf_id(k, 2)].bytes < 5 * (size_t)(prec ? 8 : 4)) return fail(h, "kernel %s: lattice_parameters must hold wilson_R, wilson_T at [3], [4]", k.name.c_str());
            a.geom = make_geom(c); a.tbc = c.tbc ? 1 : 0; a.wilson = 1;
            a.first_site = 0; a.nsites = a.geom.sites; a.block = (uint32_t)k.local0;
            size_t blocks = (a.nsites + a.block - 1) / a.block;
            a.min_blocks = (uint32_t)((k.global[0] + a.block - 1) / a.block);
            if (blocks < a.min_blocks) blocks = a.min_blocks;
            if (h->buffers[buf_id(k, 1)].capacity < blocks * 16) return fail(h, "kernel %s: measurement buffer too small (%zu partials needed)", k.name.c_str(), blocks);
            CUDA_OK(h, launch_measure(c.sun, prec, a, st));
            break;
        }
        case K_REDUCE_WILSON: {
            int size = 0, index = 0;
            if (!arg_int(k, 3, &size) || !arg_int(k, 4, &index)) return fail(h, "kernel %s: arguments (size, index) not set", k.name.c_str());
            if (!buf_ptr(h, k, 0) || !buf_ptr(h, k, 1)) return fail(h, "kernel %s: buffers not set", k.name.c_str());
            if (h->buffers[buf_id(k, 1)].capacity < ((size_t)index + 1) * 8) return fail(h, "kernel %s: result buffer too small", k.name.c_str());
            CUDA_OK(h, launch_reduce_double_x((const double2*)buf_ptr(h, k, 0), (uint32_t)size, (double*)buf_ptr(h, k, 1), (uint32_t)index, st));
            break;
        }
        // slice-resolved action (sun_measurements_cl.cl:186-706; no init.dat key of the reference switches it on: model::get_actions_diff)
        case K_ACTION_DIFF: {
            MeasureArgs a;
            memset(&a, 0, sizeof(a));
            a.table = buf_ptr(h, k, 0); a.out_action = (double2*)buf_ptr(h, k, 1); a.params = buf_ptr(h, k, 2);
            if (!a.table || !a.out_action || !a.params) return fail(h, "kernel %s: arguments (table, measurement, parameters) not set", k.name.c_str());
            if (c.biglat) return fail(h, "kernel %s: the BIGLAT slice-resolved action kernels are not implemented in qcdgpu_b200", k.name.c_str());
            a.geom = make_geom(c); a.tbc = c.tbc ? 1 : 0; a.diff_axis = k.dir;
            a.plk = c.plk_axis[k.dir] ? c.plk_axis[k.dir] : c.plk;
            if (!a.plk) return fail(h, "kernel %s: the program was created without -D PLK", k.name.c_str());
            a.block = (uint32_t)k.local0;
            if (a.block == 0 || a.block > 256) return fail(h, "kernel %s: work-group size %u not supported (1..256)", k.name.c_str(), a.block);
            a.min_blocks = (uint32_t)((k.global[0] + a.block - 1) / a.block);
            {   // every site must have a thread: slices of N_a x PLK x N4 (x, y) or N1 x PLK x N4 (z) padded sites
                const uint64_t per = (uint64_t)(k.dir == 0 ? c.n2 : c.n1) * a.plk * c.n4, nsl = (k.dir == 0) ? c.n1 : (k.dir == 1) ? c.n2 : c.n3;
                const uint32_t padded = (k.dir == 2) ? c.n2 : c.n3;
                if (a.plk < padded || (uint64_t)a.min_blocks * a.block < per * nsl)
                    return fail(h, "kernel %s: grid of %zu threads does not cover the %llu padded sites (PLK = %u)", k.name.c_str(), (size_t)k.global[0], (unsigned long long)(per * nsl), a.plk);
            }
            if (h->buffers[buf_id(k, 1)].capacity < (size_t)a.min_blocks * 16) return fail(h, "kernel %s: measurement buffer too small", k.name.c_str());
            CUDA_OK(h, launch_measure(c.sun, prec, a, st));
            break;
        }
        case K_REDUCE_ACTION_DIFF: {
            int index = 0;
            if (k.args.size() < 5 || k.args[3].type != Arg::U32X4 || !arg_int(k, 4, &index)) return fail(h, "kernel %s: arguments (param, index) not set", k.name.c_str());
            if (!buf_ptr(h, k, 0) || !buf_ptr(h, k, 1)) return fail(h, "kernel %s: buffers not set", k.name.c_str());
            const uint32_t* p = k.args[3].u4;                              // total partials, result stride
            const uint32_t n = (k.dir == 0) ? c.n1 : (k.dir == 1) ? c.n2 : c.n3;
            const uint32_t size = p[0] / n;
            if (index < 0 || h->buffers[buf_id(k, 1)].capacity < ((size_t)index + (size_t)p[1] * (n - 1) + 1) * 16) return fail(h, "kernel %s: result buffer too small", k.name.c_str());
            if (h->buffers[buf_id(k, 0)].capacity / 16 < (size_t)size * n) return fail(h, "kernel %s: measurement buffer too small", k.name.c_str());
            CUDA_OK(h, launch_reduce_slices((const double2*)buf_ptr(h, k, 0), size, n, (double2*)buf_ptr(h, k, 1), (uint32_t)index, p[1], 0u, 0u, false, st));
            break;
        }
        case K_POLYAKOV_DIFF: {
            PolyakovArgs a;
            memset(&a, 0, sizeof(a));
            a.table = buf_ptr(h, k, 0); a.out = (double2*)buf_ptr(h, k, 1); a.params = buf_ptr(h, k, 2);
            if (!a.table || !a.out || !a.params) return fail(h, "kernel %s: arguments not set", k.name.c_str());
            if (c.biglat) return fail(h, "kernel %s: the BIGLAT slice-resolved Polyakov kernels are not implemented in qcdgpu_b200", k.name.c_str());
            a.geom = make_geom(c); a.tbc = c.tbc ? 1 : 0; a.axis = k.dir;
            a.plk = c.plk_axis[k.dir] ? c.plk_axis[k.dir] : c.plk;
            if (!a.plk) return fail(h, "kernel %s: the program was created without -D PLK", k.name.c_str());
            a.block = (uint32_t)k.local0;
            if (a.block == 0 || a.block > 256) return fail(h, "kernel %s: work-group size %u not supported (1..256)", k.name.c_str(), a.block);
            a.min_blocks = (uint32_t)((k.global[0] + a.block - 1) / a.block);
            // (getK() on the host arranges that a work-group lies inside one slice when N1 == N2, suncl.cpp:2081-2091; other
            //  shapes mix neighbouring slices in a partial exactly as the reference does)
            if (h->buffers[buf_id(k, 1)].capacity < (size_t)a.min_blocks * 16) return fail(h, "kernel %s: measurement buffer too small", k.name.c_str());
            CUDA_OK(h, launch_polyakov(c.sun, prec, a, st));
            break;
        }
        case K_REDUCE_POLYAKOV_DIFF: {
            int index = 0;
            if (k.args.size() < 5 || k.args[3].type != Arg::U32X4 || !arg_int(k, 4, &index)) return fail(h, "kernel %s: arguments (param, index) not set", k.name.c_str());
            if (!buf_ptr(h, k, 0) || !buf_ptr(h, k, 1)) return fail(h, "kernel %s: buffers not set", k.name.c_str());
            if (c.pl < 3) break;                                           // #if (PL > 2) in polyakov.cl:439
            const uint32_t* p = k.args[3].u4;                              // total partials, offset of the (P^2,P^4) partials, result stride
            const uint32_t n = (k.dir == 0) ? c.n1 : (k.dir == 1) ? c.n2 : c.n3;
            const uint32_t size = p[0] / n;
            if (h->buffers[buf_id(k, 1)].capacity < ((size_t)index + (size_t)p[2] * n + 1) * 16) return fail(h, "kernel %s: result buffer too small", k.name.c_str());
            const size_t have = h->buffers[buf_id(k, 0)].capacity / 16;
            if (have < (size_t)size * n) return fail(h, "kernel %s: measurement buffer too small", k.name.c_str());
            // the extra sum over [offset, offset + size_total) re-reads the (P^2, P^4) partial region; the reference host sizes
            // lattice_measurement without it, so the range is clipped to the allocation instead of read out of bounds
            uint32_t extra = p[0];
            if ((size_t)p[1] >= have) extra = 0; else if ((size_t)p[1] + extra > have) extra = (uint32_t)(have - p[1]);
            CUDA_OK(h, launch_reduce_slices((const double2*)buf_ptr(h, k, 0), size, n, (double2*)buf_ptr(h, k, 1), (uint32_t)index, p[2], p[1], extra, !c.biglat, st));
            break;
        }
        case K_POLYAKOV: {
            PolyakovArgs a;
            memset(&a, 0, sizeof(a));
            int offset = 0;
            a.table = buf_ptr(h, k, 0); a.out = (double2*)buf_ptr(h, k, 1); a.params = buf_ptr(h, k, 2);
            if (!a.table || !a.out || !a.params || !arg_int(k, 4, &offset)) return fail(h, "kernel %s: arguments not set", k.name.c_str());
            a.geom = make_geom(c); a.tbc = c.tbc ? 1 : 0; a.offset = (uint32_t)offset; a.pl_level = c.pl;
            a.x_first = c.biglat ? 1 : 0; a.x_count = c.biglat ? c.n1 - 2 : c.n1;
            a.block = (uint32_t)k.local0;
            a.axis = -1;
            // PL_LEVEL = 3 enlarges the host's grid (suncl.cpp:2284-2285): all partial slots it will sum must be written
            a.min_blocks = (uint32_t)((k.global[0] + a.block - 1) / a.block);
            size_t blocks = ((size_t)a.x_count * a.geom.n23 + a.block - 1) / a.block;
            if (blocks < a.min_blocks) blocks = a.min_blocks;
            if (h->buffers[buf_id(k, 1)].capacity < ((c.pl > 1 ? (size_t)offset : 0) + blocks) * 16) return fail(h, "kernel %s: measurement buffer too small", k.name.c_str());
            CUDA_OK(h, launch_polyakov(c.sun, prec, a, st));
            break;
        }
        case K_REDUCE_POLYAKOV: {
            int index = 0;
            if (k.args.size() < 5 || k.args[3].type != Arg::U32X4 || !arg_int(k, 4, &index)) return fail(h, "kernel %s: arguments (param, index) not set", k.name.c_str());
            const uint32_t* p = k.args[3].u4;   // size, offset, offset2  (polyakov.cl:405-432)
            if (index < 0) return fail(h, "kernel %s: negative index", k.name.c_str());
            if (checked_reduce(h, k, buf_id(k, 0), buf_id(k, 1), 0, p[0], (size_t)index, st) < 0) return -1;
            if (c.pl >= 2) {
                if (checked_reduce(h, k, buf_id(k, 0), buf_id(k, 1), p[1], p[0], (size_t)index + p[2], st) < 0) return -1;
                h->launches++;
            }
            break;
        }
        case K_RANLUX_INIT: {
            if (!buf_ptr(h, k, 0) || !buf_ptr(h, k, 1)) return fail(h, "kernel %s: arguments (seeds, seedtable) not set", k.name.c_str());
            if (h->buffers[buf_id(k, 1)].bytes < k.global[0] * 20 * 4) return fail(h, "kernel %s: seed table too small (20 words per instance, prng.cuh)", k.name.c_str());
            CUDA_OK(h, launch_ranlux_init((const uint32_t*)buf_ptr(h, k, 0), (uint32_t*)buf_ptr(h, k, 1), (uint32_t)k.global[0], st));
            break;
        }
        case K_RANMAR_INIT: {
            if (!buf_ptr(h, k, 0) || !buf_ptr(h, k, 1)) return fail(h, "kernel %s: arguments (seeds, seedtable) not set", k.name.c_str());
            if (h->buffers[buf_id(k, 0)].bytes < k.global[0] * 2 * 16) return fail(h, "kernel %s: seeds must hold 2 uint4 per instance", k.name.c_str());
            if (h->buffers[buf_id(k, 1)].bytes < k.global[0] * 98 * 16) return fail(h, "kernel %s: seed table too small (97 + 1 float4 per instance, prngcl_ranmar.cl)", k.name.c_str());
            CUDA_OK(h, launch_ranmar_init((const uint32_t*)buf_ptr(h, k, 0), (float*)buf_ptr(h, k, 1), (uint32_t)k.global[0], st));
            break;
        }
        case K_PRODUCE: {
            int pb = buf_id(k, 1);
            int samples = 0;
            if (!buf_ptr(h, k, 0) || !valid_buf(h, pb) || !arg_int(k, 2, &samples)) return fail(h, "kernel %s: arguments (seed table, prns, samples) not set", k.name.c_str());
            {   // float4 / uint4 elements of generator state per instance
                const size_t per = (k.prn_kind == PRN_RANMAR) ? 98 : (k.prn_kind == PRN_XOR7 || k.prn_kind == PRN_RANECU || k.prn_kind == PRN_MRG32K3A) ? 2 : 0;
                if (per && h->buffers[buf_id(k, 0)].bytes < k.global[0] * per * 16)
                    return fail(h, "kernel %s: seed table too small (%zu vec4 per instance)", k.name.c_str(), per);
            }
            if (c.prng_precision == 2) return fail(h, "kernel %s: PRNG_PRECISION=2 (double PRNs) is not implemented", k.name.c_str());
            if (h->buffers[pb].bytes < (size_t)k.global[0] * samples * 16) return fail(h, "kernel %s: PRN buffer too small", k.name.c_str());
            if (h->buffers[pb].pending_partly_fused) finish_partial_pass(h, pb);   // previous pass: consumed by its x-range launches
            if (h->fused && h->buffers[pb].pending_produce == kid && k.prn_kind == PRN_RANLUX3) {
                h->buffers[pb].pending_skips++;        // the unconsumed earlier produce() becomes a jump (apply_pending_skips)
                return 0;
            }
            if (flush_pending(h, pb) < 0) return -1;   // an unconsumed earlier produce() must still advance the streams
            if (h->fused) { h->buffers[pb].pending_produce = kid; return 0; }   // decided at the consumer
            CUDA_OK(h, launch_prn_produce(k.prn_kind, (uint32_t*)buf_ptr(h, k, 0), h->buffers[pb].dptr, (uint32_t)k.global[0], (uint32_t)samples, st));
            break;
        }
        // ------------------------------------------------------------ O(N) scalar branch
        case K_ON_INIT_COLD: case K_ON_INIT_HOT: {
            Geom g = make_geom(c);
            const size_t rb = prec ? 8 : 4;
            if (!buf_ptr(h, k, 0) || h->buffers[buf_id(k, 0)].bytes < (size_t)g.sites * rb) return fail(h, "kernel %s: argument 0 (lattice_table) not set or smaller than the lattice", k.name.c_str());
            const void* prns = nullptr; const void* params = nullptr;
            if (k.kind == K_ON_INIT_HOT) {
                if (!buf_ptr(h, k, 1) || !buf_ptr(h, k, 2)) return fail(h, "kernel %s: arguments (prns, lattice_parameters) not set", k.name.c_str());
                if (flush_pending(h, buf_id(k, 1)) < 0) return -1;
                if (h->buffers[buf_id(k, 1)].bytes < (size_t)g.sites * 16) return fail(h, "kernel %s: PRN buffer too small for a hot start (%zu < %u vec4)", k.name.c_str(), h->buffers[buf_id(k, 1)].bytes / 16, g.sites);
                if (h->buffers[buf_id(k, 2)].bytes < 23 * rb) return fail(h, "kernel %s: lattice_parameters must hold 23 values", k.name.c_str());
                prns = buf_ptr(h, k, 1); params = buf_ptr(h, k, 2);
            }
            CUDA_OK(h, launch_on_init(prec, buf_ptr(h, k, 0), params, prns, g.sites, st));
            break;
        }
        case K_ON_UPDATE: {
            OnArgs a;
            memset(&a, 0, sizeof(a));
            a.table = buf_ptr(h, k, 0); a.params = buf_ptr(h, k, 1);
            if (!a.table || !a.params) return fail(h, "kernel %s: arguments (lattice_table, lattice_parameters) not set", k.name.c_str());
            if (h->buffers[buf_id(k, 1)].bytes < 23 * (size_t)(prec ? 8 : 4)) return fail(h, "kernel %s: lattice_parameters must hold 23 values", k.name.c_str());
            a.geom = make_geom(c); a.geom.tile_x = 0; a.skgroup = c.skgroup; a.odd = k.odd;
            if (c.acc_rate) {                                      // (lds, lattice_measurement) follow the PRN buffer
                a.acc = (double2*)buf_ptr(h, k, 4);
                if (!a.acc) return fail(h, "kernel %s: -D ACC_RATE needs argument 4 (lattice_measurement)", k.name.c_str());
                if (h->buffers[buf_id(k, 4)].capacity < (size_t)((a.geom.nthreads + kOnBlock - 1) / kOnBlock) * 16) return fail(h, "kernel %s: measurement buffer too small for the acceptance partials", k.name.c_str());
            }
            int pb = buf_id(k, 2);
            if (!valid_buf(h, pb)) return fail(h, "kernel %s: argument 2 (prns) not set", k.name.c_str());
            Buffer& pbuf = h->buffers[pb];
            const uint32_t need = (uint32_t)(c.nhit >> 1) + (uint32_t)(c.nhit & 1);        // vec4 a site can consume (oncl.cl:147-207)
            bool fused = false;
            if (pbuf.pending_produce) {
                Kernel& pk = h->kernels[pbuf.pending_produce];
                int samples = 0; arg_int(pk, 2, &samples);
                if (h->fused && (uint32_t)pk.global[0] == a.geom.prngstep && a.geom.nthreads <= a.geom.prngstep &&
                    (pk.prn_kind == PRN_XOR128 || pk.prn_kind == PRN_PM || pk.prn_kind == PRN_MRG32K3A || pk.prn_kind == PRN_XOR7 || pk.prn_kind == PRN_RANECU) && (uint32_t)samples >= need) {
                    a.prn_kind = pk.prn_kind; a.prn_table = (uint32_t*)buf_ptr(h, pk, 0);
                    a.prn_instances = (uint32_t)pk.global[0]; a.prn_samples = (uint32_t)samples;
                    pbuf.pending_produce = 0; fused = true;
                } else if (flush_pending(h, pb) < 0) return -1;
            }
            if (!fused) {
                a.prn_kind = PRN_BUFFERED; a.prns = pbuf.dptr;
                size_t need_v4 = (size_t)a.geom.prngstep * (need - 1) + a.geom.nthreads;
                if (pbuf.bytes < need_v4 * 16) return fail(h, "kernel %s: PRN buffer holds %zu vec4, the pass may read %zu", k.name.c_str(), pbuf.bytes / 16, need_v4);
            }
            CUDA_OK(h, launch_on_update(prec, a, st));
            break;
        }
        case K_ON_MEASURE: {
            OnArgs a;
            memset(&a, 0, sizeof(a));
            a.table = buf_ptr(h, k, 0); a.params = buf_ptr(h, k, 2);
            double2* out = (double2*)buf_ptr(h, k, 1);
            if (!a.table || !a.params || !out) return fail(h, "kernel %s: arguments (table, measurement, parameters) not set", k.name.c_str());
            if (h->buffers[buf_id(k, 2)].bytes < 23 * (size_t)(prec ? 8 : 4)) return fail(h, "kernel %s: lattice_parameters must hold 23 values", k.name.c_str());
            a.geom = make_geom(c); a.skgroup = c.skgroup;
            uint32_t step[4] = { 0, 0, 0, 0 };
            if (k.dir == 2) {
                if (k.args.size() < 5 || k.args[4].type != Arg::U32X4) return fail(h, "kernel %s: argument 4 (lattice_stepz) not set", k.name.c_str());
                memcpy(step, k.args[4].u4, 16);
                if (step[0] > c.n1 || step[1] > c.n2 || step[2] > c.n3 || step[3] > c.n4) return fail(h, "kernel %s: correlator step larger than the lattice (one wrap only, neighbours.cl:204-218)", k.name.c_str());
            }
            if (h->buffers[buf_id(k, 1)].capacity < (size_t)((a.geom.sites + kOnBlock - 1) / kOnBlock) * 16) return fail(h, "kernel %s: measurement buffer too small", k.name.c_str());
            CUDA_OK(h, launch_on_measure(prec, a, k.dir, step, out, st));
            break;
        }
        case K_ON_REDUCE: {
            int size = 0, index = 0;
            if (!arg_int(k, 3, &size) || !arg_int(k, 4, &index)) return fail(h, "kernel %s: arguments (size, index) not set", k.name.c_str());
            if (!buf_ptr(h, k, 0) || !buf_ptr(h, k, 1)) return fail(h, "kernel %s: buffers not set", k.name.c_str());
            if (h->buffers[buf_id(k, 0)].capacity < (size_t)size * 16 || h->buffers[buf_id(k, 1)].capacity < ((size_t)index + 1) * 16) return fail(h, "kernel %s: size/index outside the buffers", k.name.c_str());
            CUDA_OK(h, launch_on_reduce_add((const double2*)buf_ptr(h, k, 0), (uint32_t)size, (double2*)buf_ptr(h, k, 1), (uint32_t)index, st));
            break;
        }
        default:
            return fail(h, "kernel '%s' is not implemented in qcdgpu_b200", k.name.c_str());
    }
    h->launches++;
    return 0;
}

}  // namespace

extern "C" {

const char* qgpu_version(void) { return "qcdgpu_b200 0.1 (sm_100a)"; }

// glibc srand()/rand() are part of the reference's stream definition (random.cpp:235, 462-465): bulk helpers so that
// hosts written in languages with slow FFI can build seed tables of millions of instances quickly.
void qgpu_host_srand(unsigned seed) { srand(seed); }
void qgpu_ranlux_jump_table(uint64_t samples, uint64_t keep, unsigned skip, int from_start, uint32_t* out) {
    uint32_t t[7 * qcd::kRanluxJumpEntryWords];
    qcd::ranlux_jump_table(samples, keep, skip, t, from_start != 0);
    memcpy(out, t, sizeof(t));
}
void qgpu_host_rand_fill(uint32_t* out, size_t n, uint32_t modulo, uint32_t add) {
    for (size_t i = 0; i < n; i++) { uint32_t r = (uint32_t)rand(); if (modulo) r %= modulo; out[i] = r + add; }
}

qgpu* qgpu_create(void) { return new qgpu(); }
void qgpu_destroy(qgpu* h) {
    if (!h) return;
    if (h->initialized) qgpu_device_finalize(h, 0);
    delete h;
}
void qgpu_set_limit_max_workgroup_size(qgpu* h, unsigned limit) { h->limit_wg = limit; }
void qgpu_set_max_memory_width(qgpu* h, size_t w) { h->max_memory_width = w; if (h->initialized) h->info.max_memory_width = w; }
void qgpu_set_debug(qgpu* h, int profiling, int brief, int stage) {
    h->profiling = profiling == 1; h->profiling_mode = profiling; h->brief = brief != 0; h->show_stage = stage != 0;
}
// profiling == 2: wait for the recorded launches and add their times to the kernels' statistics (qgpu_kernel_get_execution_time)
int qgpu_profile_collect(qgpu* h) {
    if (!h->initialized) return 0;
    CUDA_OK(h, cudaSetDevice(h->device));
    for (auto& t : h->timeline) {
        CUDA_OK(h, cudaEventSynchronize(t.e1));
        float ms = 0; CUDA_OK(h, cudaEventElapsedTime(&ms, t.e0, t.e1));
        if (valid_kernel(h, t.kid)) { Kernel& k = h->kernels[t.kid]; double ns = (double)ms * 1e6; k.t_sum += ns; k.t_sq += ns * ns; k.starts++; }
        h->event_pool.push_back(t.e0); h->event_pool.push_back(t.e1);
    }
    int n = (int)h->timeline.size();
    h->timeline.clear();
    return n;
}
void qgpu_set_error_mode(qgpu* h, int r) { h->return_errors = r != 0; }
const char* qgpu_last_error(qgpu* h) { return h->last_error.c_str(); }
void qgpu_set_async(qgpu* h, int on) { h->async = on != 0; }
void qgpu_set_fused_prng(qgpu* h, int on) { h->fused = on != 0; }
void qgpu_set_tile_x(qgpu* h, int tile) { h->tile_x = tile; }
void qgpu_set_fold_reunitarisation(qgpu* h, int on) {
    h->fold_gs = on != 0;
    for (size_t i = 1; i < h->buffers.size(); i++) h->buffers[i].gs_cover.clear();
}
int qgpu_set_stream(qgpu* h, void* stream) {
    if (!h->initialized) return fail(h, "set_stream before device_initialize");
    if (h->own_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)stream;
    return 0;
}
int qgpu_kernel_set_x_range_strided(qgpu* h, int kid, int x_first, int x_count, int x_stride) {
    if (!valid_kernel(h, kid)) return fail(h, "bad kernel id %d", kid);
    Kernel& k = h->kernels[kid];
    if (k.kind != K_UPDATE && k.kind != K_OVERRELAX) return fail(h, "kernel %s has no x-range", k.name.c_str());
    if (x_stride < 1) return fail(h, "kernel %s: x-stride %d", k.name.c_str(), x_stride);
    k.x_first = x_first; k.x_count = x_count; k.x_stride = x_stride;
    return 0;
}
int qgpu_kernel_set_x_range(qgpu* h, int kid, int x_first, int x_count) { return qgpu_kernel_set_x_range_strided(h, kid, x_first, x_count, 1); }
uint64_t qgpu_launch_count(qgpu* h) { return h->launches; }
void* qgpu_stream(qgpu* h) { return (void*)h->stream; }

int qgpu_device_select(qgpu* h, unsigned platform_id, unsigned device_id) {
    if (platform_id != 0) return fail(h, "only platform 0 (CUDA) exists, got %u", platform_id);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(h, "no CUDA device: %s -- qcdgpu_b200 has no CPU fallback", cudaGetErrorString(e));
    if ((int)device_id >= n) return fail(h, "device %u requested, %d present", device_id, n);
    h->device = (int)device_id;
    return 1;
}
int qgpu_device_initialize(qgpu* h) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(h, "no CUDA device: %s -- qcdgpu_b200 has no CPU fallback", cudaGetErrorString(e));
    CUDA_OK(h, cudaSetDevice(h->device));
    cudaDeviceProp p;
    CUDA_OK(h, cudaGetDeviceProperties(&p, h->device));
    if (p.major < 10) return fail(h, "device %s is sm_%d%d; this library contains only sm_100a code", p.name, p.major, p.minor);
    CUDA_OK(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    CUDA_OK(h, cudaEventCreate(&h->ev0)); CUDA_OK(h, cudaEventCreate(&h->ev1));
    CUDA_OK(h, cudaEventCreate(&h->evp0)); CUDA_OK(h, cudaEventCreate(&h->evp1));
    CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    memset(&h->info, 0, sizeof(h->info));
    snprintf(h->info.device_name, sizeof(h->info.device_name), "%s", p.name);
    h->info.local_memory_size = p.sharedMemPerBlockOptin;
    h->info.max_constant_size = p.totalConstMem;
    h->info.max_memory_size = p.totalGlobalMem;
    // clinterface.cpp:395-401,454: max work-group size, optionally limited (QCDGPU.cpp:74 -> 64)
    size_t wg = (size_t)p.maxThreadsPerBlock;
    if (h->limit_wg && h->limit_wg < wg) wg = h->limit_wg;
    h->info.max_workgroup_size = wg;
    h->info.memory_align_factor = wg;
    // clinterface.cpp:419-421: default PRNG instance count = CL_DEVICE_IMAGE3D_MAX_WIDTH (16384 on NVIDIA's OpenCL)
    h->info.max_memory_width = h->max_memory_width ? h->max_memory_width : (size_t)p.maxTexture3D[0];
    if (!h->info.max_memory_width) h->info.max_memory_width = 32;
    h->info.sm_count = p.multiProcessorCount; h->info.cc_major = p.major; h->info.cc_minor = p.minor;
    h->initialized = true;
    return n;
}
int qgpu_device_finalize(qgpu* h, int error_code) {
    if (!h->initialized) return error_code;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (size_t i = 1; i < h->buffers.size(); i++) {
        if (h->buffers[i].dptr) cudaFree(h->buffers[i].dptr);
        if (h->buffers[i].twin) cudaFree(h->buffers[i].twin);
        if (h->buffers[i].shadow) cudaFreeHost(h->buffers[i].shadow);
        h->buffers[i] = Buffer();
    }
    if (h->ws_counters) { cudaFree(h->ws_counters); h->ws_counters = nullptr; }
    if (h->split_scratch) { cudaFree(h->split_scratch); h->split_scratch = nullptr; h->split_scratch_bytes = 0; }
    for (auto& t : h->timeline) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    h->timeline.clear(); h->event_pool.clear();
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->evp0); cudaEventDestroy(h->evp1); cudaEventDestroy(h->ev_halo);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = nullptr; h->own_stream = false; h->initialized = false;
    return error_code;
}
int qgpu_get_device_info(qgpu* h, qgpu_device_info* out) {
    if (!h->initialized) return fail(h, "device not initialized");
    *out = h->info;
    return 0;
}

int qgpu_program_create(qgpu* h, const char* /*source*/, const char* options) {
    ProgramCfg c;
    parse_options(options, c);
    if (c.on_model) {
        if (c.sun != 1) return fail(h, "program_create: -D ON_MODEL needs SUN=1 (one real per site), got SUN=%d", c.sun);
        if (c.on_sub_scheme) return fail(h, "program_create: -D ON_SUB_SCHEME (symmetric differences) is not implemented in qcdgpu_b200");
        if (c.n1exact && c.n1exact != c.n1) return fail(h, "program_create: the multi-part O(N) lattice (N1EXACT=%u != N1=%u) is not implemented in qcdgpu_b200", c.n1exact, c.n1);
        if (c.nhit < 1) return fail(h, "program_create: NHIT=%d", c.nhit);
    }
    if (c.has_lattice) {
        if (!c.on_model && c.sun != 2 && c.sun != 3) return fail(h, "program_create: SUN=%d is not supported (2 or 3)", c.sun);
        if (c.nd != 4) return fail(h, "program_create: ND=%d is not supported (4)", c.nd);
        if (!c.n1 || !c.n2 || !c.n3 || !c.n4) return fail(h, "program_create: N1..N4 missing in '%s'", c.options.c_str());
        if (c.n2 & 1) return fail(h, "program_create: N2 must be even for the checkerboard (sun_common.cl:351-380)");
        if ((uint64_t)c.n1 * c.n2 * c.n3 * c.n4 > 0xffffffffull) return fail(h, "program_create: more than 2^32 sites");
    }
    h->programs.push_back(c);
    h->active_program = (int)h->programs.size() - 1;
    return h->active_program;
}
int qgpu_program_set_active(qgpu* h, int id) {
    if (id <= 0 || id >= (int)h->programs.size()) return fail(h, "program_set_active: bad id %d", id);
    h->active_program = id; return id;
}
int qgpu_program_get_active(qgpu* h) { return h->active_program; }

int qgpu_kernel_init(qgpu* h, const char* name, unsigned dims, const size_t* global, const size_t* local) {
    if (!h->active_program) return fail(h, "kernel_init('%s') before program_create", name);
    Kernel k;
    k.name = name; k.program = h->active_program;
    const bool on_model = h->programs[k.program].on_model;
    if (on_model) { for (const NameMap& m : kNamesOn) if (k.name == m.name) { k.kind = m.kind; k.dir = m.dir; k.odd = m.odd; k.prn_kind = m.prn; } }
    else for (const NameMap& m : kNames) if (k.name == m.name) { k.kind = m.kind; k.dir = m.dir; k.odd = m.odd; k.prn_kind = m.prn; }
    if (k.kind == K_NONE) return fail(h, "kernel '%s' is not implemented in qcdgpu_b200", name);
    if (k.kind == K_PRODUCE && k.prn_kind < 0) {
        k.prn_kind = ranlux_kind(h->programs[k.program].rl_skip);
        if (k.prn_kind < 0) return fail(h, "kernel ranlux: RL_skip=%d is not one of 24/48/97/223/389", h->programs[k.program].rl_skip);
    }
    for (unsigned i = 0; i < dims && i < 3; i++) k.global[i] = global[i];
    // clinterface.cpp:1158-1169: explicit local size wins, else the (limited) kernel work-group size
    if (local && local[0]) { k.local0 = local[0]; k.local_given = true; }
    else k.local0 = (k.kind == K_UPDATE || k.kind == K_OVERRELAX) ? 128 : (h->limit_wg ? h->limit_wg : 128);
    if (k.kind == K_ON_UPDATE || k.kind == K_ON_MEASURE || k.kind == K_ON_INIT_COLD || k.kind == K_ON_INIT_HOT) { k.local0 = kOnBlock; k.local_given = false; }
    if (k.local0 > 128 && (k.kind == K_MEASURE || k.kind == K_MEASURE_PLQ || k.kind == K_POLYAKOV || k.kind == K_WILSON)) k.local0 = 128;
    h->kernels.push_back(k);
    return (int)h->kernels.size() - 1;
}
static int push_arg(qgpu* h, int kid, const Arg& a) {
    if (!valid_kernel(h, kid)) return fail(h, "bad kernel id %d", kid);
    h->kernels[kid].args.push_back(a);
    return (int)h->kernels[kid].args.size();
}
int qgpu_kernel_init_buffer(qgpu* h, int kid, int bid) {
    if (!valid_buf(h, bid)) return fail(h, "kernel_init_buffer: bad buffer id %d", bid);
    Arg a; memset(&a, 0, sizeof(a)); a.type = Arg::BUF; a.buf = bid; return push_arg(h, kid, a);
}
int qgpu_kernel_init_constant_int(qgpu* h, int kid, const int* p) { Arg a; memset(&a, 0, sizeof(a)); a.type = Arg::I32; a.i = *p; return push_arg(h, kid, a); }
int qgpu_kernel_init_constant_uint4(qgpu* h, int kid, const uint32_t p[4]) { Arg a; memset(&a, 0, sizeof(a)); a.type = Arg::U32X4; memcpy(a.u4, p, 16); return push_arg(h, kid, a); }
int qgpu_kernel_init_constant_float(qgpu* h, int kid, const float* p) { Arg a; memset(&a, 0, sizeof(a)); a.type = Arg::F32; a.f = *p; return push_arg(h, kid, a); }
int qgpu_kernel_init_constant_double(qgpu* h, int kid, const double* p) { Arg a; memset(&a, 0, sizeof(a)); a.type = Arg::F64; a.d = *p; return push_arg(h, kid, a); }
int qgpu_kernel_init_constant_reset(qgpu* h, int kid, const int* p, int argument_id) {
    // clinterface.cpp:1225-1229: sets argument `argument_id`; the reference's callers pass the value RETURNED by the
    // preceding kernel_init_constant, i.e. the index of the next, not yet set, argument (suncl.cpp:3415, 5267)
    if (!valid_kernel(h, kid)) return fail(h, "bad kernel id %d", kid);
    Kernel& k = h->kernels[kid];
    if (argument_id < 0) return fail(h, "kernel_init_constant_reset: bad argument id %d", argument_id);
    if ((size_t)argument_id >= k.args.size()) { Arg z; memset(&z, 0, sizeof(z)); z.type = Arg::I32; k.args.resize(argument_id + 1, z); }
    k.args[argument_id].type = Arg::I32; k.args[argument_id].i = *p;
    return argument_id;
}
int qgpu_kernel_get_worksize(qgpu* h, int kid) {
    if (!valid_kernel(h, kid)) return fail(h, "bad kernel id %d", kid);
    return (int)h->kernels[kid].local0;
}
int qgpu_kernel_run(qgpu* h, int kid) {
    if (!h->initialized) return fail(h, "kernel_run before device_initialize");
    if (!valid_kernel(h, kid)) return fail(h, "kernel_run: bad kernel id %d", kid);
    CUDA_OK(h, cudaSetDevice(h->device));
    if (h->profiling) CUDA_OK(h, cudaEventRecord(h->evp0, h->stream));
    qgpu::TimedLaunch tl = { kid, nullptr, nullptr };
    if (h->profiling_mode == 2) {
        for (cudaEvent_t* e : { &tl.e0, &tl.e1 }) {
            if (!h->event_pool.empty()) { *e = h->event_pool.back(); h->event_pool.pop_back(); }
            else CUDA_OK(h, cudaEventCreate(e));
        }
        CUDA_OK(h, cudaEventRecord(tl.e0, h->stream));
    }
    if (launch_kernel(h, kid) < 0) return -1;
    if (tl.e1) { CUDA_OK(h, cudaEventRecord(tl.e1, h->stream)); h->timeline.push_back(tl); }
    if (h->profiling) {
        CUDA_OK(h, cudaEventRecord(h->evp1, h->stream));
        CUDA_OK(h, cudaEventSynchronize(h->evp1));
        float ms = 0; CUDA_OK(h, cudaEventElapsedTime(&ms, h->evp0, h->evp1));
        Kernel& k = h->kernels[kid];
        double ns = (double)ms * 1e6; k.t_sum += ns; k.t_sq += ns * ns; k.starts++;
    }
    if (!h->async) CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return kid;
}
int qgpu_kernel_get_execution_time(qgpu* h, int kid, double* mean, double* dev, int* number) {
    if (!valid_kernel(h, kid)) return fail(h, "bad kernel id %d", kid);
    Kernel& k = h->kernels[kid];
    double m = k.starts ? k.t_sum / k.starts : 0.0;
    double v = k.starts > 1 ? (k.t_sq - k.starts * m * m) / (k.starts - 1) : 0.0;
    if (mean) *mean = m; if (dev) *dev = v > 0 ? sqrt(v) : 0.0; if (number) *number = k.starts;
    return 0;
}

int qgpu_buffer_init(qgpu* h, int type, size_t size, const void* host_ptr, size_t size_of) {
    if (!h->initialized) return fail(h, "buffer_init before device_initialize");
    CUDA_OK(h, cudaSetDevice(h->device));
    Buffer b;
    b.type = type; b.size = size; b.size_of = size_of; b.bytes = size * size_of; b.host_ptr = host_ptr; b.alive = true;
    if (type != QGPU_BUFFER_LDS && b.bytes) {
        // 64 KiB of zeroed slack behind every buffer: with small lattices the reference's own sizing lets
        // lattice_polyakov write its second partial block at measurement[BID + offset] past the end of
        // lattice_measurement (suncl.cpp:2272-2285 vs polyakov.cl:135-139); OpenCL runtimes tolerate it silently.
        b.capacity = b.bytes + (64u << 10);
        CUDA_OK(h, cudaMalloc(&b.dptr, b.capacity));
        CUDA_OK(h, cudaMemsetAsync(b.dptr, 0, b.capacity, h->stream));
        bool copy = host_ptr && (type == QGPU_BUFFER_INPUT || type == QGPU_BUFFER_CONSTANT || type == QGPU_BUFFER_IO);
        if (copy) CUDA_OK(h, cudaMemcpyAsync(b.dptr, host_ptr, b.bytes, cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(h, cudaStreamSynchronize(h->stream));   // host_ptr may be freed by the caller right after (blocking write, clinterface.cpp:1393-1410)
    }
    h->buffers.push_back(b);
    return (int)h->buffers.size() - 1;
}
int qgpu_buffer_write(qgpu* h, int bid) {
    if (!valid_buf(h, bid)) return fail(h, "buffer_write: bad buffer id %d", bid);
    settle_state_table(h, bid);
    Buffer& b = h->buffers[bid];
    if (!b.dptr || !b.host_ptr) return 0;
    CUDA_OK(h, cudaSetDevice(h->device));
    if (flush_pending(h, bid) < 0) return -1;
    CUDA_OK(h, cudaMemcpyAsync(b.dptr, b.host_ptr, b.bytes, cudaMemcpyHostToDevice, h->stream));
    b.gs_cover.clear();
    if (resync_twin(h, bid) < 0) return -1;
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
void* qgpu_buffer_map_range(qgpu* h, int bid, size_t off, size_t len) {
    if (!valid_buf(h, bid)) { fail(h, "buffer_map: bad buffer id %d", bid); return nullptr; }
    Buffer& b = h->buffers[bid];
    if (!b.dptr) { fail(h, "buffer_map: buffer %d has no device storage (LDS)", bid); return nullptr; }
    if (off + len > b.bytes) { fail(h, "buffer_map: range [%zu,+%zu) outside buffer of %zu bytes", off, len, b.bytes); return nullptr; }
    CUDA_OKP(h, cudaSetDevice(h->device));
    settle_state_table(h, bid);
    if (flush_pending(h, bid) < 0) return nullptr;
    if (!b.shadow) CUDA_OKP(h, cudaMallocHost(&b.shadow, b.bytes));      // pinned: D2H at full PCIe rate
    CUDA_OKP(h, cudaMemcpyAsync((char*)b.shadow + off, (char*)b.dptr + off, len, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OKP(h, cudaStreamSynchronize(h->stream));
    b.map_off = off; b.map_len = len;
    return (char*)b.shadow + off;
}
void* qgpu_buffer_map(qgpu* h, int bid) {
    if (!valid_buf(h, bid)) { fail(h, "buffer_map: bad buffer id %d", bid); return nullptr; }
    return qgpu_buffer_map_range(h, bid, 0, h->buffers[bid].bytes);
}
int qgpu_buffer_unmap(qgpu* h, int bid, void* ptr) {
    if (!valid_buf(h, bid)) return fail(h, "buffer_unmap: bad buffer id %d", bid);
    Buffer& b = h->buffers[bid];
    if (!b.shadow || !ptr) return 0;
    size_t off = (char*)ptr - (char*)b.shadow;
    if (off != b.map_off) return fail(h, "buffer_unmap: pointer does not match the mapped range");
    CUDA_OK(h, cudaSetDevice(h->device));
    settle_state_table(h, bid);
    CUDA_OK(h, cudaMemcpyAsync((char*)b.dptr + off, ptr, b.map_len, cudaMemcpyHostToDevice, h->stream));
    b.gs_cover.clear();
    if (resync_twin(h, bid) < 0) return -1;
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
int qgpu_buffer_kill(qgpu* h, int bid) {
    if (!valid_buf(h, bid)) return fail(h, "buffer_kill: bad buffer id %d", bid);
    Buffer& b = h->buffers[bid];
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    if (b.dptr) cudaFree(b.dptr);
    if (b.twin) cudaFree(b.twin);
    if (b.shadow) cudaFreeHost(b.shadow);
    b = Buffer();
    return 0;
}
size_t qgpu_buffer_size_bytes(qgpu* h, int bid) { return valid_buf(h, bid) ? h->buffers[bid].bytes : 0; }
unsigned qgpu_buffer_size_align_step(qgpu*, unsigned size, unsigned step) { return (1 + (size - 1) / step) * step; }   // clinterface.cpp:1587-1590
unsigned qgpu_buffer_size_align(qgpu* h, unsigned size) { return qgpu_buffer_size_align_step(h, size, (unsigned)h->info.max_workgroup_size); }
void* qgpu_buffer_device_ptr(qgpu* h, int bid) { if (valid_buf(h, bid)) settle_state_table(h, bid); return valid_buf(h, bid) ? h->buffers[bid].dptr : nullptr; }

int qgpu_synchronize(qgpu* h) {
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
int qgpu_timer_start(qgpu* h) { CUDA_OK(h, cudaSetDevice(h->device)); CUDA_OK(h, cudaEventRecord(h->ev0, h->stream)); return 0; }
int qgpu_timer_stop(qgpu* h, float* ms) {
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaEventRecord(h->ev1, h->stream));
    CUDA_OK(h, cudaEventSynchronize(h->ev1));
    CUDA_OK(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}
int qgpu_update_kernel_registers(qgpu* h, int program, int dir, int prn_kind) {
    if (program <= 0 || program >= (int)h->programs.size()) return -1;
    const ProgramCfg& c = h->programs[program];
    return update_kernel_regs(c.sun, prec_index(c), dir, c.tbc ? 1 : 0, prn_kind);
}

// ---------------------------------------------------------------- BigLattice halo exchange
static int planes_of(const ProgramCfg& c, int dir) { int per = (c.sun == 3) ? 3 : 1; return dir >= 0 ? per : 4 * per; }
size_t qgpu_halo_bytes(qgpu* h, int program, int dir) {
    if (program <= 0 || program >= (int)h->programs.size()) return 0;
    const ProgramCfg& c = h->programs[program];
    return (size_t)planes_of(c, dir) * c.n2 * c.n3 * c.n4 * (prec_index(c) ? 32 : 16);
}
int qgpu_halo_copy(qgpu* dst, int dprog, int dbuf, unsigned xd, qgpu* src, int sprog, int sbuf, unsigned xs, int dir) {
    if (dprog <= 0 || dprog >= (int)dst->programs.size() || sprog <= 0 || sprog >= (int)src->programs.size()) return fail(dst, "halo_copy: bad program id");
    if (!valid_buf(dst, dbuf) || !valid_buf(src, sbuf)) return fail(dst, "halo_copy: bad buffer id");
    const ProgramCfg& cd = dst->programs[dprog]; const ProgramCfg& cs = src->programs[sprog];
    if (cd.n2 != cs.n2 || cd.n3 != cs.n3 || cd.n4 != cs.n4 || cd.sun != cs.sun || prec_index(cd) != prec_index(cs)) return fail(dst, "halo_copy: geometry mismatch");
    if (xd >= cd.n1 || xs >= cs.n1) return fail(dst, "halo_copy: slice out of range");
    Geom gd = make_geom(cd), gs = make_geom(cs);
    CUDA_OK(dst, cudaSetDevice(dst->device));
    if (src != dst) {
        if (src->device != dst->device) {
            int can = 0; CUDA_OK(dst, cudaDeviceCanAccessPeer(&can, dst->device, src->device));
            if (!can) return fail(dst, "halo_copy: device %d cannot access device %d peer memory", dst->device, src->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(dst, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        // order after everything queued on the source stream
        CUDA_OK(dst, cudaSetDevice(src->device));
        CUDA_OK(dst, cudaEventRecord(src->ev_halo, src->stream));
        CUDA_OK(dst, cudaSetDevice(dst->device));
        CUDA_OK(dst, cudaStreamWaitEvent(dst->stream, src->ev_halo, 0));
    }
    int np = planes_of(cd, dir), p0 = dir >= 0 ? dir : 0, ps = dir >= 0 ? 4 : 1;
    CUDA_OK(dst, launch_copy_planes(prec_index(cd), dst->buffers[dbuf].dptr, gd.rowsize, xd, p0, ps,
                                    src->buffers[sbuf].dptr, gs.rowsize, xs, p0, ps, gd.n234, np, dst->stream));
    dst->launches++;
    if (!dst->async) CUDA_OK(dst, cudaStreamSynchronize(dst->stream));
    return 0;
}
// ---- halo exchange fused into the update kernels (one process per GPU): peer memory through CUDA IPC, flags for the hand-shake
int qgpu_buffer_ipc_export(qgpu* h, int buffer, unsigned char handle[64]) {
    if (!valid_buf(h, buffer) || !h->buffers[buffer].dptr) return fail(h, "ipc_export: bad buffer id %d", buffer);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t mh;
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaIpcGetMemHandle(&mh, h->buffers[buffer].dptr));
    memcpy(handle, &mh, 64);
    return 0;
}
void* qgpu_ipc_open(qgpu* h, const unsigned char handle[64]) {
    cudaIpcMemHandle_t mh;
    memcpy(&mh, handle, 64);
    void* p = nullptr;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        fail(h, "ipc_open: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
int qgpu_ipc_close(qgpu* h, void* ptr) {
    if (!ptr) return 0;
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, cudaIpcCloseMemHandle(ptr));
    return 0;
}
int qgpu_kernel_set_peer_halo(qgpu* h, int kid, int side, void* peer_table, unsigned src_x, unsigned dst_x, size_t peer_rowsize) {
    if (!valid_kernel(h, kid)) return fail(h, "bad kernel id %d", kid);
    Kernel& k = h->kernels[kid];
    if (k.kind != K_UPDATE && k.kind != K_OVERRELAX) return fail(h, "kernel %s stores no links: no peer halo", k.name.c_str());
    if (side < 0 || side > 1) return fail(h, "kernel %s: peer side %d (0 = first real slice, 1 = last real slice)", k.name.c_str(), side);
    const ProgramCfg& c = h->programs[k.program];
    if (peer_table && (!c.biglat || src_x < 1 || src_x + 1 >= c.n1)) return fail(h, "kernel %s: peer halo needs a -D BIGLAT program and a real source slice (got %u of %u)", k.name.c_str(), src_x, c.n1);
    k.peer[side].table = peer_table; k.peer[side].src_x = src_x; k.peer[side].dst_x = dst_x; k.peer[side].rowsize = (uint32_t)peer_rowsize;
    return 0;
}
// Parts of ONE process (several handles, one or several devices): the writer's device must reach the other part's memory, and the
// hand-shake is an event -- `waiter`'s stream continues once everything queued on `signaller`'s stream so far has completed.
int qgpu_enable_peer_access(qgpu* h, qgpu* other) {
    if (h->device == other->device) return 0;
    CUDA_OK(h, cudaSetDevice(h->device));
    int can = 0; CUDA_OK(h, cudaDeviceCanAccessPeer(&can, h->device, other->device));
    if (!can) return fail(h, "device %d cannot access device %d peer memory", h->device, other->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 0;
}
int qgpu_halo_order(qgpu* waiter, qgpu* signaller) {
    if (waiter == signaller || waiter->stream == signaller->stream) return 0;
    CUDA_OK(waiter, cudaSetDevice(signaller->device));
    CUDA_OK(waiter, cudaEventRecord(signaller->ev_halo, signaller->stream));
    CUDA_OK(waiter, cudaSetDevice(waiter->device));
    CUDA_OK(waiter, cudaStreamWaitEvent(waiter->stream, signaller->ev_halo, 0));
    return 0;
}
int qgpu_halo_signal(qgpu* h, void* flag_a, void* flag_b, unsigned long long value) {
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, launch_halo_signal((unsigned long long*)flag_a, (unsigned long long*)flag_b, value, h->stream));
    h->launches++;
    return 0;
}
int qgpu_halo_wait(qgpu* h, int flags_buffer, unsigned long long value, double timeout_s) {
    if (!valid_buf(h, flags_buffer) || h->buffers[flags_buffer].bytes < 32) return fail(h, "halo_wait: the flag buffer holds four 64-bit words (from the left, from the right, time-outs, spare)");
    CUDA_OK(h, cudaSetDevice(h->device));
    static int khz_of[64] = { 0 };      // cudaDevAttrClockRate is a driver query of about a millisecond: once per device
    int& khz = khz_of[h->device & 63];
    if (!khz) CUDA_OK(h, cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));
    if (!(timeout_s > 0.0)) timeout_s = 30.0;       // no or a nonsensical time-out: the default, never an endless spin on a lost neighbour
    const unsigned long long clocks = (unsigned long long)(timeout_s * 1e3 * (double)khz);
    CUDA_OK(h, launch_halo_wait((unsigned long long*)h->buffers[flags_buffer].dptr, value, clocks, h->stream));
    h->launches++;
    return 0;
}

int qgpu_halo_pack(qgpu* h, int program, int buffer, unsigned xs, int dir, void* packed, size_t* bytes) {
    if (program <= 0 || program >= (int)h->programs.size() || !valid_buf(h, buffer)) return fail(h, "halo_pack: bad id");
    const ProgramCfg& c = h->programs[program];
    Geom g = make_geom(c);
    if (xs >= c.n1) return fail(h, "halo_pack: slice out of range");
    int np = planes_of(c, dir), p0 = dir >= 0 ? dir : 0, ps = dir >= 0 ? 4 : 1;
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, launch_copy_planes(prec_index(c), packed, g.n234, 0, 0, 1, h->buffers[buffer].dptr, g.rowsize, xs, p0, ps, g.n234, np, h->stream));
    h->launches++;
    if (bytes) *bytes = qgpu_halo_bytes(h, program, dir);
    if (!h->async) CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
int qgpu_halo_unpack(qgpu* h, int program, int buffer, unsigned xd, int dir, const void* packed) {
    if (program <= 0 || program >= (int)h->programs.size() || !valid_buf(h, buffer)) return fail(h, "halo_unpack: bad id");
    const ProgramCfg& c = h->programs[program];
    Geom g = make_geom(c);
    if (xd >= c.n1) return fail(h, "halo_unpack: slice out of range");
    int np = planes_of(c, dir), p0 = dir >= 0 ? dir : 0, ps = dir >= 0 ? 4 : 1;
    CUDA_OK(h, cudaSetDevice(h->device));
    CUDA_OK(h, launch_copy_planes(prec_index(c), h->buffers[buffer].dptr, g.rowsize, xd, p0, ps, packed, g.n234, 0, 0, 1, g.n234, np, h->stream));
    h->launches++;
    if (!h->async) CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
