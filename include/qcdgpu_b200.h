/* qcdgpu_b200.h -- C ABI of libqcdgpu_b200.so: the CUDA (sm_100a) replacement for QCDGPU's
 * OpenCL abstraction layer `GPU_CL::GPU` ("clinterface").
 *
 * Every entry point below replaces one member of the reference class
 *     SU(N)/clinterface/clinterface.h:54-415   (implementation clinterface.cpp)
 * with the same argument meaning: integer ids for programs / kernels / buffers, positional
 * kernel arguments set once, `kernel_run(id)` to execute.  The reference's callers
 * (suncl.cpp, random.cpp, su{2,3}cpu.cpp, data_analysis.cpp) can bind to this library through
 * the thin C++ shim shown in INTEGRATION.md without being modified.
 *
 * What is NOT taken over from the reference: `program_create` does not compile OpenCL C.
 * The option string ("-D SUN=3 -D N1=.. -D ROWSIZE=.. -D PRECISION=.. -D NHIT=.." as assembled
 * in suncl.cpp:3274-3301, random.cpp:427-430) is parsed into a kernel configuration and the
 * kernel NAMES of the reference (update_odd_X, lattice_measurement_plq, ranlux, ...) are mapped
 * to hand-written CUDA kernels.  Sizes are size_t (the reference overflows `int` above 2 GiB,
 * clinterface.cpp:1296,1324).
 *
 * The O(N) tree's programs ("-D ON_MODEL -D SUN=1 -D SKGROUP=.. [-D ACC_RATE]", O(N)/suncl/suncl.cpp:1454-1482) bind the kernel
 * names of O(N)/oncl/oncl.cl and on_measurements_cl.cl (update_even, update_odd, lattice_init_hot, lattice_measurement_correlator,
 * reduce_acceptance_rate, ...) through the same entry points; see INTEGRATION.md section 3a.
 *
 * Plain C types only; no CUDA, torch or C++ types cross this boundary.
 * Thread-safety: one handle per host thread, like the reference (one GPU instance per device).
 */
#ifndef QCDGPU_B200_H
#define QCDGPU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qgpu qgpu;

/* GPU::GPU_buffer_types, clinterface.h:73-82 (same numeric values) */
enum qgpu_buffer_type {
    QGPU_BUFFER_NONE = 0, QGPU_BUFFER_GLOBAL = 1, QGPU_BUFFER_INPUT = 2, QGPU_BUFFER_IO = 3,
    QGPU_BUFFER_OUTPUT = 4, QGPU_BUFFER_CONSTANT = 5, QGPU_BUFFER_LDS = 6, QGPU_BUFFER_UAV = 7
};

/* GPU::GPU_device_info, clinterface.h:127-143 */
typedef struct qgpu_device_info {
    char     device_name[256];
    uint64_t local_memory_size;      /* CL_DEVICE_LOCAL_MEM_SIZE  -> shared memory per block (opt-in max) */
    uint64_t max_constant_size;
    uint64_t max_memory_size;        /* CL_DEVICE_MAX_MEM_ALLOC_SIZE -> total global memory */
    size_t   max_memory_width;       /* default number of PRNG instances (clinterface.cpp:419-421) */
    size_t   max_workgroup_size;     /* alignment quantum of every buffer / work size */
    size_t   memory_align_factor;
    int      sm_count;
    int      cc_major, cc_minor;
} qgpu_device_info;

/* ---- life cycle: GPU::GPU(), ~GPU(), device_select, device_initialize, device_finalize
 *      clinterface.cpp:63-123, 345-489, 566-628 */
qgpu* qgpu_create(void);
void  qgpu_destroy(qgpu* h);
int   qgpu_device_select(qgpu* h, unsigned platform_id, unsigned device_id);  /* platform must be 0; device = CUDA ordinal */
int   qgpu_device_initialize(qgpu* h);                                      /* returns number of CUDA devices */
int   qgpu_device_finalize(qgpu* h, int error_code);
int   qgpu_get_device_info(qgpu* h, qgpu_device_info* out);
/* GPU_limit_max_workgroup_size (clinterface.h:160; QCDGPU.cpp:74 sets 64).  Call before device_initialize. */
void  qgpu_set_limit_max_workgroup_size(qgpu* h, unsigned limit);
/* GPU_info.max_memory_width override (the reference takes CL_DEVICE_IMAGE3D_MAX_WIDTH). */
void  qgpu_set_max_memory_width(qgpu* h, size_t width);
/* GPU_debug.profiling / brief_report / show_stage (clinterface.h:102-109) */
void  qgpu_set_debug(qgpu* h, int profiling, int brief_report, int show_stage);
/* profiling = 2 ("timeline", not in the reference): kernel_run records an event pair per launch WITHOUT waiting; qgpu_profile_collect
 * waits for them and adds the times to the per-kernel statistics read by qgpu_kernel_get_execution_time.  Gives a kernel's duration
 * inside a run of back-to-back launches, where profiling = 1 (GPU_debug.profiling, clinterface.cpp:1251-1278) waits after every launch. */
int   qgpu_profile_collect(qgpu* h);
/* 0 (default, reference behaviour OpenCL_Check_Error clinterface.cpp:225-284): print and exit(-1) on any error.
 * 1: return a negative value and keep the message for qgpu_last_error(). */
void        qgpu_set_error_mode(qgpu* h, int return_errors);
const char* qgpu_last_error(qgpu* h);

/* ---- programs: program_create (clinterface.cpp:715-1116).  `source` is ignored and may be NULL. */
int qgpu_program_create(qgpu* h, const char* source, const char* options);
int qgpu_program_set_active(qgpu* h, int program_id);
int qgpu_program_get_active(qgpu* h);

/* ---- kernels: kernel_init*, kernel_run, kernel_get_worksize (clinterface.cpp:1132-1294) */
int qgpu_kernel_init(qgpu* h, const char* kernel_name, unsigned work_dimensions,
                     const size_t* global_size, const size_t* local_size);
int qgpu_kernel_init_buffer(qgpu* h, int kernel_id, int buffer_id);
int qgpu_kernel_init_constant_int(qgpu* h, int kernel_id, const int* host_ptr);
int qgpu_kernel_init_constant_uint4(qgpu* h, int kernel_id, const uint32_t host_ptr[4]);
int qgpu_kernel_init_constant_float(qgpu* h, int kernel_id, const float* host_ptr);
int qgpu_kernel_init_constant_double(qgpu* h, int kernel_id, const double* host_ptr);
int qgpu_kernel_init_constant_reset(qgpu* h, int kernel_id, const int* host_ptr, int argument_id);
int qgpu_kernel_run(qgpu* h, int kernel_id);
int qgpu_kernel_get_worksize(qgpu* h, int kernel_id);
/* kernel_get_execution_time (clinterface.cpp:1290): mean / deviation in ns and number of starts (profiling on) */
int qgpu_kernel_get_execution_time(qgpu* h, int kernel_id, double* mean_ns, double* deviation_ns, int* number);

/* ---- buffers: buffer_init, buffer_write, buffer_map*, buffer_unmap, buffer_kill, buffer_size_align
 *      (clinterface.cpp:1296-1592).  host_ptr stays owned by the caller and is copied at init. */
int      qgpu_buffer_init(qgpu* h, int buffer_type, size_t size, const void* host_ptr, size_t size_of);
int      qgpu_buffer_write(qgpu* h, int buffer_id);                 /* host_ptr -> device again */
void*    qgpu_buffer_map(qgpu* h, int buffer_id);                   /* device -> backend-owned host shadow */
void*    qgpu_buffer_map_range(qgpu* h, int buffer_id, size_t offset_bytes, size_t size_bytes);
int      qgpu_buffer_unmap(qgpu* h, int buffer_id, void* ptr);      /* writes the mapped range back (BIGLAT UpdateEdges) */
int      qgpu_buffer_kill(qgpu* h, int buffer_id);
size_t   qgpu_buffer_size_bytes(qgpu* h, int buffer_id);
unsigned qgpu_buffer_size_align(qgpu* h, unsigned size);
unsigned qgpu_buffer_size_align_step(qgpu* h, unsigned size, unsigned step);

/* ================= extensions with no counterpart in the reference ================= */

/* Stream-ordered execution: with async on, qgpu_kernel_run enqueues and returns; qgpu_buffer_map and
 * qgpu_synchronize wait.  (The reference blocks in every kernel_run, clinterface.cpp:1251-1278.) */
void  qgpu_set_async(qgpu* h, int on);
int   qgpu_synchronize(qgpu* h);
/* Fused PRNG: a `ranlux`/`xor128`/`pm`/`mrg32k3a`/`xor7`/`ranecu` kernel_run is deferred; if the next consumer of its PRN
 * buffer is an update kernel and PRNG_instances == PRNGSTEP, the generator runs inside the update kernel
 * with its state in registers and the buffer is never written.  Any other access materialises it first.
 * Streams are identical either way; results too (xor7 / ranecu on SU(3) fp64: to rounding). */
void  qgpu_set_fused_prng(qgpu* h, int on);
/* L2 tiling of the update passes: x-chunk width of the launch order (0 = the reference's order, -1 = automatic:
 * 8 once three x-slices of links exceed 48 MB; env QCDGPU_TILE_X overrides the automatic choice).  Results do
 * not depend on it. */
void  qgpu_set_tile_x(qgpu* h, int tile);
/* Folded reunitarisation (default on; env QCDGPU_FOLD_GS=0 turns it off).  The reference's sweep ends with a lattice_GramSchmidt
 * launch over all links (suncl.cpp:5377) although its update kernels already orthonormalise every link they write
 * (su3_update_cl.cl:672-779).  With folding on, an SU(3) update / overrelaxation kernel applies the same Gram-Schmidt function a second
 * time before storing the link, and a lattice_GramSchmidt kernel_run whose table was completely rewritten that way since the previous
 * call (all 8 (parity, direction) passes, every x-slice) returns without a launch: the table already holds what it would write.
 * Anything else that writes the table (start kernels, buffer_write, unmap) makes the next call a real launch.  During a sweep
 * neighbours are read in their twice-orthonormalised form instead of the once-orthonormalised one: a difference of rounding errors. */
void  qgpu_set_fold_reunitarisation(qgpu* h, int on);
/* Run on a caller-owned cudaStream_t instead of the handle's own stream (e.g. a torch stream, so that NCCL
 * send/recv of halo buffers issued through torch.distributed is ordered with the kernels). */
int   qgpu_set_stream(qgpu* h, void* cuda_stream);
/* Restrict an update kernel to x-slices [x_first, x_first + x_count) of the updated region (x_count < 0: all).
 * Used by the BigLattice path to update the two boundary slices first and the interior while halos travel.
 * A restricted launch always reads the PRN buffer (the generators advance once per pass, for all sites). */
int   qgpu_kernel_set_x_range(qgpu* h, int kernel_id, int x_first, int x_count);
/* The same with the slices x_first + i * x_stride, i < x_count: x_stride = Nx - 1, x_count = 2 updates BOTH boundary slices of a part in
 * one launch (two one-slice launches cost about 0.1 ms of ramp-up and tail each at 64^3 sites per slice). */
int   qgpu_kernel_set_x_range_strided(qgpu* h, int kernel_id, int x_first, int x_count, int x_stride);
/* raw device pointer of a buffer (halo exchange, interop) and raw stream handle (cudaStream_t) */
void* qgpu_buffer_device_ptr(qgpu* h, int buffer_id);
void* qgpu_stream(qgpu* h);
/* device-side timing on the launching stream (CUDA events) */
int   qgpu_timer_start(qgpu* h);
int   qgpu_timer_stop(qgpu* h, float* elapsed_ms);
/* number of kernels launched through this handle so far */
uint64_t qgpu_launch_count(qgpu* h);
/* registers per thread of the update kernel a program would use (for reports); -1 if unknown */
int   qgpu_update_kernel_registers(qgpu* h, int program_id, int direction, int prn_kind);

/* BigLattice halo exchange between x-slabs (replaces model::UpdateEdges, suncl.cpp:4005-4062).
 * Copies x-slice `x_src` of buffer `src_buffer` on handle `src` into x-slice `x_dst` of `dst_buffer` on
 * handle `dst` for the vec4 planes of direction `dir` (dir < 0: all directions).  Both buffers are link
 * tables of programs with the same N2,N3,N4.  The copy is a kernel on dst's stream reading peer memory
 * (NVLink P2P when the two handles sit on different devices) and is ordered after src's stream. */
int   qgpu_halo_copy(qgpu* dst, int dst_program, int dst_buffer, unsigned x_dst,
                     qgpu* src, int src_program, int src_buffer, unsigned x_src, int dir);
/* same, but the source is a raw device pointer to a packed slice produced by qgpu_halo_pack (for
 * transports that cannot address peer memory, e.g. NCCL send/recv between processes) */
int   qgpu_halo_pack(qgpu* h, int program, int buffer, unsigned x_src, int dir, void* packed_dev, size_t* bytes);
int   qgpu_halo_unpack(qgpu* h, int program, int buffer, unsigned x_dst, int dir, const void* packed_dev);
size_t qgpu_halo_bytes(qgpu* h, int program, int dir);
/* Halo exchange FUSED into the update kernels (one process per GPU; no counterpart in the reference, whose model::UpdateEdges,
 * suncl.cpp:4005-4062, copies the slices through mapped host memory after the pass).  The link table of a neighbouring part is
 * opened as peer memory (CUDA IPC handle of its buffer, exchanged by the host over its own transport); an update / overrelaxation
 * kernel that was given a peer then stores every link it updates in stored x-slice `src_x` (side 0: the first real slice, side 1:
 * the last one) a second time, into x-slice `dst_x` of the peer table, over NVLink -- no pack kernel, no NCCL call, no unpack kernel.
 * peer_table = NULL clears a side.  The hand-shake: qgpu_halo_signal (stream-ordered) writes `value` into one flag word of each
 * neighbour (peer pointers into their flag buffers: word 0 = "from my left neighbour", word 1 = "from my right neighbour");
 * qgpu_halo_wait (stream-ordered) holds the stream until both of MY words have reached `value`, or gives up after timeout_s and
 * counts that in word 2 of the flag buffer (four 64-bit words, zeroed by the host). */
int   qgpu_buffer_ipc_export(qgpu* h, int buffer_id, unsigned char handle[64]);
void* qgpu_ipc_open(qgpu* h, const unsigned char handle[64]);
int   qgpu_ipc_close(qgpu* h, void* peer_ptr);
int   qgpu_kernel_set_peer_halo(qgpu* h, int kernel_id, int side, void* peer_table, unsigned src_x, unsigned dst_x, size_t peer_rowsize);
/* the same peer stores for parts that live in ONE process (the reference's mode: several devices, one host program): the peer table is
 * simply the other handle's qgpu_buffer_device_ptr; qgpu_enable_peer_access lets h's device store into other's, and qgpu_halo_order
 * (an event) makes waiter's stream continue once what signaller's stream holds so far has completed. */
int   qgpu_enable_peer_access(qgpu* h, qgpu* other);
int   qgpu_halo_order(qgpu* waiter, qgpu* signaller);
int   qgpu_halo_signal(qgpu* h, void* peer_flag_a, void* peer_flag_b, unsigned long long value);
int   qgpu_halo_wait(qgpu* h, int flags_buffer_id, unsigned long long value, double timeout_s);

/* host helpers: glibc srand()/rand() in bulk (seed tables are defined through them, random.cpp:235,462-465) */
void  qgpu_host_srand(unsigned seed);
void  qgpu_host_rand_fill(uint32_t* out, size_t n, uint32_t modulo, uint32_t add);
/* RANLUX skip-ahead table (host arithmetic only; csrc/ranlux_jump_host.h).  The subtract-with-borrow recurrence of prngcl_ranlux.cl:97-563
 * / random.cpp:504-574 is a multiplication by 2^-24 mod 2^576 - 2^240 + 1 per step, so the part of a generator pass (`samples` float4)
 * behind its first `keep` samples -- or the whole pass, from_start != 0 -- is ONE modular multiplication.  out: 7 entries (`left` = numbers
 * to go before the next luxury skip on entry, / 4) of 20 words: 18 words of 2^-24(steps - 48), `left` after the pass, 1 if usable.
 * The backend builds these itself (two-phase update kernels; `ranlux` kernel_runs nobody consumed, as when a resumed run replays
 * PRNG_counter produce() calls, suncl.cpp:5350); exported so that hosts and tests can check the arithmetic. */
void  qgpu_ranlux_jump_table(uint64_t samples, uint64_t keep, unsigned skip, int from_start, uint32_t* out);

const char* qgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
