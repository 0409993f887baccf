// clinterface_cuda.cpp -- the reference-side binding: GPU_CL::GPU implemented on libqcdgpu_b200.so.
//
// This file is what a QCDGPU maintainer adds to the reference tree IN PLACE OF
// SU(N)/clinterface/clinterface.cpp.  It defines the members of the reference's own class
// GPU_CL::GPU (declared in SU(N)/clinterface/clinterface.h:54-415, which is compiled unchanged)
// by forwarding every device call to the C ABI of include/qcdgpu_b200.h.  QCDGPU.cpp, suncl.cpp,
// random.cpp, su{2,3}cpu.cpp, suncpu.cpp and data_analysis.cpp are compiled UNMODIFIED against it,
// so the reference's run loop, init.dat handling, result files and its own CPU re-measurement
// (LatticeCheckCPU -> "Verification successfully passed!") run on the CUDA backend.
//
// Build recipe: integration/Makefile (needs the reference tree and oracle/stub/CL/cl.h, a
// type-only stand-in for the Khronos header that clinterface.h includes).
//
// Only the members the callers actually use are defined (the linker is the check); the OpenCL
// program cache (.bin/.inf, MD5) has no meaning here and is not provided.
#include "clinterface/clinterface.h"
#include "qcdgpu_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <unistd.h>

namespace GPU_CL {

char GPU::current_path[FILENAME_MAX] = "\0";
GPU::GPU_debug_flags GPU::GPU_debug;

// the backend handle rides in the (otherwise unused) OpenCL context slot of the reference class
static inline qgpu* H(GPU* g) { return reinterpret_cast<qgpu*>(g->GPU_context); }

GPU::GPU(void) {
    GPU_debug.wait_for_keypress = false; GPU_debug.profiling = false; GPU_debug.brief_report = false;
    GPU_debug.show_stage = false; GPU_debug.local_run = false; GPU_debug.rebuild_binary = false;
    memset(&GPU_info, 0, sizeof(GPU_info));
    GPU_info.platform_vendor = GPU_vendor_nVidia; GPU_info.device_vendor = GPU_vendor_nVidia;
    GPU_current_kernel = GPU_current_buffer = GPU_current_program = GPU_active_program = 0;
    GPU_platform_id = GPU_device_id = 0;
    GPU_inf_max_n = 0;
    GPU_platforms_number = 1; GPU_devices_number = 0; GPU_total_devices_number = 0;
    GPU_max_work_item_dimensions = 3; GPU_max_work_item_sizes = NULL;
    GPU_platform = 0; GPU_device = 0; GPU_queue = NULL;
    CPU_timers = 32; CPU_current_timer_id = 0;
    CPU_timer = (int*)calloc(CPU_timers + 1, sizeof(int));
    GPU_limit_max_workgroup_size = 0;
    GPU_kernels = NULL; GPU_buffers = NULL; GPU_programs = NULL;   // the tables live in the backend
    cl_root_path = NULL;
    GPU_error = 0;
    MD5_finalized = false;
    GPU_context = reinterpret_cast<cl_context>(qgpu_create());
}
GPU::~GPU(void) {
    if (GPU_context) { qgpu_destroy(H(this)); GPU_context = NULL; }
    free(CPU_timer);
}
GPU::kernels_hash::kernels_hash(void) {}
GPU::kernels_hash::~kernels_hash(void) {}
GPU::buffers_hash::buffers_hash(void) {}
GPU::buffers_hash::~buffers_hash(void) {}
GPU::programs_hash::programs_hash(void) {}
GPU::programs_hash::~programs_hash(void) {}

// ---- bit converters (clinterface.cpp:298-343)
float GPU::convert_to_float(unsigned int v) { float f; memcpy(&f, &v, 4); return f; }
float GPU::convert_to_float(double v) { return (float)v; }
double GPU::convert_to_double(float v) { return (double)v; }
double GPU::convert_to_double(unsigned int v) { float f; memcpy(&f, &v, 4); return (double)f; }
double GPU::convert_to_double(unsigned int lo, unsigned int hi) { unsigned int w[2] = { lo, hi }; double d; memcpy(&d, w, 8); return d; }
unsigned int GPU::convert_to_uint_HIGH(double x) { unsigned int w[2]; memcpy(w, &x, 8); return w[1]; }
unsigned int GPU::convert_to_uint_LOW(double x) { unsigned int w[2]; memcpy(w, &x, 8); return w[0]; }
unsigned int GPU::convert_to_uint(float v) { unsigned int u; memcpy(&u, &v, 4); return u; }

// ---- device
bool GPU::device_select(unsigned int platform_id, unsigned int device_id) {
    GPU_platform_id = platform_id; GPU_device_id = device_id;
    return qgpu_device_select(H(this), 0, device_id) > 0;     // every "platform" of an init.dat maps to CUDA
}
bool GPU::device_auto_select(int, int) { return device_select(0, 0); }
int GPU::device_initialize(void) {
    int now = (int)clock();
    for (int i = 0; i < CPU_timers; i++) CPU_timer[i] = now;
    if (!getcwd(current_path, FILENAME_MAX)) current_path[0] = 0;
    if (!cl_root_path) { cl_root_path = (char*)calloc(FILENAME_MAX, 1); snprintf(cl_root_path, FILENAME_MAX, "%s/", current_path); }
    device_select(GPU_platform_id, GPU_device_id);
    qgpu_set_limit_max_workgroup_size(H(this), GPU_limit_max_workgroup_size);
    qgpu_set_debug(H(this), GPU_debug.profiling, GPU_debug.brief_report, GPU_debug.show_stage);
    if (const char* e = getenv("QCDGPU_PRNG_INSTANCES")) qgpu_set_max_memory_width(H(this), (size_t)atol(e));
    if (const char* e = getenv("QCDGPU_FUSED_PRNG")) qgpu_set_fused_prng(H(this), atoi(e));
    if (const char* e = getenv("QCDGPU_ASYNC")) qgpu_set_async(H(this), atoi(e));
    GPU_devices_number = (cl_uint)qgpu_device_initialize(H(this));
    qgpu_device_info info;
    qgpu_get_device_info(H(this), &info);
    GPU_info.device_name = strdup(info.device_name);
    GPU_info.local_memory_size = info.local_memory_size;
    GPU_info.max_constant_size = info.max_constant_size;
    GPU_info.max_memory_size = info.max_memory_size;
    GPU_info.max_memory_width = info.max_memory_width;
    GPU_info.max_workgroup_size = info.max_workgroup_size;
    GPU_info.memory_align_factor = info.memory_align_factor;
    printf("GPU_info.max_workgroup_size = %u\n", (unsigned int)GPU_info.max_workgroup_size);
    if (!GPU_limit_max_workgroup_size) GPU_limit_max_workgroup_size = (unsigned int)GPU_info.max_workgroup_size;
    return (int)GPU_devices_number;
}
int GPU::device_finalize(int error_code) { return qgpu_device_finalize(H(this), error_code); }
char* GPU::platform_get_name(cl_platform_id) { return strdup("NVIDIA CUDA (qcdgpu_b200, sm_100a)"); }
char* GPU::device_get_name(cl_device_id) { return GPU_info.device_name; }

// ---- programs / kernels / buffers: one C-ABI call each
char* GPU::source_read(const char*) { return strdup(""); }       // the kernels are precompiled CUDA; no .cl text is read
int GPU::program_create(const char* source) { return program_create(source, ""); }
int GPU::program_create(const char* source, const char* options) {
    GPU_active_program = GPU_current_program = qgpu_program_create(H(this), source, options);
    return GPU_current_program;
}
int GPU::kernel_init(const char* name, unsigned int dims, const size_t* global_size, const size_t* local_size) {
    return GPU_current_kernel = qgpu_kernel_init(H(this), name, dims, global_size, local_size);
}
int GPU::kernel_init_buffer(int kernel_id, int buffer_id) { return qgpu_kernel_init_buffer(H(this), kernel_id, buffer_id); }
int GPU::kernel_init_constant(int kernel_id, int* p) { return qgpu_kernel_init_constant_int(H(this), kernel_id, p); }
int GPU::kernel_init_constant(int kernel_id, cl_uint4* p) { return qgpu_kernel_init_constant_uint4(H(this), kernel_id, p->s); }
int GPU::kernel_init_constant(int kernel_id, float* p) { return qgpu_kernel_init_constant_float(H(this), kernel_id, p); }
int GPU::kernel_init_constant(int kernel_id, double* p) { return qgpu_kernel_init_constant_double(H(this), kernel_id, p); }
int GPU::kernel_init_constant_reset(int kernel_id, int* p, int argument_id) { return qgpu_kernel_init_constant_reset(H(this), kernel_id, p, argument_id); }
int GPU::kernel_run(int kernel_id) { return qgpu_kernel_run(H(this), kernel_id); }
int GPU::kernel_get_worksize(int kernel_id) { return qgpu_kernel_get_worksize(H(this), kernel_id); }
int GPU::buffer_init(int buffer_type, int size, void* host_ptr, int size_of) {
    return GPU_current_buffer = qgpu_buffer_init(H(this), buffer_type, (size_t)(unsigned int)size, host_ptr, (size_t)size_of);
}
unsigned int* GPU::buffer_map(int buffer_id) { return static_cast<unsigned int*>(qgpu_buffer_map(H(this), buffer_id)); }
#ifdef BIGLAT
// The members only the reference's BIGLAT build declares (clinterface.h:197-260).  `k` / `ndev` only name the files of its
// program-binary cache (clinterface.cpp:433-441, 750-790), which does not exist here.  Mapped ranges: clEnqueueMapBuffer(CL_MAP_WRITE,
// blocking) = read the range into host memory, clEnqueueUnmapMemObject = write it back; model::UpdateEdges (suncl.cpp:4005-4062)
// copies boundary slices between two such ranges.  One mapped range per buffer at a time, which is what the sequential
// (non-USE_OPENMP) BIGLAT host loop needs with two or more parts.
int GPU::device_initialize(int) { return device_initialize(); }
int GPU::program_create_ndev(const char* source, const char* options, int) { return program_create(source, options); }
unsigned int* GPU::buffer_map(int buffer_id, size_t offset, size_t size) { return static_cast<unsigned int*>(qgpu_buffer_map_range(H(this), buffer_id, offset, size)); }
void* GPU::buffer_map_part(int buffer_id, size_t start, size_t size) { return qgpu_buffer_map_range(H(this), buffer_id, start, size); }
void* GPU::buffer_map_void(int buffer_id) { return qgpu_buffer_map(H(this), buffer_id); }
void GPU::buffer_unmap(int buffer_id, void* ptr) { qgpu_buffer_unmap(H(this), buffer_id, ptr); }
void GPU::print_actual_hardware(void) { print_available_hardware(); }
#endif
cl_float4* GPU::buffer_map_float4(int buffer_id) { return static_cast<cl_float4*>(qgpu_buffer_map(H(this), buffer_id)); }
int GPU::buffer_kill(int buffer_id) { return qgpu_buffer_kill(H(this), buffer_id); }
unsigned int GPU::buffer_size_align(unsigned int size) { return buffer_size_align(size, (unsigned int)GPU_info.max_workgroup_size); }
unsigned int GPU::buffer_size_align(unsigned int size, unsigned int step) { return (1 + (size - 1) / step) * step; }

// ---- prints, timers (host only)
void GPU::print_available_hardware(void) { printf("CUDA device %u: %s\n", GPU_device_id, GPU_info.device_name ? GPU_info.device_name : "?"); }
void GPU::print_stage(const char* stage) { if (GPU_debug.show_stage) printf(">> %s\n", stage); }
void GPU::print_memory_utilized(void) {}
int GPU::print_time_detailed(void) { return 0; }
int GPU::start_timer_CPU(void) { int id = ++CPU_current_timer_id; if (id >= CPU_timers) id = 0; CPU_timer[id] = (int)clock(); return id; }
int GPU::start_timer_CPU(int timer) { int id = timer >= CPU_timers ? 0 : timer; CPU_timer[id] = (int)clock(); return id; }
int GPU::stop_timer_CPU(int timer) { int id = timer >= CPU_timers ? 0 : timer; CPU_timer[id] = (int)clock() - CPU_timer[id]; return CPU_timer[id]; }
double GPU::timer_in_seconds_CPU(int timer) { int id = timer >= CPU_timers ? 0 : timer; return CPU_timer[id] * 1.0 / CLOCKS_PER_SEC; }
double GPU::get_timer_CPU(int timer) { int id = timer >= CPU_timers ? 0 : timer; return ((int)clock() - CPU_timer[id]) * 1.0 / CLOCKS_PER_SEC; }
char* GPU::get_current_datetime(void) {
    time_t t; time(&t);
    char* r = (char*)calloc(256, 1);
    snprintf(r, 256, "%s", ctime(&t));       // "Sun Oct 18 14:58:32 2026\n": lattice_write_results slices this layout
    return r;
}

// ---- init.dat: `KEY = value   # comment` (same record type the reference's callers iterate over)
GPU::GPU_init_parameters* GPU::get_init_file(char finitf[]) {
    FILE* f = fopen(finitf, "r");
    if (!f) return NULL;
    int n = 0, cap = 16;
    GPU_init_parameters* res = (GPU_init_parameters*)calloc(cap, sizeof(GPU_init_parameters));
    char line[LENGTH];
    while (fgets(line, LENGTH, f)) {
        char* hash = strchr(line, '#');
        if (hash) *hash = 0;
        char* eq = strchr(line, '=');
        if (!eq) continue;
        *eq = 0;
        char* key = line; while (*key == ' ' || *key == '\t') key++;
        char* ke = eq - 1; while (ke >= key && (*ke == ' ' || *ke == '\t')) *ke-- = 0;
        char* val = eq + 1; while (*val == ' ' || *val == '\t') val++;
        char* ve = val + strlen(val) - 1; while (ve >= val && (*ve == ' ' || *ve == '\t' || *ve == '\n' || *ve == '\r')) *ve-- = 0;
        if (!*key) continue;
        if (n + 1 >= cap) { cap *= 2; res = (GPU_init_parameters*)realloc(res, cap * sizeof(GPU_init_parameters)); memset(res + n, 0, (cap - n) * sizeof(GPU_init_parameters)); }
        snprintf(res[n].Variable, LENGTH, "%s", key);
        snprintf(res[n].txtVarVal, LENGTH, "%s", val);
        res[n].iVarVal = atoi(val);
        res[n].fVarVal = atof(val);
#ifdef BIGLAT
        {   // "{a, b, ...}": one integer per part (clinterface.cpp:2122-2142)
            const char* q = strchr(val, '{');
            q = q ? q + 1 : val;
            for (int j = 0; j < NPARTS; j++) {
                while (*q == ' ' || *q == ',' || *q == '\t') q++;
                char* end = NULL;
                long v = strtol(q, &end, 10);
                if (end == q) break;
                res[n].ivVarVal[j] = (int)v;
                q = end;
            }
        }
#endif
        if (n) res[n - 1].final = false;
        res[n].final = true;
        n++;
    }
    fclose(f);
    if (!n) { free(res); return NULL; }
    return res;
}
void GPU::make_start_file(char* path) {
    char b[2 * FILENAME_MAX]; snprintf(b, sizeof(b), "%sstart.txt", path ? path : "");
    if (FILE* f = fopen(b, "w+")) { fprintf(f, "task started\n"); fclose(f); }
}
void GPU::make_finish_file(char* path) {
    if (!path) return;
    char b[2 * FILENAME_MAX]; snprintf(b, sizeof(b), "%sfinish.txt", path);
    if (FILE* f = fopen(b, "w+")) { fprintf(f, "task done\n"); fclose(f); }
}

}  // namespace GPU_CL
