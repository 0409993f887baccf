// on_check.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the UNMODIFIED CPU re-measurement of the reference's O(N) tree
// (O(N)/suncl/suncpu.cpp: SU::lattice_check_cpu -> lattice_action_o1_cpu,
// lattice_avr_plaquette_plq_o1_cpu, lattice_correlators_o1_cpu, lines 93-270,
// 1805-1838), compiled from where it lies, over a site table read from a file.
// The reference's `model` object is never constructed (its constructor needs an
// OpenCL device): the harness fills in exactly the public fields the checker reads.
//
//   on_check <table.bin> n1 n2 n3 n4 precision(0|1) z lambda zeta eta b skgroup cx cy cz ct
//
// prints: S  mean_phi  mean_phi2  variance  corr1  corr2   (%.17g)
// and, with PRNG given as argv[17] = XOR128 and argv[18] = randseries, argv[19] = count:
// the XOR128 host stream of O(N)/random/random.cpp:636-665 as raw IEEE-754 words.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define private public   /* the host replicas of XOR128 are private members; layout is unaffected */
#include "QCDGPU.h"
#undef private

// the handful of GPU_CL::GPU members the checker and random.cpp reference
// (bit casts, O(N)/clinterface/clinterface.cpp:2262-2330); nothing else may be called
namespace GPU_CL {
double GPU::convert_to_double(unsigned int a) { float f; memcpy(&f, &a, 4); return (double)f; }
double GPU::convert_to_double(unsigned int lo, unsigned int hi) {
    unsigned long long u = ((unsigned long long)hi << 32) | lo; double d; memcpy(&d, &u, 8); return d; }
double GPU::convert_to_double(float f) { return (double)f; }
float  GPU::convert_to_float(unsigned int a) { float f; memcpy(&f, &a, 4); return f; }
unsigned int GPU::convert_to_uint(float f) { unsigned int a; memcpy(&a, &f, 4); return a; }
double GPU::get_timer_CPU(int) { return 0.0; }
#define NOT_HERE { fprintf(stderr, "on_check: device call reached\n"); abort(); }
unsigned int* GPU::buffer_map(int) NOT_HERE
cl_float4* GPU::buffer_map_float4(int) NOT_HERE
int  GPU::kernel_run(int) NOT_HERE
int  GPU::buffer_init(int, int, void*, int) NOT_HERE
int  GPU::buffer_kill(int) NOT_HERE
int  GPU::kernel_init(const char*, unsigned int, const size_t*, const size_t*) NOT_HERE
void GPU::print_stage(const char*) {}
char* GPU::source_read(const char*) NOT_HERE
int  GPU::program_create(const char*) NOT_HERE
int  GPU::program_create(const char*, const char*) NOT_HERE
void GPU::buffer_set_name(int, char*) {}
unsigned int GPU::buffer_size_align(unsigned int n) { return n; }
int  GPU::kernel_init_buffer(int, int) NOT_HERE
int  GPU::kernel_init_constant(int, int*) NOT_HERE
}
namespace analysis_CL {
void analysis::lattice_data_analysis_joint_CPU(data_analysis*, data_analysis*, data_analysis*) {}
}

int main(int argc, char** argv) {
    using model_CL::model;
    using SUN_CPU::SU;
    if (argc >= 20 && !strcmp(argv[17], "XOR128")) {
        PRNG_CL::PRNG* p = (PRNG_CL::PRNG*)calloc(1, sizeof(PRNG_CL::PRNG));
        srand(atoi(argv[18]));
        p->XOR128_initialize_CPU();
        int n = atoi(argv[19]);
        for (int i = 0; i < n; i++) {
            float f = p->XOR128_produce_one_CPU();
            unsigned int u; memcpy(&u, &f, 4);
            printf("%08x\n", u);
        }
        return 0;
    }
    if (argc < 17) { fprintf(stderr, "usage: see header\n"); return 2; }
    model* lat = (model*)calloc(1, sizeof(model));
    lat->run = (model::run_parameters*)calloc(1, sizeof(model::run_parameters));
    int n[4] = { atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]) };
    lat->run->lattice_domain_size = (int*)calloc(4, sizeof(int)); lat->run->lattice_full_size = (int*)calloc(4, sizeof(int));
    for (int i = 0; i < 4; i++) { lat->run->lattice_domain_size[i] = n[i]; lat->run->lattice_full_size[i] = n[i]; }
    lat->run->lattice_nd = 4;
    lat->run->precision = atoi(argv[6]) ? model::model_precision_double : model::model_precision_single;
    lat->run->ON_z = atof(argv[7]); lat->run->ON_lambda = atof(argv[8]); lat->run->ON_zeta = atof(argv[9]);
    lat->run->ON_eta = atof(argv[10]); lat->run->ON_b = atof(argv[11]); lat->run->ON_SK_group = atoi(argv[12]);
    lat->run->ON_sqrt_z_lambda_zeta = 0.25 * sqrt(lat->run->ON_z / (lat->run->ON_lambda * lat->run->ON_zeta)); // suncl.cpp:1244
    lat->run->correlator_X = atoi(argv[13]); lat->run->correlator_Y = atoi(argv[14]);
    lat->run->correlator_Z = atoi(argv[15]); lat->run->correlator_T = atoi(argv[16]);
    lat->run->get_plaquettes_avr = true; lat->run->get_correlators = true;
    lat->lattice_domain_n1 = n[0];
    lat->lattice_domain_n2n3 = n[1] * n[2];
    lat->lattice_domain_n2n3n4 = n[1] * n[2] * n[3];
    lat->lattice_full_site = n[0] * n[1] * n[2] * n[3];
    lat->lattice_table_row_size = lat->lattice_full_site;
    lat->Analysis = (analysis_CL::analysis::data_analysis*)calloc(64, sizeof(analysis_CL::analysis::data_analysis));
    size_t bytes = (size_t)lat->lattice_full_site * (atoi(argv[6]) ? 8 : 4);
    unsigned int* table = (unsigned int*)malloc(bytes);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(table, 1, bytes, f) != bytes) { fprintf(stderr, "on_check: cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    lat->lattice_pointer_last = table;
    SU* cpu = new SU;
    cpu->GPU0 = (GPU_CL::GPU*)calloc(1, sizeof(GPU_CL::GPU));
    cpu->lattice_check_cpu(lat);
    printf("%.17g %.17g %.17g %.17g %.17g %.17g\n",
           lat->Analysis[DM_S_spat].CPU_last_value, lat->Analysis[DM_Plq_spat].CPU_last_value,
           lat->Analysis[DM_Plq_temp].CPU_last_value, lat->Analysis[DM_Plq_total].CPU_last_value,
           lat->Analysis[DM_Correlator1].CPU_last_value, lat->Analysis[DM_Correlator2].CPU_last_value);
    return 0;
}
