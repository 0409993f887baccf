// ckpt_main.cpp -- driver for the reference's OWN state-file code (model::lattice_save_state / lattice_load_state,
// SU(N)/suncl/suncl.cpp:1857-2079), which its stock main never reaches: QCDGPU.cpp:203 hard-codes turnoff_config_save = true and
// no init.dat key sets INIT = 0 or the name of the file to load.  This main repeats the field defaults of QCDGPU.cpp:61-236 and the
// call sequence of QCDGPU.cpp:242-262 on the reference's unmodified model class, with saving switched on and, given a second
// argument, INIT = 0 + that file as the state to continue from.  Everything else (run loop, file format, PRNG fast-forward) is
// the reference's code compiled from where it lies, over integration/clinterface_cuda.cpp + libqcdgpu_b200.so.
//
//   QCDGPU_b200_ckpt init.dat [state.qcg]        writes <OUTPUTPATH>/su3-<date>.qcg at the end of the run
//   QCDGPU_ACTIONS_DIFF=1 in the environment switches on model::get_actions_diff (action per x / y / z slice)
#include "QCDGPU.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

static char* dup_str(const char* s) {
    char* r = (char*)calloc(strlen(s) + 1, 1);
    strcpy(r, s);
    return r;
}

int main(int argc, char** argv) {
    using model_CL::model;
    if (argc < 2) { fprintf(stderr, "usage: %s init.dat [state.qcg]\n", argv[0]); return 2; }
    model* m = new (model);
    m->GPU0->GPU_limit_max_workgroup_size = 64;
    m->GPU0->GPU_debug.local_run = false;
    m->desired_platform = 0;
    m->desired_device = 0;
    m->device_select = true;
    m->GPU0->GPU_debug.rebuild_binary = false;
    m->PRNG0->PRNG_generator = PRNG_CL::PRNG::PRNG_generator_RANLUX3;
    m->PRNG0->PRNG_randseries = 54;
    m->PRNG0->PRNG_precision = PRNG_CL::PRNG::PRNG_precision_single;
    m->lattice_nd = 4;
    m->lattice_group = 3;
    for (int i = 0; i < 4; i++) { m->lattice_full_size[i] = 32; m->lattice_domain_size[i] = 32; }
    m->precision = model::model_precision_double;
    m->INIT = 1;
    m->NAV = 0;
    m->ITER = 2;
    m->NITER = 1;
    m->NHIT = 10;
    m->NHITPar = 1;
    m->BETA = 6.5;
    m->PHI = 0.0;
    m->OMEGA = 0.0;
    m->ints = model::model_start_cold;
    m->PL_level = 3;
    m->wilson_R = 7;
    m->wilson_T = 2;
    m->version = dup_str("1.5");
    m->turnoff_updates = false;
    m->turnoff_config_save = false;                  // the point of this driver
    m->turnoff_prns = false;
    m->turnoff_gramschmidt = false;
    m->get_actions_avr = true;
    m->get_plaquettes_avr = true;
    m->get_wilson_loop = false;
    m->get_Fmunu = false;
    m->get_F0mu = false;
    m->get_Fmunu1 = false;
    m->get_Fmunu5 = false;
    m->write_lattice_state_every_secs = 1.0e9;       // one file, at the end of the run (suncl.cpp:5498-5500)
    m->get_actions_diff = getenv("QCDGPU_ACTIONS_DIFF") != NULL;     // another switch without an init.dat key (QCDGPU.cpp:218)
    m->check_prngs = true;
    m->finishpath = dup_str("./");
    m->path = dup_str("./");
    m->fprefix = dup_str("su3-");
    m->fstate = dup_str(argc > 2 ? argv[2] : "none.qcg");
    m->lattice_get_init_file(argv[1]);
    if (argc > 2) m->INIT = 0;                       // continue from the state file (suncl.cpp:695, 3956, 5349-5354)
    m->lattice_init();
    m->lattice_simulate();
    m->lattice_analysis();
    m->lattice_write_results();
    printf("\n[...state driver completed...]\n");
    return 0;
}
