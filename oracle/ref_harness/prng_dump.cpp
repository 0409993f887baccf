// prng_dump.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the UNMODIFIED reference PRNG host replicas (SU(N)/random/random.cpp,
// compiled with -D CPU_RUN from where it lies) and prints the instance-0 stream
// as raw IEEE-754 words, so that tests/golden/ref_prng_*.json can pin the oracle
// and the CUDA producers bit for bit.
//
//   prng_dump <RANLUX0..4|PM|XOR7|RANECU|RANMAR> <randseries> <count>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "random/random.h"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s GEN RANDSERIES COUNT\n", argv[0]); return 2; }
    PRNG_CL::PRNG* p = new PRNG_CL::PRNG;
    p->parameters_setup((char*)"PRNG", 0, argv[1]);
    p->PRNG_randseries = atoi(argv[2]);
    int n = atoi(argv[3]);
    p->initialize_CPU();
    float* r = (float*)calloc(n, sizeof(float));
    p->produce_CPU(r, n);
    for (int i = 0; i < n; i++) {
        unsigned int u; memcpy(&u, &r[i], 4);
        printf("%08x\n", u);
    }
    return 0;
}
