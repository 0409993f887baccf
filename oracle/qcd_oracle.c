/* qcd_oracle.c -- CPU restatement of QCDGPU's gauge-update hot path (PRNG part
 * and the two precision instantiations of qcd_oracle_lattice.inc).
 *
 * TEST INFRASTRUCTURE ONLY -- see qcd_oracle.h for the usage rule and the
 * parity-pinning status.  File:line citations are relative to /root/reference/.
 */
#define _GNU_SOURCE
#include "qcd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ indexing */
/* lattice_gid_to_coords: SU(N)/suncl/sun_common.cl:41-61 */
void orc_gid_to_coords(const orc_cfg* c, uint32_t gdi, uint32_t v[4]) {
    uint32_t n23 = c->n2 * c->n3, n234 = n23 * c->n4;
    uint32_t z4 = gdi / n234;
    uint32_t z1 = gdi - z4 * n234;
    uint32_t z3 = z1 / n23;
    z1 = z1 - z3 * n23;
    uint32_t z2 = z1 / c->n2;
    z1 = z1 - z2 * c->n2;
    v[0] = z4; v[1] = z1; v[2] = z2; v[3] = z3;
}
/* lattice_even_gid / lattice_odd_gid: sun_common.cl:351-380 */
uint32_t orc_even_gid(const orc_cfg* c, uint32_t gid) {
    uint32_t gde = 2 * gid, v[4];
    orc_gid_to_coords(c, gde, v);
    return gde + ((v[0] + v[1] + v[2] + v[3]) & 1);
}
uint32_t orc_odd_gid(const orc_cfg* c, uint32_t gid) {
    uint32_t gde = 2 * gid + 1, v[4];
    orc_gid_to_coords(c, gde, v);
    return gde - (((v[0] + v[1] + v[2] + v[3]) & 1) ^ 1);
}
/* lattice_neighbours_gid{,_minus}: sun_common.cl:455-513 */
uint32_t orc_neighbour(const orc_cfg* c, uint32_t g, int dir, int sign) {
    uint32_t v[4], n[4] = { c->n1, c->n2, c->n3, c->n4 };
    orc_gid_to_coords(c, g, v);
    if (sign > 0) {
        v[dir]++;
        if (!(c->biglat && dir == 0) && v[dir] >= n[dir]) v[dir] = 0;
    } else {
        if (v[dir] == 0) v[dir] = n[dir];
        v[dir]--;
    }
    return v[1] + v[2] * c->n2 + v[3] * c->n2 * c->n3 + v[0] * c->n2 * c->n3 * c->n4;
}

/* ------------------------------------------------------------------ lattice, two precisions */
#define REAL double
#define SFX f64
#define RSIN sin
#define RCOS cos
#define RSQRT sqrt
#define RPI  3.1415926535897932384626433832795
#define RPI2 6.2831853071795864769252867665590
#include "qcd_oracle_lattice.inc"
#undef REAL
#undef SFX
#undef RSIN
#undef RCOS
#undef RSQRT
#undef RPI
#undef RPI2

#define REAL float
#define SFX f32
#define RSIN sinf
#define RCOS cosf
#define RSQRT sqrtf
#define RPI  3.141592653589793f
#define RPI2 6.283185307179586f
#include "qcd_oracle_lattice.inc"
#undef REAL
#undef SFX
#undef RSIN
#undef RCOS
#undef RSQRT
#undef RPI
#undef RPI2

/* ------------------------------------------------------------------ PRNG */
#define RL_ICONS  2147483563
#define RL_ITWO24 16777216
#define RL_TWOM24 0.000000059604644775390625f
#define RL_TWOM12 0.000244140625f

typedef struct {
    float seeds[24];
    int i24, j24, in24;
    float carry;
} rl_state;

/* RL_initialize_CPU / ranlux_init: SU(N)/random/random.cpp:504-522, prngcl_ranlux.cl:97-157 */
static void rl_seed(rl_state* s, int jseed) {
    for (int i = 0; i < 24; i++) {
        int k = jseed / 53668;
        jseed = 40014 * (jseed - k * 53668) - k * 12211;
        if (jseed < 0) jseed = jseed + RL_ICONS;
        s->seeds[i] = ((float)(jseed % RL_ITWO24)) * RL_TWOM24;
    }
    s->i24 = 23; s->j24 = 9; s->in24 = 0; s->carry = 0.0f;
    if (s->seeds[23] == 0) s->carry = RL_TWOM24;
}
static inline float rl_swb(rl_state* s) {
    float uni = s->seeds[s->j24] - s->seeds[s->i24] - s->carry;
    if (uni < 0.0f) { uni = uni + 1.0f; s->carry = RL_TWOM24; } else s->carry = 0.0f;
    s->seeds[s->i24] = uni;
    if (s->i24 == 0) s->i24 = 24;
    if (s->j24 == 0) s->j24 = 24;
    s->i24--; s->j24--;
    return uni;
}
/* RL_produce_one_CPU: random.cpp:523-574 */
static float rl_one(rl_state* s, int nskip) {
    if (s->in24 == 0) {
        s->in24 = 24;
        for (int isk = 0; isk < nskip; isk++) rl_swb(s);
    }
    float uni = rl_swb(s);
    if (uni < RL_TWOM12) {
        uni = uni + RL_TWOM24 * s->seeds[s->j24];
        if (uni == 0.0f) uni = RL_TWOM24 * RL_TWOM24;
    }
    s->in24--;
    return uni;
}

/* xor128 -- O(N)/random/random.cpp:636-652 (host) == O(N)/random/random.cl:137-173 (device):
 * the emitted word is state[0] AFTER the shift.  (The SU(N) .cl emits state[3],
 * i.e. the same sequence three steps ahead, prngcl_xor128.cl:124-131, and
 * disagrees with its own host replica; the O(N) pair is the pinned one.) */
static inline uint32_t xor128_step(uint32_t s[4]) {
    uint32_t t = (s[0] ^ (s[0] << 11));
    s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
    s[3] = (s[3] ^ (s[3] >> 19)) ^ (t ^ (t >> 8));
    return s[0];
}
/* Park-Miller, one lane: prngcl_pm.cl:86-102 (device) / random.cpp:896-910 (host) */
static inline uint32_t pm_step(uint32_t seed) {
    uint32_t t1 = 16807u * (seed % 127773u);
    uint32_t t2 = 2836u * (seed / 127773u);
    if (t1 > t2) return t1 - t2;
    return t1 - t2 + 2147483647u;
}
/* MRG32k3a: prngcl_mrg32k3a.cl:73-97.  No host replica in the reference. */
static inline float mrg_step(uint32_t s1[3], uint32_t s2[3]) {
    int64_t x1n = 1403580LL * (int64_t)s1[1] - 810728LL * (int64_t)s1[0];
    int64_t k = x1n / 4294967087LL;
    x1n -= k * 4294967087LL;
    if (x1n < 0) x1n += 4294967087LL;
    s1[0] = s1[1]; s1[1] = s1[2]; s1[2] = (uint32_t)x1n;
    int64_t x2n = 527612LL * (int64_t)s2[2] - 1370589LL * (int64_t)s2[0];
    k = x2n / 4294944443LL;
    x2n -= k * 4294944443LL;
    if (x2n < 0) x2n += 4294944443LL;
    s2[0] = s2[1]; s2[1] = s2[2]; s2[2] = (uint32_t)x2n;
    const double m = 2.3283065492957276882397381481401e-10;
    if (x1n > x2n) return (float)((double)(x1n - x2n) * m);
    return (float)((double)(x1n - x2n + 4294967087LL) * m);
}

/* XOR7 -- prngcl_xor7.cl:59-75 (device: two uint4 shifting by one word per step) == random.cpp:975-986
 * (host: ring of eight words with a moving index).  s[0] oldest, s[7] newest. */
static inline uint32_t xor7_step(uint32_t s[8]) {
    uint32_t t, y;
    t = s[7]; t = t ^ (t << 13); y = t ^ (t << 9);
    t = s[4]; y ^= t ^ (t << 7);
    t = s[3]; y ^= t ^ (t >> 3);
    t = s[1]; y ^= t ^ (t >> 10);
    t = s[0]; t = t ^ (t >> 7); y ^= t ^ (t << 24);
    for (int k = 0; k < 7; k++) s[k] = s[k + 1];
    s[7] = y;
    return y;
}
/* RANECU, one lane: prngcl_ranecu.cl:77-103 (device, unsigned compares) */
static inline uint32_t ranecu_step(uint32_t* a, uint32_t* b) {
    uint32_t q = *a / 53668u;
    uint32_t t1 = 40014u * (*a - q * 53668u), t2 = q * 12211u;
    *a = (t1 > t2) ? (t1 - t2) : (t1 - t2 + 2147483563u);
    q = *b / 52774u;
    t1 = 40692u * (*b - q * 52774u); t2 = q * 3791u;
    *b = (t1 > t2) ? (t1 - t2) : (t1 - t2 + 2147483399u);
    return (*a > *b) ? (*a - *b) : (*a - *b + 2147483562u);
}
/* RANMAR: seeds of one lane (ranmar_init, prngcl_ranmar.cl:65-101 == RANMAR_initialize_CPU, random.cpp:773-799) */
static void ranmar_seed_lane(uint32_t seed1, uint32_t seed2, float out[97]) {
    uint32_t i = ((seed1 / 177u) % 177u) + 2u, j = (seed1 % 177u) + 2u, k = ((seed2 / 169u) % 178u) + 1u, l = seed2 % 169u;
    for (int n = 0; n < 97; n++) {
        uint32_t sv = 0u, t = 8388608u;
        for (int m = 0; m < 24; m++) {
            uint32_t u = (i * j) % 179u;
            u = (u * k) % 179u;
            i = j; j = k; k = u;
            l = (53u * l + 1u) % 169u;
            if (((l * u) % 64u) >= 32u) sv += t;
            t >>= 1;
        }
        out[n] = ((float)sv) / 16777216.0f;
    }
}
/* one RANMAR instance: 97 seeds x 4 lanes, the two indices and the arithmetic-sequence value (prngcl_ranmar.cl:103-133) */
typedef struct { float u[97][4]; uint32_t i97, j97; float c; } rm_state;
static void ranmar_step4(rm_state* r, float o[4]) {
    const float CD = 7654321.0f / 16777216.0f, CM = 16777213.0f / 16777216.0f;
    for (int k = 0; k < 4; k++) {
        float uni = r->u[r->i97][k] - r->u[r->j97][k];
        if (uni < 0.0f) uni = uni + 1.0f;
        r->u[r->i97][k] = uni;
        o[k] = uni;
    }
    if (r->i97 == 0) r->i97 = 97;
    if (r->j97 == 0) r->j97 = 97;
    r->i97--; r->j97--;
    r->c -= CD;
    if (r->c < 0.0f) r->c += CM;
    for (int k = 0; k < 4; k++) {
        o[k] = o[k] - r->c;
        if (o[k] < 0.0f) o[k] = o[k] + 1.0f;
    }
}

struct orc_prng {
    int generator;
    uint32_t instances;
    uint32_t table_size;
    int nskip;
    rl_state* rl;
    uint32_t* seeds;     /* raw words handed to the device */
    uint32_t n_seed_words;
    uint32_t* st;        /* xor128/pm: 4 words per instance; mrg: 8 words per instance */
    rm_state* rm;        /* ranmar */
};

static int rl_nskip_of(int generator) {
    switch (generator) {
        case ORC_PRNG_RANLUX0: return 24;
        case ORC_PRNG_RANLUX1: return 48;
        case ORC_PRNG_RANLUX2: return 97;
        case ORC_PRNG_RANLUX3: return 223;
        case ORC_PRNG_RANLUX4: return 389;
        default: return -1;
    }
}

orc_prng* orc_prng_create(int generator, uint32_t randseries, uint32_t instances, uint32_t table_size) {
    orc_prng* p = (orc_prng*)calloc(1, sizeof(orc_prng));
    p->generator = generator; p->instances = instances;
    if (table_size < instances) table_size = instances;
    p->table_size = table_size;
    p->nskip = rl_nskip_of(generator);
    srand(randseries);                                   /* random.cpp:235 */
    if (p->nskip >= 0) {
        /* random.cpp:508 then 462-465 */
        p->n_seed_words = instances;
        p->seeds = (uint32_t*)calloc(instances, sizeof(uint32_t));
        p->seeds[0] = (uint32_t)(rand() % 2147483647);
        for (uint32_t i = 1; i < instances; i++) p->seeds[i] = (uint32_t)rand();
        p->rl = (rl_state*)calloc(instances, sizeof(rl_state));
        for (uint32_t i = 0; i < instances; i++) rl_seed(&p->rl[i], (int)p->seeds[i]);
    } else if (generator == ORC_PRNG_XOR128) {
        /* random.cpp:670-676 then 651-663: table rows 1..table_size-1 all drawn */
        p->n_seed_words = 4 * table_size;
        p->seeds = (uint32_t*)calloc(p->n_seed_words, sizeof(uint32_t));
        for (uint32_t i = 0; i < 4 * table_size; i++) p->seeds[i] = (uint32_t)rand();
        p->st = (uint32_t*)malloc(p->n_seed_words * sizeof(uint32_t));
        memcpy(p->st, p->seeds, p->n_seed_words * sizeof(uint32_t));
    } else if (generator == ORC_PRNG_PM) {
        /* random.cpp:892-895 then 866-877 */
        p->n_seed_words = 4 * table_size;
        p->seeds = (uint32_t*)calloc(p->n_seed_words, sizeof(uint32_t));
        for (uint32_t i = 0; i < 4 * table_size; i++) p->seeds[i] = (uint32_t)(rand() % 2147483647);
        p->st = (uint32_t*)malloc(p->n_seed_words * sizeof(uint32_t));
        memcpy(p->st, p->seeds, p->n_seed_words * sizeof(uint32_t));
    } else if (generator == ORC_PRNG_MRG32K3A) {
        /* no reference host code: seed table = uint4 seed1[instances] then uint4 seed2[instances]
         * (prngcl_mrg32k3a.cl:126-127), every word rand() % 2147483647 + 1 (non-zero, < m2) */
        p->n_seed_words = 8 * instances;
        p->seeds = (uint32_t*)calloc(p->n_seed_words, sizeof(uint32_t));
        for (uint32_t i = 0; i < 8 * instances; i++) p->seeds[i] = (uint32_t)(rand() % 2147483647) + 1u;
        p->st = (uint32_t*)malloc(p->n_seed_words * sizeof(uint32_t));
        memcpy(p->st, p->seeds, p->n_seed_words * sizeof(uint32_t));
    } else if (generator == ORC_PRNG_XOR7) {
        /* random.cpp:967-974 (eight words of the host replica), then 934-950: rows 1 .. 2I-1 drawn, rows 0 and I = the replica's state */
        uint32_t st8[8];
        for (int k = 0; k < 8; k++) st8[k] = (uint32_t)rand();
        p->n_seed_words = 8 * instances;
        p->seeds = (uint32_t*)calloc(p->n_seed_words, sizeof(uint32_t));
        for (uint32_t i = 4; i < 8 * instances; i++) p->seeds[i] = (uint32_t)rand();
        for (int k = 0; k < 4; k++) { p->seeds[k] = st8[k]; p->seeds[4 * (size_t)instances + k] = st8[4 + k]; }
        p->st = (uint32_t*)malloc(p->n_seed_words * sizeof(uint32_t));
        memcpy(p->st, p->seeds, p->n_seed_words * sizeof(uint32_t));
    } else if (generator == ORC_PRNG_RANECU) {
        /* random.cpp:1044-1048, then 1017-1027 */
        uint32_t j1 = (uint32_t)(rand() % 2147483647), j2 = (uint32_t)(rand() % 2147483647);
        p->n_seed_words = 8 * instances;
        p->seeds = (uint32_t*)calloc(p->n_seed_words, sizeof(uint32_t));
        for (uint32_t i = 0; i < 8 * instances; i++) p->seeds[i] = (uint32_t)(rand() % 2147483647);
        p->seeds[0] = j1; p->seeds[4 * (size_t)instances] = j2;
        p->st = (uint32_t*)malloc(p->n_seed_words * sizeof(uint32_t));
        memcpy(p->st, p->seeds, p->n_seed_words * sizeof(uint32_t));
    } else if (generator == ORC_PRNG_RANMAR) {
        /* random.cpp:776-777, then 726-744: seeds4[0] / seeds4[I] = the replica's pair in all four lanes */
        uint32_t s1 = (uint32_t)(rand() % 31328), s2 = (uint32_t)(rand() % 30081);
        p->n_seed_words = 8 * instances;
        p->seeds = (uint32_t*)calloc(p->n_seed_words, sizeof(uint32_t));
        for (int k = 0; k < 4; k++) { p->seeds[k] = s1; p->seeds[4 * (size_t)instances + k] = s2; }
        for (uint32_t i = 1; i < instances; i++) {
            for (int k = 0; k < 4; k++) p->seeds[4 * (size_t)i + k] = (uint32_t)(rand() % 31328);
            for (int k = 0; k < 4; k++) p->seeds[4 * ((size_t)instances + i) + k] = (uint32_t)(rand() % 30081);
        }
        p->rm = (rm_state*)calloc(instances, sizeof(rm_state));
        for (uint32_t i = 0; i < instances; i++) {
            float lane[97];
            for (int k = 0; k < 4; k++) {
                ranmar_seed_lane(p->seeds[4 * (size_t)i + k], p->seeds[4 * ((size_t)instances + i) + k], lane);
                for (int n = 0; n < 97; n++) p->rm[i].u[n][k] = lane[n];
            }
            p->rm[i].i97 = 96; p->rm[i].j97 = 32; p->rm[i].c = 362436.0f / 16777216.0f;
        }
    } else {
        free(p); return NULL;
    }
    return p;
}
void orc_prng_destroy(orc_prng* p) {
    if (!p) return;
    free(p->rl); free(p->seeds); free(p->st); free(p->rm); free(p);
}
const uint32_t* orc_prng_seeds(const orc_prng* p, uint32_t* n_words) {
    if (n_words) *n_words = p->n_seed_words;
    return p->seeds;
}
/* kernels ranlux / xor128 / pm / mrg32k3a: element (GID + i*GID_SIZE) of the PRN buffer */
void orc_prng_produce(orc_prng* p, uint32_t samples, float* out) {
    uint32_t I = p->instances;
    for (uint32_t g = 0; g < I; g++) {
        for (uint32_t i = 0; i < samples; i++) {
            float* o = out + 4 * ((size_t)g + (size_t)i * I);
            if (p->nskip >= 0) {
                for (int k = 0; k < 4; k++) o[k] = rl_one(&p->rl[g], p->nskip);
            } else if (p->generator == ORC_PRNG_XOR128) {
                uint32_t* s = p->st + 4 * g;
                for (int k = 0; k < 4; k++) o[k] = ((float)xor128_step(s)) / 4294967296.0f;
            } else if (p->generator == ORC_PRNG_PM) {
                uint32_t* s = p->st + 4 * g;
                for (int k = 0; k < 4; k++) { s[k] = pm_step(s[k]); o[k] = ((float)s[k]) / 2147483648.0f; }
            } else if (p->generator == ORC_PRNG_XOR7) {
                uint32_t s8[8];
                memcpy(s8, p->st + 4 * g, 16); memcpy(s8 + 4, p->st + 4 * ((size_t)I + g), 16);
                for (int k = 0; k < 4; k++) o[k] = ((float)xor7_step(s8)) / 4294967296.0f;
                memcpy(p->st + 4 * g, s8, 16); memcpy(p->st + 4 * ((size_t)I + g), s8 + 4, 16);
            } else if (p->generator == ORC_PRNG_RANECU) {
                uint32_t* a = p->st + 4 * g;
                uint32_t* b = p->st + 4 * ((size_t)I + g);
                for (int k = 0; k < 4; k++) o[k] = ((float)ranecu_step(a + k, b + k)) / 2147483648.0f;
            } else if (p->generator == ORC_PRNG_RANMAR) {
                ranmar_step4(&p->rm[g], o);
            } else {
                uint32_t* s1 = p->st + 4 * g;
                uint32_t* s2 = p->st + 4 * ((size_t)I + g);
                for (int k = 0; k < 4; k++) o[k] = mrg_step(s1, s2);
            }
        }
    }
}
/* scalar host replicas of instance 0 -- the streams PRNG::check() compares against
 * (random.cpp:339-396) */
void orc_prng_host_stream(int generator, uint32_t randseries, uint32_t n, float* out) {
    srand(randseries);
    int nskip = rl_nskip_of(generator);
    if (nskip >= 0) {
        rl_state s; rl_seed(&s, rand() % 2147483647);
        for (uint32_t i = 0; i < n; i++) out[i] = rl_one(&s, nskip);
    } else if (generator == ORC_PRNG_XOR128) {
        uint32_t s[4]; for (int k = 0; k < 4; k++) s[k] = (uint32_t)rand();
        for (uint32_t i = 0; i < n; i++) out[i] = ((float)xor128_step(s)) / 4294967295.0f;
    } else if (generator == ORC_PRNG_PM) {
        /* random.cpp:896-910: signed arithmetic, double division */
        int seed = rand() % 2147483647;
        for (uint32_t i = 0; i < n; i++) {
            unsigned int hi = (unsigned int)(seed / 127773), lo = (unsigned int)(seed % 127773);
            int test = (int)(16807u * lo - 2836u * hi);
            seed = (test > 0) ? test : test + 2147483647;
            out[i] = (float)(((double)seed) / 2147483647.0);
        }
    } else if (generator == ORC_PRNG_XOR7) {
        /* random.cpp:965-986 */
        uint32_t s8[8]; for (int k = 0; k < 8; k++) s8[k] = (uint32_t)rand();
        for (uint32_t i = 0; i < n; i++) out[i] = (float)((float)xor7_step(s8) / 4294967296.0);
    } else if (generator == ORC_PRNG_RANECU) {
        /* random.cpp:1044-1068: signed arithmetic */
        int j1 = rand() % 2147483647, j2 = rand() % 2147483647;
        for (uint32_t i = 0; i < n; i++) {
            int k = j1 / 53668;
            j1 = 40014 * (j1 - k * 53668) - k * 12211;
            if (j1 < 0) j1 += 2147483563;
            k = j2 / 52774;
            j2 = 40692 * (j2 - k * 52774) - k * 3791;
            if (j2 < 0) j2 += 2147483399;
            int z = j1 - j2;
            if (z < 1) z += 2147483562;
            out[i] = (float)((float)z) / ((float)2147483648.0);
        }
    } else if (generator == ORC_PRNG_RANMAR) {
        /* random.cpp:773-816: one lane */
        uint32_t s1 = (uint32_t)(rand() % 31328), s2 = (uint32_t)(rand() % 30081);
        float u[97];
        ranmar_seed_lane(s1, s2, u);
        int i97 = 96, j97 = 32;
        float c = (float)(362436.0 / 16777216.0);
        for (uint32_t i = 0; i < n; i++) {
            float uni = u[i97] - u[j97];
            if (uni < 0.0) uni = uni + 1.0f;
            u[i97] = uni;
            if (i97 == 0) i97 = 97;
            if (j97 == 0) j97 = 97;
            i97--; j97--;
            c = (float)c - (float)(7654321.0 / 16777216.0);
            if (c < 0.0) c = (float)c + (float)(16777213.0 / 16777216.0);
            uni = uni - c;
            if (uni < 0.0) uni = uni + 1.0f;
            out[i] = uni;
        }
    }
}
