/* qcd_oracle.h -- CPU restatement of QCDGPU's gauge-update hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or
 * executed by the product (qcdgpu_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liboracle.so, and only as the checker / the CPU baseline.
 *
 * Parity status: PINNED.
 *   - lattice arithmetic (site indexing, SoA layout, staple, Cabibbo-Marinari /
 *     Kennedy-Pendleton update, plaquette, action) is pinned against the
 *     UNMODIFIED reference compiled in its CPU_RUN configuration
 *     (oracle/_ref/QCDGPU_cpu, recipe in oracle/Makefile) through the
 *     deterministic INTS=2 start: tests/golden/ref_cpurun_*.json.
 *   - RANLUX / Park-Miller streams are pinned against the reference's host
 *     replicas as executed by that same binary (tests/golden/ref_prng_*.json);
 *     XOR128 against the O(N)-tree host replica restated here and the
 *     xorshift128 known answer; MRG32k3a (no host code in the reference) against
 *     L'Ecuyer's published known answer -- "parity unpinned by the reference".
 *   - Polyakov loop, twist (TBC) and the hot start have no CPU_RUN counterpart
 *     (reference suncl.cpp:763-766); they are restated from the .cl sources and
 *     pinned only by invariants (gauge invariance, unitarity, cold-start values).
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef QCD_ORACLE_H
#define QCD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Compile-time physics of one OpenCL program, i.e. the -D option string
 * assembled in SU(N)/suncl/suncl.cpp:3274-3299. */
typedef struct orc_cfg {
    int32_t  sun;          /* SUN: 2 or 3 */
    uint32_t n1, n2, n3, n4;
    uint32_t rowsize;      /* ROWSIZE  (suncl.cpp:2272) */
    uint32_t prngstep;     /* PRNGSTEP (suncl.cpp:2273,3301) */
    int32_t  nhit;         /* NHIT */
    int32_t  nhitpar;      /* NHITPar */
    int32_t  tbc;          /* -D TBC present (PHI or OMEGA != 0) */
    int32_t  gid_upd;      /* -D GID_UPD (INTS=2): 0 off, 1 OpenCL flavour,
                              2 CPU_RUN flavour (pseudo-randoms rounded to float,
                              suncpp/su2/update_su2.cpp:71-100) */
    int32_t  biglat;       /* -D BIGLAT: x carries halo slices, no wrap in +x */
    uint32_t left_sites;   /* LEFT_SITES */
    uint32_t full_sites;   /* FULL_SITES */
    int32_t  pl;           /* PL level for the Polyakov kernel */
} orc_cfg;

/* ---- lattice kernels, fp64 table (PRECISION=1) and fp32 table (PRECISION=0/2).
 * `table` is the device-layout SoA table: element (gindex + plane*ROWSIZE),
 * 4 reals each (su3_matrix_memory.cl:43-55).  `params` = lattice_parameters
 * [0]=beta/N [1]=phi [2]=omega.  `prns` = float4 PRN buffer. */
#define ORC_DECL(SFX, REAL)                                                              \
    void orc_init_cold_##SFX(const orc_cfg*, REAL* table);                               \
    void orc_init_gid_##SFX(const orc_cfg*, REAL* table);                                \
    void orc_init_hot_##SFX(const orc_cfg*, REAL* table, const float* prns, int dir);    \
    void orc_gramschmidt_##SFX(const orc_cfg*, REAL* table);                             \
    void orc_update_##SFX(const orc_cfg*, REAL* table, const REAL* params,               \
                          const float* prns, int odd, int dir);                          \
    void orc_update_range_##SFX(const orc_cfg*, REAL* table, const REAL* params,         \
                          const float* prns, int odd, int dir, uint32_t g0, uint32_t g1);\
    void orc_overrelax_##SFX(const orc_cfg*, REAL* table, int odd, int dir);             \
    void orc_measure_action_##SFX(const orc_cfg*, const REAL* table, const REAL* params, \
                                  double out[2]);                                        \
    void orc_measure_plq_##SFX(const orc_cfg*, const REAL* table, const REAL* params,    \
                               double out[2]);                                           \
    void orc_polyakov_##SFX(const orc_cfg*, const REAL* table, const REAL* params,       \
                            double out[4]);                                              \
    double orc_link_retrace_ustaple_##SFX(const orc_cfg*, const REAL* table,             \
                                          const REAL* params, uint32_t gindex, int dir);
ORC_DECL(f64, double)
ORC_DECL(f32, float)
#undef ORC_DECL

/* ---- index maps (integer, bit-exact): sun_common.cl:41-66, 351-380 */
uint32_t orc_even_gid(const orc_cfg*, uint32_t gid);
uint32_t orc_odd_gid(const orc_cfg*, uint32_t gid);
void     orc_gid_to_coords(const orc_cfg*, uint32_t gindex, uint32_t xyzt[4]);
uint32_t orc_neighbour(const orc_cfg*, uint32_t gindex, int dir, int sign);

/* ---- PRNG host logic + streams */
enum { ORC_PRNG_XOR128 = 1, ORC_PRNG_RANLUX0 = 2, ORC_PRNG_RANLUX1 = 3, ORC_PRNG_RANLUX2 = 4,
       ORC_PRNG_RANLUX3 = 5, ORC_PRNG_RANLUX4 = 6, ORC_PRNG_RANMAR = 8, ORC_PRNG_PM = 9, ORC_PRNG_XOR7 = 10, ORC_PRNG_RANECU = 11,
       ORC_PRNG_MRG32K3A = 12 };   /* convert_generator_to_uint, random.cpp:108-121 */

typedef struct orc_prng orc_prng;
/* PRNG::initialize (random.cpp:216-277): srand(randseries), per-generator seed
 * tables with glibc rand(); does NOT perform the trailing produce(). */
orc_prng* orc_prng_create(int generator, uint32_t randseries, uint32_t instances, uint32_t seed_table_size);
void      orc_prng_destroy(orc_prng*);
/* one PRNG::produce(): every instance emits `samples` float4; out[GID + i*instances] */
void      orc_prng_produce(orc_prng*, uint32_t samples, float* out);
/* raw per-instance input seeds as handed to the device (for seed-table checks) */
const uint32_t* orc_prng_seeds(const orc_prng*, uint32_t* n_words);
/* scalar host replica of instance 0 (RL_produce_one_CPU & friends) */
void      orc_prng_host_stream(int generator, uint32_t randseries, uint32_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
