/* on_oracle.h -- CPU restatement of the O(N) scalar branch of QCDGPU (reference tree O(N)/).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as qcd_oracle.h): only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline legs may load liboracle.so; the product never does.
 *
 * Parity status: PINNED for the measurements (action, field mean / mean square, correlators):
 * tests/golden/ref_on_check_*.json hold the output of the UNMODIFIED reference checker
 * (O(N)/suncl/suncpu.cpp:93-270, 1805-1838, compiled from where it lies by oracle/Makefile
 * target `ref`, harness oracle/ref_harness/on_check.cpp) on seeded tables.  The Metropolis
 * update (O(N)/oncl/oncl.cl:73-370) has no CPU counterpart in the reference: it is restated
 * from the .cl source and shares the pinned action function; XOR128 is pinned by the
 * reference's host replica (O(N)/random/random.cpp:636-665, tests/golden/ref_prng_XOR128_*.json).
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef ON_ORACLE_H
#define ON_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* the -D options of O(N)/suncl/suncl.cpp:1454-1482 that the kernels use */
typedef struct orc_on_cfg {
    uint32_t n1, n2, n3, n4;   /* N1 (== N1EXACT: single part), N2, N3, N4 */
    uint32_t prngstep;         /* PRNGSTEP = align(SITES/2) */
    uint32_t nhit;             /* NHIT */
    uint32_t skgroup;          /* SKGROUP */
} orc_on_cfg;

/* model_lattice_init, O(N)/suncl/suncl.cpp:1244-1253 + 1987-1994: fills params[16..22] (at least 23 doubles) */
void orc_on_params(double z, double lambda, double zeta, double eta, double b, double* params);

uint32_t orc_on_even_gid(const orc_on_cfg*, uint32_t gid);
uint32_t orc_on_odd_gid(const orc_on_cfg*, uint32_t gid);

#define ORC_ON_DECL(SFX, REAL)                                                                         \
    void orc_on_init_cold_##SFX(const orc_on_cfg*, REAL* table);                                       \
    void orc_on_init_hot_##SFX(const orc_on_cfg*, REAL* table, const float* prns, const REAL* params); \
    void orc_on_update_##SFX(const orc_on_cfg*, REAL* table, const REAL* params, const float* prns,    \
                             int odd, double* acc);                                                    \
    void orc_on_measure_action_##SFX(const orc_on_cfg*, const REAL* table, const REAL* params, double out[2]); \
    void orc_on_measure_plq_##SFX(const orc_on_cfg*, const REAL* table, const REAL* params, double out[2]);    \
    void orc_on_measure_corr_##SFX(const orc_on_cfg*, const REAL* table, const REAL* params,           \
                                   const uint32_t step[4], double out[2]);
ORC_ON_DECL(f64, double)
ORC_ON_DECL(f32, float)
#undef ORC_ON_DECL

#ifdef __cplusplus
}
#endif
#endif
