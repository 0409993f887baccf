/* on_oracle.c -- see on_oracle.h.  TEST INFRASTRUCTURE ONLY. */
#include "on_oracle.h"
#include <math.h>
#include <stddef.h>

static inline uint32_t on_sites(const orc_on_cfg* c) { return c->n1 * c->n2 * c->n3 * c->n4; }
/* lattice_gid_to_coords / lattice_coords_to_gid: O(N)/kernel/neighbours.cl:40-66; x[] = (x, y, z, t) */
static inline void on_coords(const orc_on_cfg* c, uint32_t gdi, uint32_t x[4]) {
    const uint32_t n23 = c->n2 * c->n3, n234 = n23 * c->n4;
    uint32_t z4 = gdi / n234, z1 = gdi - z4 * n234;
    uint32_t z3 = z1 / n23; z1 -= z3 * n23;
    uint32_t z2 = z1 / c->n2; z1 -= z2 * c->n2;
    x[0] = z4; x[1] = z1; x[2] = z2; x[3] = z3;
}
static inline uint32_t on_gid(const orc_on_cfg* c, const uint32_t x[4]) {
    return x[1] + x[2] * c->n2 + x[3] * c->n2 * c->n3 + x[0] * c->n2 * c->n3 * c->n4;
}
static inline uint32_t on_plus(const orc_on_cfg* c, const uint32_t x[4], int dir) {
    const uint32_t n[4] = { c->n1, c->n2, c->n3, c->n4 };
    uint32_t y[4] = { x[0], x[1], x[2], x[3] };
    y[dir]++; if (y[dir] >= n[dir]) y[dir] = 0;
    return on_gid(c, y);
}
/* lattice_even_gid / lattice_odd_gid: neighbours.cl:93-117 */
static inline uint32_t on_even_gid(const orc_on_cfg* c, uint32_t gid) {
    uint32_t gde = 2 * gid, x[4];
    on_coords(c, gde, x);
    return gde + ((x[0] + x[1] + x[2] + x[3]) & 1u);
}
static inline uint32_t on_odd_gid(const orc_on_cfg* c, uint32_t gid) {
    uint32_t gde = 2 * gid + 1, x[4];
    on_coords(c, gde, x);
    return gde - (((x[0] + x[1] + x[2] + x[3]) & 1u) ^ 1u);
}
uint32_t orc_on_even_gid(const orc_on_cfg* c, uint32_t gid) { return on_even_gid(c, gid); }
uint32_t orc_on_odd_gid(const orc_on_cfg* c, uint32_t gid) { return on_odd_gid(c, gid); }

void orc_on_params(double z, double lambda, double zeta, double eta, double b, double* p) {
    double max_u;
    if (eta > 0.0) {
        double u1 = 1.0 - exp((1.0 - sqrt(1.0 + 4.0 * sqrt(2.0) * eta / sqrt(b))) / eta);
        double u2 = 1.0 - exp((1.0 + sqrt(1.0 + 4.0 * sqrt(2.0) * eta / sqrt(b))) / eta);
        if (u1 < 0.0) u1 = u2;
        max_u = u1;
    } else {
        max_u = 1.0 - exp(-1.9999999999 / sqrt(b));
    }
    p[16] = z; p[17] = lambda; p[18] = zeta; p[19] = eta; p[20] = b;
    p[21] = 0.25 * sqrt(z / (lambda * zeta));
    p[22] = max_u;
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL double
#define FN(name) CAT(name, _f64)
#define LOGR(x) log(x)
#define EXPR(x) exp(x)
#include "on_oracle.inc"
#undef REAL
#undef FN
#undef LOGR
#undef EXPR

#define REAL float
#define FN(name) CAT(name, _f32)
#define LOGR(x) logf(x)
#define EXPR(x) expf(x)
#include "on_oracle.inc"
