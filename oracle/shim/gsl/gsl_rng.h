// TEST INFRASTRUCTURE ONLY (oracle build). GSL RNG surface used by src/utility.h:55-78.
// The default path (-I n, zero initial velocities) never draws a number; trap if it does.
#ifndef NPSL_SHIM_GSL_RNG_H
#define NPSL_SHIM_GSL_RNG_H
#include <cstdio>
#include <cstdlib>
struct gsl_rng_type { int unused; };
struct gsl_rng { int unused; };
static const gsl_rng_type npsl_shim_rng_type = {0};
static const gsl_rng_type *const gsl_rng_default = &npsl_shim_rng_type;
inline void gsl_rng_env_setup() {}
inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { return new gsl_rng(); }
inline void gsl_rng_free(gsl_rng *r) { delete r; }
inline double gsl_rng_uniform(gsl_rng *) {
    std::fprintf(stderr, "oracle shim: gsl_rng_uniform reached (counterions are out of scope)\n");
    std::abort();
}
#endif
