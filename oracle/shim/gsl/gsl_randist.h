// TEST INFRASTRUCTURE ONLY. src/functions.cpp:29 (thermal velocity init, not on the default path).
#ifndef NPSL_SHIM_GSL_RANDIST_H
#define NPSL_SHIM_GSL_RANDIST_H
#include <gsl/gsl_rng.h>
inline double gsl_ran_gaussian(gsl_rng *, double) {
    std::fprintf(stderr, "oracle shim: gsl_ran_gaussian reached\n");
    std::abort();
}
#endif
