// TEST INFRASTRUCTURE ONLY. src/functions.h:102 (SHAKE, only with -G V).
#ifndef NPSL_SHIM_GSL_POLY_H
#define NPSL_SHIM_GSL_POLY_H
#include <cstdio>
#include <cstdlib>
inline int gsl_poly_solve_cubic(double, double, double, double *, double *, double *) {
    std::fprintf(stderr, "oracle shim: gsl_poly_solve_cubic reached (-G V is a 'next' row)\n");
    std::abort();
}
#endif
