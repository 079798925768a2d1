// TEST INFRASTRUCTURE ONLY. src/functions.h:102 (SHAKE, reached with -G V) calls gsl_poly_solve_cubic.
// GSL is not in the image; this is a restatement of the algorithm GSL documents for that function (the
// trigonometric / Cardano solution of x^3 + a x^2 + b x + c = 0 in the form of Numerical Recipes 5.6):
// real roots returned in ascending order, return value = number of real roots (1 or 3).
#ifndef NPSL_SHIM_GSL_POLY_H
#define NPSL_SHIM_GSL_POLY_H
#include <cmath>
inline int gsl_poly_solve_cubic(double a, double b, double c, double *x0, double *x1, double *x2) {
    const double q = a * a - 3 * b;
    const double r = 2 * a * a * a - 9 * a * b + 27 * c;
    const double Q = q / 9, R = r / 54;
    const double Q3 = Q * Q * Q, R2 = R * R;
    const double CR2 = 729 * r * r, CQ3 = 2916 * q * q * q;
    if (R == 0 && Q == 0) {
        *x0 = *x1 = *x2 = -a / 3;
        return 3;
    }
    if (CR2 == CQ3) {  // two of the three real roots coincide
        const double sqrtQ = std::sqrt(Q);
        if (R > 0) {
            *x0 = -2 * sqrtQ - a / 3;
            *x1 = *x2 = sqrtQ - a / 3;
        } else {
            *x0 = *x1 = -sqrtQ - a / 3;
            *x2 = 2 * sqrtQ - a / 3;
        }
        return 3;
    }
    if (R2 < Q3) {  // three distinct real roots
        const double sgnR = R >= 0 ? 1 : -1;
        const double ratio = sgnR * std::sqrt(R2 / Q3);
        const double theta = std::acos(ratio);
        const double norm = -2 * std::sqrt(Q);
        double r0 = norm * std::cos(theta / 3) - a / 3;
        double r1 = norm * std::cos((theta + 2.0 * M_PI) / 3) - a / 3;
        double r2 = norm * std::cos((theta - 2.0 * M_PI) / 3) - a / 3;
        if (r0 > r1) { double t = r0; r0 = r1; r1 = t; }
        if (r1 > r2) {
            double t = r1; r1 = r2; r2 = t;
            if (r0 > r1) { t = r0; r0 = r1; r1 = t; }
        }
        *x0 = r0; *x1 = r1; *x2 = r2;
        return 3;
    }
    const double sgnR = R >= 0 ? 1 : -1;
    const double A = -sgnR * std::pow(std::fabs(R) + std::sqrt(R2 - Q3), 1.0 / 3.0);
    const double B = Q / A;
    *x0 = A + B - a / 3;
    return 1;
}
#endif
