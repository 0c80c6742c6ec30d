/* Stand-in for <gsl/gsl_math.h> (GSL is not installed in this image).  The reference's
 * exact-solution sources include it only for M_PI (src/verification/tests/exactTestsABCD.c:23,
 * exactTestsFG.cc:25). */
#ifndef ORACLE_SHIM_GSL_MATH_H
#define ORACLE_SHIM_GSL_MATH_H
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
#endif
