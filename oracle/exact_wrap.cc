/*
 * exact_wrap.cc -- TEST INFRASTRUCTURE ONLY.
 * C-callable wrappers around the reference's own exact solutions, compiled UNMODIFIED from
 * /root/reference/src/verification/tests/{exactTestsABCD.c,exactTestsFG.cc} into
 * oracle/_ref/libpism_exact.so by oracle/Makefile (target `ref`).  Used to (a) validate the
 * numpy restatements in pism_b200/verification.py and (b) generate tests/golden fixtures.
 */
#include <vector>

#include "exactTestsABCD.h"
#include "exactTestsFG.hh"

extern "C" {

int ref_exactC(double t, double r, double *H, double *M) {
  struct TestABCDParameters P = exactC(t, r);
  *H = P.H;
  *M = P.M;
  return P.error_code;
}

int ref_exactB(double t, double r, double *H, double *M) {
  struct TestABCDParameters P = exactB(t, r);
  *H = P.H;
  *M = P.M;
  return P.error_code;
}

/* exactFG(t, r, z[0..Mz), Cp): outputs H, M and the five columns. */
int ref_exactFG(double t, double r, int Mz, const double *z, double Cp, double *H, double *M, double *T, double *U,
                double *w, double *Sig, double *Sigc) {
  try {
    std::vector<double> zz(z, z + Mz);
    pism::TestFGParameters P = pism::exactFG(t, r, zz, Cp);
    *H = P.H;
    *M = P.M;
    for (int k = 0; k < Mz; ++k) {
      T[k] = P.T[k];
      U[k] = P.U[k];
      w[k] = P.w[k];
      Sig[k] = P.Sig[k];
      Sigc[k] = P.Sigc[k];
    }
  } catch (...) {
    return 1;
  }
  return 0;
}
}
