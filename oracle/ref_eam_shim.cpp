// oracle/ref_eam_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry point around the reference's own EAM class (compiled from /root/reference, see
// oracle/Makefile).  A fresh EAM object per call: the class caches cell lists between calls
// (EAM.cpp:44-48,92-107) and the cached branch is dead weight for the result.
#include "eon/potentials/EAM/EAM.h"

extern "C" int ref_eam_force(long n, const double *R, const int *Z, double *F, double *U, const double *box) {
  try {
    Parameters params;
    EAM pot(params);
    double var = 0.0;
    pot.force(n, R, Z, F, U, &var, box);
    return 0;
  } catch (int code) {
    return code;
  } catch (...) {
    return -1;
  }
}
