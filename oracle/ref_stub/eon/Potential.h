// oracle/ref_stub/eon/Potential.h -- TEST INFRASTRUCTURE ONLY.
//
// Ten-line stand-in for the reference's include/eon/Potential.h so that the reference's own
// client/potentials/EAM/EAM.cpp can be compiled, unmodified and from where it lies under
// /root/reference, into oracle/_ref/libeam_ref.so (recipe: oracle/Makefile).  The real header
// pulls in Eigen, quill and the Parameters tree, none of which exist in this image.  Only what
// EAM.h touches is declared: the base-class constructor, the virtual force() signature
// (include/eon/Potential.h:55-57) and the PotType tag (include/eon/potentials/EAM/EAM.h:29-31).
#pragma once
#include <cstddef>

enum class PotType { EAM_AL };
class Parameters {};

class Potential {
public:
  Potential(PotType, const Parameters &) {}
  virtual ~Potential() = default;
  virtual void force(long nAtoms, const double *positions, const int *atomicNrs, double *forces,
                     double *energy, double *variance, const double *box) = 0;
};
