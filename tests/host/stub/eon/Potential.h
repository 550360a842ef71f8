// tests/host/stub/eon/Potential.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's include/eon/Potential.h:23-123 with the same virtual interface, so that
// eon_b200/host/B200Potential.hpp can be compiled and driven in this image (the real header needs Eigen,
// quill and the Parameters tree).  Signatures and defaults are the reference's.
#pragma once
#define EONB200_HOST_STUB 1
#include <atomic>
#include <cstddef>

namespace eonc {
enum class PotType { UNKNOWN = 0, EMT, EXT_POT, LJ, LJCLUSTER, MORSE_PT, CUH2, TIP4P, TIP4P_PT, TIP4P_H, SPCE, EAM_AL };

class Potential {
protected:
  PotType ptype;

public:
  std::atomic<size_t> forceCallCounter;
  explicit Potential(PotType a_ptype) : ptype{a_ptype}, forceCallCounter{0} {}
  virtual ~Potential() = default;
  void virtual force(long nAtoms, const double *positions, const int *atomicNrs, double *forces, double *energy,
                     double *variance, const double *box) = 0;
  [[nodiscard]] PotType getType() const { return this->ptype; }
  [[nodiscard]] virtual bool isThreadSafe() const noexcept { return true; }
  [[nodiscard]] virtual bool isSharedInstanceThreadSafe() const noexcept { return isThreadSafe(); }
  [[nodiscard]] virtual bool needsPerImageInstance() const noexcept { return false; }
  [[nodiscard]] virtual bool supportsBatchEvaluation() const noexcept { return false; }
  virtual void forceBatch(long nSystems, long nAtoms, const double *const *positions, const int *const *atomicNrs,
                          double *const *forces, double *energies, double *variances, const double *const *boxes) {
    for (long i = 0; i < nSystems; i++) {
      double var = 0;
      force(nAtoms, positions[i], atomicNrs[i], forces[i], &energies[i], &var, boxes[i]);
      if (variances) variances[i] = var;
      forceCallCounter++;
    }
  }
};
}  // namespace eonc
using eonc::Potential;
