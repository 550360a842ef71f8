// tests/host/harness.cpp -- TEST INFRASTRUCTURE: drives eon_b200/host/B200Potential.hpp the way eonclient does.
//
//   harness <lib> <kind> <nSystems> <nAtoms> [neb] < input > output
// input (text): for every system 3N positions, then N atomic numbers, then 9 box values.
// It calls forceBatch with arrays of per-system pointers exactly like client/ImprovedDimer.cpp:176 /
// client/NudgedElasticBand.cpp:515, then force() on system 0 like client/Matter.cpp:523, and prints
// energies and forces with %.17g.  With "neb" the systems are the images of one band (both endpoints included) and
// nebBandForces replaces the two calls: prints the climbing image, the energies and the projected forces.
// With "bench <steps> [pin] [conservative]" it times forceBatch instead: the systems read are jittered into four
// position sets (+-0.05 A, lists stay valid), every system in its own heap block like the images of a dimer or a band
// (client/Matter.cpp holds one AtomMatrix per Matter), three warm-up calls, then <steps> calls cycling the sets;
// prints "e2e_ms_per_step <mean> <min>" (wall clock around the plugin call: host buffers in, host buffers out).
// Exit code 3 = potential could not be created (no GPU).
#include "B200Potential.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  const std::string lib = argv[1];
  const int kind = std::atoi(argv[2]);
  const long S = std::atol(argv[3]), N = std::atol(argv[4]);
  std::vector<std::vector<double>> R(S, std::vector<double>(3 * N)), F(S, std::vector<double>(3 * N, -777.0)),
      B(S, std::vector<double>(9));
  std::vector<std::vector<int>> Z(S, std::vector<int>(N));
  for (long s = 0; s < S; s++) {
    for (auto &x : R[s]) if (scanf("%lf", &x) != 1) return 2;
    for (auto &z : Z[s]) if (scanf("%d", &z) != 1) return 2;
    for (auto &b : B[s]) if (scanf("%lf", &b) != 1) return 2;
  }
  std::shared_ptr<Potential> pot;
  int repeat = 1;  // "check - pin repeat=K": the batch is evaluated K times (later calls find the lists in place and, with
                   // pin, take the direct path for systems scattered over the heap) and the forces are poisoned in between
  const bool bench = argc > 6 && std::string(argv[5]) == "bench";
  try {
    const eonc::PotType tag = kind == EONB200_POT_EAM ? eonc::PotType::EAM_AL
                              : kind == EONB200_POT_MORSE_PT ? eonc::PotType::MORSE_PT : eonc::PotType::LJ;
    eonc::B200Options opt;
    opt.library = lib;
    for (int k = 7; k < argc; k++) {
      if (!strncmp(argv[k], "repeat=", 7)) repeat = std::max(1, std::atoi(argv[k] + 7));
      if (!strcmp(argv[k], "pin")) opt.pinHostBuffers = true;
      if (!strcmp(argv[k], "conservative")) opt.eamConservative = true;
      if (!strncmp(argv[k], "device=", 7)) opt.device = std::atoi(argv[k] + 7);
    }
    pot = std::make_shared<eonc::B200Potential>(tag, kind, opt);
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  if (!pot->supportsBatchEvaluation() || pot->isSharedInstanceThreadSafe()) return 4;
  std::vector<const double *> pR(S), pB(S);
  std::vector<const int *> pZ(S);
  std::vector<double *> pF(S);
  for (long s = 0; s < S; s++) {
    pR[s] = R[s].data();
    pZ[s] = Z[s].data();
    pF[s] = F[s].data();
    pB[s] = B[s].data();
  }
  std::vector<double> E(S), V(S, -1.0);
  if (bench) {
    const int steps = std::max(1, std::atoi(argv[6]));
    constexpr int kSets = 4;
    std::vector<std::vector<std::vector<double>>> sets(kSets, R);
    unsigned long long lcg = 88172645463325252ULL;
    for (int q = 1; q < kSets; q++)
      for (auto &sys : sets[q])
        for (auto &x : sys) {
          lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
          x += (static_cast<double>(lcg >> 11) / 9007199254740992.0 - 0.5) * 0.1;
        }
    std::vector<std::vector<const double *>> pRs(kSets, std::vector<const double *>(S));
    for (int q = 0; q < kSets; q++)
      for (long s = 0; s < S; s++) pRs[q][s] = sets[q][s].data();
    for (int k = 0; k < 3 + kSets; k++)
      pot->forceBatch(S, N, pRs[k % kSets].data(), pZ.data(), pF.data(), E.data(), V.data(), pB.data());
    double sum = 0.0, best = 1e300;
    for (int k = 0; k < steps; k++) {
      const auto t0 = std::chrono::steady_clock::now();
      pot->forceBatch(S, N, pRs[k % kSets].data(), pZ.data(), pF.data(), E.data(), V.data(), pB.data());
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      sum += ms;
      best = std::min(best, ms);
    }
    printf("e2e_ms_per_step %.6f %.6f\n", sum / steps, best);
    double chk = 0.0;
    for (double e : E) chk += e;
    printf("energy_sum %.12g\n", chk);
    return 0;
  }
  if (argc > 5 && std::string(argv[5]) == "neb") {
    auto *bp = static_cast<eonc::B200Potential *>(pot.get());
    std::vector<double> band(S * 3 * N), proj((S - 2) * 3 * N, -777.0);
    for (long s = 0; s < S; s++) std::copy(R[s].begin(), R[s].end(), band.begin() + s * 3 * N);
    EonB200NebParams prm = eonc::B200Potential::nebDefaults();
    prm.climbing_image = 1;
    int ci = -1;
    bp->nebBandForces(1, S - 2, N, band.data(), Z[0].data(), nullptr, B[0].data(), true, prm, proj.data(), E.data(),
                      nullptr, &ci);
    printf("neb %d calls %zu\n", ci, pot->forceCallCounter.load());
    for (double e : E) printf("%.17g\n", e);
    for (double f : proj) printf("%.17g\n", f);
    return 0;
  }
  try {
    for (int k = 0; k < repeat; k++) {
      for (auto &f : F) std::fill(f.begin(), f.end(), -777.0);
      std::fill(E.begin(), E.end(), -777.0);
      pot->forceBatch(S, N, pR.data(), pZ.data(), pF.data(), E.data(), V.data(), pB.data());
    }
  } catch (int code) {
    printf("thrown %d\n", code);
    return 0;
  }
  std::vector<double> F0(3 * N);
  double E0 = 0, var0 = -1;
  pot->force(N, R[0].data(), Z[0].data(), F0.data(), &E0, &var0, B[0].data());
  pot->forceCallCounter++;  // the caller's increment, client/Matter.cpp:525
  printf("calls %zu\n", pot->forceCallCounter.load());
  for (long s = 0; s < S; s++) {
    printf("%.17g %.17g\n", E[s], V[s]);
    for (double f : F[s]) printf("%.17g\n", f);
  }
  printf("%.17g %.17g\n", E0, var0);
  for (double f : F0) printf("%.17g\n", f);
  return 0;
}
