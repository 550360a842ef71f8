// B200Potential.hpp -- the eonclient-side shim: an eonc::Potential subclass over libeonb200.so.
//
// This is the file a maintainer drops into client/potentials/B200/ of the reference tree (see
// INTEGRATION.md).  It follows the reference's own dlopen C-ABI precedent
//     include/eon/potentials/Metatomic/metatomic_c_abi.h:30-79
//     client/potentials/Metatomic/MetatomicDynPot.cpp:39-70   (create -> handle, rc != 0 -> runtime_error)
//     include/eon/potentials/Metatomic/MetatomicLoader.h:18-52 (function-pointer table + ABI version check)
// and overrides exactly the virtuals of include/eon/Potential.h:55-57,62-122 that route callers
// (client/Matter.cpp:523, client/ImprovedDimer.cpp:160-193, client/Dimer.cpp:170-210,
// client/NudgedElasticBand.cpp:475-521) through one batched device evaluation.
//
// It compiles against the real <eon/Potential.h> inside an eOn build, and against the small stand-in
// tests/host/stub/eon/Potential.h in this repository's host harness (the real header needs Eigen,
// quill and the Parameters tree, none of which exist in this image).
#pragma once

#include "eon/Potential.h"

#include "eonb200.h"

#include <dlfcn.h>

#include <array>
#include <mutex>
#include <stdexcept>
#include <string>

namespace eonc {

// Function-pointer table over libeonb200.so (the analogue of MetatomicLoader).
class B200Loader {
public:
  static B200Loader &instance() {
    static B200Loader l;
    return l;
  }
  // Search order: explicit path, $EONB200_LIBRARY, then the dynamic linker's default path
  // (the reference adds `[Potential] potentials_path` / $EON_POTENTIALS_PATH in PluginLoader.cpp).
  void require_loaded(const std::string &path = std::string()) {
    std::lock_guard<std::mutex> g(m_mu);
    if (m_handle) return;
    std::string p = path;
    if (p.empty()) {
      if (const char *env = std::getenv("EONB200_LIBRARY")) p = env;
    }
    if (p.empty()) p = "libeonb200.so";
    m_handle = dlopen(p.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!m_handle) throw std::runtime_error(std::string("B200Loader: cannot load ") + p + ": " + dlerror());
    bind(abi_version, "eonb200_abi_version");
    if (abi_version() != EONB200_ABI_VERSION) {
      dlclose(m_handle);
      m_handle = nullptr;
      throw std::runtime_error("B200Loader: libeonb200.so ABI version mismatch");
    }
    bind(config_default, "eonb200_config_default");
    bind(create, "eonb200_create");
    bind(destroy, "eonb200_destroy");
    bind(last_error, "eonb200_last_error");
    bind(force, "eonb200_force");
    bind(force_batch, "eonb200_force_batch");
    bind(force_calls, "eonb200_force_calls");
    bind(neb_params_default, "eonb200_neb_params_default");
    bind(neb_forces, "eonb200_neb_forces");
    bind(neb_band_forces, "eonb200_neb_band_forces");
  }

  decltype(&eonb200_abi_version) abi_version{nullptr};
  decltype(&eonb200_config_default) config_default{nullptr};
  decltype(&eonb200_create) create{nullptr};
  decltype(&eonb200_destroy) destroy{nullptr};
  decltype(&eonb200_last_error) last_error{nullptr};
  decltype(&eonb200_force) force{nullptr};
  decltype(&eonb200_force_batch) force_batch{nullptr};
  decltype(&eonb200_force_calls) force_calls{nullptr};
  decltype(&eonb200_neb_params_default) neb_params_default{nullptr};
  decltype(&eonb200_neb_forces) neb_forces{nullptr};
  decltype(&eonb200_neb_band_forces) neb_band_forces{nullptr};

private:
  template <typename F>
  void bind(F &fn, const char *name) {
    void *s = dlsym(m_handle, name);
    if (!s) throw std::runtime_error(std::string("B200Loader: missing symbol ") + name);
    fn = reinterpret_cast<F>(s);
  }
  void *m_handle{nullptr};
  std::mutex m_mu;
};

// What a [Potential] section may set beyond the potential's name (INTEGRATION.md lists the config keys).
struct B200Options {
  int device = 0;
  bool pinHostBuffers = false;   // page-lock the caller's position / force blocks on first use (EonB200Config)
  bool eamConservative = false;  // in-tree EAM with F = -grad U instead of the sign of EAM.cpp:203-215
  double skin = -1.0;            // Verlet skin of the cached lists; < 0 keeps the library default
  double cutoff = -1.0;          // LJ / Morse cutoff; < 0 keeps the default
  std::string library;           // path of libeonb200.so; empty: EONB200_LIBRARY or the loader path
};

// eonc::Potential for PotType::{LJ, LJCLUSTER, MORSE_PT} and the in-tree EAM class, evaluated on a B200.
class B200Potential : public Potential {
public:
  // `kind` is an EONB200_POT_* value; `ptype` the reference tag the instance reports through
  // getType() / PotRegistry (include/eon/BaseStructures.h:36-71).
  B200Potential(PotType ptype, int kind, int device = 0, const std::string &library = std::string())
      : B200Potential(ptype, kind, with(device, library)) {}
  B200Potential(PotType ptype, int kind, const B200Options &opt) : Potential(ptype) {
    auto &l = B200Loader::instance();
    l.require_loaded(opt.library);
    EonB200Config cfg;
    l.config_default(&cfg, kind);
    cfg.device = opt.device;
    cfg.pin_host_buffers = opt.pinHostBuffers ? 1 : 0;
    cfg.eam_conservative = opt.eamConservative ? 1 : 0;
    if (opt.skin >= 0.0) cfg.skin = opt.skin;
    if (opt.cutoff > 0.0) cfg.cutoff = opt.cutoff;
    std::array<char, 1024> err{};
    m_handle = l.create(&cfg, err.data(), err.size());
    if (!m_handle) throw std::runtime_error(std::string("B200Potential: create failed: ") + err.data());
  }
  ~B200Potential() override {
    if (m_handle) B200Loader::instance().destroy(m_handle);
  }
  B200Potential(const B200Potential &) = delete;
  B200Potential &operator=(const B200Potential &) = delete;

  // Potential::force (include/eon/Potential.h:55-57).  Counters are the caller's job here
  // (client/Matter.cpp:525-526).
  void force(long nAtoms, const double *positions, const int *atomicNrs, double *forces, double *energy,
             double *variance, const double *box) override {
    check(B200Loader::instance().force(m_handle, nAtoms, positions, atomicNrs, forces, energy, variance, box));
  }

  // Potential::forceBatch (include/eon/Potential.h:108-122): one H2D, one launch sequence, one D2H.
  // Counters are the callee's job in the batch path (Potential.h:119-120, RgpotAdapter.h:91-92).
  void forceBatch(long nSystems, long nAtoms, const double *const *positions, const int *const *atomicNrs,
                  double *const *forces, double *energies, double *variances, const double *const *boxes) override {
    check(B200Loader::instance().force_batch(m_handle, nSystems, nAtoms, positions, atomicNrs, forces, energies,
                                             variances, boxes));
    for (long i = 0; i < nSystems; i++) {
      forceCallCounter++;
#ifndef EONB200_HOST_STUB
      PotRegistry::get().on_force_call(ptype);
#endif
    }
  }

  // Callers test supportsBatchEvaluation() first (ImprovedDimer.cpp:160, NudgedElasticBand.cpp:475) and
  // must never thread one device handle (isSharedInstanceThreadSafe gates std::thread use).
  [[nodiscard]] bool supportsBatchEvaluation() const noexcept override { return true; }
  [[nodiscard]] bool isThreadSafe() const noexcept override { return false; }
  [[nodiscard]] bool isSharedInstanceThreadSafe() const noexcept override { return false; }
  [[nodiscard]] bool needsPerImageInstance() const noexcept override { return false; }

  // ---- beyond eonc::Potential: the NEB band algebra (client/NudgedElasticBand.cpp:551-646) ----------------------
  static EonB200NebParams nebDefaults() {
    EonB200NebParams p;
    B200Loader::instance().neb_params_default(&p);
    return p;
  }
  // Tangent, springs, projection, climbing image of nBands bands from positions / Matter forces / energies
  // ([nBands][nImages + 2][3 nAtoms], [nBands][nImages + 2]) the caller already holds.
  void nebForces(long nBands, long nImages, long nAtoms, long nFreeAtoms, const double *positions, const double *forces,
                 const double *energies, const double *box, const EonB200NebParams &prm, double *projected,
                 double *tangents, int *climbing) {
    check(B200Loader::instance().neb_forces(m_handle, nBands, nImages, nAtoms, nFreeAtoms, positions, forces, energies,
                                            box, &prm, projected, tangents, climbing));
  }
  // The whole updateForces step in one call: potential on every image, Matter::getForces masking, band algebra.
  // Counts one force call per image like forceBatch does (endpoints included: they are re-evaluated).
  void nebBandForces(long nBands, long nImages, long nAtoms, const double *positions, const int *atomicNrs,
                     const int *fixed, const double *box, bool removeNetForce, const EonB200NebParams &prm,
                     double *projected, double *energies, double *tangents, int *climbing) {
    check(B200Loader::instance().neb_band_forces(m_handle, nBands, nImages, nAtoms, positions, atomicNrs, fixed, box,
                                                 removeNetForce ? 1 : 0, &prm, projected, energies, tangents, climbing));
    for (long i = 0; i < nBands * (nImages + 2); i++) {
      forceCallCounter++;
#ifndef EONB200_HOST_STUB
      PotRegistry::get().on_force_call(ptype);
#endif
    }
  }

  [[nodiscard]] EonB200Pot *handle() const noexcept { return m_handle; }

private:
  static B200Options with(int device, const std::string &library) {
    B200Options o;
    o.device = device;
    o.library = library;
    return o;
  }
  void check(int rc) {
    if (rc == 0) return;
    if (rc == 14324) throw 14324;  // client/potentials/EAM/EAM.cpp:116-123
    if (rc == 100) throw 100;      // legacy "potential failed" -> status 11 (client/SaddleSearchJob.cpp:116-123)
    throw std::runtime_error(std::string("B200Potential: rc=") + std::to_string(rc) + ": " +
                             B200Loader::instance().last_error(m_handle));
  }
  EonB200Pot *m_handle{nullptr};
};

}  // namespace eonc
