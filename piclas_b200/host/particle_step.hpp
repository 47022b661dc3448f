// particle_step.hpp — compiled host side of the drop-in boundary (C++17, no dependency beyond include/piclas_gpu.h).
//
// The reference's host is compiled Fortran whose time step calls argument-less subroutines on module globals
// (timedisc_TimeStepPoissonByBorisLeapfrog.f90:93-279).  No Fortran compiler exists in the build image, so this header is
// the compiled mirror of that call sequence above the C ABI — one method per reference call site, named after it; the
// ISO_C_BINDING module a PICLas build links is piclas_b200/fortran/mod_particle_gpu.f90 and declares the same entry points.
// Errors: every ABI call returns nonzero on failure; the Fortran glue turns that into CALL abort(__STAMP__, msg)
// (globals.f90:322-397), here it is pgpu::Abort carrying piclas_gpu_last_error().
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "piclas_gpu.h"

namespace pgpu {

struct Abort : std::runtime_error {
  explicit Abort(const std::string& where) : std::runtime_error(where + ": " + piclas_gpu_last_error()) {}
};

class ParticleStep {
 public:
  // after InitParticleMesh + InitializeDeposition (piclaslib.f90:177): the host's tables go to the device once
  ParticleStep(const pgpu_mesh_t& mesh, const pgpu_params_t& params)
      : nDOF_((int64_t)mesh.nElems * (mesh.N + 1) * (mesh.N + 1) * (mesh.N + 1)), nNodes_(mesh.nUniqueGlobalNodes),
        nSpecies_(params.nSpecies), cvwm_(params.DepositionType == PGPU_DEPO_CVWM) {
    if (piclas_gpu_init(&mesh, &params)) throw Abort("piclas_gpu_init");
  }
  ~ParticleStep() { piclas_gpu_finalize(); }   // FinalizePiclas
  ParticleStep(const ParticleStep&) = delete;
  ParticleStep& operator=(const ParticleStep&) = delete;

  int64_t nDOF() const { return nDOF_; }
  int64_t nNodes() const { return nNodes_; }

  // after ParticleRestart / initial ParticleInserting: PartState(1:6,1:n), PartSpecies, PEM%GlobalElemID, PDM%ParticleInside,
  // PDM%IsNewPart as they lie in the host's memory (Fortran order == double[n][6])
  void UploadParticles(int64_t n, const double* PartState, const int32_t* PartSpecies, const int32_t* GlobalElemID,
                       const int32_t* ParticleInside = nullptr, const int32_t* IsNewPart = nullptr, const double* PartPosRef = nullptr,
                       const int64_t* ids = nullptr, bool append = false) {
    if (piclas_gpu_upload_particles(n, PartState, PartSpecies, GlobalElemID, ParticleInside, IsNewPart, PartPosRef, ids, append ? 1 : 0))
      throw Abort("piclas_gpu_upload_particles");
  }
  // CALL Deposition()  (:93) -> PS_N(iElem)%PartSource(1:4,i,j,k) packed element after element, NodeSource for cell_volweight_mean
  void Deposition(double* PartSource, double* NodeSource = nullptr) {
    if (piclas_gpu_deposit(PartSource, cvwm_ ? NodeSource : nullptr)) throw Abort("piclas_gpu_deposit");
  }
  // PartSource(4,:) only: all CalcSourceHDG reads (equations/poisson/equation.f90:1043)
  void ChargeDensity(double* rho) {
    if (piclas_gpu_get_charge(rho)) throw Abort("piclas_gpu_get_charge");
  }
  // after CALL HDG(time,iter)  (:99): U_N(iElem)%E(1:3,i,j,k) packed
  void SetField(const double* E) {
    if (piclas_gpu_set_field(E)) throw Abort("piclas_gpu_set_field");
  }
  // :109-215 + :270  LastPartPos = PartState, InterpolateFieldToParticle, push loop, PerformTracking, UpdateNextFreePosition.
  // Returns NbrOfLostParticles of the step (counted, not fatal: particle_triatracking.f90:302-306).
  int32_t PushAndTrack(double dt, int64_t iter) {
    int32_t nLost = 0;
    if (piclas_gpu_push_track(dt, iter, &nLost)) throw Abort("piclas_gpu_push_track");
    return nLost;
  }
  int64_t NumParticles() const { return piclas_gpu_num_particles(); }   // PDM%ParticleVecLength after compaction
  // before PerformAnalyze / WriteStateToHDF5 (timedisc.f90:399-410): compacted 1..ParticleVecLength, sorted by element
  int64_t DownloadParticles(int64_t nmax, double* PartState, int32_t* PartSpecies, int32_t* GlobalElemID, double* PartPosRef = nullptr,
                            int64_t* ids = nullptr) {
    int64_t nOut = 0;
    if (piclas_gpu_download_particles(nmax, PartState, PartSpecies, GlobalElemID, PartPosRef, ids, &nOut))
      throw Abort("piclas_gpu_download_particles");
    return nOut;
  }
  // PartSource to the host beside the rest of the step (the HDG source needs the charge only); complete after PartSourceWait()
  void PartSourceAsync(double* PartSourcePinned) {
    if (piclas_gpu_get_partsource_async(PartSourcePinned)) throw Abort("piclas_gpu_get_partsource_async");
  }
  void PartSourceWait() {
    if (piclas_gpu_partsource_wait()) throw Abort("piclas_gpu_partsource_wait");
  }
  // CalcKineticEnergy / CalcNumPartsOfSpec (particle_analyze_tools.f90:709-842)
  void KineticEnergy(double* Ekin, int64_t* nPart) {
    if (piclas_gpu_kinetic_energy(Ekin, nPart)) throw Abort("piclas_gpu_kinetic_energy");
  }
  // initial emission with SpaceIC = sin_deviation (1) / cos_distribution (2), velocityDistribution = constant
  // (particle_emission_tools.f90:1235-1371 + SinglePointToElement); returns the particles this rank accepted
  int64_t EmitLattice(int32_t SpaceIC, int32_t iSpec, const int32_t maxParticleNumber[3], double Amplitude, double WaveNumber,
                      const double velocity[3], bool append = false) {
    int64_t n = 0;
    if (piclas_gpu_emit_lattice(SpaceIC, iSpec, maxParticleNumber, Amplitude, WaveNumber, velocity, append ? 1 : 0, &n)) throw Abort("piclas_gpu_emit_lattice");
    return n;
  }
  int nSpecies() const { return nSpecies_; }

  // One pass of TimeStepPoissonByBorisLeapfrog (or TimeStepPoisson with TimeDiscMethod 509): the HDG solve stays with the host and is
  // passed in as HDG(PartSource, E), which reads the packed source and writes the packed field.
  template <class FieldSolve>
  int32_t TimeStepPoissonByBorisLeapfrog(double dt, int64_t iter, FieldSolve&& HDG, std::vector<double>& PartSource, std::vector<double>& E,
                                         double* NodeSource = nullptr) {
    PartSource.resize((size_t)nDOF_ * 4);
    E.resize((size_t)nDOF_ * 3);
    Deposition(PartSource.data(), NodeSource);   // :93
    HDG(PartSource, E);                          // :99
    SetField(E.data());
    return PushAndTrack(dt, iter);               // :109-270
  }

 private:
  int64_t nDOF_, nNodes_;
  int nSpecies_;
  bool cvwm_;
};

}  // namespace pgpu
