// run_case.cpp — compiled host driver: runs the particle time step of a case file through the C ABI and writes the result.
//   run_case <case file> <result file>
//   run_case --echo <case file> <copy>      (reads and rewrites the file: format check, needs no device)
// The case file (piclas_b200/casefile.py) holds what the Fortran host owns at run time: the particle-mesh tables, the parsed
// parameters, the particles and the field.  The field solve is the host's (hdg/hdg.f90:734) and is not part of this path: the
// driver hands the same E back after every deposition, as a frozen-field HDG would.
// Exit code 0 on success; 2 with the library's message on stderr when the ABI aborts (no CUDA device, unsupported setup, ...).
#include <cstdio>
#include <exception>
#include <vector>

#include "particle_step.hpp"
#include "pgpu_case.hpp"

int main(int argc, char** argv) {
  if (argc == 4 && std::string(argv[1]) == "--echo") {
    try {
      pgpu::CaseFile::read(argv[2]).write(argv[3]);
      return 0;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "run_case: %s\n", e.what());
      return 2;
    }
  }
  if (argc != 3) {
    std::fprintf(stderr, "usage: run_case <case file> <result file> | run_case --echo <case file> <copy>\n");
    return 1;
  }
  try {
    const pgpu::CaseFile c = pgpu::CaseFile::read(argv[1]);
    pgpu_mesh_t mesh;
    pgpu_params_t prm;
    pgpu::fill_mesh(c, mesh);
    pgpu::fill_params(c, prm);
    const pgpu::Array& PS = c.get("part.PartState", 'd');
    const int64_t n = PS.shape.at(0);
    const double dt = c.get("run.dt", 'd').as<double>()[0];
    const int nsteps = c.get("run.nsteps", 'i').as<int32_t>()[0];
    const bool cvwm = prm.DepositionType == PGPU_DEPO_CVWM && prm.DoDeposition;

    pgpu::ParticleStep step(mesh, prm);
    step.UploadParticles(n, PS.as<double>(), c.get("part.PartSpecies", 'i').as<int32_t>(), c.get("part.GlobalElemID", 'i').as<int32_t>(),
                         nullptr, c.get("part.IsNewPart", 'i').as<int32_t>(), nullptr,
                         c.has("part.ids") ? c.get("part.ids", 'l').as<int64_t>() : nullptr);
    const pgpu::Array& Efile = c.get("field.E", 'd');
    if (Efile.count() != step.nDOF() * 3) throw std::runtime_error("field.E does not have 3 values per degree of freedom");
    std::vector<double> PartSource, E, NodeSource(cvwm ? (size_t)step.nNodes() * 4 : 0);
    int64_t nLost = 0;
    for (int it = 0; it < nsteps; ++it) {
      if (prm.DoDeposition) {
        nLost += step.TimeStepPoissonByBorisLeapfrog(
            dt, it, [&](const std::vector<double>&, std::vector<double>& Eout) { Eout.assign(Efile.as<double>(), Efile.as<double>() + Efile.count()); },
            PartSource, E, cvwm ? NodeSource.data() : nullptr);
      } else {
        step.SetField(Efile.as<double>());
        nLost += step.PushAndTrack(dt, it);
      }
    }
    const int64_t np = step.NumParticles();
    std::vector<double> outPS((size_t)np * 6);
    std::vector<int32_t> outSpec((size_t)np), outElem((size_t)np);
    std::vector<int64_t> outIds(c.has("part.ids") ? (size_t)np : 0);
    const int64_t got = step.DownloadParticles(np, outPS.data(), outSpec.data(), outElem.data(), nullptr, outIds.empty() ? nullptr : outIds.data());
    std::vector<double> Ekin((size_t)step.nSpecies());
    std::vector<int64_t> nPart((size_t)step.nSpecies());
    step.KineticEnergy(Ekin.data(), nPart.data());

    pgpu::CaseFile r;
    r.put("PartState", 'd', {got, 6}, outPS.data());
    r.put("PartSpecies", 'i', {got}, outSpec.data());
    r.put("GlobalElemID", 'i', {got}, outElem.data());
    if (!outIds.empty()) r.put("ids", 'l', {got}, outIds.data());
    if (!PartSource.empty()) r.put("PartSource", 'd', {step.nDOF(), 4}, PartSource.data());
    if (cvwm) r.put("NodeSource", 'd', {step.nNodes(), 4}, NodeSource.data());
    r.put("Ekin", 'd', {(int64_t)Ekin.size()}, Ekin.data());
    r.put("nPart", 'l', {(int64_t)nPart.size()}, nPart.data());
    r.put("nLost", 'l', {1}, &nLost);
    r.write(argv[2]);
    std::printf("run_case: %lld particles, %d steps, %lld lost\n", (long long)got, nsteps, (long long)nLost);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "run_case: %s\n", e.what());
    return 2;
  }
}
