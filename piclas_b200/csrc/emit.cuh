// emit.cuh — initial particle emission on the device for the two lattice position types of the PIC-Poisson tutorials:
//   sin_deviation     SetParticlePositionSinDeviation     particle_emission_tools.f90:1235-1296
//   cos_distribution  SetParticlePositionCosDistribution  particle_emission_tools.f90:1299-1371
// followed by the localisation of every position (SetParticlePosition -> SinglePointToElement(doHALO=F),
// particle_position_and_velocity.f90:434, particle_localization.f90:81-190) and a constant velocity (velocityDistribution =
// constant: VeloIC * VeloVecIC).  One thread per lattice point; the staged records take the same way into the particle arrays
// as an upload from the host (k_aos_to_soa, sort by element).
#pragma once
#include "ref.cuh"

constexpr int EMIT_SIN_DEVIATION = 1, EMIT_COS_DISTRIBUTION = 2;

struct EmitSpec {
  int kind;
  int nx, ny, nz;                 // maxParticleNumberX / Y / Z
  double amplitude, wavenumber;   // Part-Species-Init-Amplitude / WaveNumber
  double velo[3];
  double lo[3], len[3];           // GEO%x/y/zminglob, ABS(max - min)
  int species;                    // 1-based
  int firstLocal, lastLocal;      // offsetElem + 1, offsetElem + nElems
};

// SinglePointToElement (particle_localization.f90:81-190) for TriaTracking (ParticleInsideQuad3D) and RefMapping
// (MAXVAL(ABS(xi)) <= ElemEpsOneCell; doEmission_opt is absent at the call of the emission).  doHALO = F: elements of other ranks
// are passed over.  Returns the global element id or -1.
// G: element tables in global memory (device) / plain host arrays (tests/device_track_host.cpp).
template <bool REF, bool G = true>
__device__ int single_point_to_element(const RefTables& T, const TriaElem* __restrict__ tria, const double x[3], int firstLocal, int lastLocal) {
  int Cell[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    Cell[d] = (int)ceil((x[d] - cst.xyzminglob[d]) / cst.FIBGMdeltas[d]);
    Cell[d] = max(min(cst.FIBGMmax[d], Cell[d]), cst.FIBGMmin[d]);
  }
  const int ni = cst.FIBGMmax[0] - cst.FIBGMmin[0] + 1, nj = cst.FIBGMmax[1] - cst.FIBGMmin[1] + 1;
  const size_t cell = (size_t)(Cell[0] - cst.FIBGMmin[0]) + (size_t)ni * ((size_t)(Cell[1] - cst.FIBGMmin[1]) + (size_t)nj * (size_t)(Cell[2] - cst.FIBGMmin[2]));
  const int nB = T.FIBGM_nElems[cell];
  // candidates within reach (squared distance to the barycentre <= ElemRadius2NGeo), nearest first, equal distances in list order
  // (InsertionSort, utils.f90:52-101, is stable)
  const LocateKey key{T, (size_t)T.FIBGM_offsetElem[cell], x};
  SortedVisit sv;
  for (int i = next_in_sorted_order(nB, key, -1., sv); i >= 0; i = next_in_sorted_order(nB, key, -1., sv)) {
    const int e = T.FIBGM_Element[T.FIBGM_offsetElem[cell] + i];
    if (e < firstLocal || e > lastLocal) continue;
    bool in;
    if (REF) {
      double xi[3];
      // GetPositionInRefElem without ForceMode aborts the reference run when the Newton iteration leaves the element by more
      // than 1.5; here such a candidate is not accepted
      const int rc = position_in_ref_elem(T.geo + (e - 1), x, xi, false, true);
      in = (rc & 1) != 0 && maxabs3(xi) <= T.ElemEpsOneCell[e - 1];
    } else {
      uint32_t mask;
      in = inside_quad3d_mask<G>(tria + (e - 1), x, mask);
    }
    if (in) return e;
  }
  return -1;
}

// position number t of the reference's loop nest (i outermost, k innermost), expressions in the order they are written there
__device__ __forceinline__ void emit_lattice_position(const EmitSpec& s, int64_t t, double x[3]) {
  const int kz = (int)(t % s.nz) + 1, jy = (int)((t / s.nz) % s.ny) + 1, ix = (int)(t / ((int64_t)s.nz * s.ny)) + 1;
  const double x_step = s.len[0] / (double)s.nx, y_step = s.len[1] / (double)s.ny, z_step = s.len[2] / (double)s.nz;
  if (s.kind == EMIT_SIN_DEVIATION) {
    const double pilen = 2.0 * 3.141592653589793238 / s.len[0];
    const double x_pos = ((double)ix * x_step - x_step * 0.5);
    x[0] = (s.lo[0] + x_pos) + s.amplitude * sin((s.wavenumber * pilen) * x_pos);
  } else {
    // inverse of the cumulative distribution F(x) = x + a / w sin(w x) by Newton's method (:1341-1345)
    const double a = s.amplitude, w = s.wavenumber;
    const double x_uniform = ((double)ix * x_step - x_step * 0.5);
    double x_pos = x_uniform;
    int guard = 0;
    while (fabs((x_pos + (a / w) * sin(w * x_pos)) - x_uniform) > 1.e-12 && ++guard < 1000)
      x_pos = x_pos - ((x_pos + (a / w) * sin(w * x_pos)) - x_uniform) / (1. + a * cos(w * x_pos));
    x[0] = s.lo[0] + x_pos;
  }
  x[1] = (s.lo[1] + (double)jy * y_step) - y_step * 0.5;
  x[2] = (s.lo[2] + (double)kz * z_step) - z_step * 0.5;
}

template <bool REF>
__global__ void k_emit_lattice(RefTables T, const TriaElem* __restrict__ tria, EmitSpec s, int64_t t0, int64_t m, double* __restrict__ ps,
                               int32_t* __restrict__ spec, int32_t* __restrict__ elem, int32_t* __restrict__ inside, int32_t* __restrict__ isnew,
                               int64_t* __restrict__ ids, unsigned long long* __restrict__ nAccepted) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t t = t0 + i;
  double x[3];
  emit_lattice_position(s, t, x);
  const int e = single_point_to_element<REF>(T, tria, x, s.firstLocal, s.lastLocal);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    ps[i * 6 + d] = x[d];
    ps[i * 6 + 3 + d] = s.velo[d];
  }
  spec[i] = s.species;
  elem[i] = e > 0 ? e : 1;
  inside[i] = e > 0 ? 1 : 0;   // LocateParticleInElement: ElemID = -1 -> RemoveParticle
  isnew[i] = 1;                // PDM%isNewPart = T (particle_localization.f90:72)
  if (ids) ids[i] = t;
  if (e > 0) atomicAdd(nAccepted, 1ull);
}
