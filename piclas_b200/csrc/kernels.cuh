// kernels.cuh — the particle-step kernels (sm_100a): deposition (cell_volweight_mean) and interpolate+push+track.
//
// Launch geometry: particles are sorted by local element; one CTA works on one element at a time (grid-stride over
// elements, grid = a multiple of the SM count), stages that element's geometry / field tile / corner table in shared
// memory and lets its threads stride over the element's particle segment.  Particle state is SoA FP64, read and
// written fully coalesced.
#pragma once
#include "math.cuh"

constexpr int STEP_NT = 128;  // threads per CTA of the per-element kernels

__device__ __forceinline__ void stage_words(void* dst, const void* src, int nbytes) {
  // cooperative copy of a 16-byte aligned record into shared memory with 128-bit loads
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d = reinterpret_cast<int4*>(dst);
  for (int i = threadIdx.x; i < nbytes / 16; i += blockDim.x) d[i] = __ldg(s + i);
}

// ---- Newton mapping for all particles of an element + optional CVWM accumulation ------------------------------------------
// DepositionMethod_CVWM particle loop, pic_depo_method.f90:471-544.  elemAcc[e][node 0..7 (CGNS)][0..3] receives the
// element-local sums of TSource*weight; xi and the SucRefPos flag are cached for the interpolation of the same step.
__global__ void __launch_bounds__(STEP_NT) k_deposit_cvwm(PartBuf pb, const int64_t* __restrict__ elemOff, int nElems, int offsetElem,
                                                          const GeoElem* __restrict__ geo, const TriaElem* __restrict__ tria,
                                                          double* __restrict__ elemAcc, int* __restrict__ errFlag) {
  __shared__ GeoElem sg;
  __shared__ TriaElem st;
  __shared__ double red[STEP_NT / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int64_t p0 = elemOff[e], p1 = elemOff[e + 1];
    double acc[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) acc[a] = 0.;
    if (p1 > p0) {
      __syncthreads();
      stage_words(&sg, geo + (offsetElem + e), sizeof(GeoElem));
      stage_words(&st, tria + (offsetElem + e), sizeof(TriaElem));
      __syncthreads();
      for (int64_t p = p0 + threadIdx.x; p < p1; p += STEP_NT) {
        const double x[3] = {pb.x[0][p], pb.x[1][p], pb.x[2][p]};
        uint8_t meta = pb.meta[p];
        const int spec = meta & META_SPEC_MASK;
        double xi[3];
        const int r = position_in_ref_elem(&sg, x, xi, true, true);
        const bool suc = (r & 1) != 0;
        pb.xi[0][p] = xi[0];
        pb.xi[1][p] = xi[1];
        pb.xi[2][p] = xi[2];
        meta = suc ? (meta & ~META_XIFAIL) : (meta | META_XIFAIL);
        pb.meta[p] = meta;
        const double q = cst.ChargeIC[spec];
        if (!(fabs(q) > 0.0)) continue;  // isDepositParticle
        const double Charge = q * cst.MPF[spec];
        const double T[4] = {pb.v[0][p] * Charge, pb.v[1][p] * Charge, pb.v[2][p] * Charge, Charge};
        double w[8];
        if (suc) {
          const double a1 = 0.5 * (xi[0] + 1.0), a2 = 0.5 * (xi[1] + 1.0), a3 = 0.5 * (xi[2] + 1.0);
          w[0] = ((1 - a1) * (1 - a2)) * (1 - a3);
          w[1] = ((a1) * (1 - a2)) * (1 - a3);
          w[2] = ((a1) * (a2)) * (1 - a3);
          w[3] = ((1 - a1) * (a2)) * (1 - a3);
          w[4] = ((1 - a1) * (1 - a2)) * (a3);
          w[5] = ((a1) * (1 - a2)) * (a3);
          w[6] = ((a1) * (a2)) * (a3);
          w[7] = ((1 - a1) * (a2)) * (a3);
#pragma unroll
          for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[n * 4 + c] = acc[n * 4 + c] + (T[c] * w[n]);
        } else {
          // inverse-distance fallback, :512-538.  CGNS corner n is tensor node cns[n]
          const int cns[8] = {0, 1, 3, 2, 4, 5, 7, 6};
          bool hit = false;
          for (int n = 0; n < 8 && !hit; ++n) {
            const double* c = st.corner[cns[n]];
            const double d0 = c[0] - x[0], d1 = c[1] - x[1], d2 = c[2] - x[2];
            const double norm = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
            if (norm > 0.) w[n] = 1. / norm;
            else {
              for (int j = 0; j < 8; ++j) w[j] = 0.;
              w[n] = 1.0;
              hit = true;
            }
          }
          // after the EXIT the entries behind the hit keep the 0. written by PartDistDepo(:) = 0. (already set above)
          double DistSum = 0.;
          for (int n = 0; n < 8; ++n) DistSum = DistSum + w[n];
          for (int n = 0; n < 8; ++n)
            for (int c = 0; c < 4; ++c) acc[n * 4 + c] = acc[n * 4 + c] + w[n] / DistSum * T[c];
        }
      }
    }
    // deterministic block reduction of the 32 accumulators (fixed shuffle tree, then warps in order)
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 32; ++a) {
      double v = acc[a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      double s = red[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < STEP_NT / 32; ++w) s = s + red[w][threadIdx.x];
      elemAcc[(size_t)e * 32 + threadIdx.x] = s;
    }
  }
  (void)errFlag;
}

// S[n][c] = sum over the (element, corner) pairs adjacent to unique node n, fixed order (ascending element, corner)
__global__ void k_node_sum(const int32_t* __restrict__ adjOff, const int32_t* __restrict__ adj, const double* __restrict__ elemAcc,
                           double* __restrict__ S, int nNodes) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = t >> 2, c = t & 3;
  if (n >= nNodes) return;
  double s = 0.;
  for (int a = adjOff[n]; a < adjOff[n + 1]; ++a) s = s + elemAcc[(size_t)adj[a] * 4 + c];  // adj = elemLocal*8 + corner
  S[(size_t)n * 4 + c] = s;
}

// NodeSource[n] = (S[n] + sum of periodic partners' S) / NodeVolume[n]   (pic_depo_method.f90:499-510, :692-695)
__global__ void k_node_final(const double* __restrict__ S, const int32_t* __restrict__ pN, const int32_t* __restrict__ pOff,
                             const int32_t* __restrict__ pNodes, const double* __restrict__ NodeVolume, double* __restrict__ NodeSource,
                             int nNodes, int periodic) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = t >> 2, c = t & 3;
  if (n >= nNodes) return;
  double s = S[(size_t)n * 4 + c];
  if (periodic) {
    const int cnt = pN[n], off = pOff[n];
    for (int j = 0; j < cnt; ++j) s = s + S[(size_t)(pNodes[off + j] - 1) * 4 + c];
  }
  const double vol = NodeVolume[n];
  if (vol > 0.) s = s / vol;
  NodeSource[(size_t)n * 4 + c] = s;
}

// PartSource(1:4,kk,ll,mm) of every local element from its 8 node values (pic_depo_method.f90:710-733)
template <int NP>
__global__ void k_nodes_to_dofs(const double* __restrict__ NodeSource, const int32_t* __restrict__ elemNodeU /*[nElems][8] 0-based unique*/,
                                double* __restrict__ PartSource, int nElems) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  constexpr int ND = NP * NP * NP;
  const size_t total = (size_t)nElems * ND * 4;
  if (t >= total) return;
  const int c = t & 3;
  const size_t dof = t >> 2;
  const int e = (int)(dof / ND);
  const int r = (int)(dof % ND);
  const int kk = r % NP, ll = (r / NP) % NP, mm = r / (NP * NP);
  const double a1 = cst.cvwFac[kk], a2 = cst.cvwFac[ll], a3 = cst.cvwFac[mm];
  const int32_t* en = elemNodeU + (size_t)e * 8;
  double ns[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) ns[n] = NodeSource[(size_t)en[n] * 4 + c];
  double v = ((ns[0] * (1 - a1)) * (1 - a2)) * (1 - a3);
  v = v + ((ns[1] * (a1)) * (1 - a2)) * (1 - a3);
  v = v + ((ns[2] * (a1)) * (a2)) * (1 - a3);
  v = v + ((ns[3] * (1 - a1)) * (a2)) * (1 - a3);
  v = v + ((ns[4] * (1 - a1)) * (1 - a2)) * (a3);
  v = v + ((ns[5] * (a1)) * (1 - a2)) * (a3);
  v = v + ((ns[6] * (a1)) * (a2)) * (a3);
  v = v + ((ns[7] * (1 - a1)) * (a2)) * (a3);
  PartSource[t] = v;
}

// ---- interpolate + push + TriaTracking ---------------------------------------------------------------------------------------
// timedisc_TimeStepPoissonByBorisLeapfrog.f90:109-211 for the particles of one element per CTA iteration.
// Writes the new state in place and the sort key of the new owner element.
template <int NP>
__global__ void __launch_bounds__(STEP_NT) k_push_track_tria(PartBuf pb, const int64_t* __restrict__ elemOff, int nElems,
                                                             int offsetElem, const GeoElem* __restrict__ geo,
                                                             const TriaElem* __restrict__ tria, const double* __restrict__ E,
                                                             const double* __restrict__ Elem_xGP, const int32_t* __restrict__ elemRank,
                                                             uint32_t* __restrict__ keys, double dt, int xiValid,
                                                             int* __restrict__ counters /*[0]=lost,[1]=error code*/) {
  constexpr int ND = NP * NP * NP;
  __shared__ GeoElem sg;
  __shared__ TriaElem st;
  __shared__ __align__(16) double sE[ND * 3];
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int64_t p0 = elemOff[e], p1 = elemOff[e + 1];
    if (p1 <= p0) continue;
    const int gElem = offsetElem + e + 1;
    __syncthreads();
    if (!xiValid) stage_words(&sg, geo + (gElem - 1), sizeof(GeoElem));
    stage_words(&st, tria + (gElem - 1), sizeof(TriaElem));
    for (int i = threadIdx.x; i < ND * 3; i += STEP_NT) sE[i] = __ldg(E + (size_t)e * ND * 3 + i);
    __syncthreads();
    for (int64_t p = p0 + threadIdx.x; p < p1; p += STEP_NT) {
      double x[3] = {pb.x[0][p], pb.x[1][p], pb.x[2][p]};
      double v[3] = {pb.v[0][p], pb.v[1][p], pb.v[2][p]};
      uint8_t meta = pb.meta[p];
      const int spec = meta & META_SPEC_MASK;
      bool isNew = (meta & META_ISNEW) != 0;
      double lp[3] = {x[0], x[1], x[2]};  // LastPartPos
      double F[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) F[c] = 0.;
      const double q = cst.ChargeIC[spec];
      if (cst.DoInterpolation && fabs(q) > 0.0) {  // isInterpolateParticle
        double xi[3];
        bool suc;
        if (xiValid) {
          xi[0] = pb.xi[0][p]; xi[1] = pb.xi[1][p]; xi[2] = pb.xi[2][p];
          suc = !(meta & META_XIFAIL);
        } else {
          suc = (position_in_ref_elem(&sg, x, xi, false, true) & 1) != 0;
        }
        double f3[3];
        if (!suc && cst.DepositionType == PGPU_DEPO_CVWM)
          field_inverse_distance<NP>(x, sE, Elem_xGP + (size_t)(gElem - 1) * ND * 3, f3);
        else
          evaluate_field<NP>(xi, sE, f3);
#pragma unroll
        for (int c = 0; c < 6; ++c) F[c] = cst.externalField[c];
        F[0] = F[0] + f3[0];
        F[1] = F[1] + f3[1];
        F[2] = F[2] + f3[2];
        F[3] = F[3] + 0.; F[4] = F[4] + 0.; F[5] = F[5] + 0.;
      }
      push_particle(x, v, F, spec, isNew, dt);
      // PerformTracking -> SingleParticleTriaTracking3D
      int newElem = gElem;
      double det[6][2];
      int status = TRK_OK;
      if (!inside_quad3d(&st, x, det)) status = tria_track_walk(tria, x, lp, newElem, det);
      uint32_t key;
      if (status == TRK_OK) {
        const int rk = elemRank[newElem - 1];
        key = (rk == cst.myRank) ? (uint32_t)(newElem - 1 - offsetElem) : (uint32_t)(nElems + rk);
      } else {
        key = (uint32_t)(nElems + cst.nRanks);  // removed
        newElem = 0;
        if (status == TRK_LOST) atomicAdd(&counters[0], 1);
        else if (status != TRK_REMOVED) atomicMax(&counters[1], status);
      }
      pb.x[0][p] = x[0]; pb.x[1][p] = x[1]; pb.x[2][p] = x[2];
      pb.v[0][p] = v[0]; pb.v[1][p] = v[1]; pb.v[2][p] = v[2];
      pb.elem[p] = newElem;
      pb.meta[p] = (uint8_t)((meta & META_SPEC_MASK) | (isNew ? META_ISNEW : 0));
      keys[p] = key;
    }
  }
}

// keys for freshly uploaded particles (sorted into element order before the first step)
__global__ void k_keys_from_elem(const int32_t* __restrict__ elem, const int32_t* __restrict__ elemRank, uint32_t* __restrict__ keys,
                                 int64_t n, int nElems, int offsetElem, int myRank, int nRanks) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = elem[i];
  uint32_t key;
  if (g < 1) key = (uint32_t)(nElems + nRanks);
  else {
    const int rk = elemRank[g - 1];
    key = (rk == myRank) ? (uint32_t)(g - 1 - offsetElem) : (uint32_t)(nElems + rk);
  }
  keys[i] = key;
}
