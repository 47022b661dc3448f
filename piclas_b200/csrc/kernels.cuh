// kernels.cuh — the particle-step kernels (sm_100a): deposition (cell_volweight_mean) and interpolate+push+track.
//
// Launch geometry: particles are sorted by local element; one CTA works on one element at a time (grid-stride over
// elements, grid = a multiple of the SM count), stages that element's geometry / field tile / corner table in shared
// memory and lets its threads stride over the element's particle segment.  Particle state is SoA FP64, read and
// written fully coalesced.
//
// Structure of the step (r1 profile: one fused kernel was instruction-cache bound — 20k SASS lines, 66 % no_inst
// stalls, SIMD efficiency 45 %, profiles/r1_v1_push_track_stalls.txt):
//   k_interp_push   all particles: field evaluation, push, ParticleInsideQuad3D in the own element (the common exit of
//                   SingleParticleTriaTracking3D).  Particles that left the element are appended to a leaver list.
//   k_track_leavers one thread per leaver: the element walk of SingleParticleTriaTracking3D on the global tables.
// Loops that do not need unrolling are kept rolled so that each kernel's hot loop stays inside the instruction caches.
#pragma once
#include "math.cuh"
#include "fastmath.cuh"

constexpr int STEP_NT = 128;  // threads per CTA of the per-element kernels
#ifndef KA_MINB
#define KA_MINB 4   // min resident CTAs/SM of k_interp_push (register cap = 65536 / (128 * KA_MINB))
#endif
#ifndef DEP_MINB
#define DEP_MINB 4
#endif

__device__ __forceinline__ void stage_words(void* dst, const void* src, int nbytes) {
  // cooperative copy of a 16-byte aligned record into shared memory with 128-bit loads
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d = reinterpret_cast<int4*>(dst);
  for (int i = threadIdx.x; i < nbytes / 16; i += blockDim.x) d[i] = __ldg(s + i);
}

// ---- asynchronous staging of the next particle tile (cp.async, one 8-byte copy per thread and array) ---------------------------
// The per-element kernels run 16 warps per SM (128 registers per thread), too few to hide the DRAM latency of the
// particle loads behind arithmetic (r1v5 profile: 32 % of k_interp_push's stall samples sat on the first use of the
// loaded particle).  Each thread therefore copies its particle of the NEXT loop iteration into a private shared-memory
// slot while it works on the current one; slots are private to the thread, so no barrier is needed.
__device__ __forceinline__ void cp_async8(double* smemDst, const double* gmemSrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// ---- Newton mapping for all particles of an element + CVWM accumulation -----------------------------------------------------
// DepositionMethod_CVWM particle loop, pic_depo_method.f90:471-544.  elemAcc[e][node 0..7 (CGNS)][0..3] receives the
// element-local sums of TSource*weight; xi and the SucRefPos flag are cached for the interpolation of the same step.
// Two accumulation paths, both reduced over the CTA in a fixed order (deterministic, independent of scheduling):
//   * general: the 32 per-thread accumulators live in shared memory ([a][thread], conflict free) so that the Newton
//     iteration keeps the register file;
//   * restructured arithmetic on an affine element (CTA-uniform): closed-form xi, accumulators in registers, no calls in
//     the loop; the (never observed) particles whose closed-form xi is far outside the element take the general path
//     in a second sweep.
typedef double DepAcc[32][STEP_NT];

// one particle through the general path.  x, xi: position and (output) reference position
__device__ __forceinline__ void deposit_particle_general(const PartBuf& pb, int64_t p, const GeoElem* sg, const double (*corner)[3],
                                                         DepAcc& sAcc, int tid) {
  double* __restrict__ const PF = pb.f;
  double* __restrict__ const PXI = pb.xif;
  const int64_t PS_ = pb.stride;
  const double x[3] = {PF[p], PF[1 * PS_ + p], PF[2 * PS_ + p]};
  const uint8_t meta = pb.meta[p];
  const int spec = meta & META_SPEC_MASK;
  double xi[3];
  const bool suc = (position_in_ref_elem(sg, x, xi, true, true) & 1) != 0;
  PXI[p] = xi[0];
  PXI[1 * PS_ + p] = xi[1];
  PXI[2 * PS_ + p] = xi[2];
  const uint8_t nmeta = suc ? (meta & ~META_XIFAIL) : (meta | META_XIFAIL);
  if (nmeta != meta) pb.meta[p] = nmeta;
  const double q = cst.ChargeIC[spec];
  if (!(fabs(q) > 0.0)) return;  // isDepositParticle
  const double Charge = q * cst.MPF[spec];
  const double T[4] = {PF[3 * PS_ + p] * Charge, PF[4 * PS_ + p] * Charge, PF[5 * PS_ + p] * Charge, Charge};
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
      for (int c = 0; c < 4; ++c) sAcc[n * 4 + c][tid] = sAcc[n * 4 + c][tid] + (T[c] * w[n]);
  } else {
    // inverse-distance fallback, :512-538.  CGNS corner n is tensor node cns[n]
    const int cns[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    bool hit = false;
    for (int n = 0; n < 8 && !hit; ++n) {
      const double* c = corner[cns[n]];
      const double d0 = c[0] - x[0], d1 = c[1] - x[1], d2 = c[2] - x[2];
      const double norm = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
      if (norm > 0.) w[n] = 1. / norm;
      else {
        for (int j = 0; j < 8; ++j) w[j] = 0.;
        w[n] = 1.0;
        hit = true;
      }
    }
    double DistSum = 0.;
    for (int n = 0; n < 8; ++n) DistSum = DistSum + w[n];
    for (int n = 0; n < 8; ++n)
      for (int c = 0; c < 4; ++c) sAcc[n * 4 + c][tid] = sAcc[n * 4 + c][tid] + w[n] / DistSum * T[c];
  }
}

__device__ __noinline__ void deposit_particle_cold(const PartBuf* pb, int64_t p, const GeoElem* sg, const double (*corner)[3],
                                                   DepAcc* sAcc, int tid) {
  deposit_particle_general(*pb, p, sg, corner, *sAcc, tid);
}

// closed-form reference position on an affine element (fastmath.cuh ref_position_fast); false: take the general path
__device__ __forceinline__ bool affine_xi(const AffElem* af, const double x[3], double xi[3]) {
  const double r0 = x[0] - af->x0[0], r1 = x[1] - af->x0[1], r2 = x[2] - af->x0[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) xi[d] = fma(af->A[d][0], r0, fma(af->A[d][1], r1, af->A[d][2] * r2)) - 1.0;
  return fabs(xi[0]) <= 1.5 && fabs(xi[1]) <= 1.5 && fabs(xi[2]) <= 1.5;
}

template <bool FAST>
__global__ void __launch_bounds__(STEP_NT, DEP_MINB) k_deposit_cvwm(PartBuf pb, const int64_t* __restrict__ elemOff, int nElems,
                                                             int offsetElem, const GeoElem* __restrict__ geo,
                                                             const TriaElem* __restrict__ tria, const AffElem* __restrict__ aff,
                                                             double* __restrict__ elemAcc) {
  double* __restrict__ const PF = pb.f;
  double* __restrict__ const PXI = pb.xif;
  const int64_t PS_ = pb.stride;
  __shared__ GeoElem sg;
  __shared__ AffElem sa;
  __shared__ double corner[8][3];
  __shared__ DepAcc sAcc;
  __shared__ double sP[FAST ? 2 : 1][FAST ? 6 : 1][FAST ? STEP_NT : 1];   // x, v of the current / next particle of every thread
  const int tid = threadIdx.x;
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int64_t p0 = elemOff[e], p1 = elemOff[e + 1];
    bool staged = false;
    if (p1 > p0) {
      __syncthreads();
      if (FAST) stage_words(&sa, aff + (offsetElem + e), sizeof(AffElem));
      __syncthreads();
    }
    if (FAST && p1 > p0 && sa.affine != 0.0) {
      double acc[32];
#pragma unroll
      for (int a = 0; a < 32; ++a) acc[a] = 0.;
      int nGeneral = 0;
      int64_t p = p0 + tid;
      int stage = 0;
      uint8_t meta = 0, metaNext = 0;
      if (p < p1) {
#pragma unroll
        for (int a = 0; a < 6; ++a) cp_async8(&sP[0][a][tid], PF + a * PS_ + p);
        meta = pb.meta[p];
      }
      cp_async_commit();
      for (; p < p1; p += STEP_NT, stage ^= 1, meta = metaNext) {
        const int64_t pn = p + STEP_NT;
        if (pn < p1) {
#pragma unroll
          for (int a = 0; a < 6; ++a) cp_async8(&sP[stage ^ 1][a][tid], PF + a * PS_ + pn);
          metaNext = pb.meta[pn];
        }
        cp_async_commit();
        cp_async_wait_prev();
        const double x[3] = {sP[stage][0][tid], sP[stage][1][tid], sP[stage][2][tid]};
        double xi[3];
        if (!affine_xi(&sa, x, xi)) { ++nGeneral; continue; }
        PXI[p] = xi[0];
        PXI[1 * PS_ + p] = xi[1];
        PXI[2 * PS_ + p] = xi[2];
        if (meta & META_XIFAIL) pb.meta[p] = meta & ~META_XIFAIL;
        const int spec = meta & META_SPEC_MASK;
        const double q = cst.ChargeIC[spec];
        if (!(fabs(q) > 0.0)) continue;  // isDepositParticle
        const double Charge = q * cst.MPF[spec];
        const double T[4] = {sP[stage][3][tid] * Charge, sP[stage][4][tid] * Charge, sP[stage][5][tid] * Charge, Charge};
        const double a1 = 0.5 * (xi[0] + 1.0), a2 = 0.5 * (xi[1] + 1.0), a3 = 0.5 * (xi[2] + 1.0);
        const double b1 = 1 - a1, b2 = 1 - a2, b3 = 1 - a3;
        const double w[8] = {(b1 * b2) * b3, (a1 * b2) * b3, (a1 * a2) * b3, (b1 * a2) * b3,
                             (b1 * b2) * a3, (a1 * b2) * a3, (a1 * a2) * a3, (b1 * a2) * a3};
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[n * 4 + c] = fma(T[c], w[n], acc[n * 4 + c]);
      }
#pragma unroll
      for (int a = 0; a < 32; ++a) sAcc[a][tid] = acc[a];
      if (__syncthreads_or(nGeneral)) {
        stage_words(&sg, geo + (offsetElem + e), sizeof(GeoElem));
        if (tid < 24) corner[tid / 3][tid % 3] = tria[offsetElem + e].corner[tid / 3][tid % 3];
        __syncthreads();
        for (int64_t p = p0 + tid; p < p1; p += STEP_NT) {
          const double x[3] = {PF[p], PF[1 * PS_ + p], PF[2 * PS_ + p]};
          double xi[3];
          if (!affine_xi(&sa, x, xi)) deposit_particle_cold(&pb, p, &sg, corner, &sAcc, tid);
        }
      }
      staged = true;
    }
    if (!staged) {
#pragma unroll
      for (int a = 0; a < 32; ++a) sAcc[a][tid] = 0.;
      if (p1 > p0) {
        stage_words(&sg, geo + (offsetElem + e), sizeof(GeoElem));
        if (tid < 24) corner[tid / 3][tid % 3] = tria[offsetElem + e].corner[tid / 3][tid % 3];
        __syncthreads();
        for (int64_t p = p0 + tid; p < p1; p += STEP_NT) {
          asm volatile("" ::: "memory");
          if (FAST) deposit_particle_cold(&pb, p, &sg, corner, &sAcc, tid);   // non-affine element: keep the hot loop's registers
          else deposit_particle_general(pb, p, &sg, corner, sAcc, tid);
        }
      }
    }
    // deterministic block reduction: accumulator a is summed over threads by warp a/8.. in a fixed tree
    __syncthreads();
    {
      const int lane = tid & 31, warp = tid >> 5;  // 4 warps, each reduces 8 accumulators
#pragma unroll 1
      for (int a = warp * 8; a < warp * 8 + 8; ++a) {
        double v = ((sAcc[a][lane] + sAcc[a][lane + 32]) + sAcc[a][lane + 64]) + sAcc[a][lane + 96];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) elemAcc[(size_t)e * 32 + a] = v;
      }
    }
    __syncthreads();
  }
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

// field tile in shared memory: sE[((k*NP + j)*3 + c)*NP + i]  (rows of NP contiguous i-values per component)
template <int NP>
__device__ __forceinline__ void evaluate_field_tile(const double xi[3], const double* __restrict__ sE, double out[3]) {
  double L0[NP], L1[NP], L2[NP];
  lagrange_polys<NP>(xi[0], cst.xGP, cst.wBary, L0);
  lagrange_polys<NP>(xi[1], cst.xGP, cst.wBary, L1);
  lagrange_polys<NP>(xi[2], cst.xGP, cst.wBary, L2);
  double o0 = 0., o1 = 0., o2 = 0.;
#pragma unroll 1
  for (int k = 0; k < NP; ++k) {
    double lz = L2[0];
#pragma unroll
    for (int q = 1; q < NP; ++q) lz = (k == q) ? L2[q] : lz;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double lez = L1[j] * lz;
      const double* row = sE + ((k * NP + j) * 3) * NP;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        o0 = fma(row[i] * L0[i], lez, o0);  // U_OUT + U_IN*L_xi(1,i)*L_Eta_Zeta, eval_xyz.f90:207-215
        o1 = fma(row[NP + i] * L0[i], lez, o1);
        o2 = fma(row[2 * NP + i] * L0[i], lez, o2);
      }
    }
  }
  out[0] = o0;
  out[1] = o1;
  out[2] = o2;
}

// ---- interpolate + push + own-element inside test --------------------------------------------------------------------------------
// timedisc_TimeStepPoissonByBorisLeapfrog.f90:109-198 and the first iteration of SingleParticleTriaTracking3D
// (particle_triatracking.f90:203-218) for the particles of one element per CTA iteration.
// Stayers: x, v written in place, key = own element.  Leavers: v written in place, the pushed position goes to xNew
// (the idle half of the double buffer) while pb.x keeps LastPartPos; their index is appended to leaverIdx.
// REF (TrackingMethod = refmapping): the reference coordinates are the stored PartPosRef (pic_interpolation_tools.f90:243), the
// pushed position of every particle goes to xn and pb.x keeps LastPartPos for k_track_ref (no inside test here).
template <int NP, bool FAST, bool REF>
__global__ void __launch_bounds__(STEP_NT, KA_MINB) k_interp_push(PartBuf pb, double* __restrict__ xn0, double* __restrict__ xn1,
                                                            double* __restrict__ xn2, const int64_t* __restrict__ elemOff, int nElems,
                                                            int offsetElem, const GeoElem* __restrict__ geo,
                                                            const TriaElem* __restrict__ tria, const PlaneElem* __restrict__ planes,
                                                            const AffElem* __restrict__ aff, const double* __restrict__ E,
                                                            const double* __restrict__ Elem_xGP, uint32_t* __restrict__ keys,
                                                            uint32_t* __restrict__ leaverIdx, double dt, int xiValid,
                                                            int* __restrict__ counters /*[0]=lost,[1]=error,[2]=nLeavers*/) {
  constexpr int ND = NP * NP * NP;
  double* __restrict__ const PF = pb.f;
  const double* __restrict__ const PXI = pb.xif;
  const int64_t PS_ = pb.stride;
  __shared__ GeoElem sg;
  __shared__ TriaElem st;
  __shared__ PlaneElem sp;
  __shared__ AffElem sa;
  __shared__ __align__(16) double sE[ND * 3];
  __shared__ double sP[2][9][STEP_NT];   // x, v, xi of the current / next particle of every thread (cp.async staging)
  const int tid = threadIdx.x;
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int64_t p0 = elemOff[e], p1 = elemOff[e + 1];
    if (p1 <= p0) continue;
    const int gElem = offsetElem + e + 1;
    __syncthreads();
    if (!xiValid && !REF) {
      stage_words(&sg, geo + (gElem - 1), sizeof(GeoElem));
      if (FAST) stage_words(&sa, aff + (gElem - 1), sizeof(AffElem));
    }
    if (FAST && !REF) stage_words(&sp, planes + (gElem - 1), sizeof(PlaneElem));
    if (!REF) stage_words(&st, tria + (gElem - 1), sizeof(TriaElem));
    for (int t = threadIdx.x; t < ND * 3; t += STEP_NT) {
      const int c = t % 3, node = t / 3;
      const int i = node % NP, kj = node / NP;
      sE[(kj * 3 + c) * NP + i] = __ldg(E + (size_t)e * ND * 3 + t);
    }
    __syncthreads();
    const bool useXi = REF || xiValid;
    int64_t p = p0 + tid;
    int stage = 0;
    uint8_t meta = 0, metaNext = 0;
    if (p < p1) {
#pragma unroll
      for (int a = 0; a < 6; ++a) cp_async8(&sP[0][a][tid], PF + a * PS_ + p);
      if (useXi) {
#pragma unroll
        for (int a = 0; a < 3; ++a) cp_async8(&sP[0][6 + a][tid], PXI + a * PS_ + p);
      }
      meta = pb.meta[p];
    }
    cp_async_commit();
    for (; p < p1; p += STEP_NT, stage ^= 1, meta = metaNext) {
      const int64_t pn = p + STEP_NT;
      if (pn < p1) {   // next particle of this thread: copies overlap the arithmetic below
#pragma unroll
        for (int a = 0; a < 6; ++a) cp_async8(&sP[stage ^ 1][a][tid], PF + a * PS_ + pn);
        if (useXi) {
#pragma unroll
          for (int a = 0; a < 3; ++a) cp_async8(&sP[stage ^ 1][6 + a][tid], PXI + a * PS_ + pn);
        }
        metaNext = pb.meta[pn];
      }
      cp_async_commit();
      cp_async_wait_prev();   // (memory clobber: also keeps the shared-memory tiles out of the registers, no hoisting over the loop)
      double x[3] = {sP[stage][0][tid], sP[stage][1][tid], sP[stage][2][tid]};
      double v[3] = {sP[stage][3][tid], sP[stage][4][tid], sP[stage][5][tid]};
      const int spec = meta & META_SPEC_MASK;
      bool isNew = (meta & META_ISNEW) != 0;
      double F[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) F[c] = 0.;
      const double q = cst.ChargeIC[spec];
      if (cst.DoInterpolation && fabs(q) > 0.0) {  // isInterpolateParticle
        double xi[3];
        bool suc;
        if (REF) {
          xi[0] = sP[stage][6][tid]; xi[1] = sP[stage][7][tid]; xi[2] = sP[stage][8][tid];
          suc = true;
        } else if (xiValid) {
          xi[0] = sP[stage][6][tid]; xi[1] = sP[stage][7][tid]; xi[2] = sP[stage][8][tid];
          suc = !(meta & META_XIFAIL);
        } else if (FAST) {
          suc = ref_position_fast(&sa, &sg, x, xi, false);
        } else {
          suc = (position_in_ref_elem(&sg, x, xi, false, true) & 1) != 0;
        }
        double f3[3];
        if (!suc && cst.DepositionType == PGPU_DEPO_CVWM)
          field_inverse_distance<NP>(x, E + (size_t)e * ND * 3, Elem_xGP + (size_t)(gElem - 1) * ND * 3, f3);
        else if (FAST)
          evaluate_field_fast<NP>(xi, sE, f3);
        else
          evaluate_field_tile<NP>(xi, sE, f3);
#pragma unroll
        for (int c = 0; c < 6; ++c) F[c] = cst.externalField[c];
        F[0] = F[0] + f3[0];
        F[1] = F[1] + f3[1];
        F[2] = F[2] + f3[2];
        F[3] = F[3] + 0.; F[4] = F[4] + 0.; F[5] = F[5] + 0.;
      }
      if (FAST) push_particle_fast(x, v, F, spec, isNew, dt);
      else push_particle(x, v, F, spec, isNew, dt);
      PF[3 * PS_ + p] = v[0]; PF[4 * PS_ + p] = v[1]; PF[5 * PS_ + p] = v[2];
      const uint8_t nmeta = (uint8_t)(meta & META_SPEC_MASK);  // IsNewPart and the xi flag are consumed
      if (nmeta != meta) pb.meta[p] = nmeta;
      if (REF) {
        xn0[p] = x[0]; xn1[p] = x[1]; xn2[p] = x[2];
        continue;
      }
      uint32_t mask;
      const bool inElem = FAST ? inside_fast(&sp, &st, x, mask) : inside_quad3d_mask(&st, x, mask);
      if (inElem) {
        PF[p] = x[0]; PF[1 * PS_ + p] = x[1]; PF[2 * PS_ + p] = x[2];
        keys[p] = (uint32_t)e;
      } else {
        xn0[p] = x[0]; xn1[p] = x[1]; xn2[p] = x[2];
        keys[p] = mask;  // handed to k_track_leavers, which overwrites it with the final key
        const unsigned act = __activemask();   // warp-aggregated append: one atomic per warp, consecutive slots
        const int ln = threadIdx.x & 31, leader = __ffs(act) - 1;
        int slot0 = 0;
        if (ln == leader) slot0 = atomicAdd(&counters[2], __popc(act));
        slot0 = __shfl_sync(act, slot0, leader);
        leaverIdx[slot0 + __popc(act & ((1u << ln) - 1u))] = (uint32_t)p;
      }
    }
  }
}

// ---- SingleParticleTriaTracking3D for the particles that left their element (particle_triatracking.f90:137-484) --------------------
// One thread per leaver, element records read from global memory (L2 resident).  All loops are rolled and the candidate
// triangles are visited through a bit mask so that the lanes of a warp run the same through-side test at the same time.
// (r1 profile of the first version, which looped per thread until done: every warp ran 2-3 rounds for 1.1 hops per particle,
// 12 of 32 lanes active; a shared-memory re-queue with block barriers was slower still.)
constexpr int LV_NT = 128;
#ifndef LV_MINB
#define LV_MINB 4
#endif

template <bool FAST>
__global__ void __launch_bounds__(LV_NT, LV_MINB) k_track_leavers(PartBuf pb, const double* __restrict__ xn0, const double* __restrict__ xn1,
                                                         const double* __restrict__ xn2, const uint32_t* __restrict__ leaverIdx,
                                                         const TriaElem* __restrict__ tria, const PlaneElem* __restrict__ planes,
                                                         const int32_t* __restrict__ elemRank, uint32_t* __restrict__ keys, int nElems,
                                                         int offsetElem, int* __restrict__ counters) {
  double* __restrict__ const PF = pb.f;
  const int64_t PS_ = pb.stride;
  const int nLeavers = counters[2];
  const int lane = threadIdx.x & 31;
  bool active = false;
  int p = 0, ElemID = 0, guard = 0;
  uint32_t mask = 0;
  double x[3] = {0., 0., 0.}, lp[3] = {0., 0., 0.};
  int dE0 = 0, dE1 = 0, dE2 = 0, dE3 = 0, dE4 = 0, dE5 = 0;  // DoneLastElem(1:4,1:6); entry 0 is the most recent crossing
  int dS0 = 0, dS1 = 0, dS2 = 0, dS3 = 0, dS4 = 0, dS5 = 0;
  int dT0 = 0, dT1 = 0, dT2 = 0, dT3 = 0, dT4 = 0, dT5 = 0;
  // persistent warps: a lane performs one element crossing ("hop") per round and fetches the next leaver as soon as its
  // particle is localised, so that lanes stay busy while a neighbour lane walks through several elements
  while (true) {
    {
      const unsigned need = __ballot_sync(0xffffffffu, !active);
      if (need) {
        int first = 0;
        if (lane == __ffs(need) - 1) first = atomicAdd(&counters[3], __popc(need));
        first = __shfl_sync(0xffffffffu, first, __ffs(need) - 1);
        const int mine = first + __popc(need & ((1u << lane) - 1u));
        if (!active && mine < nLeavers) {
          active = true;
          p = (int)leaverIdx[mine];
          x[0] = xn0[p]; x[1] = xn1[p]; x[2] = xn2[p];
          lp[0] = PF[p]; lp[1] = PF[1 * PS_ + p]; lp[2] = PF[2 * PS_ + p];  // LastPartPos
          ElemID = pb.elem[p];                                               // LastGlobalElemID
          mask = keys[p];                                                    // det <= 0 triangles of the start element
          guard = 0;
          dE0 = dE1 = dE2 = dE3 = dE4 = dE5 = 0;
          dS0 = dS1 = dS2 = dS3 = dS4 = dS5 = 0;
          dT0 = dT1 = dT2 = dT3 = dT4 = dT5 = 0;
        }
      }
      if (__ballot_sync(0xffffffffu, active) == 0) break;
    }
    {
      int status = -1;  // -1: needs another hop
      if (active) {
        const TriaElem* te = tria + (ElemID - 1);
        // 2b) crossed triangles among those with det <= 0
        double V[3] = {x[0] - lp[0], x[1] - lp[1], x[2] - lp[2]};
        const double len = sqrt((V[0] * V[0] + V[1] * V[1]) + V[2] * V[2]);
        if (fabs(len) > 0.) {
          V[0] = V[0] / len; V[1] = V[1] / len; V[2] = V[2] / len;
        }
        uint32_t thr = 0;
        int nThrough = 0;
        uint32_t cand = mask;
#pragma unroll 1
        while (cand) {
          const int b = __ffs(cand) - 1;
          cand &= cand - 1;
          if (through_side_check_fast<true>(te, lp, V, b >> 1, (b & 1) + 1)) {
            thr |= 1u << b;
            ++nThrough;
          }
        }
        int side = -1, tri = 0;
        if (nThrough == 0) {
          status = TRK_LOST;
        } else if (nThrough == 1) {
          const int b = __ffs(thr) - 1;
          side = b >> 1;
          tri = (b & 1) + 1;
        } else {
          // several candidate sides: the one crossed first has the largest |det(PartPos)/det(LastPartPos)| (:309-405)
          int second = 0;
          double minRatio = 0;
          uint32_t c2 = thr;
#pragma unroll 1
          while (c2) {
            const int b = __ffs(c2) - 1;
            c2 &= c2 - 1;
            const int s = b >> 1, t = (b & 1) + 1;
            const int gs = te->sideID[s];
            const bool treated = (dE1 == ElemID && dS1 == gs && dT1 == t) || (dE2 == ElemID && dS2 == gs && dT2 == t) ||
                                 (dE3 == ElemID && dS3 == gs && dT3 == t) || (dE4 == ElemID && dS4 == gs && dT4 == t) ||
                                 (dE5 == ElemID && dS5 == gs && dT5 == t);
            if (treated) continue;
            double detM;
            if (!through_side_lastpos_check<true>(te, lp, s, t, detM)) continue;
            double d1, d2;
            side_dets<true>(te, x, s, d1, d2);
            const double dS = (t == 1) ? d1 : d2;
            if (detM == 0 && dS == 0) continue;
            if (detM == 0 && minRatio == 0) {
              ++second; side = s; tri = t;
            } else {
              if (detM == 0) continue;
              const double ratio = dS / detM;
              if (ratio < minRatio) {
                minRatio = ratio;
                ++second; side = s; tri = t;
              }
            }
          }
          if (second == 0) status = TRK_LOST;
        }
        if (status == -1) {
          // 3) boundary interaction or step into the neighbour
          const int gside = te->sideID[side];
          const int bc = te->bcid[side];
          const int oldElem = ElemID;
          if (bc > 0) {
            const int kind = cst.bc_kind[bc - 1];
            if (kind == PGPU_BC_OPEN) status = TRK_REMOVED;
            else if (kind != PGPU_BC_PERIODIC) status = TRK_ERR_BC;
            else {
              const double alpha = intersection_with_wall<true>(te, lp, V, side, tri);
              const int pvid = cst.bc_alpha[bc - 1];  // PeriodicBoundary, particle_boundary_condition.f90:224-284
              const int pv = (pvid < 0 ? -pvid : pvid) - 1;
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                lp[d] = lp[d] + V[d] * alpha;
                lp[d] = lp[d] + copysign(cst.PeriodicVectors[pv][d], (double)pvid);
                x[d] = lp[d] + (len - alpha) * V[d];
              }
            }
          }
          if (status == -1) {
            ElemID = te->nbElem[side];
            dE5 = dE4; dS5 = dS4; dT5 = dT4;
            dE4 = dE3; dS4 = dS3; dT4 = dT3;
            dE3 = dE2; dS3 = dS2; dT3 = dT2;
            dE2 = dE1; dS2 = dS1; dT2 = dT1;
            dE1 = dE0; dS1 = dS0; dT1 = dT0;
            dE0 = oldElem; dS0 = gside; dT0 = tri;
            if (ElemID < 1) status = TRK_ERR_ELEM;
            else {
              // 2a) inside test in the new element
              const bool inNew = FAST ? inside_fast<true>(planes + (ElemID - 1), tria + (ElemID - 1), x, mask)
                                      : inside_quad3d_mask<true>(tria + (ElemID - 1), x, mask);
              if (inNew) status = TRK_OK;
              else if (++guard > 100000) status = TRK_ERR_LOOP;
            }
          }
        }
        if (status != -1) {
          uint32_t key;
          int newElem = ElemID;
          if (status == TRK_OK) {
            const int rk = elemRank[newElem - 1];
            key = (rk == cst.myRank) ? (uint32_t)(newElem - 1 - offsetElem) : (uint32_t)(nElems + rk);
          } else {
            key = (uint32_t)(nElems + cst.nRanks);  // removed
            newElem = 0;
            if (status == TRK_LOST) atomicAdd(&counters[0], 1);
            else if (status != TRK_REMOVED) atomicMax(&counters[1], status);
          }
          PF[p] = x[0]; PF[1 * PS_ + p] = x[1]; PF[2 * PS_ + p] = x[2];
          pb.elem[p] = newElem;
          keys[p] = key;
        }
      }
      if (active && status != -1) active = false;
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
