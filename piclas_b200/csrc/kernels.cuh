// kernels.cuh — the particle-step kernels (sm_100a): deposition (cell_volweight_mean) and interpolate+push+track.
//
// Launch geometry: particles are sorted by local element; one CTA works on one element at a time (grid-stride over
// elements, grid = a multiple of the SM count), stages that element's geometry / field tile / corner table in shared
// memory and lets its threads stride over the element's particle segment.  Particle state is SoA FP64, read and
// written fully coalesced.
//
// Structure of the step (r1 profile: one fused kernel was instruction-cache bound — 20k SASS lines, 66 % no_inst
// stalls, SIMD efficiency 45 %, profiles/r1_v1_push_track_stalls.txt):
//   k_interp_push   phase 1, all particles of a sweep: field evaluation, push, ParticleInsideQuad3D in the own element (the
//                   common exit of SingleParticleTriaTracking3D); leavers are queued in shared memory.
//                   phase 2 (restructured arithmetic), dense warps over the queue: the first element crossing on the staged
//                   records of the element and its six face neighbours — exit side from the side planes where that is safe
//                   (queue A), the determinant tests of ParticleThroughSideCheck3DFast otherwise (queue B).
//   k_track_leavers one thread per particle that is not localised after that (edge / corner crossings, long flights, or all
//                   leavers in reference-order arithmetic): the element walk on the global tables, persistent warps.
// Loops that do not need unrolling are kept rolled so that each kernel's hot loop stays inside the instruction caches.
#pragma once
#include <type_traits>
#include "math.cuh"
#include "fastmath.cuh"

constexpr int STEP_NT = 128;  // threads per CTA of the per-element kernels
#ifndef KA_MINB
#define KA_MINB 3   // min resident CTAs/SM of k_interp_push (register cap = 65536 / (128 * KA_MINB)); 3 -> 168 registers, no spills
#endif
#ifndef DEP_MINB
#define DEP_MINB 4
#endif
#ifndef KA_WARPQ
#define KA_WARPQ 0   // 1: leaver queues private to each warp (no block barriers between the phases of k_interp_push)
#endif
#ifndef KA_CHUNKF
#define KA_CHUNKF 16   // particles per thread and phase-1 sweep of k_interp_push (bounds the shared-memory leaver queue); halved for
#endif                 // N > 4, where the field tile needs the shared memory.  Measured 4 / 8 / 16 at 1907 particles per element:
                       // 6.53 / 5.81 / 5.57 ms (fewer block barriers per element)

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
template <int NT>
__device__ __forceinline__ void deposit_particle_general(const PartBuf& pb, int64_t p, const GeoElem* sg, const double (*corner)[3],
                                                         double (&sAcc)[32][NT], int tid) {
  double* __restrict__ const PF = pb.f;
  double* __restrict__ const PXI = pb.xif;
  const int64_t PS_ = pb.stride;
  const double x[3] = {PF[p], PF[1 * PS_ + p], PF[2 * PS_ + p]};
  const uint8_t meta = pb.meta[p];
  const int spec = meta & META_SPEC_MASK;
  double xi[3];
  const bool suc = (position_in_ref_elem(sg, x, xi, true, true) & 1) != 0;
  if (PXI) {   // reference position cached for the interpolation of the same step (sorted layout only; the bins keep none)
    PXI[p] = xi[0];
    PXI[1 * PS_ + p] = xi[1];
    PXI[2 * PS_ + p] = xi[2];
    const uint8_t nmeta = suc ? (meta & ~META_XIFAIL) : (meta | META_XIFAIL);
    if (nmeta != meta) pb.meta[p] = nmeta;
  }
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
                                                             double* __restrict__ elemAcc, int storeAffineXi) {
  // storeAffineXi = 0: on affine elements the closed-form xi is not written back; k_interp_push recomputes it from the
  // same position (9 multiply-adds instead of 24 B written here and read there)
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
        if (storeAffineXi) {
          PXI[p] = xi[0];
          PXI[1 * PS_ + p] = xi[1];
          PXI[2 * PS_ + p] = xi[2];
        }
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

// ---- one element crossing of SingleParticleTriaTracking3D (particle_triatracking.f90:220-470) ------------------------------------
// DoneLastElem(1:4,1:6): the last six crossings (element, global side, triangle); entry 0 is the most recent one
struct HopHist {
  int e0, e1, e2, e3, e4, e5, s0, s1, s2, s3, s4, s5, t0, t1, t2, t3, t4, t5;
  __device__ __forceinline__ void clear() { e0 = e1 = e2 = e3 = e4 = e5 = s0 = s1 = s2 = s3 = s4 = s5 = t0 = t1 = t2 = t3 = t4 = t5 = 0; }
  __device__ __forceinline__ void push(int e, int s, int t) {
    e5 = e4; s5 = s4; t5 = t4;
    e4 = e3; s4 = s3; t4 = t3;
    e3 = e2; s3 = s2; t3 = t2;
    e2 = e1; s2 = s1; t2 = t1;
    e1 = e0; s1 = s0; t1 = t0;
    e0 = e; s0 = s; t0 = t;
  }
  __device__ __forceinline__ bool treated(int e, int s, int t) const {
    return (e1 == e && s1 == s && t1 == t) || (e2 == e && s2 == s && t2 == t) || (e3 == e && s3 == s && t3 == t) ||
           (e4 == e && s4 == s && t4 == t) || (e5 == e && s5 == s && t5 == t);
  }
};

// The particle sits in element ElemID (record te) whose inside test failed with `mask` (triangles with det <= 0); x is the
// pushed position, lp LastPartPos.  Finds the crossed side, applies the boundary condition or steps into the neighbour and
// runs the inside test there.  Returns TRK_OK (localised in ElemID), -1 (another hop is needed from ElemID / mask), or the
// removal / error status.  G: te and the records of the new element are read from global memory with one element per lane;
// otherwise te is a shared-memory copy and planeOf(side) returns the staged planes of the neighbour behind local side `side`.
template <bool G>
__device__ __forceinline__ void load_plane4(const double* __restrict__ q, double& a, double& b, double& c, double& d) {
  if (G) asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(q));
  else { a = q[0]; b = q[1]; c = q[2]; d = q[3]; }
}

// Exit side of a convex element with planar sides from the side planes (restructured arithmetic).  Applies when the pushed
// position is beyond exactly one side plane: the flight then leaves through that side, and the reference's through-side
// tests single out the triangle the crossing point lies in.  Every decision is taken with a margin of PlaneElem::tol (start
// point clearly inside, crossing point clearly inside the side polygon and clearly off the triangle diagonal); otherwise
// false is returned and the determinant tests of ParticleThroughSideCheck3DFast decide.
template <bool G>
__device__ __forceinline__ bool exit_side_planar(const PlaneElem* __restrict__ pc, const double x[3], const double lp[3], uint32_t mask,
                                                 int& side, int& tri) {
  // The pushed position may lie beyond several side planes (flight through an edge or corner region): the flight leaves through the
  // side whose plane it reaches first.  Every decision keeps a margin of PlaneElem::tol (start point clearly inside each crossed
  // plane, end point clearly beyond, crossing point clearly inside all other side planes — which also separates the first from
  // the second crossing — and clearly off the triangle diagonal); otherwise false: the determinant tests decide.
  const uint32_t ns = (mask | (mask >> 1)) & 0x555u;   // bit 2s: side s has a triangle with det <= 0
  if (ns == 0u) return false;
  if ((G ? __ldg(&pc->planar) : pc->planar) == 0u) return false;
  const double tol = G ? __ldg(&pc->tol) : pc->tol;
  double ol[6], ox[6];
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    double a, b, c, d;
    load_plane4<G>(pc->pl[2 * o], a, b, c, d);
    ol[o] = fma(a, lp[0], fma(b, lp[1], fma(c, lp[2], -d)));
    ox[o] = fma(a, x[0], fma(b, x[1], fma(c, x[2], -d)));
  }
  int s = -1;
  double alpha = 2.0;
  bool ok = true;
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    if ((ns >> (2 * o)) & 1u) {
      if (!(ol[o] > tol && ox[o] < -tol)) ok = false;
      const double a = ol[o] / (ol[o] - ox[o]);   // crossing point = lp + a (x - lp)
      if (a < alpha) { alpha = a; s = o; }
    }
  }
  if (!ok || s < 0) return false;
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    const double oc = fma(alpha, ox[o] - ol[o], ol[o]);
    if (o != s && !(oc > tol)) ok = false;
  }
  double a, b, c, d;
  load_plane4<G>(pc->dg[s], a, b, c, d);
  const double gl = fma(a, lp[0], fma(b, lp[1], fma(c, lp[2], -d)));
  const double gx = fma(a, x[0], fma(b, x[1], fma(c, x[2], -d)));
  const double gc = fma(alpha, gx - gl, gl);
  if (!(fabs(gc) > tol)) ok = false;
  side = s;
  tri = gc > 0. ? 1 : 2;
  return ok;
}

// MODE 0: determinant tests only; 1: exit-side shortcut only (returns HOP_NO_SHORTCUT when it does not apply); 2: shortcut, then
// the determinant tests if it does not apply.
constexpr int HOP_NO_SHORTCUT = -2;
// reflectV(n): mirrors the particle's velocity at a wall with outward unit normal n (PerfectReflection, wall at rest).
template <bool FAST, bool G, int MODE, class PlaneOf, class ReflectV>
__device__ __forceinline__ int tria_hop(const TriaElem* __restrict__ te, const TriaElem* __restrict__ tria, const PlaneElem* __restrict__ plCur,
                                        PlaneOf planeOf, ReflectV reflectV, double x[3], double lp[3], int& ElemID, uint32_t& mask,
                                        HopHist& h) {
  int side = -1, tri = 0;
  bool shortcut = false;
  if (FAST && MODE != 0) shortcut = exit_side_planar<G>(plCur, x, lp, mask, side, tri);
  if (MODE == 1 && !shortcut) return HOP_NO_SHORTCUT;
  // unit vector and length of the flight (particle_triatracking.f90:171-173): needed by the determinant tests and the BCs
  double V[3] = {x[0] - lp[0], x[1] - lp[1], x[2] - lp[2]};
  double len = 0.;
  bool haveV = false;
  auto flight = [&]() {
    if (haveV) return;
    haveV = true;
    len = sqrt((V[0] * V[0] + V[1] * V[1]) + V[2] * V[2]);
    if (fabs(len) > 0.) {
      V[0] = V[0] / len; V[1] = V[1] / len; V[2] = V[2] / len;
    }
  };
  if (MODE != 1 && !shortcut) {
  flight();
  // 2b) crossed triangles among those with det <= 0
  uint32_t thr = 0;
  int nThrough = 0;
  uint32_t cand = mask;
#pragma unroll 1
  while (cand) {
    const int b = __ffs(cand) - 1;
    cand &= cand - 1;
    if (through_side_check_fast<G>(te, lp, V, b >> 1, (b & 1) + 1)) {
      thr |= 1u << b;
      ++nThrough;
    }
  }
  if (nThrough == 0) return TRK_LOST;
  if (nThrough == 1) {
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
      if (h.treated(ElemID, te->sideID[s], t)) continue;
      double detM;
      if (!through_side_lastpos_check<G>(te, lp, s, t, detM)) continue;
      double d1, d2;
      side_dets<G>(te, x, s, d1, d2);
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
    if (second == 0) return TRK_LOST;
  }
  }  // determinant tests
  // 3) boundary interaction or step into the neighbour
  const int gside = te->sideID[side];
  const int bc = te->bcid[side];
  const int oldElem = ElemID;
  // shortcut-only instantiation: boundary sides (few) are left to the general one, so that the warps of the common case do
  // not run through the periodic-shift arithmetic (square root, divisions, IntersectionWithWall) for a single lane
  if (MODE == 1 && bc > 0) return HOP_NO_SHORTCUT;
  if (bc > 0) {
    const int kind = cst.bc_kind[bc - 1];
    if (kind == PGPU_BC_OPEN) return TRK_REMOVED;
    if (kind == PGPU_BC_REFLECTIVE) {
      // GetBoundaryInteraction case 2 -> SurfaceModelling -> PerfectReflection (surfacemodel_tools.f90:81-256): specular wall at
      // rest.  The particle stays in ElemID; DoneLastElem is cleared (particle_triatracking.f90:426-427).
      flight();
      const double alpha = intersection_with_wall<G>(te, lp, V, side, tri);
      double n[3];
      triangle_normal<G>(te, side, tri, n);
      reflectV(n);
      const double tn = (V[0] * n[0] + V[1] * n[1]) + V[2] * n[2];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        lp[d] = lp[d] + V[d] * alpha;          // point of impact
        V[d] = V[d] - 2. * tn * n[d];
        x[d] = lp[d] + V[d] * (len - alpha);
      }
      h.clear();
      const bool inSame = FAST ? inside_fast<G>(plCur, te, x, mask) : inside_quad3d_mask<G>(te, x, mask);
      return inSame ? TRK_OK : -1;
    }
    if (kind != PGPU_BC_PERIODIC) return TRK_ERR_BC;
    flight();
    const double alpha = intersection_with_wall<G>(te, lp, V, side, tri);
    const int pvid = cst.bc_alpha[bc - 1];  // PeriodicBoundary, particle_boundary_condition.f90:224-284
    const int pv = (pvid < 0 ? -pvid : pvid) - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lp[d] = lp[d] + V[d] * alpha;
      lp[d] = lp[d] + copysign(cst.PeriodicVectors[pv][d], (double)pvid);
      x[d] = lp[d] + (len - alpha) * V[d];
    }
  }
  ElemID = te->nbElem[side];
  h.push(oldElem, gside, tri);
  if (ElemID < 1) return TRK_ERR_ELEM;
  // 2a) inside test in the new element
  const bool inNew = FAST ? inside_fast<G>(planeOf(side, ElemID), tria + (ElemID - 1), x, mask)
                          : inside_quad3d_mask<G>(tria + (ElemID - 1), x, mask);
  return inNew ? TRK_OK : -1;
}

// end of the walk of one particle: final position, element and sort key (emigrants: nElems + rank, removed: nElems + nRanks)
__device__ __forceinline__ void tria_finish(const PartBuf& pb, int64_t p, int status, int ElemID, const double x[3],
                                            const int32_t* __restrict__ elemRank, uint32_t* __restrict__ keys, int nElems, int offsetElem,
                                            int* __restrict__ counters) {
  uint32_t key;
  int newElem = ElemID;
  if (status == TRK_OK) {
    const int rk = (cst.nRanks == 1) ? cst.myRank : elemRank[newElem - 1];
    key = (rk == cst.myRank) ? (uint32_t)(newElem - 1 - offsetElem) : (uint32_t)(nElems + rk);
  } else {
    key = (uint32_t)(nElems + cst.nRanks);  // removed
    newElem = 0;
    if (status == TRK_LOST) atomicAdd(&counters[0], 1);
    else if (status != TRK_REMOVED) atomicMax(&counters[1], status);
  }
  pb.f[p] = x[0]; pb.f[1 * pb.stride + p] = x[1]; pb.f[2 * pb.stride + p] = x[2];
  pb.elem[p] = newElem;
  keys[p] = key;
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
                                                            uint32_t* __restrict__ leaverIdx, uint32_t* __restrict__ leaverHistE,
                                                            uint32_t* __restrict__ leaverHistS, const int32_t* __restrict__ elemRank,
                                                            double dt, int xiValid,
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
  constexpr bool QUEUE = FAST && !REF;
  __shared__ PlaneElem spn[QUEUE ? 6 : 1];   // planes of the six face neighbours: first element crossing without global loads
  // leaver queues, private to each warp (no block barriers between the phases): queue A grows from the front, B from the
  // back; A entries may move to B.  Entry: (offset in the chunk) << 12 | det <= 0 mask of the 12 triangles
#if KA_WARPQ
  constexpr int NQ = STEP_NT / 32, QSTRIDE = 32;
#define KA_QSYNC() __syncwarp()
  const int qg = threadIdx.x >> 5, qid = threadIdx.x & 31;
#else
  constexpr int NQ = 1, QSTRIDE = STEP_NT;
#define KA_QSYNC() __syncthreads()
  const int qg = 0, qid = threadIdx.x;
#endif
  constexpr int KA_CHUNK = (NP <= 5 ? KA_CHUNKF : (KA_CHUNKF + 1) / 2) * STEP_NT;
  constexpr int QCAP = 2 * KA_CHUNK / NQ;
  __shared__ uint32_t sQw[QUEUE ? NQ : 1][QUEUE ? QCAP : 1];
  __shared__ int sQnw[NQ], sQnBw[NQ];
  uint32_t* const sQ = sQw[QUEUE ? qg : 0];
  int& sQn = sQnw[qg];
  int& sQnB = sQnBw[qg];
  const int tid = threadIdx.x;
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int64_t p0 = elemOff[e], p1 = elemOff[e + 1];
    if (p1 <= p0) continue;
    const int gElem = offsetElem + e + 1;
    __syncthreads();
    if (!xiValid && !REF) stage_words(&sg, geo + (gElem - 1), sizeof(GeoElem));
    if (FAST && !REF) stage_words(&sa, aff + (gElem - 1), sizeof(AffElem));
    if (FAST && !REF) {
      stage_words(&sp, planes + (gElem - 1), sizeof(PlaneElem));
      constexpr int W = (int)(sizeof(PlaneElem) / 16);
      for (int t = tid; t < 6 * W; t += STEP_NT) {
        const int sd = t / W, w = t - sd * W;
        const int nb = __ldg(&tria[gElem - 1].nbElem[sd]);
        if (nb >= 1) reinterpret_cast<int4*>(&spn[sd])[w] = __ldg(reinterpret_cast<const int4*>(planes + (nb - 1)) + w);
      }
    }
    if (!REF) stage_words(&st, tria + (gElem - 1), sizeof(TriaElem));
    for (int t = threadIdx.x; t < ND * 3; t += STEP_NT) {
      const int c = t % 3, node = t / 3;
      const int i = node % NP, kj = node / NP;
      sE[(kj * 3 + c) * NP + i] = __ldg(E + (size_t)e * ND * 3 + t);
    }
    __syncthreads();
    // restructured arithmetic on an affine element: xi is recomputed from x in closed form (k_deposit_cvwm did not store it)
    const bool affineXi = FAST && !REF && sa.affine != 0.0;
    const bool useXi = REF || (xiValid && !affineXi);
    // the segment is worked off in chunks: phase 1 runs interpolation, push and the own-element inside test for every particle
    // of the chunk and queues the leavers in shared memory; phase 2 does their first element crossing with dense warps on the
    // staged records (restructured arithmetic only; otherwise the leavers go straight to k_track_leavers)
    for (int64_t c0 = p0; c0 < p1; c0 += KA_CHUNK) {
    const int64_t c1 = (c0 + KA_CHUNK < p1) ? c0 + KA_CHUNK : p1;
    if (QUEUE) {
      if (qid == 0) { sQn = 0; sQnB = 0; }
      KA_QSYNC();
    }
    int64_t p = c0 + tid;
    int stage = 0;
    uint8_t meta = 0, metaNext = 0;
    if (p < c1) {
#pragma unroll
      for (int a = 0; a < 6; ++a) cp_async8(&sP[0][a][tid], PF + a * PS_ + p);
      if (useXi) {
#pragma unroll
        for (int a = 0; a < 3; ++a) cp_async8(&sP[0][6 + a][tid], PXI + a * PS_ + p);
      }
      meta = pb.meta[p];
    }
    cp_async_commit();
    for (; p < c1; p += STEP_NT, stage ^= 1, meta = metaNext) {
      const int64_t pn = p + STEP_NT;
      if (pn < c1) {   // next particle of this thread: copies overlap the arithmetic below
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
        } else if (affineXi && affine_xi(&sa, x, xi)) {
          suc = true;
        } else if (xiValid) {
          if (useXi) { xi[0] = sP[stage][6][tid]; xi[1] = sP[stage][7][tid]; xi[2] = sP[stage][8][tid]; }
          else { xi[0] = PXI[p]; xi[1] = PXI[1 * PS_ + p]; xi[2] = PXI[2 * PS_ + p]; }   // general path of the deposition
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
        const unsigned act = __activemask();   // warp-aggregated append: one atomic per warp, consecutive slots
        const int ln = threadIdx.x & 31, leader = __ffs(act) - 1;
        int slot0 = 0;
        if (QUEUE) {
          // queue A (front): beyond exactly one side plane of a planar-sided element -> exit-side shortcut; queue B (back): the rest
          const bool isA = sp.planar != 0u && __popc((mask | (mask >> 1)) & 0x555u) == 1;
          const unsigned balA = __ballot_sync(act, isA);
          const unsigned grp = isA ? balA : (act & ~balA);
          const int gl = __ffs(grp) - 1;
          if (ln == gl) slot0 = atomicAdd(isA ? &sQn : &sQnB, __popc(grp));
          slot0 = __shfl_sync(act, slot0, gl);
          int slot = slot0 + __popc(grp & ((1u << ln) - 1u));
          if (!isA) slot = QCAP - 1 - slot;
          sQ[slot] = ((uint32_t)(p - c0) << 12) | mask;
        } else {
          keys[p] = mask;  // handed to k_track_leavers, which overwrites it with the final key
          if (ln == leader) slot0 = atomicAdd(&counters[2], __popc(act));
          slot0 = __shfl_sync(act, slot0, leader);
          const int slot = slot0 + __popc(act & ((1u << ln) - 1u));
          leaverIdx[slot] = (uint32_t)p;
          leaverHistE[slot] = 0u;   // no crossing done yet
          leaverHistS[slot] = 0u;
        }
      }
    }
    if (QUEUE) {
      // phase 2: first element crossing of the chunk's leavers on the staged records, queue A then queue B
      // pushed position and LastPartPos of a queued leaver come back through the (now idle) cp.async staging slots
      auto fetch = [&](int stg, uint32_t qe) {
        const int64_t q = c0 + (qe >> 12);
        cp_async8(&sP[stg][0][tid], xn0 + q);
        cp_async8(&sP[stg][1][tid], xn1 + q);
        cp_async8(&sP[stg][2][tid], xn2 + q);
#pragma unroll
        for (int a = 0; a < 3; ++a) cp_async8(&sP[stg][3 + a][tid], PF + a * PS_ + q);
      };
      auto crossing = [&](int64_t q, uint32_t mask, int stg, auto modeTag) -> bool {
        constexpr int MODE = decltype(modeTag)::value;
        double x[3] = {sP[stg][0][tid], sP[stg][1][tid], sP[stg][2][tid]};
        double lp[3] = {sP[stg][3][tid], sP[stg][4][tid], sP[stg][5][tid]};   // LastPartPos
        int ElemID = gElem;
        HopHist h;
        h.clear();
        const int status = tria_hop<true, false, MODE>(
            &st, tria, &sp, [&](int sd, int) { return (const PlaneElem*)&spn[sd]; },
            [&](const double n[3]) {   // the velocity has been written already: mirror it in place
              const double v0 = PF[3 * PS_ + q], v1 = PF[4 * PS_ + q], v2 = PF[5 * PS_ + q];
              const double vn = (v0 * n[0] + v1 * n[1]) + v2 * n[2];
              PF[3 * PS_ + q] = v0 - 2. * vn * n[0]; PF[4 * PS_ + q] = v1 - 2. * vn * n[1]; PF[5 * PS_ + q] = v2 - 2. * vn * n[2];
            },
            x, lp, ElemID, mask, h);
        if (status == HOP_NO_SHORTCUT) return false;
        if (status != -1) {
          tria_finish(pb, q, status, ElemID, x, elemRank, keys, nElems, offsetElem, counters);
          return true;
        }
        // not localised in the face neighbour: the walk continues in k_track_leavers from ElemID
        PF[q] = lp[0]; PF[1 * PS_ + q] = lp[1]; PF[2 * PS_ + q] = lp[2];   // LastPartPos after a periodic shift
        xn0[q] = x[0]; xn1[q] = x[1]; xn2[q] = x[2];
        pb.elem[q] = ElemID;
        keys[q] = mask;
        const unsigned act = __activemask();
        const int ln = threadIdx.x & 31, leader = __ffs(act) - 1;
        int slot0 = 0;
        if (ln == leader) slot0 = atomicAdd(&counters[2], __popc(act));
        slot0 = __shfl_sync(act, slot0, leader);
        const int slot = slot0 + __popc(act & ((1u << ln) - 1u));
        leaverIdx[slot] = (uint32_t)q;
        leaverHistE[slot] = (uint32_t)h.e0;                       // DoneLastElem entry of the crossing done here
        leaverHistS[slot] = (uint32_t)(h.s0 * 2 + (h.t0 == 2 ? 1 : 0));
        return true;
      };
      KA_QSYNC();
      const int nA = sQn;
      {
        int i = qid, stg = 0;
        if (i < nA) fetch(0, sQ[i]);
        cp_async_commit();
        for (; i < nA; i += QSTRIDE, stg ^= 1) {
          if (i + QSTRIDE < nA) fetch(stg ^ 1, sQ[i + QSTRIDE]);
          cp_async_commit();
          cp_async_wait_prev();
          const uint32_t qe = sQ[i];
          if (!crossing(c0 + (qe >> 12), qe & 0xfffu, stg, std::integral_constant<int, 1>()))   // margin not met: determinant tests
            sQ[QCAP - 1 - atomicAdd(&sQnB, 1)] = qe;
        }
      }
      KA_QSYNC();
      const int nB = sQnB;
      if (qid == 0) { atomicAdd(&counters[4], nA); atomicAdd(&counters[5], nB); }   // diagnostics (PICLAS_GPU_DEBUG)
      {
        int i = qid, stg = 0;
        if (i < nB) fetch(0, sQ[QCAP - 1 - i]);
        cp_async_commit();
        for (; i < nB; i += QSTRIDE, stg ^= 1) {
          if (i + QSTRIDE < nB) fetch(stg ^ 1, sQ[QCAP - 1 - (i + QSTRIDE)]);
          cp_async_commit();
          cp_async_wait_prev();
          const uint32_t qe = sQ[QCAP - 1 - i];
          crossing(c0 + (qe >> 12), qe & 0xfffu, stg, std::integral_constant<int, 0>());
        }
      }
      KA_QSYNC();
    }
    }  // chunks
  }
}

// ---- SingleParticleTriaTracking3D for the particles that left their element (particle_triatracking.f90:137-484) --------------------
// One thread per leaver, element records read from global memory (L2 resident).  All loops are rolled and the candidate
// triangles are visited through a bit mask so that the lanes of a warp run the same through-side test at the same time.
// (r1 profile of the first version, which looped per thread until done: every warp ran 2-3 rounds for 1.1 hops per particle,
// 12 of 32 lanes active; a shared-memory re-queue with block barriers was slower still.)
// With the restructured arithmetic k_interp_push has already done the first crossing; the list then holds only the particles
// that were not localised in a face neighbour (edge / corner crossings, multi-element flights).
constexpr int LV_NT = 128;
#ifndef LV_MINB
#define LV_MINB 5
#endif

template <bool FAST>
__global__ void __launch_bounds__(LV_NT, LV_MINB) k_track_leavers(PartBuf pb, const double* __restrict__ xn0, const double* __restrict__ xn1,
                                                         const double* __restrict__ xn2, const uint32_t* __restrict__ leaverIdx,
                                                         const uint32_t* __restrict__ leaverHistE, const uint32_t* __restrict__ leaverHistS,
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
  HopHist h;
  h.clear();
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
          ElemID = pb.elem[p];                                               // element the walk has reached
          mask = keys[p];                                                    // det <= 0 triangles of that element
          guard = 0;
          h.clear();
          const uint32_t he = leaverHistE[mine], hs = leaverHistS[mine];
          h.e0 = (int)he; h.s0 = (int)(hs >> 1); h.t0 = he ? (int)(hs & 1u) + 1 : 0;
        }
      }
      if (__ballot_sync(0xffffffffu, active) == 0) break;
    }
    if (active) {
      int status = tria_hop<FAST, true, FAST ? 2 : 0>(
          tria + (ElemID - 1), tria, FAST ? planes + (ElemID - 1) : nullptr, [&](int, int ne) { return planes + (ne - 1); },
          [&](const double n[3]) {
            const double v0 = PF[3 * PS_ + p], v1 = PF[4 * PS_ + p], v2 = PF[5 * PS_ + p];
            const double vn = (v0 * n[0] + v1 * n[1]) + v2 * n[2];
            PF[3 * PS_ + p] = v0 - 2. * vn * n[0]; PF[4 * PS_ + p] = v1 - 2. * vn * n[1]; PF[5 * PS_ + p] = v2 - 2. * vn * n[2];
          },
          x, lp, ElemID, mask, h);
      if (status == -1 && ++guard > 100000) status = TRK_ERR_LOOP;
      if (status != -1) {
        tria_finish(pb, p, status, ElemID, x, elemRank, keys, nElems, offsetElem, counters);
        active = false;
      }
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
