// ref.cuh — RefMapping tracking (TrackingMethod = refmapping) on sm_100a, one thread per particle.
//
// Reference: ParticleRefTracking particle_reftracking.f90:30-414, ParticleBCTracking :417-694 (tail recursion -> loop),
// ComputePlanarRectIntersection particle_intersection.f90:515-684, PARTHASMOVED particle_localization.f90:447-469,
// GetBoundaryInteraction (REFMAPPING branch) + PeriodicBoundary particle_boundary_condition.f90:35-284,
// InsertionSort utils.f90:52-101.  CartesianPeriodic = F; sides must be PLANAR_RECT (checked at init).
// Arithmetic is in the reference's operation order (the unit is compiled with --fmad=false), so ownership and the
// stored reference coordinates PartPosRef are bitwise equal to the CPU restatement.
#pragma once
#include "math.cuh"

struct RefTables {
  const GeoElem* geo;
  const int32_t* ElemToBCSides;   // [nGlobalElems][2]
  const double* SideBCMetrics;    // [nBCSidesTotal][7]
  const int32_t* SideInfo;        // [nSides][sideInfoSize]
  int sideInfoSize;
  const double *SideNormVec, *SideDistance, *BaseVectors0, *BaseVectors1, *BaseVectors2;
  const double *ElemBary, *ElemRadius, *ElemRadius2, *ElemEpsOneCell;
  const int32_t *FIBGM_nElems, *FIBGM_offsetElem, *FIBGM_Element;
};

#define REF_MAX_HITS 16
#define REF_MAX_BGM 32
#define REF_ALMOSTZERO(x) (fabs(x) <= 2.22e-16)

__device__ __forceinline__ double maxabs3(const double v[3]) { return fmax(fabs(v[0]), fmax(fabs(v[1]), fabs(v[2]))); }

__device__ __forceinline__ int side_flip(const RefTables& T, int SideID) {
  const int32_t* si = T.SideInfo + (size_t)(SideID - 1) * T.sideInfoSize;
  return (si[1] > 0) ? 0 : (si[3] % 10);
}

// ComputePlanarRectIntersection: returns alpha (-1 = no intersection)
__device__ __forceinline__ double planar_rect_intersection(const RefTables& T, const double traj[3], double len, const double lp[3],
                                                           int flip, int SideID) {
  const double* nv = T.SideNormVec + (size_t)(SideID - 1) * 3;
  double n0 = nv[0], n1 = nv[1], n2 = nv[2], locDistance = T.SideDistance[SideID - 1];
  if (flip != 0) { n0 = -n0; n1 = -n1; n2 = -n2; locDistance = -locDistance; }
  const double coeffA = (n0 * traj[0] + n1 * traj[1]) + n2 * traj[2];
  const double locSideDistance = locDistance - ((lp[0] * n0 + lp[1] * n1) + lp[2] * n2);
  if (REF_ALMOSTZERO(coeffA)) return -1.;
  const double alpha = locSideDistance / coeffA;
  if (locSideDistance < -100 * EPSMACH) return -1.;
  const double alphaNorm = alpha / len;
  if ((alphaNorm > 1.0) || (alphaNorm < -(100. * EPSMACH))) return -1.;
  const double* b0 = T.BaseVectors0 + (size_t)(SideID - 1) * 3;
  const double* b1 = T.BaseVectors1 + (size_t)(SideID - 1) * 3;
  const double* b2 = T.BaseVectors2 + (size_t)(SideID - 1) * 3;
  double P0[3], P1[3], P2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double inter = lp[d] + alpha * traj[d];
    P0[d] = -0.25 * b0[d] + inter;
    P1[d] = 0.25 * b1[d];
    P2[d] = 0.25 * b2[d];
  }
  const double A1 = (P1[0] * P1[0] + P1[1] * P1[1]) + P1[2] * P1[2];
  const double B1 = (P2[0] * P1[0] + P2[1] * P1[1]) + P2[2] * P1[2];
  const double C1 = (P1[0] * P0[0] + P1[1] * P0[1]) + P1[2] * P0[2];
  const double A2 = B1;
  const double B2 = (P2[0] * P2[0] + P2[1] * P2[1]) + P2[2] * P2[2];
  const double C2 = (P2[0] * P0[0] + P2[1] * P0[1]) + P2[2] * P0[2];
  double sdet = A1 * B2 - A2 * B1;
  sdet = 1.0 / sdet;
  const double epsLoc = 1.0 + 100. * EPSMACH;
  const double xi = (B2 * C1 - B1 * C2) * sdet;
  if (fabs(xi) > epsLoc) return -1.;
  const double eta = (-A2 * C1 + A1 * C2) * sdet;
  if (fabs(eta) > epsLoc) return -1.;
  return alpha;
}

// ParticleBCTracking.  status: TRK_OK (continue with the Newton check of the caller unless done), TRK_REMOVED, TRK_ERR_*
// done = PartisDone.  ElemID in/out.
__device__ int bc_tracking(const RefTables& T, double x[3], double lp[3], int& ElemID, bool& done, int& outElem) {
  double len0 = 0.;
  done = false;
  for (int iCount = 0; iCount < 100000; ++iCount) {
    const int nloc = T.ElemToBCSides[(size_t)(ElemID - 1) * 2];
    const int first = T.ElemToBCSides[(size_t)(ElemID - 1) * 2 + 1];   // 0-based offset: sides first+1 .. first+nloc (1-based)
    double traj[3] = {x[0] - lp[0], x[1] - lp[1], x[2] - lp[2]};
    double len = sqrt((traj[0] * traj[0] + traj[1] * traj[1]) + traj[2] * traj[2]);
    if (REF_ALMOSTZERO(len / T.ElemRadius[ElemID - 1])) {
      outElem = ElemID;
      done = true;
      return TRK_OK;
    }
    traj[0] = traj[0] / len; traj[1] = traj[1] / len; traj[2] = traj[2] / len;
    len0 = fmax(len0, len);
    bool doTracing = true, doubleCheck = false, recurse = false;
    while (doTracing) {
      // hits in visiting order; the reference sorts (alpha, side) of all listed sides with a stable insertion sort and walks the
      // entries with alpha > -1: sorting the hits alone gives the same sequence
      double hitAlpha[REF_MAX_HITS];
      int hitSide[REF_MAX_HITS];
      int nInter = 0;
      for (int il = 0; il < nloc; ++il) {
        const double* bm = T.SideBCMetrics + (size_t)(first + il) * 7;
        if (bm[2] > len0) break;
        const int SideID = (int)bm[0];
        const double a = planar_rect_intersection(T, traj, len, lp, side_flip(T, SideID), SideID);
        if (a > -1.0) {
          if (nInter >= REF_MAX_HITS) return TRK_ERR_LOOP;
          hitAlpha[nInter] = a;
          hitSide[nInter] = SideID;
          ++nInter;
        }
      }
      if (nInter == 0) {
        doTracing = false;
      } else {
        for (int i = 1; i < nInter; ++i) {  // InsertionSort
          int j = i - 1;
          const double tr = hitAlpha[i];
          const int ti = hitSide[i];
          while (j >= 0) {
            if (hitAlpha[j] <= tr) break;
            hitAlpha[j + 1] = hitAlpha[j];
            hitSide[j + 1] = hitSide[j];
            --j;
          }
          hitAlpha[j + 1] = tr;
          hitSide[j + 1] = ti;
        }
        bool reflected = false;
        for (int ih = 0; ih < nInter; ++ih) {
          const int SideID = hitSide[ih];
          const int flip = side_flip(T, SideID);
          const int OldElemID = ElemID;
          const double alpha = hitAlpha[ih];
          const double* nv = T.SideNormVec + (size_t)(SideID - 1) * 3;
          double n0 = nv[0], n1 = nv[1], n2 = nv[2];
          if (flip != 0) { n0 = -n0; n1 = -n1; n2 = -n2; }
          reflected = false;
          if (!(((n0 * traj[0] + n1 * traj[1]) + n2 * traj[2]) <= 0.)) {
            reflected = true;
            const int32_t* si = T.SideInfo + (size_t)(SideID - 1) * T.sideInfoSize;
            const int bc = si[4];
            const int kind = cst.bc_kind[bc - 1];
            if (kind == PGPU_BC_OPEN) return TRK_REMOVED;
            if (kind != PGPU_BC_PERIODIC) return TRK_ERR_BC;
            const int pvid = cst.bc_alpha[bc - 1];
            const int pv = (pvid < 0 ? -pvid : pvid) - 1;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              lp[d] = lp[d] + traj[d] * alpha;
              lp[d] = lp[d] + copysign(cst.PeriodicVectors[pv][d], (double)pvid);
              x[d] = lp[d] + (len - alpha) * traj[d];
            }
            len = len - alpha;
            ElemID = si[2];  // SIDE_NBELEMID
          }
          if (ElemID != OldElemID) {
            if (cst.nPeriodicVectors > 0) {
              const int on = T.ElemToBCSides[(size_t)(OldElemID - 1) * 2], of = T.ElemToBCSides[(size_t)(OldElemID - 1) * 2 + 1];
              double mx = -HUGE_D;
              for (int k = 0; k < on; ++k) mx = fmax(mx, T.SideBCMetrics[(size_t)(of + k) * 7 + 2]);
              len0 = mx;
            }
            recurse = true;
            break;
          }
          if (reflected) break;
        }
        if (recurse) break;
        if (!reflected) {
          if (!doubleCheck) doubleCheck = true;
          else doTracing = false;
        }
      }
    }
    if (!recurse) return TRK_OK;
    if (ElemID < 1) return TRK_ERR_ELEM;
  }
  return TRK_ERR_LOOP;
}

// GetPositionInRefElem without isSuccessful / ForceMode (iMode = 2): never aborts
__device__ __forceinline__ void ref_newton(const RefTables& T, const double x[3], double xi[3], int ElemID) {
  position_in_ref_elem(T.geo + (ElemID - 1), x, xi, false, false);
}

// ParticleRefTracking for one particle: x pushed position, lp LastPartPos, xi PartPosRef (in: old, out: new), elem in/out
__device__ int ref_tracking(const RefTables& T, double x[3], double lp[3], double xi[3], int& elem) {
  const int LastElem = elem;
  int ElemID = LastElem;
  bool done = false;
  if (T.ElemToBCSides[(size_t)(ElemID - 1) * 2] > 0) {
    int outElem = ElemID;
    const int st = bc_tracking(T, x, lp, ElemID, done, outElem);
    if (st != TRK_OK) return st;
    if (done) { elem = outElem; return TRK_OK; }
    ref_newton(T, x, xi, ElemID);
    if (maxabs3(xi) < 1.0) { elem = ElemID; return TRK_OK; }
  } else {
    ref_newton(T, x, xi, ElemID);
    if (maxabs3(xi) < 1.0) { elem = ElemID; return TRK_OK; }
  }
  // relocate through the background mesh cell of the particle
  int oldElemID = LastElem;
  int Cell[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    Cell[d] = max((int)floor((x[d] - cst.xyzminglob[d]) / cst.FIBGMdeltas[d]), 0) + 1;
    Cell[d] = min(cst.FIBGMmax[d], Cell[d]);
  }
  const int ni = cst.FIBGMmax[0] - cst.FIBGMmin[0] + 1, nj = cst.FIBGMmax[1] - cst.FIBGMmin[1] + 1;
  const size_t cell = (size_t)(Cell[0] - cst.FIBGMmin[0]) + (size_t)ni * ((size_t)(Cell[1] - cst.FIBGMmin[1]) + (size_t)nj * (size_t)(Cell[2] - cst.FIBGMmin[2]));
  const int nBGM = T.FIBGM_nElems[cell];
  if (nBGM > REF_MAX_BGM) return TRK_ERR_LOOP;
  double Distance[REF_MAX_BGM];
  int List[REF_MAX_BGM];
  if (nBGM > 1) {
    for (int i = 0; i < nBGM; ++i) {
      const int e = T.FIBGM_Element[T.FIBGM_offsetElem[cell] + i];
      List[i] = e;
      if (e == oldElemID) Distance[i] = -HUGE_D;
      else {
        const double* b = T.ElemBary + (size_t)(e - 1) * 3;
        const double d0 = x[0] - b[0], d1 = x[1] - b[1], d2 = x[2] - b[2];
        Distance[i] = (d0 * d0 + d1 * d1) + d2 * d2;
        if (Distance[i] > T.ElemRadius2[e - 1]) Distance[i] = -HUGE_D;
      }
    }
    for (int i = 1; i < nBGM; ++i) {  // InsertionSort
      int j = i - 1;
      const double tr = Distance[i];
      const int ti = List[i];
      while (j >= 0) {
        if (Distance[j] <= tr) break;
        Distance[j + 1] = Distance[j];
        List[j + 1] = List[j];
        --j;
      }
      Distance[j + 1] = tr;
      List[j + 1] = ti;
    }
  } else if (nBGM == 1) {
    Distance[0] = 0.;
    List[0] = T.FIBGM_Element[T.FIBGM_offsetElem[cell]];
  }
  const double OldXi[3] = {xi[0], xi[1], xi[2]};
  double newXi[3] = {HUGE_D, HUGE_D, HUGE_D};
  int newElemID = -1;
  for (int i = 0; i < nBGM; ++i) {
    if (Distance[i] == -HUGE_D) continue;
    ElemID = List[i];
    ref_newton(T, x, xi, ElemID);
    if (maxabs3(xi) < 1.0) { elem = ElemID; return TRK_OK; }
    if (maxabs3(xi) < maxabs3(newXi)) { newXi[0] = xi[0]; newXi[1] = xi[1]; newXi[2] = xi[2]; newElemID = ElemID; }
  }
  int Test;
  if (maxabs3(OldXi) < maxabs3(newXi)) {
    xi[0] = OldXi[0]; xi[1] = OldXi[1]; xi[2] = OldXi[2];
    Test = oldElemID;
  } else {
    xi[0] = newXi[0]; xi[1] = newXi[1]; xi[2] = newXi[2];
    Test = newElemID;
  }
  if (Test < 1) return TRK_ERR_ELEM;
  elem = Test;
  if (maxabs3(xi) > T.ElemEpsOneCell[Test - 1]) {
    if (T.ElemToBCSides[(size_t)(Test - 1) * 2] <= 0) return TRK_ERR_ELEM;  // tolerance issue with internal element: abort
    int outElem = Test;
    const int st = bc_tracking(T, x, lp, Test, done, outElem);
    if (st != TRK_OK) return st;
    if (done) { elem = outElem; return TRK_OK; }
    ref_newton(T, x, xi, Test);
    if (maxabs3(xi) > T.ElemEpsOneCell[Test - 1]) return TRK_ERR_ELEM;    // LocateParticleInElement fallback not built
    elem = Test;
  }
  return TRK_OK;
}

// PerformTracking (REFMAPPING) for all particles: pb.x holds LastPartPos, xn the pushed position (k_interp_push, REF mode)
__global__ void __launch_bounds__(128) k_track_ref(PartBuf pb, const double* __restrict__ xn0, const double* __restrict__ xn1,
                                                   const double* __restrict__ xn2, int64_t n, RefTables T,
                                                   const int32_t* __restrict__ elemRank, uint32_t* __restrict__ keys, int nElems,
                                                   int offsetElem, int* __restrict__ counters) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    double x[3] = {xn0[p], xn1[p], xn2[p]};
    double lp[3] = {pb.x[0][p], pb.x[1][p], pb.x[2][p]};
    double xi[3] = {pb.xi[0][p], pb.xi[1][p], pb.xi[2][p]};
    int elem = pb.elem[p];
    const int status = ref_tracking(T, x, lp, xi, elem);
    uint32_t key;
    if (status == TRK_OK) {
      const int rk = elemRank[elem - 1];
      key = (rk == cst.myRank) ? (uint32_t)(elem - 1 - offsetElem) : (uint32_t)(nElems + rk);
    } else {
      key = (uint32_t)(nElems + cst.nRanks);
      elem = 0;
      if (status == TRK_LOST) atomicAdd(&counters[0], 1);
      else if (status != TRK_REMOVED) atomicMax(&counters[1], status);
    }
    pb.x[0][p] = x[0]; pb.x[1][p] = x[1]; pb.x[2][p] = x[2];
    pb.xi[0][p] = xi[0]; pb.xi[1][p] = xi[1]; pb.xi[2][p] = xi[2];
    pb.elem[p] = elem;
    keys[p] = key;
  }
}

// PartPosRef of uploaded particles when the host did not supply it (GetPositionInRefElem as at emission)
__global__ void k_init_posref(PartBuf pb, int64_t p0, int64_t n, const GeoElem* __restrict__ geo) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = p0 + i;
  const int e = pb.elem[p];
  if (e < 1) return;
  const double x[3] = {pb.x[0][p], pb.x[1][p], pb.x[2][p]};
  double xi[3];
  position_in_ref_elem(geo + (e - 1), x, xi, false, false);
  pb.xi[0][p] = xi[0]; pb.xi[1][p] = xi[1]; pb.xi[2][p] = xi[2];
}
