// ref.cuh — RefMapping tracking (TrackingMethod = refmapping) on sm_100a, one thread per particle.
//
// Reference: ParticleRefTracking particle_reftracking.f90:30-414, ParticleBCTracking :417-694 (tail recursion -> loop),
// ComputePlanarRectIntersection particle_intersection.f90:515-684, PARTHASMOVED particle_localization.f90:447-469,
// GetBoundaryInteraction (REFMAPPING branch) + PeriodicBoundary particle_boundary_condition.f90:35-284,
// ComputeBiLinearIntersection :855-1240 (PLANAR_NONRECT and BILINEAR sides) with QuadraticSolver utils.f90:401-465,
// CalcNormAndTangBilinear particle_surfaces.f90:404-440, LocateParticleInElement particle_localization.f90:47-190,
// InsertionSort utils.f90:52-101.  CartesianPeriodic = F; curved sides (NGeo > 1) are rejected at init.
// Arithmetic is in the reference's operation order (the unit is compiled with --fmad=false), so ownership and the
// stored reference coordinates PartPosRef are bitwise equal to the CPU restatement.
#pragma once
#include "math.cuh"
#include "select.cuh"

struct RefTables {
  const GeoElem* geo;
  const int32_t* ElemToBCSides;   // [nGlobalElems][2]
  const double* SideBCMetrics;    // [nBCSidesTotal][7]
  const int32_t* SideInfo;        // [nSides][sideInfoSize]
  int sideInfoSize;
  const double *SideNormVec, *SideDistance, *BaseVectors0, *BaseVectors1, *BaseVectors2, *BaseVectors3;
  const int32_t* SideType;        // [nSides] 0 PLANAR_RECT, 1 PLANAR_NONRECT, 2 BILINEAR
  const double *ElemBary, *ElemRadius, *ElemRadius2, *ElemEpsOneCell;
  const int32_t *FIBGM_nElems, *FIBGM_offsetElem, *FIBGM_Element;
};

#define REF_MAX_HITS 16
#define REF_MAX_BGM 32   // candidates of a FIBGM cell kept as a sorted per-thread list; fuller cells are visited by repeated selection
#define REF_ALMOSTZERO(x) (fabs(x) <= 2.22e-16)
#define REF_ALMOSTEQUAL(x, y) (fabs((x) - (y)) <= fmax(fabs(x), fabs(y)) * 4.441e-16)

// Sort keys of the two relocation searches for cells with more than REF_MAX_BGM elements (select.cuh).
// ParticleRefTracking :192-216: squared distance to the barycentre; the old element and elements out of reach are skipped.
struct RelocKey {
  const RefTables& T;
  size_t off;
  const double* x;
  int oldElemID;
  __host__ __device__ double operator()(int i) const {
    const int e = T.FIBGM_Element[off + i];
    if (e == oldElemID) return -HUGE_D;
    const double* b = T.ElemBary + (size_t)(e - 1) * 3;
    const double d0 = x[0] - b[0], d1 = x[1] - b[1], d2 = x[2] - b[2];
    const double D = (d0 * d0 + d1 * d1) + d2 * d2;
    return (D <= T.ElemRadius2[e - 1]) ? D : -HUGE_D;   // (a NaN distance is out of reach)
  }
};
// SinglePointToElement (particle_localization.f90:81-190): elements out of reach carry -1 and are skipped.
struct LocateKey {
  const RefTables& T;
  size_t off;
  const double* x;
  __host__ __device__ double operator()(int i) const {
    const int e = T.FIBGM_Element[off + i];
    const double* b = T.ElemBary + (size_t)(e - 1) * 3;
    const double d0 = x[0] - b[0], d1 = x[1] - b[1], d2 = x[2] - b[2];
    const double D2 = (d0 * d0 + d1 * d1) + d2 * d2;
    return (D2 <= T.ElemRadius2[e - 1]) ? D2 : -1.;
  }
};

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

// ComputeBiLinearIntersection (refmapping).  Returns alpha (-1: no intersection), xi/eta of the intersection; err is set when the
// reference aborts ("Invalid intersection with bilinear side").  alpha2 <= -1: not given.
__device__ __noinline__ double bilinear_intersection(const RefTables& T, const double traj[3], double len, const double lp[3], int SideID,
                                                     double alpha2, double& xiOut, double& etaOut, bool& err) {
  xiOut = -2.0; etaOut = -2.0;
  const double* b0 = T.BaseVectors0 + (size_t)(SideID - 1) * 3;
  const double* b1 = T.BaseVectors1 + (size_t)(SideID - 1) * 3;
  const double* b2 = T.BaseVectors2 + (size_t)(SideID - 1) * 3;
  const double* b3 = T.BaseVectors3 + (size_t)(SideID - 1) * 3;
  double BC[4][3], NC[4][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    BC[0][d] = 0.25 * b3[d];
    BC[1][d] = 0.25 * b1[d];
    BC[2][d] = 0.25 * b2[d];
    BC[3][d] = 0.25 * b0[d];
  }
  const double* nv = T.SideNormVec + (size_t)(SideID - 1) * 3;
  const double scaleFac = (traj[0] * nv[0] + traj[1] * nv[1]) + traj[2] * nv[2];
  if (fabs(scaleFac) < 100. * EPSMACH) return -1.;
#pragma unroll
  for (int d = 0; d < 3; ++d) BC[3][d] = BC[3][d] - lp[d];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double sp = (BC[c][0] * traj[0] + BC[c][1] * traj[1]) + BC[c][2] * traj[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) NC[c][d] = BC[c][d] - sp * traj[d];
  }
  double A1[4], A2[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { A1[c] = NC[c][2] - NC[c][0]; A2[c] = NC[c][2] - NC[c][1]; }
  const double A = A1[0] * A2[2] - A2[0] * A1[2];
  const double B = A1[0] * A2[3] - A2[0] * A1[3] + A1[1] * A2[2] - A2[1] * A1[2];
  const double C = A1[1] * A2[3] - A2[1] * A1[3];
  // QuadraticSolver
  int nRoot;
  double eta[2] = {0., 0.}, xi[2] = {0., 0.}, t[2] = {-1., -1.};
  if (A != 0. && B == 0. && C == 0.) nRoot = 1;
  else if (A != 0.) {
    const double radicant = (0.5 * B / A) * (0.5 * B / A) - (C / A);
    if (radicant < 0.) nRoot = 0;
    else {
      nRoot = 2;
      eta[0] = -0.5 * (B / A) - copysign(1., B / A) * sqrt(radicant);
      eta[1] = (C / A) / eta[0];
    }
  } else if (B != 0.) { nRoot = 1; eta[0] = -C / B; }
  else nRoot = 0;
  if (nRoot == 0) return -1.;
  int InterType = 0;
  for (int r = 0; r < nRoot; ++r) {
    if (!(fabs(eta[r]) <= 1.0)) continue;
    {  // ComputeXi
      const double a = eta[r] * A2[0] + A2[1];
      const double b = eta[r] * (A2[0] - A1[0]) + A2[1] - A1[1];
      if (fabs(b) >= fabs(a)) {
        if (REF_ALMOSTZERO(fabs(b))) { err = true; return -1.; }
        xi[r] = (-eta[r] * (A2[2] - A1[2]) - (A2[3] - A1[3])) / b;
      } else {
        xi[r] = (-eta[r] * A2[2] - A2[3]) / a;
      }
    }
    if (!(fabs(xi[r]) <= 1.0)) continue;
    double tt = 0.;   // ComputeSurfaceDistance2
#pragma unroll
    for (int d = 0; d < 3; ++d) tt = tt + (xi[r] * eta[r] * BC[0][d] + xi[r] * BC[1][d] + eta[r] * BC[2][d] + BC[3][d]) * traj[d];
    t[r] = tt;
    if (alpha2 > -1.0 && REF_ALMOSTEQUAL(t[r], alpha2)) t[r] = -1.0;
    const double alphaNorm = t[r] / len;
    if (alphaNorm <= 1.0 && alphaNorm >= 0.) InterType += r + 1;
  }
  if (InterType == 0) return -1.;
  int k = (InterType == 2) ? 1 : 0;
  if (InterType == 3) {
    const int32_t* si = T.SideInfo + (size_t)(SideID - 1) * T.sideInfoSize;
    if (si[4] > 0) k = (t[0] < t[1]) ? 0 : 1;   // the first of the two intersections
    else { xiOut = 0.; etaOut = 0.; return -1.; }
  }
  xiOut = xi[k];
  etaOut = eta[k];
  return t[k];
}

// CalcNormAndTangBilinear (nVec), oriented like SideNormVec (out of the master element)
__device__ __forceinline__ void bilinear_normal(const RefTables& T, double xi, double eta, int SideID, double n[3]) {
  const double* b1 = T.BaseVectors1 + (size_t)(SideID - 1) * 3;
  const double* b2 = T.BaseVectors2 + (size_t)(SideID - 1) * 3;
  const double* b3 = T.BaseVectors3 + (size_t)(SideID - 1) * 3;
  double a[3], b[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    b[d] = xi * 0.25 * b3[d] + 0.25 * b2[d];
    a[d] = eta * 0.25 * b3[d] + 0.25 * b1[d];
  }
  const double c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
  const double l = sqrt((c0 * c0 + c1 * c1) + c2 * c2);
  n[0] = c0 / l; n[1] = c1 / l; n[2] = c2 / l;
  const double* nv = T.SideNormVec + (size_t)(SideID - 1) * 3;
  if ((n[0] * nv[0] + n[1] * nv[1]) + n[2] * nv[2] < 0.) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}

// ParticleBCTracking.  status: TRK_OK (continue with the Newton check of the caller unless done), TRK_REMOVED, TRK_ERR_*
// done = PartisDone.  ElemID in/out.
__device__ int bc_tracking(const RefTables& T, const PartBuf& pb, int64_t p, double x[3], double lp[3], int& ElemID, bool& done,
                           int& outElem) {
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
    double alphaOld = -1.0;
    while (doTracing) {
      // hits in visiting order; the reference sorts (alpha, side) of all listed sides with a stable insertion sort and walks the
      // entries with alpha > -1: sorting the hits alone gives the same sequence
      double hitAlpha[REF_MAX_HITS], hitXi[REF_MAX_HITS], hitEta[REF_MAX_HITS];
      int hitSide[REF_MAX_HITS];
      int nInter = 0;
      for (int il = 0; il < nloc; ++il) {
        const double* bm = T.SideBCMetrics + (size_t)(first + il) * 7;
        if (bm[2] > len0) break;
        const int SideID = (int)bm[0];
        double a, hx = -2., he = -2.;
        if (T.SideType[SideID - 1] == 0) a = planar_rect_intersection(T, traj, len, lp, side_flip(T, SideID), SideID);
        else {   // PLANAR_NONRECT, BILINEAR; the double check excludes the intersection found before (alpha2 = alphaOld)
          bool err = false;
          a = bilinear_intersection(T, traj, len, lp, SideID, doubleCheck ? alphaOld : -2., hx, he, err);
          if (err) return TRK_ERR_ELEM;
        }
        if (a > -1.0) {
          if (nInter >= REF_MAX_HITS) return TRK_ERR_LOOP;
          hitAlpha[nInter] = a;
          hitSide[nInter] = SideID;
          hitXi[nInter] = hx;
          hitEta[nInter] = he;
          ++nInter;
        }
      }
      if (nInter == 0) {
        doTracing = false;
      } else {
        for (int i = 1; i < nInter; ++i) {  // InsertionSort
          int j = i - 1;
          const double tr = hitAlpha[i], tx = hitXi[i], te = hitEta[i];
          const int ti = hitSide[i];
          while (j >= 0) {
            if (hitAlpha[j] <= tr) break;
            hitAlpha[j + 1] = hitAlpha[j];
            hitSide[j + 1] = hitSide[j];
            hitXi[j + 1] = hitXi[j];
            hitEta[j + 1] = hitEta[j];
            --j;
          }
          hitAlpha[j + 1] = tr;
          hitSide[j + 1] = ti;
          hitXi[j + 1] = tx;
          hitEta[j + 1] = te;
        }
        alphaOld = hitAlpha[0];
        bool reflected = false;
        for (int ih = 0; ih < nInter; ++ih) {
          const int SideID = hitSide[ih];
          const int flip = side_flip(T, SideID);
          const int OldElemID = ElemID;
          const double alpha = hitAlpha[ih];
          const double* nv = T.SideNormVec + (size_t)(SideID - 1) * 3;
          double n0 = nv[0], n1 = nv[1], n2 = nv[2];
          if (T.SideType[SideID - 1] == 2) {   // BILINEAR: normal at the intersection point
            double nb[3];
            bilinear_normal(T, hitXi[ih], hitEta[ih], SideID, nb);
            n0 = nb[0]; n1 = nb[1]; n2 = nb[2];
          }
          if (flip != 0) { n0 = -n0; n1 = -n1; n2 = -n2; }
          reflected = false;
          if (!(((n0 * traj[0] + n1 * traj[1]) + n2 * traj[2]) <= 0.)) {
            reflected = true;
            const int32_t* si = T.SideInfo + (size_t)(SideID - 1) * T.sideInfoSize;
            const int bc = si[4];
            const int kind = cst.bc_kind[bc - 1];
            if (kind == PGPU_BC_OPEN) return TRK_REMOVED;
            if (kind == PGPU_BC_REFLECTIVE) {
              // PerfectReflection (surfacemodel_tools.f90:81-256), wall at rest: mirrored velocity and trajectory, the rest of the
              // flight from the point of impact; the particle stays in ElemID and the side loop starts again
              const double nn[3] = {n0, n1, n2};
              double v[3] = {pb.v[0][p], pb.v[1][p], pb.v[2][p]};
              const double vn = (v[0] * nn[0] + v[1] * nn[1]) + v[2] * nn[2];
              const double tn = (traj[0] * nn[0] + traj[1] * nn[1]) + traj[2] * nn[2];
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                pb.v[d][p] = v[d] - 2. * vn * nn[d];
                lp[d] = lp[d] + traj[d] * alpha;
                traj[d] = traj[d] - 2. * tn * nn[d];
                x[d] = lp[d] + traj[d] * (len - alpha);
                traj[d] = x[d] - lp[d];
              }
              len = sqrt((traj[0] * traj[0] + traj[1] * traj[1]) + traj[2] * traj[2]);
              if (REF_ALMOSTZERO(len)) len = 0.0;
              else { traj[0] = traj[0] / len; traj[1] = traj[1] / len; traj[2] = traj[2] / len; }
            } else {
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
// relocated: the LocateParticleInElement fallback found the element (the reference then sets PDM%isNewPart)
__device__ int ref_tracking(const RefTables& T, const PartBuf& pb, int64_t p, double x[3], double lp[3], double xi[3], int& elem,
                            bool& relocated) {
  relocated = false;
  const int LastElem = elem;
  int ElemID = LastElem;
  bool done = false;
  if (T.ElemToBCSides[(size_t)(ElemID - 1) * 2] > 0) {
    int outElem = ElemID;
    const int st = bc_tracking(T, pb, p, x, lp, ElemID, done, outElem);
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
  const bool bigCell = nBGM > REF_MAX_BGM;   // the list does not fit: same visiting order by repeated selection (select.cuh)
  double Distance[REF_MAX_BGM];
  int List[REF_MAX_BGM];
  if (bigCell) {
  } else if (nBGM > 1) {
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
  if (bigCell) {
    const RelocKey key{T, (size_t)T.FIBGM_offsetElem[cell], x, oldElemID};
    SortedVisit sv;
    for (int i = next_in_sorted_order(nBGM, key, -HUGE_D, sv); i >= 0; i = next_in_sorted_order(nBGM, key, -HUGE_D, sv)) {
      ElemID = T.FIBGM_Element[T.FIBGM_offsetElem[cell] + i];
      ref_newton(T, x, xi, ElemID);
      if (maxabs3(xi) < 1.0) { elem = ElemID; return TRK_OK; }
      if (maxabs3(xi) < maxabs3(newXi)) { newXi[0] = xi[0]; newXi[1] = xi[1]; newXi[2] = xi[2]; newElemID = ElemID; }
    }
  }
  for (int i = 0; i < (bigCell ? 0 : nBGM); ++i) {
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
    const int st = bc_tracking(T, pb, p, x, lp, Test, done, outElem);
    if (st != TRK_OK) return st;
    if (done) { elem = outElem; return TRK_OK; }
    ref_newton(T, x, xi, Test);
    if (maxabs3(xi) > T.ElemEpsOneCell[Test - 1]) {
      // LocateParticleInElement(doHALO = T) -> SinglePointToElement: cell by CEILING, all elements of the cell whose radius reaches
      // the particle, nearest barycentre first, accepted with MAXVAL(ABS(xi)) <= ElemEpsOneCell
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Cell[d] = (int)ceil((x[d] - cst.xyzminglob[d]) / cst.FIBGMdeltas[d]);
        Cell[d] = max(min(cst.FIBGMmax[d], Cell[d]), cst.FIBGMmin[d]);
      }
      const size_t cell2 = (size_t)(Cell[0] - cst.FIBGMmin[0]) + (size_t)ni * ((size_t)(Cell[1] - cst.FIBGMmin[1]) + (size_t)nj * (size_t)(Cell[2] - cst.FIBGMmin[2]));
      const int nB = T.FIBGM_nElems[cell2];
      if (nB > REF_MAX_BGM) {   // same search by repeated selection
        const LocateKey key{T, (size_t)T.FIBGM_offsetElem[cell2], x};
        SortedVisit sv;
        int found = -1;
        for (int i = next_in_sorted_order(nB, key, -1., sv); i >= 0; i = next_in_sorted_order(nB, key, -1., sv)) {
          const int e = T.FIBGM_Element[T.FIBGM_offsetElem[cell2] + i];
          ref_newton(T, x, xi, e);
          if (maxabs3(xi) <= T.ElemEpsOneCell[e - 1]) { found = e; break; }
        }
        if (found < 1) return TRK_ERR_ELEM;   // 'Particle not inside of Element' or no candidate accepted
        ref_newton(T, x, xi, found);
        elem = found;
        relocated = true;
        return TRK_OK;
      }
      double mx = -1.;
      for (int i = 0; i < nB; ++i) {
        const int e = T.FIBGM_Element[T.FIBGM_offsetElem[cell2] + i];
        const double* b = T.ElemBary + (size_t)(e - 1) * 3;
        const double d0 = x[0] - b[0], d1 = x[1] - b[1], d2 = x[2] - b[2];
        const double D2 = (d0 * d0 + d1 * d1) + d2 * d2;
        Distance[i] = (D2 <= T.ElemRadius2[e - 1]) ? D2 : -1.;
        List[i] = e;
        mx = fmax(mx, Distance[i]);
      }
      if (nB < 1 || REF_ALMOSTEQUAL(mx, -1.)) return TRK_ERR_ELEM;   // 'Particle not inside of Element'
      for (int i = 1; i < nB; ++i) {  // InsertionSort
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
      int found = -1;
      for (int i = 0; i < nB; ++i) {
        if (REF_ALMOSTEQUAL(Distance[i], -1.)) continue;
        ref_newton(T, x, xi, List[i]);
        if (maxabs3(xi) <= T.ElemEpsOneCell[List[i] - 1]) { found = List[i]; break; }
      }
      if (found < 1) return TRK_ERR_ELEM;
      ref_newton(T, x, xi, found);
      elem = found;
      relocated = true;
      return TRK_OK;
    }
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
    bool relocated = false;
    const int status = ref_tracking(T, pb, p, x, lp, xi, elem, relocated);
    if (relocated) pb.meta[p] = pb.meta[p] | META_ISNEW;
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
