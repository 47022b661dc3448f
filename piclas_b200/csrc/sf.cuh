// sf.cuh — shape-function deposition (shape_function, shape_function_cc, shape_function_adaptive with smoothing) on sm_100a.
//
// Reference: DepositionMethod_SF pic_depo_method.f90:851-1003, calcSfSource and helpers in pic_depo_shapefunction_tools.f90.
// The reference scatters every particle (and its periodic images) onto all DOFs within r_sf.  Here the same sum is formed
// as a GATHER so that it is deterministic without floating-point atomics:
//   k_sf_prepare : one thread per particle computes its final source vector Fac(1:4) — for the charge-conserving variants
//                  this is the reference's own per-particle traversal (FIBGM cells -> elements -> DOFs, same order) that
//                  yields totalCharge, hence the same normalisation factor to the last bit;
//   k_sf_gather  : one CTA per target element, one thread per DOF; the particles of every (source element, periodic case)
//                  pair that can reach the element are staged through shared memory and summed in a fixed order.
#pragma once
#include "math.cuh"

struct SFTables {
  const int32_t *FIBGM_nElems, *FIBGM_offsetElem, *FIBGM_Element, *ElemToBGM;
  const double *ElemBary, *ElemRadius, *Elem_xGP, *ElemsJ, *SFElemr2;
  const int32_t* target;    // [nTargets] global element id of every gather target: the local elements, then halo elements
  const int32_t* candOff;   // [nTargets+1]
  const int32_t* candSrc;   // local source element
  const uint8_t* candCase;  // periodic case (1-based), 0 = unshifted position
};

__device__ __forceinline__ double sf_norm(const double v[3]) {
  switch (cst.dim_sf) {
    case 1: return fabs(v[cst.dim_sf_dir - 1]);
    case 2: return sqrt(v[cst.dim_sf_dir1 - 1] * v[cst.dim_sf_dir1 - 1] + v[cst.dim_sf_dir2 - 1] * v[cst.dim_sf_dir2 - 1]);
    default: return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
  }
}
__device__ __forceinline__ double sf_radius2(const double v[3]) {
  switch (cst.dim_sf) {
    case 1: return v[cst.dim_sf_dir - 1] * v[cst.dim_sf_dir - 1];
    case 2: return v[cst.dim_sf_dir1 - 1] * v[cst.dim_sf_dir1 - 1] + v[cst.dim_sf_dir2 - 1] * v[cst.dim_sf_dir2 - 1];
    default: return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
  }
}
__device__ __forceinline__ double sf_pv(int I, int iVec) { return cst.PeriodicVectors[iVec - 1][I - 1]; }

// GetPartPosShifted (:1061-1124); iCase 1-based, iCase == 0 returns the position itself
__device__ __forceinline__ void sf_shifted(int iCase, const double x[3], double out[3]) {
  if (iCase == 0) { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; return; }
  const int* cm = cst.sfCase[iCase - 1];
  if (cst.dim_sf == 1) {
    out[0] = out[1] = out[2] = 0.;
    const int d = cst.dim_sf_dir;
    out[d - 1] = x[d - 1] + cm[0] * sf_pv(d, d);
  } else if (cst.dim_sf == 2) {
    const int d = cst.dim_sf_dir, d1 = cst.dim_sf_dir1, d2 = cst.dim_sf_dir2;
    out[d - 1] = x[d - 1];
    out[d1 - 1] = x[d1 - 1] + cm[0] * sf_pv(d1, cst.dim_periodic_vec1);
    out[d2 - 1] = x[d2 - 1] + cm[0] * sf_pv(d2, cst.dim_periodic_vec1);
    if (cst.dim_periodic_vec2 > 0) {
      out[d1 - 1] = out[d1 - 1] + cm[1] * sf_pv(d1, cst.dim_periodic_vec2);
      out[d2 - 1] = out[d2 - 1] + cm[1] * sf_pv(d2, cst.dim_periodic_vec2);
    }
  } else {
#pragma unroll
    for (int I = 1; I <= 3; ++I) out[I - 1] = ((x[I - 1] + cm[0] * sf_pv(I, 1)) + cm[1] * sf_pv(I, 2)) + cm[2] * sf_pv(I, 3);
  }
}

__device__ __forceinline__ double sf_kernel(double S) {
  double S1 = S * S;
  for (int e = 3; e <= cst.alpha_sf; ++e) S1 = S * S1;
  return S1;
}

// calcTotalChargePeriodic_cc / the totalCharge part of depoChargeOnDOFsSFChargeCon for one (shifted) position: returns the sum
// of wGP^3 * Fac * S1 / sJ over all DOFs within the radius, visiting FIBGM cells, elements and DOFs in the reference's order.
// The ChargeSFDone bookkeeping is replaced by the equivalent rule "an element is treated in the first cell of the loop order
// that contains it", evaluated from its ElemToBGM box.
__device__ double sf_total_charge(const SFTables& T, const double P[3], double Fac, double r_sf, double r2_sf, double r2_sf_inv,
                                  double total) {
  int kmax = (int)ceil((P[0] + r_sf - cst.xyzminglob[0]) / cst.FIBGMdeltas[0]);
  int kmin = (int)floor((P[0] - r_sf - cst.xyzminglob[0]) / cst.FIBGMdeltas[0] + 1);
  int lmax = (int)ceil((P[1] + r_sf - cst.xyzminglob[1]) / cst.FIBGMdeltas[1]);
  int lmin = (int)floor((P[1] - r_sf - cst.xyzminglob[1]) / cst.FIBGMdeltas[1] + 1);
  int mmax = (int)ceil((P[2] + r_sf - cst.xyzminglob[2]) / cst.FIBGMdeltas[2]);
  int mmin = (int)floor((P[2] - r_sf - cst.xyzminglob[2]) / cst.FIBGMdeltas[2] + 1);
  if (cst.dim_sf == 2) {
    if (cst.dim_sf_dir == 1) { kmax = cst.FIBGMmax[0]; kmin = cst.FIBGMmin[0]; }
    else if (cst.dim_sf_dir == 2) { lmax = cst.FIBGMmax[1]; lmin = cst.FIBGMmin[1]; }
    else { mmax = cst.FIBGMmax[2]; mmin = cst.FIBGMmin[2]; }
  } else if (cst.dim_sf == 1) {
    if (cst.dim_sf_dir == 1) { lmax = cst.FIBGMmax[1]; lmin = cst.FIBGMmin[1]; mmax = cst.FIBGMmax[2]; mmin = cst.FIBGMmin[2]; }
    else if (cst.dim_sf_dir == 2) { kmax = cst.FIBGMmax[0]; kmin = cst.FIBGMmin[0]; mmax = cst.FIBGMmax[2]; mmin = cst.FIBGMmin[2]; }
    else { kmax = cst.FIBGMmax[0]; kmin = cst.FIBGMmin[0]; lmax = cst.FIBGMmax[1]; lmin = cst.FIBGMmin[1]; }
  }
  kmax = min(kmax, cst.FIBGMmax[0]); kmin = max(kmin, cst.FIBGMmin[0]);
  lmax = min(lmax, cst.FIBGMmax[1]); lmin = max(lmin, cst.FIBGMmin[1]);
  mmax = min(mmax, cst.FIBGMmax[2]); mmin = max(mmin, cst.FIBGMmin[2]);
  const int ni = cst.FIBGMmax[0] - cst.FIBGMmin[0] + 1, nj = cst.FIBGMmax[1] - cst.FIBGMmin[1] + 1;
  const int NP = cst.N + 1;
  for (int kk = kmin; kk <= kmax; ++kk)
    for (int ll = lmin; ll <= lmax; ++ll)
      for (int mm = mmin; mm <= mmax; ++mm) {
        const size_t c = (size_t)(kk - cst.FIBGMmin[0]) + (size_t)ni * ((size_t)(ll - cst.FIBGMmin[1]) + (size_t)nj * (size_t)(mm - cst.FIBGMmin[2]));
        const int cnt = T.FIBGM_nElems[c], off = T.FIBGM_offsetElem[c];
        for (int ppp = 0; ppp < cnt; ++ppp) {
          const int g = T.FIBGM_Element[off + ppp];
          const int32_t* bx = T.ElemToBGM + (size_t)(g - 1) * 6;
          if (kk != max(kmin, bx[0]) || ll != max(lmin, bx[2]) || mm != max(mmin, bx[4])) continue;  // ChargeSFDone
          const double* b = T.ElemBary + (size_t)(g - 1) * 3;
          const double dv[3] = {P[0] - b[0], P[1] - b[1], P[2] - b[2]};
          if (sf_norm(dv) > (r_sf + T.ElemRadius[g - 1])) continue;
          const double* xgp = T.Elem_xGP + (size_t)(g - 1) * NP * NP * NP * 3;
          const double* sJ = T.ElemsJ + (size_t)(g - 1) * NP * NP * NP;
          for (int m = 0; m < NP; ++m)
            for (int l = 0; l < NP; ++l)
              for (int k = 0; k < NP; ++k) {
                const int r = (m * NP + l) * NP + k;
                const double dd[3] = {P[0] - xgp[r * 3], P[1] - xgp[r * 3 + 1], P[2] - xgp[r * 3 + 2]};
                const double radius2 = sf_radius2(dd);
                if (radius2 <= r2_sf) {
                  const double S1 = sf_kernel(1. - r2_sf_inv * radius2);
                  total = total + (((cst.wGP[k] * cst.wGP[l]) * cst.wGP[m]) * Fac) * S1 / sJ[r];
                }
              }
        }
      }
  return total;
}

// final source vector of every particle: Fac(1:4) as handed to depoChargeOnDOFsSF / scaled by alpha (calcSfSource :30-164)
__global__ void k_sf_prepare(PartBuf pb, int64_t n, SFTables T, double* __restrict__ f0, double* __restrict__ f1,
                             double* __restrict__ f2, double* __restrict__ f3, int* __restrict__ err) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int spec = pb.meta[p] & META_SPEC_MASK;
  const double q = cst.ChargeIC[spec];
  double Fac[4] = {0., 0., 0., 0.};
  if (fabs(q) > 0.0) {  // isDepositParticle
    const double Charge = q * cst.MPF[spec];
    Fac[0] = pb.v[0][p] * Charge; Fac[1] = pb.v[1][p] * Charge; Fac[2] = pb.v[2][p] * Charge; Fac[3] = Charge;
    const double x[3] = {pb.x[0][p], pb.x[1][p], pb.x[2][p]};
    double r_sf = cst.r_sf, r2_sf = cst.r2_sf, r2_sf_inv = cst.r2_sf_inv;
    if (cst.DepositionType == PGPU_DEPO_SF_ADAPTIVE) {
      const int g = pb.elem[p];
      r_sf = T.SFElemr2[(size_t)(g - 1) * 2];
      r2_sf = T.SFElemr2[(size_t)(g - 1) * 2 + 1];
      r2_sf_inv = 1. / r2_sf;
    }
    if (cst.DepositionType == PGPU_DEPO_SF) {
#pragma unroll
      for (int c = 0; c < 4; ++c) Fac[c] = Fac[c] * cst.w_sf;
    } else if (cst.nSFCases > 1) {
      double total = 0.;
      for (int iCase = 1; iCase <= cst.nSFCases; ++iCase) {
        double xs[3];
        sf_shifted(iCase, x, xs);
        total = sf_total_charge(T, xs, Fac[3], r_sf, r2_sf, r2_sf_inv, total);
      }
      if (!cst.sfDepo3D) total = total / cst.dimFactorSF;
      const double f4 = Fac[3];
#pragma unroll
      for (int c = 0; c < 4; ++c) Fac[c] = Fac[c] * f4 / total;
      if (!(total != 0.)) atomicMax(err, 1);
    } else {
      // depoChargeOnDOFsSFChargeCon: alpha = Fac(4)/totalCharge, deposited value alpha*(Fac*S1)
      const double total = sf_total_charge(T, x, Fac[3], r_sf, r2_sf, r2_sf_inv, 0.);
      if (total != 0.) {
        const double alpha = Fac[3] / total;
#pragma unroll
        for (int c = 0; c < 4; ++c) Fac[c] = alpha * Fac[c];
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) Fac[c] = 0.;  // nUsedElems == 0: nothing is deposited
      }
    }
  }
  f0[p] = Fac[0]; f1[p] = Fac[1]; f2[p] = Fac[2]; f3[p] = Fac[3];
}

constexpr int SF_CHUNK = 128;

// halo sum of the DOF contributions received from the other ranks (pic_depo_method.f90:983-996), rank after rank
__global__ void k_sf_add_halo(double* __restrict__ PartSource, const double* __restrict__ recv, const int32_t* __restrict__ recvElem,
                              const int32_t* __restrict__ recvOff, const int32_t* __restrict__ recvIdx, int nd4) {
  // one CTA per receiving element; the blocks of the sending ranks are added in rank order (recvIdx is sorted)
  const int e = recvElem[blockIdx.x];
  for (int t = threadIdx.x; t < nd4; t += blockDim.x) {
    double a = PartSource[(size_t)e * nd4 + t];
    for (int i = recvOff[blockIdx.x]; i < recvOff[blockIdx.x + 1]; ++i) a = a + recv[(size_t)recvIdx[i] * nd4 + t];
    PartSource[(size_t)e * nd4 + t] = a;
  }
}

// PartSource of one target element per CTA; thread t < ND owns DOF t (k fastest)
// targets 0..nElems-1 are the local elements; further targets are elements of other ranks reached by local particles
// (the reference's SendBuffer for ShapeMapping, pic_depo_shapefunction_tools.f90:948-964)
template <bool FUSED>
__global__ void k_sf_gather(PartBuf pb, const int64_t* __restrict__ elemOff, int nElems, int offsetElem, SFTables T,
                            const double* __restrict__ f0, const double* __restrict__ f1, const double* __restrict__ f2,
                            const double* __restrict__ f3, double* __restrict__ PartSource) {
  __shared__ double sx[SF_CHUNK][3];
  __shared__ double sf4[SF_CHUNK][4];
  __shared__ double sBox[2][3];            // bounding box of the target's Gauss points
  __shared__ uint32_t sPass[SF_CHUNK / 32];  // staged particles within reach of that box (the others cannot touch any DOF)
  const int NP = cst.N + 1, ND = NP * NP * NP;
  const int t = threadIdx.x;
  const int alpha = cst.alpha_sf;
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int g = T.target[e];
    double xd[3] = {0., 0., 0.};
    if (t < ND) {
      const double* xg = T.Elem_xGP + ((size_t)(g - 1) * ND + t) * 3;
      xd[0] = xg[0]; xd[1] = xg[1]; xd[2] = xg[2];
    }
    __syncthreads();
    if (t < 3) {
      const double* xg = T.Elem_xGP + (size_t)(g - 1) * ND * 3 + t;
      double lo = xg[0], hi = xg[0];
      for (int n = 1; n < ND; ++n) { lo = fmin(lo, xg[3 * n]); hi = fmax(hi, xg[3 * n]); }
      sBox[0][t] = lo;
      sBox[1][t] = hi;
    }
    __syncthreads();
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
    for (int ci = T.candOff[e]; ci < T.candOff[e + 1]; ++ci) {
      const int s = T.candSrc[ci];
      const int iCase = T.candCase[ci];
      double r2_sf = cst.r2_sf, r2_sf_inv = cst.r2_sf_inv;
      if (cst.DepositionType == PGPU_DEPO_SF_ADAPTIVE) {
        r2_sf = T.SFElemr2[(size_t)(offsetElem + s) * 2 + 1];
        r2_sf_inv = 1. / r2_sf;
      }
      const int64_t p0 = elemOff[s], p1 = elemOff[s + 1];
      for (int64_t c0 = p0; c0 < p1; c0 += SF_CHUNK) {
        const int m = (int)min((int64_t)SF_CHUNK, p1 - c0);
        __syncthreads();
        for (int i0 = 0; i0 < SF_CHUNK; i0 += blockDim.x) {   // every warp walks whole 32-particle groups: one mask word each
          const int i = i0 + t;
          bool pass = false;
          if (i < m) {
            const double x[3] = {pb.x[0][c0 + i], pb.x[1][c0 + i], pb.x[2][c0 + i]};
            double xs[3];
            sf_shifted(iCase, x, xs);
            sx[i][0] = xs[0]; sx[i][1] = xs[1]; sx[i][2] = xs[2];
            sf4[i][0] = f0[c0 + i]; sf4[i][1] = f1[c0 + i]; sf4[i][2] = f2[c0 + i]; sf4[i][3] = f3[c0 + i];
            // distance from the particle to the box of the DOFs (in the directions the shape function spans)
            double gap[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) gap[d] = fmax(fmax(sBox[0][d] - xs[d], xs[d] - sBox[1][d]), 0.);
            pass = sf_radius2(gap) <= r2_sf * (1. + 1e-12);
          }
          const unsigned bal = __ballot_sync(0xffffffffu, pass);
          if ((t & 31) == 0 && i0 + t < SF_CHUNK) sPass[(i0 + t) >> 5] = bal;
        }
        __syncthreads();
        if (FUSED) {
          // restructured arithmetic, 3-D shape function: fused multiply-adds, exponent unrolled, no per-pair switches
          // (same polynomial, same particle order; differences O(1e-16) of the element's source)
          if (t < ND) {
            for (int w = 0; w < SF_CHUNK / 32; ++w) {
              uint32_t bits = sPass[w];
              while (bits) {
                const int i = w * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const double d0 = sx[i][0] - xd[0], d1 = sx[i][1] - xd[1], d2 = sx[i][2] - xd[2];
                const double radius2 = fma(d2, d2, fma(d1, d1, d0 * d0));
                if (radius2 <= r2_sf) {
                  const double S = fma(-r2_sf_inv, radius2, 1.);
                  double S1 = S * S;
                  for (int ex = 3; ex <= alpha; ++ex) S1 = S * S1;
                  a0 = fma(S1, sf4[i][0], a0);
                  a1 = fma(S1, sf4[i][1], a1);
                  a2 = fma(S1, sf4[i][2], a2);
                  a3 = fma(S1, sf4[i][3], a3);
                }
              }
            }
          }
        } else if (t < ND) {
          for (int w = 0; w < SF_CHUNK / 32; ++w) {
            uint32_t bits = sPass[w];
            while (bits) {   // ascending particle order: the sums keep the order of the unfiltered loop
              const int i = w * 32 + __ffs(bits) - 1;
              bits &= bits - 1;
              const double dd[3] = {sx[i][0] - xd[0], sx[i][1] - xd[1], sx[i][2] - xd[2]};
              const double radius2 = sf_radius2(dd);
              if (radius2 <= r2_sf) {
                const double S1 = sf_kernel(1. - r2_sf_inv * radius2);
                a0 = a0 + S1 * sf4[i][0];
                a1 = a1 + S1 * sf4[i][1];
                a2 = a2 + S1 * sf4[i][2];
                a3 = a3 + S1 * sf4[i][3];
              }
            }
          }
        }
      }
    }
    if (t < ND) {
      double* ps = PartSource + ((size_t)e * ND + t) * 4;
      ps[0] = a0; ps[1] = a1; ps[2] = a2; ps[3] = a3;
    }
  }
}
