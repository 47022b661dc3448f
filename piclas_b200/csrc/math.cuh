// math.cuh — per-particle device functions of the particle step, sm_100a.
//
// Arithmetic contract (DESIGN.md "Arithmetic"): the translation unit is compiled with --fmad=false, so the
// compiler never fuses; every expression below is written in the reference's evaluation order, and the only
// fused operations are the explicit fma() calls in the three tensor-product accumulations (field evaluation,
// Newton residual, Newton Jacobian) — the places where gfortran -O3 -march=native contracts in the reference
// build (cmake/SetCompiler.cmake:138,149).  Element ownership is decided exclusively by unfused arithmetic.
#pragma once
#include "common.cuh"

__constant__ ConstTables cst;

#define EPSMACH 2.220446049250313e-16   /* EPSILON(0.) with REAL == 64 bit, globals_vars.f90:50 */
#define HUGE_D 1.7976931348623157e308

// ---- basis.f90:1011-1035 ALMOSTEQUAL_UNITY, :1223-1264 LagrangeInterpolationPolys ----------------------------------
__device__ __forceinline__ bool almost_equal_unity(double x, double y) {
  const double d = fabs(x - y);
  if (x == 0. || y == 0.) return d <= 2. * EPSMACH;
  return (d <= EPSMACH * fabs(x)) && (d <= EPSMACH * fabs(y));
}

template <int NP>  // NP = N_in + 1 nodes
__device__ __forceinline__ void lagrange_polys(double x, const double* __restrict__ xGP, const double* __restrict__ wBary,
                                               double* L) {
  bool hit = false;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const bool e = almost_equal_unity(x, xGP[i]);
    L[i] = e ? 1. : 0.;
    hit |= e;
  }
  if (hit) return;
  double s = 0.;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    L[i] = wBary[i] / (x - xGP[i]);
    s = s + L[i];
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) L[i] = L[i] / s;
}

// ---- eval_xyz.f90:448-495 getDet / getInv ----------------------------------------------------------------------------
__device__ __forceinline__ double get_det(const double M[3][3]) {
  return (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * M[2][2] + (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * M[2][0] +
         (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * M[2][1];
}
__device__ __forceinline__ void get_inv(const double M[3][3], double sdet, double R[3][3]) {
  R[0][0] = (M[1][1] * M[2][2] - M[1][2] * M[2][1]) * sdet;
  R[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * sdet;
  R[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * sdet;
  R[1][0] = (M[1][2] * M[2][0] - M[1][0] * M[2][2]) * sdet;
  R[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * sdet;
  R[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * sdet;
  R[2][0] = (M[1][0] * M[2][1] - M[1][1] * M[2][0]) * sdet;
  R[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * sdet;
  R[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * sdet;
}

// ---- eval_xyz.f90:498-612 GetRefNewtonStartValue (guesses 1, 3, 4; guess 2 needs Elem_xGP and is mapped by the host
// layer to an error at init) -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void newton_start_value(const GeoElem* __restrict__ g, const double x[3], double xi[3]) {
  const int guess = cst.RefMappingGuess;
  if (guess == 1) {
    const double epsOne = 1.0 + cst.RefMappingEps;
    const double P0 = x[0] - g->bary[0], P1 = x[1] - g->bary[1], P2 = x[2] - g->bary[2];
    double lin[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) lin[d] = ((P0 * g->xez[d][0] + P1 * g->xez[d][1]) + P2 * g->xez[d][2]) * g->slen[d];
#pragma unroll
    for (int d = 0; d < 3; ++d) xi[d] = 0.5 * (lin[d] - lin[d + 3]);
    const double mx = fmax(fabs(xi[0]), fmax(fabs(xi[1]), fabs(xi[2])));
    if (mx > epsOne) {
#pragma unroll
      for (int d = 0; d < 3; ++d) xi[d] = fmax(fmin(1.0, xi[d]), -1.0);
    }
  } else if (guess == 3) {
    double d0 = x[0] - g->XCL[0][0], d1 = x[1] - g->XCL[0][1], d2 = x[2] - g->XCL[0][2];
    double win = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    xi[0] = xi[1] = xi[2] = cst.XiCL[0];
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) {
          const double* p = g->XCL[i + 2 * j + 4 * k];
          const double dX = fabs(x[0] - p[0]);
          if (dX > win) continue;
          const double dY = fabs(x[1] - p[1]);
          if (dY > win) continue;
          const double dZ = fabs(x[2] - p[2]);
          if (dZ > win) continue;
          const double dist = sqrt((dX * dX + dY * dY) + dZ * dZ);
          if (dist < win) {
            win = dist;
            xi[0] = cst.XiCL[i];
            xi[1] = cst.XiCL[j];
            xi[2] = cst.XiCL[k];
          }
        }
  } else {
    xi[0] = xi[1] = xi[2] = 0.;
  }
}

// ---- eval_xyz.f90:298-445 RefElemNewton at NGeo == 1 ----------------------------------------------------------------------
// returns: bit0 = isSuccessful (always 1 when !hasSuccess), bit1 = abort requested (Mode 1 without isSuccessful)
__device__ __forceinline__ int ref_elem_newton(const GeoElem* __restrict__ g, const double x_in[3], double xi[3], int mode,
                                               bool hasSuccess) {
  double Lag[3][2];
  double F[3];
  int result = 1;
  lagrange_polys<2>(xi[0], cst.XiCL, cst.wBaryCL, Lag[0]);
  lagrange_polys<2>(xi[1], cst.XiCL, cst.wBaryCL, Lag[1]);
  lagrange_polys<2>(xi[2], cst.XiCL, cst.wBaryCL, Lag[2]);
  F[0] = -x_in[0];
  F[1] = -x_in[1];
  F[2] = -x_in[2];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double buff = Lag[1][j] * Lag[2][k];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double* p = g->XCL[i + 2 * j + 4 * k];
#pragma unroll
        for (int d = 0; d < 3; ++d) F[d] = fma(p[d] * Lag[0][i], buff, F[d]);  // F+XCL*Lag(1,i)*buff, :343-349
      }
    }
  double deltaXi2 = (fabs(F[0]) < EPSMACH && fabs(F[1]) < EPSMACH && fabs(F[2]) < EPSMACH) ? 0. : 1.;
  double Norm_F = (F[0] * F[0] + F[1] * F[1]) + F[2] * F[2];
  double Norm_F_old;
  int it = 0;
  const double eps = cst.RefMappingEps;
  while (deltaXi2 > eps && it < 100) {
    ++it;
    double Jac[3][3], sJac[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Jac[r][c] = 0.;
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double buff = Lag[1][j] * Lag[2][k];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double buff2 = Lag[0][i] * buff;
          const double(*dx)[3] = g->dXCL[i + 2 * j + 4 * k];  // [nn][dd]
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Jac[r][c] = fma(dx[r][c], buff2, Jac[r][c]);  // Jac(r,c)+=dXCL(c,r,..)*buff2
        }
      }
    double sdet = get_det(Jac);
    if (sdet > 0.) sdet = 1. / sdet;
    get_inv(Jac, sdet, sJac);
    double dXi[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) dXi[r] = (sJac[r][0] * F[0] + sJac[r][1] * F[1]) + sJac[r][2] * F[2];
    deltaXi2 = (dXi[0] * dXi[0] + dXi[1] * dXi[1]) + dXi[2] * dXi[2];
    const double xo0 = xi[0], xo1 = xi[1], xo2 = xi[2];
    Norm_F_old = Norm_F;
    Norm_F = Norm_F * 2.;
    double lambda = 1.0;
    int iArmijo = 1;
    while (Norm_F > Norm_F_old * (1. - 0.0001 * lambda) && iArmijo <= 8) {
      xi[0] = xo0 - lambda * dXi[0];
      xi[1] = xo1 - lambda * dXi[1];
      xi[2] = xo2 - lambda * dXi[2];
      lagrange_polys<2>(xi[0], cst.XiCL, cst.wBaryCL, Lag[0]);
      lagrange_polys<2>(xi[1], cst.XiCL, cst.wBaryCL, Lag[1]);
      lagrange_polys<2>(xi[2], cst.XiCL, cst.wBaryCL, Lag[2]);
      F[0] = -x_in[0];
      F[1] = -x_in[1];
      F[2] = -x_in[2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double buff = Lag[1][j] * Lag[2][k];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double buff2 = Lag[0][i] * buff;
            const double* p = g->XCL[i + 2 * j + 4 * k];
#pragma unroll
            for (int d = 0; d < 3; ++d) F[d] = fma(p[d], buff2, F[d]);  // F+XCL*buff2, :404-413
          }
        }
      lambda = 0.2 * lambda;
      ++iArmijo;
      Norm_F = (F[0] * F[0] + F[1] * F[1]) + F[2] * F[2];
    }
    if (fabs(xi[0]) > 1.5 || fabs(xi[1]) > 1.5 || fabs(xi[2]) > 1.5) {
      if (hasSuccess) {
        result = 0;
        break;
      } else if (mode == 1) {
        result = 1 | 2;
        break;
      } else
        break;
    }
  }
  return result;
}

// eval_xyz.f90:35-123 GetPositionInRefElem without DoReUseMap
struct RefPos { double xi0, xi1, xi2; int status; };
__device__ __noinline__ RefPos position_in_ref_elem_worker(const GeoElem* __restrict__ g, double x0, double x1, double x2, bool forceMode,
                                                           bool hasSuccess) {
  const double x[3] = {x0, x1, x2};
  double xi[3];
  newton_start_value(g, x, xi);
  RefPos r;
  r.status = ref_elem_newton(g, x, xi, forceMode ? 1 : 2, hasSuccess);
  r.xi0 = xi[0]; r.xi1 = xi[1]; r.xi2 = xi[2];
  return r;
}
// the caller's x/xi stay in registers: only scalars cross the call
__device__ __forceinline__ int position_in_ref_elem(const GeoElem* __restrict__ g, const double x[3], double xi[3], bool forceMode,
                                                    bool hasSuccess) {
  const RefPos r = position_in_ref_elem_worker(g, x[0], x[1], x[2], forceMode, hasSuccess);
  xi[0] = r.xi0; xi[1] = r.xi1; xi[2] = r.xi2;
  return r.status;
}

// ---- eval_xyz.f90:167-295 EvaluateFieldAtRefPos (E only; PP_nVar == 1) -----------------------------------------------------
// U: the element's field tile [(k*NP+j)*NP+i][3]
template <int NP>
__device__ __forceinline__ void evaluate_field(const double xi[3], const double* __restrict__ U, double out[3]) {
  double L0[NP], L1[NP], L2[NP];
  lagrange_polys<NP>(xi[0], cst.xGP, cst.wBary, L0);
  lagrange_polys<NP>(xi[1], cst.xGP, cst.wBary, L1);
  lagrange_polys<NP>(xi[2], cst.xGP, cst.wBary, L2);
  double o0 = 0., o1 = 0., o2 = 0.;
#pragma unroll
  for (int k = 0; k < NP; ++k)
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double lez = L1[j] * L2[k];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const double* u = U + ((k * NP + j) * NP + i) * 3;
        o0 = fma(u[0] * L0[i], lez, o0);  // U_OUT + U_IN*L_xi(1,i)*L_Eta_Zeta, :207-215
        o1 = fma(u[1] * L0[i], lez, o1);
        o2 = fma(u[2] * L0[i], lez, o2);
      }
    }
  out[0] = o0;
  out[1] = o1;
  out[2] = o2;
}

// pic_interpolation_tools.f90:458-572 GetEMFieldDW: inverse-distance weighting over the element's Gauss points
// (only reached when the Newton mapping failed and the deposition is cell_volweight_mean)
struct Vec3 { double a, b, c; };
template <int NP>
__device__ __noinline__ Vec3 field_inverse_distance_worker(double p0, double p1, double p2, const double* __restrict__ U,
                                                           const double* __restrict__ xgp /* Elem_xGP tile [(k*NP+j)*NP+i][3] */) {
  const double pos[3] = {p0, p1, p2};
  // two passes instead of a temporary array; weights are recomputed bit-identically
  int hk = -1, hl = -1, hm = -1;  // last exact hit (norm == 0)
  double DistSum = 0.0;
  for (int k = 0; k < NP; ++k)
    for (int l = 0; l < NP; ++l)
      for (int m = 0; m < NP; ++m) {
        const double* gp = xgp + ((m * NP + l) * NP + k) * 3;
        const double d0 = gp[0] - pos[0], d1 = gp[1] - pos[1], d2 = gp[2] - pos[2];
        const double norm = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
        if (norm > 0.) {
          DistSum = DistSum + 1. / norm;
        } else {
          hk = k; hl = l; hm = m;
          DistSum = 1.;
          break;  // EXIT leaves the m loop only (:548)
        }
      }
  double o0 = 0., o1 = 0., o2 = 0.;
  for (int k = 0; k < NP; ++k)
    for (int l = 0; l < NP; ++l)
      for (int m = 0; m < NP; ++m) {
        // weight of (k,l,m) as left in PartDistDepo by the first loop
        double w;
        const int lin = (k * NP + l) * NP + m, hlin = (hk * NP + hl) * NP + hm;
        if (hk >= 0 && lin < hlin) w = 0.;                       // zeroed by PartDistDepo(:,:,:) = 0
        else if (hk >= 0 && lin == hlin) w = 1.;
        else if (hk >= 0 && k == hk && l == hl && m > hm) w = 0.;  // skipped by the EXIT, still zero
        else {
          const double* gp = xgp + ((m * NP + l) * NP + k) * 3;
          const double d0 = gp[0] - pos[0], d1 = gp[1] - pos[1], d2 = gp[2] - pos[2];
          const double norm = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
          w = (norm > 0.) ? 1. / norm : 1.;
        }
        const double* u = U + ((m * NP + l) * NP + k) * 3;
        const double ww = w / DistSum;
        o0 = o0 + ww * u[0];
        o1 = o1 + ww * u[1];
        o2 = o2 + ww * u[2];
      }
  Vec3 r;
  r.a = o0; r.b = o1; r.c = o2;
  return r;
}
template <int NP>
__device__ __forceinline__ void field_inverse_distance(const double pos[3], const double* __restrict__ U,
                                                       const double* __restrict__ xgp, double out[3]) {
  const Vec3 r = field_inverse_distance_worker<NP>(pos[0], pos[1], pos[2], U, xgp);
  out[0] = r.a; out[1] = r.b; out[2] = r.c;
}

// ---- particle push: timedisc_TimeStepPoissonByBorisLeapfrog.f90:128-198 (508), timedisc_TimeStepPoisson.f90:111-181 (509)
__device__ __forceinline__ void push_particle(double x[3], double v[3], const double F[6], int spec0, bool& isNew, double dt) {
  const double q = cst.ChargeIC[spec0], mass = cst.MassIC[spec0];
  const bool isPush = fabs(q) > 0.0;
  const double c2_inv = cst.c2_inv;
  if (cst.TimeDiscMethod == PGPU_TIMEDISC_BORIS_LEAPFROG) {
    if (isNew) {
      if (isPush && cst.DoInterpolation) {
        const double qmt = q / mass;  // PartRHS_NR, particle_rhs.f90:206-228
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = v[d] - ((F[d] * qmt) * dt) * 0.5;
      }
      isNew = false;
    }
    if (isPush && cst.DoInterpolation) {
      const double c_1 = (q * dt) / (mass * 2.);
      const double gamma = 1. / sqrt(1 - (((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]) * c2_inv));
      double vm[3], t[3], vp[3], vn[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) vm[d] = v[d] * gamma + c_1 * F[d];
      const double gamma_minus = sqrt(1 + ((vm[0] * vm[0] + vm[1] * vm[1]) + vm[2] * vm[2]) * c2_inv);
      const double Bn = sqrt((F[3] * F[3] + F[4] * F[4]) + F[5] * F[5]);  // VECNORM3D
      if (fabs(Bn) > 0.0) {
        const double invL = 1. / Bn;  // UNITVECTOR
        const double u[3] = {F[3] * invL, F[4] * invL, F[5] * invL};
        const double tn = tan(c_1 / gamma_minus * Bn);
#pragma unroll
        for (int d = 0; d < 3; ++d) t[d] = tn * u[d];
        vp[0] = vm[0] + (vm[1] * t[2] - vm[2] * t[1]);
        vp[1] = vm[1] + (vm[2] * t[0] - vm[0] * t[2]);
        vp[2] = vm[2] + (vm[0] * t[1] - vm[1] * t[0]);
        const double fac = 2.0 / (1. + ((t[0] * t[0] + t[1] * t[1]) + t[2] * t[2]));
        vn[0] = (vm[0] + fac * (vp[1] * t[2] - vp[2] * t[1])) + c_1 * F[0];
        vn[1] = (vm[1] + fac * (vp[2] * t[0] - vp[0] * t[2])) + c_1 * F[1];
        vn[2] = (vm[2] + fac * (vp[0] * t[1] - vp[1] * t[0])) + c_1 * F[2];
      } else {
        // B = 0: UNITVECTOR(0) = 0 and TAN(0) = 0, so t_vec = 0, v_prime = v_plus = v_minus exactly
#pragma unroll
        for (int d = 0; d < 3; ++d) vn[d] = vm[d] + c_1 * F[d];
      }
      const double s = sqrt(1 + ((vn[0] * vn[0] + vn[1] * vn[1]) + vn[2] * vn[2]) * c2_inv);
#pragma unroll
      for (int d = 0; d < 3; ++d) v[d] = vn[d] / s;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = x[d] + v[d] * dt;
  } else {
    double Pt[3] = {0., 0., 0.};
    if (cst.DoInterpolation && isPush) {
      const double qmt = q / mass;
#pragma unroll
      for (int d = 0; d < 3; ++d) Pt[d] = F[d] * qmt;
    }
    if (isNew) {
      if (isPush) {
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = v[d] - (Pt[d] * dt) * 0.5;
      }
      isNew = false;
    }
    if (isPush) {
#pragma unroll
      for (int d = 0; d < 3; ++d) v[d] = v[d] + Pt[d] * dt;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = x[d] + v[d] * dt;
  }
}

// ---- element-record access -------------------------------------------------------------------------------------------------------
// The tria records are read either from a CTA's shared-memory copy (G = false) or, in the leaver walk, straight from global
// memory with every lane on a different element (G = true).  There each 8-byte load is its own L1 wavefront (r1v6 profile: the
// walk ran at 82 % of the LSU wavefront peak), so a corner (x, y, z, pad) is fetched with ONE 256-bit load and the four node
// indices of a side with one 32-bit load.
template <bool G>
__device__ __forceinline__ void load_corner(const TriaElem* __restrict__ te, int n, double c[3]) {
  if (G) {
    double w;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c[0]), "=d"(c[1]), "=d"(c[2]), "=d"(w) : "l"(&te->corner[n][0]));
  } else {
    c[0] = te->corner[n][0]; c[1] = te->corner[n][1]; c[2] = te->corner[n][2];
  }
}
// ElemSideNodeID(1:4,side) packed: byte n = node n
template <bool G>
__device__ __forceinline__ uint32_t load_side_nodes(const TriaElem* __restrict__ te, int s) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(&te->sideNode[s][0]);
  return G ? __ldg(q) : *q;
}

// ---- particle_intersection.f90:167-280 ParticleThroughSideCheck3DFast (regular side) -------------------------------------------
template <bool G = false>
__device__ __forceinline__ bool through_side_check_fast(const TriaElem* __restrict__ te, const double lp[3], const double V[3],
                                                        int s /*0-based local side*/, int tri /*1|2*/) {
  double Ax[3], Ay[3], Az[3];
  const uint32_t sn = load_side_nodes<G>(te, s);
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    const int node = (n == 0) ? 0 : n + tri - 1;
    double c[3];
    load_corner<G>(te, (sn >> (8 * node)) & 0xffu, c);
    Ax[n] = c[0] - lp[0]; Ay[n] = c[1] - lp[1]; Az[n] = c[2] - lp[2];
  }
  const double Vx = V[0], Vy = V[1], Vz = V[2];
  bool through = true;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int o = (r + 2) % 3;  // rows use A(3), A(1), A(2) as third factor
    const double a = (Ay[r] * Vz - Az[r] * Vy) * Ax[o];
    const double b = (Az[r] * Vx - Ax[r] * Vz) * Ay[o];
    const double c = (Ax[r] * Vy - Ay[r] * Vx) * Az[o];
    const double det = (a + b) + c;
    double mn = HUGE_D;
    const double fa = fabs(a), fb = fabs(b), fc = fabs(c);
    if (fa > 0.0 && fa < mn) mn = fa;
    if (fb > 0.0 && fb < mn) mn = fb;
    if (fc > 0.0 && fc < mn) mn = fc;
    const double minComp = -EPSMACH * mn;
    if (!(det >= minComp)) through = false;
  }
  return through;
}

// ---- particle_intersection.f90:430-512 ParticleThroughSideLastPosCheck (regular side) -------------------------------------------
template <bool G = false>
__device__ __forceinline__ bool through_side_lastpos_check(const TriaElem* __restrict__ te, const double lp[3], int s, int tri,
                                                           double& det) {
  double Ax[3], Ay[3], Az[3];
  const uint32_t sn = load_side_nodes<G>(te, s);
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    const int node = (n == 0) ? 0 : n + tri - 1;
    double c[3];
    load_corner<G>(te, (sn >> (8 * node)) & 0xffu, c);
    Ax[n] = c[0] - lp[0]; Ay[n] = c[1] - lp[1]; Az[n] = c[2] - lp[2];
  }
  det = ((Ay[0] * Az[1] - Az[0] * Ay[1]) * Ax[2] + (Az[0] * Ax[1] - Ax[0] * Az[1]) * Ay[2]) +
        (Ax[0] * Ay[1] - Ay[0] * Ax[1]) * Az[2];
  return !((det < 0) || (det != det));
}

// ---- particle_intersection.f90:79-164 IntersectionWithWall: returns TrackInfo%alpha ------------------------------------------------
template <bool G = false>
__device__ __forceinline__ double intersection_with_wall(const TriaElem* __restrict__ te, const double lp[3], const double V[3],
                                                         int s, int tri) {
  const uint32_t sn = load_side_nodes<G>(te, s);
  double n0[3], n1[3], n2[3];
  load_corner<G>(te, sn & 0xffu, n0);
  load_corner<G>(te, (sn >> (8 * tri)) & 0xffu, n1);
  load_corner<G>(te, (sn >> (8 * (tri + 1))) & 0xffu, n2);
  const double xN = n0[0], yN = n0[1], zN = n0[2];
  const double v1x = n1[0] - xN, v1y = n1[1] - yN, v1z = n1[2] - zN;
  const double v2x = n2[0] - xN, v2y = n2[1] - yN, v2z = n2[2] - zN;
  double nx = v1y * v2z - v1z * v2y;
  double ny = v1z * v2x - v1x * v2z;
  double nz = v1x * v2y - v1y * v2x;
  const double nVal = sqrt((nx * nx + ny * ny) + nz * nz);
  nx = nx / nVal; ny = ny / nVal; nz = nz / nVal;
  const double bx = lp[0] - xN, by = lp[1] - yN, bz = lp[2] - zN;
  const double bn = (bx * nx + by * ny) + bz * nz;
  const double ax = bx - nx * bn, ay = by - ny * bn, az = bz - nz * bn;
  const double t0 = ay * bz - az * by, t1 = az * bx - ax * bz, t2 = ax * by - ay * bx;
  double dist = sqrt(((t0 * t0 + t1 * t1) + t2 * t2) / ((ax * ax + ay * ay) + az * az));
  if (dist != dist) dist = sqrt((bx * bx + by * by) + bz * bz);
  double alpha = (V[0] * nx + V[1] * ny) + V[2] * nz;
  if (fabs(alpha) > 0.) alpha = dist / alpha;
  return alpha;
}

// ---- particle_surfaces.f90:256-401 CalcNormAndTangTriangle, nVec only: outward unit normal of triangle tri of local side s --------
template <bool G = false>
__device__ __forceinline__ void triangle_normal(const TriaElem* __restrict__ te, int s, int tri, double n[3]) {
  const uint32_t sn = load_side_nodes<G>(te, s);
  double p0[3], p1[3], p2[3];
  load_corner<G>(te, sn & 0xffu, p0);
  load_corner<G>(te, (sn >> (8 * tri)) & 0xffu, p1);
  load_corner<G>(te, (sn >> (8 * (tri + 1))) & 0xffu, p2);
  const double a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
  const double b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
  double nx = -a1 * b2 + a2 * b1;   // NV (inwards)
  double ny = -a2 * b0 + a0 * b2;
  double nz = -a0 * b1 + a1 * b0;
  const double nVal = sqrt((nx * nx + ny * ny) + nz * nz);
  n[0] = -nx / nVal;
  n[1] = -ny / nVal;
  n[2] = -nz / nVal;
}

// determinants of the two triangles of local side s (particle_mesh_tools.f90:187-199)
template <bool G = false>
__device__ __forceinline__ void side_dets(const TriaElem* __restrict__ te, const double x[3], int s, double& d1, double& d2) {
  double A[4][3];
  const uint32_t sn = load_side_nodes<G>(te, s);
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    double c[3];
    load_corner<G>(te, (sn >> (8 * n)) & 0xffu, c);
    A[n][0] = c[0] - x[0];
    A[n][1] = c[1] - x[1];
    A[n][2] = c[2] - x[2];
  }
  const double c0 = A[0][1] * A[2][2] - A[0][2] * A[2][1];
  const double c1 = A[0][2] * A[2][0] - A[0][0] * A[2][2];
  const double c2 = A[0][0] * A[2][1] - A[0][1] * A[2][0];
  d1 = (c0 * A[1][0] + c1 * A[1][1]) + c2 * A[1][2];
  d1 = -d1;
  d2 = (c0 * A[3][0] + c1 * A[3][1]) + c2 * A[3][2];
}

// ParticleInsideQuad3D (rolled over the six sides).  Returns InElementCheck; mask bit 2*s+t-1 is set when the
// determinant of triangle t of local side s+1 is <= 0 (the triangles SingleParticleTriaTracking3D then examines).
template <bool G = false>
__device__ __forceinline__ bool inside_quad3d_mask(const TriaElem* __restrict__ te, const double x[3], uint32_t& mask) {
  bool inElem = true;
  const unsigned conc = te->concave;
  uint32_t m = 0;
#pragma unroll 1
  for (int s = 0; s < 6; ++s) {
    double d1, d2;
    side_dets<G>(te, x, s, d1, d2);
    const bool neg = (d1 < 0) || (d2 < 0);
    const bool pos = !(d1 < 0) || !(d2 < 0);
    if ((conc >> s) & 1u) {
      if (!pos) inElem = false;
    } else {
      if (neg) inElem = false;
    }
    if (d1 <= 0.0) m |= 1u << (2 * s);
    if (d2 <= 0.0) m |= 2u << (2 * s);
  }
  mask = m;
  return inElem;
}

enum { TRK_OK = 0, TRK_LOST = 1, TRK_REMOVED = 2, TRK_ERR_BC = 3, TRK_ERR_ELEM = 4, TRK_ERR_LOOP = 5 };

