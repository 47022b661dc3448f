// fastmath.cuh — restructured arithmetic of the particle step (params.arithmetic == 1).
//
// Same algorithm and the same decisions as math.cuh, with the floating-point work re-associated where the north star's
// tolerance allows it (positions, velocities, deposited charge <= 1e-12 relative; element ownership unchanged):
//   * Lagrange basis in product form (no divisions, exact at the nodes without the node-hit branch),
//   * sum-factorised field evaluation (252 instead of 448 FP64 instructions at N = 3),
//   * sign of the ParticleInsideQuad3D determinants from precomputed triangle planes, falling back to the determinant
//     itself whenever the point lies within PlaneElem::tol of a plane (so ownership is decided by the same sign),
//   * closed-form reference coordinates on affine elements (the Newton iteration of the reference converges to the same
//     point; RefElemNewton is still used on every non-affine element),
//   * rsqrt instead of 1/sqrt in the B = 0 Boris push.
#pragma once
#include "math.cuh"

template <int NP>
__device__ __forceinline__ void lagrange_fast(double x, double* L) {
  double d[NP], suf[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) d[i] = x - cst.xGP[i];
  suf[NP - 1] = 1.;
#pragma unroll
  for (int i = NP - 2; i >= 0; --i) suf[i] = suf[i + 1] * d[i + 1];
  double pre = 1.;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    L[i] = (cst.wBary[i] * pre) * suf[i];
    pre = pre * d[i];
  }
}

// field tile sE[((k*NP + j)*3 + c)*NP + i]
template <int NP>
__device__ __forceinline__ void evaluate_field_fast(const double xi[3], const double* __restrict__ sE, double out[3]) {
  double L0[NP], L1[NP], L2[NP];
  lagrange_fast<NP>(xi[0], L0);
  lagrange_fast<NP>(xi[1], L1);
  lagrange_fast<NP>(xi[2], L2);
  double o0 = 0., o1 = 0., o2 = 0.;
#pragma unroll 1
  for (int k = 0; k < NP; ++k) {
    double lz = L2[0];
#pragma unroll
    for (int q = 1; q < NP; ++q) lz = (k == q) ? L2[q] : lz;
    double s0 = 0., s1 = 0., s2 = 0.;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double* row = sE + ((k * NP + j) * 3) * NP;
      double t0 = row[0] * L0[0], t1 = row[NP] * L0[0], t2 = row[2 * NP] * L0[0];
#pragma unroll
      for (int i = 1; i < NP; ++i) {
        t0 = fma(row[i], L0[i], t0);
        t1 = fma(row[NP + i], L0[i], t1);
        t2 = fma(row[2 * NP + i], L0[i], t2);
      }
      s0 = fma(t0, L1[j], s0);
      s1 = fma(t1, L1[j], s1);
      s2 = fma(t2, L1[j], s2);
    }
    o0 = fma(s0, lz, o0);
    o1 = fma(s1, lz, o1);
    o2 = fma(s2, lz, o2);
  }
  out[0] = o0;
  out[1] = o1;
  out[2] = o2;
}

// cold paths kept out of line so that their registers do not count against the hot loop
struct PushOut { double x0, x1, x2, v0, v1, v2; int isNew; };
__device__ __noinline__ PushOut push_particle_cold(double x0, double x1, double x2, double v0, double v1, double v2, double F0, double F1,
                                                   double F2, double F3, double F4, double F5, int spec0, int isNewIn, double dt) {
  double x[3] = {x0, x1, x2}, v[3] = {v0, v1, v2};
  const double F[6] = {F0, F1, F2, F3, F4, F5};
  bool isNew = isNewIn != 0;
  push_particle(x, v, F, spec0, isNew, dt);
  PushOut o;
  o.x0 = x[0]; o.x1 = x[1]; o.x2 = x[2]; o.v0 = v[0]; o.v1 = v[1]; o.v2 = v[2]; o.isNew = isNew ? 1 : 0;
  return o;
}
template <bool G>
__device__ __noinline__ uint32_t inside_exact_cold(const TriaElem* __restrict__ te, double x0, double x1, double x2) {
  const double x[3] = {x0, x1, x2};
  uint32_t mask;
  const bool in = inside_quad3d_mask<G>(te, x, mask);
  return mask | (in ? 0x80000000u : 0u);
}

// Boris-Leapfrog / Leapfrog push with fused multiply-adds; the B != 0 rotation is delegated to the reference-order code
__device__ __forceinline__ void push_particle_fast(double x[3], double v[3], const double F[6], int spec0, bool& isNew, double dt) {
  const double Bn2 = (F[3] * F[3] + F[4] * F[4]) + F[5] * F[5];
  if (cst.TimeDiscMethod != PGPU_TIMEDISC_BORIS_LEAPFROG || Bn2 > 0.0) {
    const PushOut o = push_particle_cold(x[0], x[1], x[2], v[0], v[1], v[2], F[0], F[1], F[2], F[3], F[4], F[5], spec0, isNew ? 1 : 0, dt);
    x[0] = o.x0; x[1] = o.x1; x[2] = o.x2; v[0] = o.v0; v[1] = o.v1; v[2] = o.v2;
    isNew = o.isNew != 0;
    return;
  }
  const double q = cst.ChargeIC[spec0], mass = cst.MassIC[spec0];
  const bool isPush = fabs(q) > 0.0;
  if (isNew) {
    if (isPush && cst.DoInterpolation) {
      const double h = -0.5 * dt * (q / mass);
#pragma unroll
      for (int d = 0; d < 3; ++d) v[d] = fma(F[d], h, v[d]);
    }
    isNew = false;
  }
  if (isPush && cst.DoInterpolation) {
    const double c_1 = (q * dt) / (mass * 2.);
    const double c2_inv = cst.c2_inv;
    const double gamma = rsqrt(fma(-fma(v[0], v[0], fma(v[1], v[1], v[2] * v[2])), c2_inv, 1.0));
    double vn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) vn[d] = fma(2. * c_1, F[d], v[d] * gamma);  // v_minus + c_1 E (t_vec = 0) + c_1 E
    const double s = rsqrt(fma(fma(vn[0], vn[0], fma(vn[1], vn[1], vn[2] * vn[2])), c2_inv, 1.0));
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] = vn[d] * s;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) x[d] = fma(v[d], dt, x[d]);
}

// ParticleInsideQuad3D through the triangle planes.  Exact fallback (determinants) within tol of any plane.
template <bool G = false>
__device__ __forceinline__ bool inside_fast(const PlaneElem* __restrict__ pl, const TriaElem* __restrict__ te, const double x[3],
                                            uint32_t& mask) {
  uint32_t neg = 0;
  bool ambiguous = false;
  const double tol = G ? __ldg(&pl->tol) : pl->tol;
  const uint32_t c2 = G ? __ldg(&pl->concave2) : pl->concave2;
  const uint32_t planar = G ? __ldg(&pl->planar) : pl->planar;
  if (planar) {
    // both triangles of a side lie in one plane (to 1e-12, far inside tol): one evaluation per side decides both signs
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      double n0, n1, n2, d;
      if (G) asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(n0), "=d"(n1), "=d"(n2), "=d"(d) : "l"(&pl->pl[2 * s][0]));
      else { n0 = pl->pl[2 * s][0]; n1 = pl->pl[2 * s][1]; n2 = pl->pl[2 * s][2]; d = pl->pl[2 * s][3]; }
      const double dist = fma(n0, x[0], fma(n1, x[1], fma(n2, x[2], -d)));
      ambiguous |= fabs(dist) <= tol;
      if (dist < 0.) neg |= 3u << (2 * s);
    }
  } else {
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      double n0, n1, n2, d;
      if (G) asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(n0), "=d"(n1), "=d"(n2), "=d"(d) : "l"(&pl->pl[t][0]));
      else { n0 = pl->pl[t][0]; n1 = pl->pl[t][1]; n2 = pl->pl[t][2]; d = pl->pl[t][3]; }
      const double dist = fma(n0, x[0], fma(n1, x[1], fma(n2, x[2], -d)));
      ambiguous |= fabs(dist) <= tol;
      if (dist < 0.) neg |= 1u << t;
    }
  }
  if (ambiguous) {
    const uint32_t r = inside_exact_cold<G>(te, x[0], x[1], x[2]);
    mask = r & 0x7fffffffu;
    return (r >> 31) != 0;
  }
  mask = neg;
  const uint32_t lo = 0x555u;
  const uint32_t any = (neg | (neg >> 1)) & lo, both = (neg & (neg >> 1)) & lo;
  return ((any & ~c2) | (both & c2)) == 0;
}

// reference coordinates: closed form on affine elements, RefElemNewton otherwise.  Returns SucRefPos.
__device__ __forceinline__ bool ref_position_fast(const AffElem* __restrict__ af, const GeoElem* __restrict__ g, const double x[3],
                                                  double xi[3], bool forceMode) {
  if (af->affine != 0.0) {
    const double r0 = x[0] - af->x0[0], r1 = x[1] - af->x0[1], r2 = x[2] - af->x0[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) xi[d] = fma(af->A[d][0], r0, fma(af->A[d][1], r1, af->A[d][2] * r2)) - 1.0;
    if (fabs(xi[0]) <= 1.5 && fabs(xi[1]) <= 1.5 && fabs(xi[2]) <= 1.5) return true;
  }
  return (position_in_ref_elem(g, x, xi, forceMode, true) & 1) != 0;
}
