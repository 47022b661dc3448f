// device_math_host.cpp — the device math headers of the CUDA path (piclas_b200/csrc/math.cuh, fastmath.cuh) compiled for the
// HOST with a handful of macro shims, so that the restructured arithmetic (params.arithmetic = 1) can be compared with the
// reference-order arithmetic (arithmetic = 0) of the same headers on many random inputs where no GPU exists.  Test
// infrastructure (tests/test_device_math_on_host.py); the -m gpu parity tests remain the check of the kernels themselves.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <random>
#undef __device__
#undef __host__
#undef __global__
#undef __constant__
#undef __forceinline__
#undef __noinline__
#undef __align__
#undef __launch_bounds__
#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __noinline__
#define __align__(n) alignas(n)
#define __launch_bounds__(...)
#define asm(...)   /* inline PTX sits behind `if (G)` template flags that are false here */
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
#include "fastmath.cuh"

static double rel(double a, double b, double scale) { return std::fabs(a - b) / scale; }

#ifdef DEVMATH_SHARED
// ---- C entry points for tests/test_device_math_on_host.py: the reference-order device functions next to the oracle -------------
#include <cstring>
extern "C" {
void dm_set_newton_consts(const double* XiCL, const double* wBaryCL, double RefMappingEps, int RefMappingGuess) {
  for (int i = 0; i < 2; ++i) { cst.XiCL[i] = XiCL[i]; cst.wBaryCL[i] = wBaryCL[i]; }
  cst.RefMappingEps = RefMappingEps;
  cst.RefMappingGuess = RefMappingGuess;
}
void dm_lagrange(int NP, double x, const double* xGP, const double* wBary, double* L) {
  switch (NP) {
    case 2: lagrange_polys<2>(x, xGP, wBary, L); break;
    case 3: lagrange_polys<3>(x, xGP, wBary, L); break;
    case 4: lagrange_polys<4>(x, xGP, wBary, L); break;
    case 6: lagrange_polys<6>(x, xGP, wBary, L); break;
    default: lagrange_polys<8>(x, xGP, wBary, L); break;
  }
}
// GetPositionInRefElem on one element whose tables are copied into a GeoElem exactly as piclas_gpu_init does
void dm_position_in_ref_elem(const double* XCL24, const double* dXCL72, const double* bary3, const double* xez18, const double* slen6,
                             int64_t n, const double* x, int forceMode, double* xi, int32_t* status) {
  GeoElem g;
  std::memset(&g, 0, sizeof g);
  std::memcpy(g.XCL, XCL24, 24 * 8);
  std::memcpy(g.dXCL, dXCL72, 72 * 8);
  std::memcpy(g.bary, bary3, 3 * 8);
  std::memcpy(g.xez, xez18, 18 * 8);
  std::memcpy(g.slen, slen6, 6 * 8);
  for (int64_t i = 0; i < n; ++i) status[i] = position_in_ref_elem(&g, x + 3 * i, xi + 3 * i, forceMode != 0, true);
}
// ParticleInsideQuad3D on one element: corners (tensor order), side -> corner map, concave mask as in TriaElem
void dm_inside_quad3d(const double* corner24, const int32_t* sideNode24, int concaveMask, int64_t n, const double* x, int32_t* inside,
                      uint32_t* mask) {
  TriaElem t;
  std::memset(&t, 0, sizeof t);
  for (int c = 0; c < 8; ++c) for (int d = 0; d < 3; ++d) t.corner[c][d] = corner24[3 * c + d];
  for (int s = 0; s < 6; ++s) for (int k = 0; k < 4; ++k) t.sideNode[s][k] = (uint8_t)sideNode24[4 * s + k];
  t.concave = (uint8_t)concaveMask;
  for (int64_t i = 0; i < n; ++i) inside[i] = inside_quad3d_mask<false>(&t, x + 3 * i, mask[i]) ? 1 : 0;
}
}
#else
int main() {
  constexpr int NP = 4;
  const double xg[4] = {-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526};   // Gauss, N = 3
  for (int i = 0; i < 4; ++i) {
    cst.xGP[i] = xg[i];
    double w = 1;
    for (int j = 0; j < 4; ++j) if (j != i) w *= (xg[i] - xg[j]);
    cst.wBary[i] = 1.0 / w;
  }
  cst.ChargeIC[0] = -1.60217653e-19; cst.MassIC[0] = 9.1093826e-31; cst.c2_inv = 1.0 / (299792458.0 * 299792458.0);
  cst.ChargeIC[1] = 0.0; cst.MassIC[1] = 1.0;
  cst.DoInterpolation = 1;
  std::mt19937_64 rng(1);
  std::uniform_real_distribution<double> U(-1, 1);
  double wL = 0, wF = 0, wPx = 0, wPv = 0;
  for (int t = 0; t < 100000; ++t) {
    const double xi[3] = {U(rng), U(rng), U(rng)};
    double La[NP], Lb[NP];
    lagrange_polys<NP>(xi[0], cst.xGP, cst.wBary, La);        // barycentric form with the node-hit branch (basis.f90:1223-1264)
    lagrange_fast<NP>(xi[0], Lb);                             // product form
    for (int i = 0; i < NP; ++i) wL = std::fmax(wL, std::fabs(La[i] - Lb[i]));
    // random field tile: reference layout U[((k*NP+j)*NP+i)*3+c], staged layout sE[((k*NP+j)*3+c)*NP+i]
    double Uref[NP * NP * NP * 3], sE[NP * NP * NP * 3];
    for (int k = 0; k < NP; ++k) for (int j = 0; j < NP; ++j) for (int i = 0; i < NP; ++i) for (int c = 0; c < 3; ++c) {
      const double v = U(rng);
      Uref[((k * NP + j) * NP + i) * 3 + c] = v;
      sE[((k * NP + j) * 3 + c) * NP + i] = v;
    }
    double a[3], b[3];
    evaluate_field<NP>(xi, Uref, a);
    evaluate_field_fast<NP>(xi, sE, b);
    for (int c = 0; c < 3; ++c) wF = std::fmax(wF, std::fabs(a[c] - b[c]));
    // push: Boris-Leapfrog with B = 0 (the restructured form) and Leapfrog / neutral / B != 0 (delegated to the reference order)
    for (int mode = 0; mode < 4; ++mode) {
      cst.TimeDiscMethod = (mode == 1) ? PGPU_TIMEDISC_LEAPFROG : PGPU_TIMEDISC_BORIS_LEAPFROG;
      const int spec = (mode == 2) ? 1 : 0;
      const double vs = 0.3 * 299792458.0, dt = 1e-9;
      double x1[3] = {U(rng), U(rng), U(rng)}, v1[3] = {vs * U(rng), vs * U(rng), vs * U(rng)};
      double F[6] = {1e4 * U(rng), 1e4 * U(rng), 1e4 * U(rng), 0, 0, 0};
      if (mode == 3) { F[3] = 1e-2 * U(rng); F[4] = 1e-2 * U(rng); F[5] = 1e-2 * U(rng); }
      double x2[3] = {x1[0], x1[1], x1[2]}, v2[3] = {v1[0], v1[1], v1[2]};
      bool n1 = (t & 1) != 0, n2 = n1;
      push_particle(x1, v1, F, spec, n1, dt);
      push_particle_fast(x2, v2, F, spec, n2, dt);
      if (n1 || n2) { std::printf("IsNewPart not cleared\n"); return 1; }
      for (int d = 0; d < 3; ++d) { wPx = std::fmax(wPx, rel(x1[d], x2[d], 1.0)); wPv = std::fmax(wPv, rel(v1[d], v2[d], vs)); }
    }
  }
  std::printf("lagrange %.3e field %.3e push_x %.3e push_v %.3e\n", wL, wF, wPx, wPv);
  return (wL <= 1e-14 && wF <= 1e-13 && wPx <= 1e-14 && wPv <= 1e-14) ? 0 : 1;
}
#endif
