// device_track_host.cpp — the element-crossing code of the CUDA path (tria_hop, piclas_b200/csrc/kernels.cuh, with the
// determinant tests of math.cuh) compiled for the HOST and driven particle by particle, so that the reference-order
// (params.arithmetic = 0) TriaTracking of the device — neighbour walk, periodic shift, specular reflection — can be run on
// the reference's own tracking checks where no GPU exists.  Test infrastructure (tests/test_device_math_on_host.py).  The
// loop around tria_hop below stands in for k_interp_push's inside test of the own element and k_track_leavers' walk; the
// kernels themselves are what the -m gpu tests check.  tria_hop is instantiated with G = false (records read through plain
// loads, as from a CTA's shared-memory copy); G = true differs only in fetching them with 256-bit PTX loads.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#undef __device__
#undef __host__
#undef __global__
#undef __constant__
#undef __shared__
#undef __forceinline__
#undef __noinline__
#undef __align__
#undef __launch_bounds__
#define __device__
#define __host__
#define __global__
#define __constant__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define asm(...)   /* inline PTX sits behind template flags that are false here, or in kernels that are never instantiated */
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static const uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static const dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __syncthreads_or(int p) { return p; }
static inline unsigned __activemask() { return 1u; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p += v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline int max(int a, int b) { return a > b ? a : b; }   // CUDA's integer overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int __double2hiint(double d) { long long v; std::memcpy(&v, &d, 8); return (int)(v >> 32); }
#include "kernels.cuh"
#include "ref.cuh"
#include "hint.cuh"
#include "emit.cuh"

static int dt_build_tables(int nG, int elemInfoSize, int sideInfoSize, const int32_t* ElemInfo, const int32_t* SideInfo, const double* NodeCoords,
                           const int32_t* ElemSideNodeID, const int32_t* ConcaveElemSide, int nBCs, const int32_t* bc_kind,
                           const int32_t* bc_alpha, int nPV, const double* PeriodicVectors, int fast, std::vector<TriaElem>& tria,
                           std::vector<PlaneElem>& planes) {
  tria.assign((size_t)nG, TriaElem());
  for (int e = 0; e < nG; ++e) {
    const int32_t* ei = ElemInfo + (size_t)e * elemInfoSize;
    const int firstSide = ei[2], firstNode = ei[4];                      // ELEM_FIRSTSIDEIND, ELEM_FIRSTNODEIND
    if (ei[3] - firstSide != 6 || ei[5] - firstNode != 8) return -1;
    TriaElem& t = tria[e];
    std::memset(&t, 0, sizeof t);
    for (int c = 0; c < 8; ++c)
      for (int d = 0; d < 3; ++d) t.corner[c][d] = NodeCoords[(size_t)(firstNode + c) * 3 + d];
    for (int s = 0; s < 6; ++s) {
      const int sid = firstSide + s + 1;
      const int32_t* si = SideInfo + (size_t)(sid - 1) * sideInfoSize;
      t.nbElem[s] = si[2];                                                 // SIDE_NBELEMID
      t.sideID[s] = sid;
      t.bcid[s] = (uint8_t)si[4];                                          // SIDE_BCID
      if (ConcaveElemSide[(size_t)e * 6 + s]) t.concave |= (uint8_t)(1u << s);
      for (int k = 0; k < 4; ++k) t.sideNode[s][k] = (uint8_t)(ElemSideNodeID[((size_t)e * 6 + s) * 4 + k] - firstNode);
    }
  }
  cst.nBCs = nBCs;
  for (int b = 0; b < nBCs; ++b) { cst.bc_kind[b] = bc_kind[b]; cst.bc_alpha[b] = bc_alpha[b]; }
  cst.nPeriodicVectors = nPV;
  for (int p = 0; p < nPV; ++p) for (int d = 0; d < 3; ++d) cst.PeriodicVectors[p][d] = PeriodicVectors[3 * p + d];
  planes.assign(fast ? (size_t)nG : 0, PlaneElem());
  for (int e = 0; fast && e < nG; ++e) {
    const TriaElem& t = tria[e];
    PlaneElem& pl = planes[e];
    std::memset(&pl, 0, sizeof pl);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < 8; ++c)
      for (int d = 0; d < 3; ++d) { lo[d] = std::fmin(lo[d], t.corner[c][d]); hi[d] = std::fmax(hi[d], t.corner[c][d]); }
    const double diam = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
    pl.tol = 1e-8 * diam;
    for (int s = 0; s < 6; ++s)
      for (int tr = 0; tr < 2; ++tr) {
        const double* P1 = t.corner[t.sideNode[s][0]];
        const double* Pb = t.corner[t.sideNode[s][tr + 1]];
        const double* Pc = t.corner[t.sideNode[s][tr + 2]];
        const double u[3] = {Pb[0] - P1[0], Pb[1] - P1[1], Pb[2] - P1[2]}, w[3] = {Pc[0] - P1[0], Pc[1] - P1[1], Pc[2] - P1[2]};
        const double N[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
        const double len = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
        if (!(len > 0.)) return -1;
        const int k = 2 * s + tr;
        for (int d = 0; d < 3; ++d) pl.pl[k][d] = -N[d] / len;
        pl.pl[k][3] = pl.pl[k][0] * P1[0] + pl.pl[k][1] * P1[1] + pl.pl[k][2] * P1[2];
      }
    for (int s = 0; s < 6; ++s) pl.concave2 |= (uint32_t)((t.concave >> s) & 1u) << (2 * s);
    bool planar = true;
    for (int k = 0; k < 12; ++k)
      for (int c = 0; c < 8; ++c)
        if (pl.pl[k][0] * t.corner[c][0] + pl.pl[k][1] * t.corner[c][1] + pl.pl[k][2] * t.corner[c][2] - pl.pl[k][3] < -1e-12 * diam) planar = false;
    for (int s = 0; s < 6; ++s) {
      const double *a = pl.pl[2 * s], *b = pl.pl[2 * s + 1];
      if (std::fabs(a[0] - b[0]) > 1e-12 || std::fabs(a[1] - b[1]) > 1e-12 || std::fabs(a[2] - b[2]) > 1e-12 || std::fabs(a[3] - b[3]) > 1e-12 * diam)
        planar = false;
      const double* P1 = t.corner[t.sideNode[s][0]];
      const double* P2 = t.corner[t.sideNode[s][1]];
      const double* P3 = t.corner[t.sideNode[s][2]];
      const double dgn[3] = {P3[0] - P1[0], P3[1] - P1[1], P3[2] - P1[2]};
      const double mm[3] = {a[1] * dgn[2] - a[2] * dgn[1], a[2] * dgn[0] - a[0] * dgn[2], a[0] * dgn[1] - a[1] * dgn[0]};
      const double ml = std::sqrt(mm[0] * mm[0] + mm[1] * mm[1] + mm[2] * mm[2]);
      if (!(ml > 0.)) { planar = false; continue; }
      const double sgn = (mm[0] * (P2[0] - P1[0]) + mm[1] * (P2[1] - P1[1]) + mm[2] * (P2[2] - P1[2])) >= 0. ? 1. : -1.;
      for (int d = 0; d < 3; ++d) pl.dg[s][d] = sgn * mm[d] / ml;
      pl.dg[s][3] = pl.dg[s][0] * P1[0] + pl.dg[s][1] * P1[1] + pl.dg[s][2] * P1[2];
    }
    pl.planar = planar ? 1u : 0u;
  }
  return 0;
}
extern "C" {
// Tria records from the host tables, as piclas_gpu_init builds them (piclas_gpu.cu, "per-element records"), then for every
// particle: inside test of the own element at the pushed position, walk while tria_hop asks for another crossing.
// x: pushed positions (in/out: periodic shifts, reflections), lp: LastPartPos (in/out), v (in/out: reflections), elem (in/out,
// 1-based; 0 = removed), status out (TRK_*).  Returns the largest number of crossings one particle needed, -1 on bad tables.
// fast != 0: the restructured arithmetic (params.arithmetic = 1) - inside test through the triangle planes with the exact
// fallback, exit-side shortcut on planar convex elements - on PlaneElem records built as piclas_gpu_init builds them (that
// builder is host code inside piclas_gpu_init and is restated here; the device functions under test are the originals).
int dt_tria_track(int nG, int elemInfoSize, int sideInfoSize, const int32_t* ElemInfo, const int32_t* SideInfo, const double* NodeCoords,
                  const int32_t* ElemSideNodeID, const int32_t* ConcaveElemSide, int nBCs, const int32_t* bc_kind,
                  const int32_t* bc_alpha, int nPV, const double* PeriodicVectors, int64_t n, double* x, double* lp, double* v,
                  int32_t* elem, int32_t* status, int fast) {
  std::vector<TriaElem> tria;
  std::vector<PlaneElem> planes;
  if (dt_build_tables(nG, elemInfoSize, sideInfoSize, ElemInfo, SideInfo, NodeCoords, ElemSideNodeID, ConcaveElemSide, nBCs, bc_kind, bc_alpha, nPV,
                      PeriodicVectors, fast, tria, planes)) return -1;
  int maxHops = 0;
  for (int64_t i = 0; i < n; ++i) {
    double* xi = x + 3 * i; double* li = lp + 3 * i; double* vi = v + 3 * i;
    int ElemID = elem[i];
    uint32_t mask = 0;
    int st = TRK_OK;
    auto reflect = [&](const double nrm[3]) {
      const double vn = (vi[0] * nrm[0] + vi[1] * nrm[1]) + vi[2] * nrm[2];
      vi[0] = vi[0] - 2. * vn * nrm[0]; vi[1] = vi[1] - 2. * vn * nrm[1]; vi[2] = vi[2] - 2. * vn * nrm[2];
    };
    const bool insideOwn = fast ? inside_fast<false>(&planes[ElemID - 1], &tria[ElemID - 1], xi, mask)
                                : inside_quad3d_mask<false>(&tria[ElemID - 1], xi, mask);       // particle_triatracking.f90:203-218
    if (!insideOwn) {
      HopHist h;
      h.clear();
      st = -1;
      int hops = 0;
      while (st == -1) {
        if (fast)
          st = tria_hop<true, false, 2>(&tria[ElemID - 1], tria.data(), &planes[ElemID - 1],
                                        [&](int, int ne) { return (const PlaneElem*)&planes[ne - 1]; }, reflect, xi, li, ElemID, mask, h);
        else
          st = tria_hop<false, false, 0>(&tria[ElemID - 1], tria.data(), (const PlaneElem*)nullptr,
                                         [&](int, int) { return (const PlaneElem*)nullptr; }, reflect, xi, li, ElemID, mask, h);
        if (st == -1 && ++hops > 100000) st = TRK_ERR_LOOP;
      }
      if (hops > maxHops) maxHops = hops;
    }
    status[i] = st;
    elem[i] = (st == TRK_OK) ? ElemID : 0;
  }
  return maxHops;
}
// k_far_hint's per-record decision (far_hint_record of csrc/hint.cuh) on the host.  PushElem / HintNb are filled as build_push_elems
// (csrc/bins_host.inc) fills the fields the hint reads; that builder is host code inside the library and is restated here, the device
// function under test is the original.  For every flight lp -> x starting in elem: fin[i] = element the hint settles the particle
// in (0: left to the exact walk), x (and lp) displaced by the periodic vectors crossed.  Returns the number of settled records.
int64_t dt_tria_hint(int nG, int elemInfoSize, int sideInfoSize, const int32_t* ElemInfo, const int32_t* SideInfo, const double* NodeCoords,
                     const int32_t* ElemSideNodeID, const int32_t* ConcaveElemSide, int nBCs, const int32_t* bc_kind,
                     const int32_t* bc_alpha, int nPV, const double* PeriodicVectors, int64_t n, double* x, double* lp,
                     const int32_t* elem, int32_t* fin) {
  std::vector<TriaElem> tria;
  std::vector<PlaneElem> planes;
  if (dt_build_tables(nG, elemInfoSize, sideInfoSize, ElemInfo, SideInfo, NodeCoords, ElemSideNodeID, ConcaveElemSide, nBCs, bc_kind, bc_alpha, nPV,
                      PeriodicVectors, 1, tria, planes)) return -1;
  auto shift_code = [&](const TriaElem& t, int s, uint32_t& code) {
    code = 0;
    if (t.nbElem[s] < 1) return false;
    if (t.bcid[s] == 0) return true;
    if (bc_kind[t.bcid[s] - 1] != PGPU_BC_PERIODIC) return false;
    const int pvid = bc_alpha[t.bcid[s] - 1];
    const int pv = pvid < 0 ? -pvid : pvid;
    if (pv < 1 || pv > nPV) return false;
    code = (uint32_t)pv | (pvid < 0 ? 16u : 0u);
    return true;
  };
  std::vector<PushElem> pe((size_t)nG);
  std::vector<HintNb> hn((size_t)nG);
  std::memset(pe.data(), 0, pe.size() * sizeof(PushElem));
  std::memset(hn.data(), 0, hn.size() * sizeof(HintNb));
  for (int e = 0; e < nG; ++e) {
    const PlaneElem& pl = planes[e];
    const TriaElem& t = tria[e];
    PushElem& P = pe[e];
    P.planar = pl.planar;
    P.tol = pl.tol;
    for (int s = 0; s < 6; ++s)
      for (int c = 0; c < 4; ++c) { P.pl[s][c] = pl.pl[2 * s][c]; P.dg[s][c] = pl.dg[s][c]; }
    if (!pl.planar) continue;
    for (int s = 0; s < 6; ++s) {
      uint32_t c1 = 0;
      if (!shift_code(t, s, c1) || !planes[t.nbElem[s] - 1].planar) continue;
      const int nb = t.nbElem[s];
      P.nbValid |= 1u << s;
      P.nbtol[s] = planes[nb - 1].tol;
      for (int o = 0; o < 6; ++o)
        for (int c = 0; c < 4; ++c) P.nbpl[s][o][c] = planes[nb - 1].pl[2 * o][c];
      hn[e].nb[s] = nb;
      hn[e].sh1 |= c1 << (5 * s);
      const TriaElem& t1 = tria[nb - 1];
      for (int o = 0; o < 6; ++o) {
        uint32_t c2 = 0;
        if (!shift_code(t1, o, c2) || !planes[t1.nbElem[o] - 1].planar) continue;
        hn[e].nbnb[s][o] = t1.nbElem[o];
        hn[e].sh2[s] |= c2 << (5 * o);
      }
    }
  }
  int64_t settled = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int ge = elem[i];
    bool moved = false;
    fin[i] = pe[ge - 1].planar ? far_hint_record<false>(pe[ge - 1], hn[ge - 1], tria.data(), planes.data(), ge, x + 3 * i, lp + 3 * i, moved) : 0;
    if (fin[i] > 0) ++settled;
  }
  return settled;
}
// The emission of csrc/emit.cuh on the host: emit_lattice_position for every lattice point and single_point_to_element (TriaTracking:
// ParticleInsideQuad3D; RefMapping: Newton + ElemEpsOneCell) among the elements first..last.  X [n][3], elem [n] (-1: not accepted).
int dt_emit_lattice(const pgpu_mesh_t* m, const pgpu_params_t* p, int kind, int nx, int ny, int nz, double amplitude, double wavenumber,
                    int first, int last, double* X, int32_t* elem) {
  const int nG = m->nGlobalElems;
  std::vector<TriaElem> tria;
  std::vector<PlaneElem> planes;
  if (dt_build_tables(nG, m->elemInfoSize, m->sideInfoSize, m->ElemInfo, m->SideInfo, m->NodeCoords, m->ElemSideNodeID, m->ConcaveElemSide, m->nBCs,
                      m->bc_kind, m->bc_alpha, m->nPeriodicVectors, m->PeriodicVectors, 0, tria, planes)) return -1;
  std::vector<GeoElem> geo((size_t)nG);
  for (int e = 0; e < nG; ++e) {
    GeoElem& ge = geo[e];
    std::memset(&ge, 0, sizeof ge);
    std::memcpy(ge.XCL, m->XCL_NGeo + (size_t)e * 24, 24 * 8);
    std::memcpy(ge.dXCL, m->dXCL_NGeo + (size_t)e * 72, 72 * 8);
    std::memcpy(ge.bary, m->ElemBaryNGeo + (size_t)e * 3, 3 * 8);
    std::memcpy(ge.xez, m->XiEtaZetaBasis + (size_t)e * 18, 18 * 8);
    std::memcpy(ge.slen, m->slenXiEtaZetaBasis + (size_t)e * 6, 6 * 8);
  }
  for (int i = 0; i < 2; ++i) { cst.XiCL[i] = m->XiCL_NGeo[i]; cst.wBaryCL[i] = m->wBaryCL_NGeo[i]; }
  cst.RefMappingEps = p->RefMappingEps;
  cst.RefMappingGuess = p->RefMappingGuess;
  for (int d = 0; d < 3; ++d) {
    cst.FIBGMdeltas[d] = m->FIBGMdeltas[d]; cst.xyzminglob[d] = m->xyzminglob[d];
    cst.FIBGMmin[d] = m->FIBGMmin[d]; cst.FIBGMmax[d] = m->FIBGMmax[d];
  }
  RefTables T;
  std::memset(&T, 0, sizeof T);
  T.geo = geo.data(); T.ElemBary = m->ElemBaryNGeo; T.ElemRadius2 = m->ElemRadius2NGeo; T.ElemEpsOneCell = m->ElemEpsOneCell;
  T.FIBGM_nElems = m->FIBGM_nElems; T.FIBGM_offsetElem = m->FIBGM_offsetElem; T.FIBGM_Element = m->FIBGM_Element;
  EmitSpec s;
  std::memset(&s, 0, sizeof s);
  s.kind = kind; s.nx = nx; s.ny = ny; s.nz = nz; s.amplitude = amplitude; s.wavenumber = wavenumber;
  for (int d = 0; d < 3; ++d) { s.lo[d] = m->xyzminglob[d]; s.len[d] = std::fabs(m->xyzmaxglob[d] - m->xyzminglob[d]); }
  s.species = 1; s.firstLocal = first; s.lastLocal = last;
  const bool ref = p->TrackingMethod == 1;   // REFMAPPING (piclas.h:357)
  const int64_t n = (int64_t)nx * ny * nz;
  for (int64_t t = 0; t < n; ++t) {
    emit_lattice_position(s, t, X + 3 * t);
    elem[t] = ref ? single_point_to_element<true, false>(T, tria.data(), X + 3 * t, first, last)
                  : single_point_to_element<false, false>(T, tria.data(), X + 3 * t, first, last);
  }
  return 0;
}
// ParticleRefTracking of the device (ref_tracking of csrc/ref.cuh: Newton in the old element, BC-side intersections, periodic shift
// and reflection, FIBGM relocation incl. the repeated-selection path, LocateParticleInElement fallback) for n particles on the host.
// The tables are the host's (pgpu_mesh_t as handed to piclas_gpu_init); the GeoElem records and the constant table are filled as
// piclas_gpu_init fills them.  x pushed positions, lp LastPartPos, v, xi PartPosRef, elem: all in/out.
int dt_ref_track(const pgpu_mesh_t* m, const pgpu_params_t* p, int64_t n, double* x, double* lp, double* v, double* xi, int32_t* elem,
                 int32_t* status, int32_t* relocated) {
  const int nG = m->nGlobalElems;
  std::vector<GeoElem> geo((size_t)nG);
  for (int e = 0; e < nG; ++e) {
    GeoElem& ge = geo[e];
    std::memset(&ge, 0, sizeof ge);
    std::memcpy(ge.XCL, m->XCL_NGeo + (size_t)e * 24, 24 * 8);
    std::memcpy(ge.dXCL, m->dXCL_NGeo + (size_t)e * 72, 72 * 8);
    std::memcpy(ge.bary, m->ElemBaryNGeo + (size_t)e * 3, 3 * 8);
    std::memcpy(ge.xez, m->XiEtaZetaBasis + (size_t)e * 18, 18 * 8);
    std::memcpy(ge.slen, m->slenXiEtaZetaBasis + (size_t)e * 6, 6 * 8);
  }
  for (int i = 0; i < 2; ++i) { cst.XiCL[i] = m->XiCL_NGeo[i]; cst.wBaryCL[i] = m->wBaryCL_NGeo[i]; }
  cst.RefMappingEps = p->RefMappingEps;
  cst.RefMappingGuess = p->RefMappingGuess;
  cst.nBCs = m->nBCs;
  for (int b = 0; b < m->nBCs; ++b) { cst.bc_kind[b] = m->bc_kind[b]; cst.bc_alpha[b] = m->bc_alpha[b]; }
  cst.nPeriodicVectors = m->nPeriodicVectors;
  for (int k = 0; k < m->nPeriodicVectors; ++k) for (int d = 0; d < 3; ++d) cst.PeriodicVectors[k][d] = m->PeriodicVectors[k * 3 + d];
  for (int d = 0; d < 3; ++d) {
    cst.FIBGMdeltas[d] = m->FIBGMdeltas[d]; cst.xyzminglob[d] = m->xyzminglob[d];
    cst.FIBGMmin[d] = m->FIBGMmin[d]; cst.FIBGMmax[d] = m->FIBGMmax[d];
  }
  RefTables T;
  T.geo = geo.data(); T.ElemToBCSides = m->ElemToBCSides; T.SideBCMetrics = m->SideBCMetrics; T.SideInfo = m->SideInfo;
  T.sideInfoSize = m->sideInfoSize; T.SideNormVec = m->SideNormVec; T.SideDistance = m->SideDistance;
  T.BaseVectors0 = m->BaseVectors0; T.BaseVectors1 = m->BaseVectors1; T.BaseVectors2 = m->BaseVectors2; T.BaseVectors3 = m->BaseVectors3;
  T.SideType = m->SideType; T.ElemBary = m->ElemBaryNGeo; T.ElemRadius = m->ElemRadiusNGeo; T.ElemRadius2 = m->ElemRadius2NGeo;
  T.ElemEpsOneCell = m->ElemEpsOneCell; T.FIBGM_nElems = m->FIBGM_nElems; T.FIBGM_offsetElem = m->FIBGM_offsetElem;
  T.FIBGM_Element = m->FIBGM_Element;
  // velocities as the structure of arrays bc_tracking reflects in place
  std::vector<double> v0((size_t)n), v1((size_t)n), v2((size_t)n);
  for (int64_t i = 0; i < n; ++i) { v0[i] = v[3 * i]; v1[i] = v[3 * i + 1]; v2[i] = v[3 * i + 2]; }
  PartBuf pb;
  std::memset(&pb, 0, sizeof pb);
  pb.v[0] = v0.data(); pb.v[1] = v1.data(); pb.v[2] = v2.data();
  int worst = 0;
  for (int64_t i = 0; i < n; ++i) {
    int e = elem[i];
    bool rel = false;
    const int st = ref_tracking(T, pb, i, x + 3 * i, lp + 3 * i, xi + 3 * i, e, rel);
    status[i] = st;
    relocated[i] = rel ? 1 : 0;
    elem[i] = (st == TRK_OK) ? e : 0;
    if (st > worst) worst = st;
  }
  for (int64_t i = 0; i < n; ++i) { v[3 * i] = v0[i]; v[3 * i + 1] = v1[i]; v[3 * i + 2] = v2[i]; }
  return worst;
}
// DepositionMethod_CVWM per particle on the host (deposit_particle_general of csrc/kernels.cuh: GetPositionInRefElem with ForceMode,
// the eight trilinear weights in CGNS corner order or the inverse-distance fallback): accumulates every particle of element e into
// acc[e][corner][1:4] in particle order and returns the reference positions.  Corners in CGNS order as the kernel hands them over.
int dt_cvwm_accumulate(const pgpu_mesh_t* m, const pgpu_params_t* p, int64_t n, const double* PartState, const int32_t* spec,
                       const int32_t* elem, double* acc /*[nElems][8][4]*/, double* xiOut /*[n][3]*/, int32_t* failed) {
  for (int i = 0; i < 2; ++i) { cst.XiCL[i] = m->XiCL_NGeo[i]; cst.wBaryCL[i] = m->wBaryCL_NGeo[i]; }
  cst.RefMappingEps = p->RefMappingEps;
  cst.RefMappingGuess = p->RefMappingGuess;
  for (int s = 0; s < p->nSpecies; ++s) { cst.ChargeIC[s] = p->ChargeIC[s]; cst.MPF[s] = p->MacroParticleFactor[s]; }
  std::vector<double> f((size_t)6 * n), xif((size_t)3 * n);
  std::vector<uint8_t> meta((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    for (int a = 0; a < 6; ++a) f[(size_t)a * n + i] = PartState[6 * i + a];
    meta[i] = (uint8_t)((spec[i] - 1) & META_SPEC_MASK);
  }
  PartBuf pb;
  std::memset(&pb, 0, sizeof pb);
  pb.f = f.data(); pb.xif = xif.data(); pb.stride = n; pb.meta = meta.data();
  static DepAcc sAcc;
  for (int64_t i = 0; i < n; ++i) {
    const int e = elem[i] - 1;
    GeoElem g;
    std::memset(&g, 0, sizeof g);
    std::memcpy(g.XCL, m->XCL_NGeo + (size_t)e * 24, 24 * 8);
    std::memcpy(g.dXCL, m->dXCL_NGeo + (size_t)e * 72, 72 * 8);
    std::memcpy(g.bary, m->ElemBaryNGeo + (size_t)e * 3, 3 * 8);
    std::memcpy(g.xez, m->XiEtaZetaBasis + (size_t)e * 18, 18 * 8);
    std::memcpy(g.slen, m->slenXiEtaZetaBasis + (size_t)e * 6, 6 * 8);
    const int firstNode = m->ElemInfo[(size_t)e * m->elemInfoSize + 4];
    double corner[8][3];
    for (int c = 0; c < 8; ++c) for (int d = 0; d < 3; ++d) corner[c][d] = m->NodeCoords[(size_t)(firstNode + c) * 3 + d];
    for (int a = 0; a < 32; ++a) sAcc[a][0] = acc[(size_t)e * 32 + a];
    deposit_particle_general(pb, i, &g, corner, sAcc, 0);
    for (int a = 0; a < 32; ++a) acc[(size_t)e * 32 + a] = sAcc[a][0];
    for (int d = 0; d < 3; ++d) xiOut[3 * i + d] = xif[(size_t)d * n + i];
    failed[i] = (meta[i] & META_XIFAIL) ? 1 : 0;
  }
  return 0;
}
}
