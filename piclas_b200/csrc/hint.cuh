// hint.cuh — per-element records of the binned push (PushElem) and the plane-decided walk of the far records (k_far_hint).
// No kernels and no inline PTX here: tests/device_track_host.cpp compiles this header for the host and runs far_hint_record against the
// exact walk.
#pragma once
#include "kernels.cuh"

// Everything k_bin_push stages for one element besides the field tile: one contiguous record -> one bulk (TMA) copy.
struct __align__(16) PushElem {
  double pl[6][4];        // own side planes: inward unit normal, offset (valid when planar)
  double dg[6][4];        // per side: plane through the triangle diagonal (as PlaneElem::dg)
  double nbpl[6][6][4];   // side planes of the neighbour behind side s
  double tol;             // 1e-8 element diameters
  double nbtol[6];
  double x0[3], A[3][3];  // closed-form reference coordinates of an affine element: xi = A (x - x0) - 1
  int32_t nbLocal[6];     // local index of the neighbour that takes movers through side s directly, -1: far list
  int32_t nbBox[6];       // inbox of that neighbour = its local side facing this element
  uint32_t planar;        // convex element with planar sides: exit side from the side planes
  uint32_t affine;
  uint32_t nbValid;       // bit s: nbpl[s] / nbtol[s] hold the planes of an inner, planar neighbour (any rank)
  uint32_t shiftCode;     // 5 bits per side with nbLocal >= 0: periodic vector the crossing adds (0: inner side), see build_push_elems
};
static_assert(sizeof(PushElem) % 16 == 0, "PushElem is copied with cp.async.bulk (16-byte granules)");

// second ring of an element: global ids (0: none / boundary side / non-planar) of the neighbours and of their neighbours
struct __align__(16) HintNb {
  int32_t nb[6];
  int32_t nbnb[6][6];
  uint32_t sh1;       // 5 bits per side: periodic vector added when crossing into nb[s] (as PushElem::shiftCode)
  uint32_t sh2[6];    // the same for the crossing nb[s] -> nbnb[s][o]
  int32_t pad[3];
};
static_assert(sizeof(HintNb) % 16 == 0, "HintNb is staged with 16-byte loads");

// -1: undecided, 0: clearly inside, 1: leaves through `side`; the crossing point is lp + (num / den) (x - lp), den > 0 (the
// diagonal test hint_diag is still due).  Division-free: with a_o = ol_o / (ol_o - ox_o) and positive denominators, a_1 < a_2 is
// ol_1 den_2 < ol_2 den_1, and a distance at the crossing point, ol + a_s (ox - ol) > tol, is ol den_s + ol_s (ox - ol) > tol den_s.
// den_s >= 1e4 tol (the flight crosses the plane by more than 1e-4 element diameters) keeps the margin tol den_s four orders of
// magnitude above the rounding error of the products; shorter crossings are left to the exact walk.
template <class PL>
__device__ __forceinline__ int hint_exit(PL pl, double tol, const double x[3], const double lp[3], int& side, double& num, double& den) {
  double ol[6], ox[6];
  uint32_t neg = 0;
  bool amb = false;
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    double a, b, c, d;
    pl(o, a, b, c, d);
    ox[o] = fma(a, x[0], fma(b, x[1], fma(c, x[2], -d)));
    ol[o] = fma(a, lp[0], fma(b, lp[1], fma(c, lp[2], -d)));
    amb |= fabs(ox[o]) <= tol;
    neg |= ((uint32_t)__double2hiint(ox[o]) >> 31) << o;
  }
  if (amb) return -1;
  if (neg == 0u) return 0;
  int s = -1;
  num = 1.0;
  den = 0.0;   // a = num / den = +inf
  bool ok = true;
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    const bool cr = (neg >> o) & 1u;
    const double dn = ol[o] - ox[o];
    if (cr && !(ol[o] > tol)) ok = false;
    if (cr && ol[o] * den < num * dn) { num = ol[o]; den = dn; s = o; }
  }
  if (!ok || s < 0 || !(den >= 1e4 * tol)) return -1;
  const double lim = tol * den;
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    const double oc = fma(ol[o], den, num * (ox[o] - ol[o]));
    if (o != s && !(oc > lim)) ok = false;
  }
  side = s;
  return ok ? 1 : -1;
}
// crossing point clearly off the triangle diagonal of the side (plane a, b, c, d as PlaneElem::dg)
__device__ __forceinline__ bool hint_diag(double a, double b, double c, double d, double num, double den, double tol, const double x[3],
                                          const double lp[3]) {
  const double gl = fma(a, lp[0], fma(b, lp[1], fma(c, lp[2], -d)));
  const double gx = fma(a, x[0], fma(b, x[1], fma(c, x[2], -d)));
  return fabs(fma(gl, den, num * (gx - gl))) > tol * den;
}

// One far record: up to three plane-decided crossings from the element's staged records (k_far_hint).  Returns the global id of the
// element the particle comes to rest in, 0 when any decision falls within the margins or leaves the planar / inner / periodic
// case (the exact walk then takes the record from the start).  x, lp are displaced by the periodic vectors crossed (moved).
// G: element tables in global memory (device) / plain host arrays (tests/device_track_host.cpp).
template <bool G>
__device__ __forceinline__ int far_hint_record(const PushElem& pe, const HintNb& hn, const TriaElem* __restrict__ tria,
                                               const PlaneElem* __restrict__ planes, int ge, double x[3], double lp[3], bool& moved) {
  int fin = 0;
  // periodic side: the flight goes on behind the partner side, both of its end points displaced by the periodic vector
  auto periodic = [&](uint32_t sc) {
    if (!sc) return;
    const int pv = (int)(sc & 15u) - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double v = (sc & 16u) ? -cst.PeriodicVectors[pv][d] : cst.PeriodicVectors[pv][d];
      x[d] += v;
      lp[d] += v;
    }
    moved = true;
  };
  int s0 = 0, s1 = 0, s2 = 0;
  double n0 = 0., q0 = 0., n1 = 0., q1 = 0., n2 = 0., q2 = 0.;   // crossing parameters num / den of the crossings
  const int r0 = hint_exit([&](int o, double& a, double& b, double& c, double& d) { a = pe.pl[o][0]; b = pe.pl[o][1]; c = pe.pl[o][2]; d = pe.pl[o][3]; },
                           pe.tol, x, lp, s0, n0, q0);
  if (r0 == 0) fin = ge;
  else if (r0 == 1 && hn.nb[s0] > 0 && hint_diag(pe.dg[s0][0], pe.dg[s0][1], pe.dg[s0][2], pe.dg[s0][3], n0, q0, pe.tol, x, lp)) {
    const int nb1 = hn.nb[s0];
    periodic((hn.sh1 >> (5 * s0)) & 31u);
    const double(*npl)[4] = pe.nbpl[s0];
    const double tol1 = pe.nbtol[s0];
    const int r1 = hint_exit([&](int o, double& a, double& b, double& c, double& d) { a = npl[o][0]; b = npl[o][1]; c = npl[o][2]; d = npl[o][3]; },
                             tol1, x, lp, s1, n1, q1);
    if (r1 == 0) fin = nb1;
    else if (r1 == 1 && hn.nbnb[s0][s1] > 0) {
      // the diagonal plane of the neighbour's side and the planes of the element behind it: independent loads, one round trip
      const int nb2 = hn.nbnb[s0][s1];
      const PlaneElem* p2 = planes + (nb2 - 1);
      double ga, gb, gc, gd;
      load_plane4<G>((planes + (nb1 - 1))->dg[s1], ga, gb, gc, gd);
      const bool offDiag = hint_diag(ga, gb, gc, gd, n1, q1, tol1, x, lp);   // before the displacement of the second crossing
      periodic((hn.sh2[s0] >> (5 * s1)) & 31u);
      const double tol2 = G ? __ldg(&p2->tol) : p2->tol;
      const int r2 = hint_exit([&](int o, double& a, double& b, double& c, double& d) { load_plane4<G>(p2->pl[2 * o], a, b, c, d); },
                               tol2, x, lp, s2, n2, q2);
      if (r2 == 0 && offDiag) fin = nb2;
      else if (r2 == 1 && offDiag) {
        // third crossing (flight through a corner region): inner sides only
        const TriaElem* t2 = tria + (nb2 - 1);
        const int nb3 = t2->nbElem[s2];
        if (t2->bcid[s2] == 0 && nb3 >= 1) {
          const PlaneElem* p3 = planes + (nb3 - 1);
          load_plane4<G>(p2->dg[s2], ga, gb, gc, gd);
          const double tol3 = G ? __ldg(&p3->tol) : p3->tol;
          bool in = (G ? __ldg(&p3->planar) : p3->planar) != 0u;
#pragma unroll
          for (int o = 0; o < 6; ++o) {
            double a, b, c, d;
            load_plane4<G>(p3->pl[2 * o], a, b, c, d);
            if (!(fma(a, x[0], fma(b, x[1], fma(c, x[2], -d))) > tol3)) in = false;
          }
          if (in && hint_diag(ga, gb, gc, gd, n2, q2, tol2, x, lp)) fin = nb3;
        }
      }
    }
  }
  return fin;
}
