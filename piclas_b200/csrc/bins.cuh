// bins.cuh — binned particle storage and the kernels that work on it (TriaTracking + cell_volweight_mean, sm_100a).
//
// Round 1 kept the particle SoA globally sorted by element and re-sorted it every step: 31 % of the step went into moving
// all particles although 60 % of them keep their element.  Here every local element owns a REGION of
// the particle arrays with slack:
//
//     [ main: capMain slots ][ side inboxes, parity 0: 6 x capIn ][ side inboxes, parity 1: 6 x capIn ]
//
//   * stayers are compacted in place at the front of `main` by the CTA that pushes the element (stable, deterministic);
//   * a particle that crosses ONE side plane into a face neighbour owned by this rank is written by the same kernel straight
//     into that neighbour's inbox for this side (slot = its rank among the element's movers through that side: deterministic,
//     no atomics, no sort).  Inboxes are double buffered by step parity: the neighbour's CTA may still be reading the other half;
//   * everything else — edge / corner crossings, boundary sides, long flights, emigrants to other ranks, non-planar elements,
//     reference-order arithmetic, and any particle that finds its target region full — goes to the FAR list, is walked by
//     SingleParticleTriaTracking3D from the start (k_far_walk), sorted by destination (the round-1 radix sort, now on a few
//     per cent of the particles) and handed to the destination element as a range of a compact, element-sorted POOL.
// An element's particles are therefore the concatenation of eight ranges in fixed order: main, inbox 1..6, pool range.
// UpdateNextFreePosition (particle_tools.f90:57-207) becomes the in-place compaction; the order inside an element is a
// deterministic function of the previous order, so the deposition sums are reproducible from run to run.
#pragma once
#include "kernels.cuh"

constexpr int BIN_NT = 128;          // threads per CTA of the per-element kernels
constexpr int BIN_PPT = 2;           // particles per thread and sweep (one 128-bit copy per array and thread)
constexpr int BIN_CHUNK = BIN_NT * BIN_PPT;
constexpr int BIN_NRANGE = 8;        // main, six side inboxes, pool
#ifndef KB_MINB
#define KB_MINB 3                    // resident CTAs / SM of k_bin_push (register cap 65536 / (128 * KB_MINB))
#endif

struct BinView {
  const int64_t* base;      // [nElems + 1] first slot of the element's region (multiple of 32)
  const int32_t* capMain;   // [nElems]
  const int32_t* capIn;     // [nElems] capacity of each side inbox
  int32_t* nMain;           // [nElems]
  int32_t* nIn;             // [2][nElems][8] particles in the side inboxes (entries 6, 7 unused)
  const int64_t* poolOff[2];// [nElems + ...] element ranges of the two pools (segment offsets of the sorted far keys)
  int nElems;
};

__device__ __forceinline__ int64_t bin_inbox_base(const BinView& b, int e, int parity, int box) {
  return b.base[e] + b.capMain[e] + (int64_t)(parity * 6 + box) * b.capIn[e];
}

// far list: particles that are not delivered by the push kernel itself.  Aliases the (idle) sorted buffers.
struct FarBuf {
  double* x[3];    // pushed position (in), final position (out: periodic shifts, reflections)
  double* lp[3];   // LastPartPos
  double* v[3];
  int32_t* elem;   // in: global element the walk starts in; out: final global element (0 = removed)
  uint8_t* meta;
  int64_t* id;     // optional
  uint32_t* src;   // unique origin slot: tie-break that makes the order inside a destination deterministic
};

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
  uint32_t pad[2];
};
static_assert(sizeof(PushElem) % 16 == 0, "PushElem is copied with cp.async.bulk (16-byte granules)");

// ---- 1-D bulk copies (TMA) with mbarrier completion ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smemDst, const void* gmemSrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smemDst)), "l"(gmemSrc) : "memory");
}

// ---- element ranges ---------------------------------------------------------------------------------------------------------
// start[r]: first slot of range r (r < 7: bins arrays, r == 7: pool arrays); pre[r]: first virtual index of range r, every range
// padded to an even length so that a thread's pair of particles never straddles two ranges and stays 16-byte aligned.
struct ElemRanges {
  int64_t start[BIN_NRANGE];
  int cnt[BIN_NRANGE];
  int pre[BIN_NRANGE + 1];
};

__device__ __forceinline__ void load_ranges(ElemRanges& R, const BinView& bv, int e, int cur, int tid) {
  if (tid < BIN_NRANGE) {
    int64_t s;
    int c;
    if (tid == 0) { s = bv.base[e]; c = bv.nMain[e]; }
    else if (tid < 7) { s = bin_inbox_base(bv, e, cur, tid - 1); c = bv.nIn[((size_t)cur * bv.nElems + e) * 8 + (tid - 1)]; }
    else { const int64_t* po = cur ? bv.poolOff[1] : bv.poolOff[0]; s = po[e]; c = (int)(po[e + 1] - s); }
    R.start[tid] = s;
    R.cnt[tid] = c;
  }
}
__device__ __forceinline__ void prefix_ranges(ElemRanges& R) {   // one thread
  int p = 0;
#pragma unroll
  for (int r = 0; r < BIN_NRANGE; ++r) { R.pre[r] = p; p += (R.cnt[r] + 1) & ~1; }
  R.pre[BIN_NRANGE] = p;
}
// virtual index (even for pairs) -> range, slot, number of live particles from v on in that range
__device__ __forceinline__ void resolve(const ElemRanges& R, int v, int& r, int64_t& slot, int& left) {
  r = 0;
#pragma unroll
  for (int k = 1; k < BIN_NRANGE; ++k) r += (v >= R.pre[k]) ? 1 : 0;
  const int o = v - R.pre[r];
  slot = R.start[r] + o;
  left = R.cnt[r] - o;
}

// ---- layout construction -------------------------------------------------------------------------------------------------------
// capacities from the element populations (segment offsets of the sorted arrays): main = n + slack, side inbox = fraction of n
__global__ void k_bin_plan(const int64_t* __restrict__ elemOff, int nElems, double mainSlack, double inFrac, int32_t* __restrict__ capMain,
                           int32_t* __restrict__ capIn, int64_t* __restrict__ size) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nElems) return;
  const int64_t n = elemOff[e + 1] - elemOff[e];
  const int64_t cm = ((int64_t)((double)n * (1.0 + mainSlack)) + 64 + 31) & ~(int64_t)31;
  const int64_t ci = ((int64_t)((double)n * inFrac) + 24 + 31) & ~(int64_t)31;
  capMain[e] = (int32_t)cm;
  capIn[e] = (int32_t)ci;
  size[e] = cm + 12 * ci;
}

// exclusive scan of int64 values by one CTA; out has n + 1 entries (out[n] = total)
__global__ void __launch_bounds__(1024) k_scan_i64(const int64_t* __restrict__ in, int n, int64_t* __restrict__ out) {
  __shared__ int64_t sh[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int64_t v = i < n ? in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) out[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

// sorted arrays -> bins: the element's segment becomes its main range; inboxes and pool empty
__global__ void k_bin_fill(PartBuf src, const int64_t* __restrict__ elemOff, PartBuf bins, BinView bv) {
  for (int e = blockIdx.x; e < bv.nElems; e += gridDim.x) {
    const int64_t p0 = elemOff[e], n = elemOff[e + 1] - p0, b0 = bv.base[e];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
      for (int a = 0; a < 6; ++a) bins.f[a * bins.stride + b0 + i] = src.f[a * src.stride + p0 + i];
      bins.meta[b0 + i] = src.meta[p0 + i];
      if (bins.id) bins.id[b0 + i] = src.id[p0 + i];
    }
    if (threadIdx.x == 0) bv.nMain[e] = (int32_t)n;
    if (threadIdx.x < 16) bv.nIn[((size_t)(threadIdx.x >> 3) * bv.nElems + e) * 8 + (threadIdx.x & 7)] = 0;
  }
}

// particles per element (all eight ranges) -> cnt[e]
__global__ void k_bin_count(BinView bv, int cur, int64_t* __restrict__ cnt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= bv.nElems) return;
  int64_t n = bv.nMain[e];
  for (int s = 0; s < 6; ++s) n += bv.nIn[((size_t)cur * bv.nElems + e) * 8 + s];
  const int64_t* po = cur ? bv.poolOff[1] : bv.poolOff[0];
  n += po[e + 1] - po[e];
  cnt[e] = n;
}

// bins -> sorted arrays (element order, inside an element the fixed range order); elemOff from k_scan_i64 of k_bin_count
__global__ void __launch_bounds__(BIN_NT) k_bin_gather(PartBuf bins, PartBuf pool, BinView bv, int cur, const int64_t* __restrict__ elemOff,
                                                        int offsetElem, PartBuf dst) {
  __shared__ ElemRanges R;
  for (int e = blockIdx.x; e < bv.nElems; e += gridDim.x) {
    __syncthreads();
    load_ranges(R, bv, e, cur, threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {   // unpadded prefix: the sorted arrays are dense
      int p = 0;
      for (int r = 0; r < BIN_NRANGE; ++r) { R.pre[r] = p; p += R.cnt[r]; }
      R.pre[BIN_NRANGE] = p;
    }
    __syncthreads();
    const int total = R.pre[BIN_NRANGE];
    const int64_t d0 = elemOff[e];
    for (int v = threadIdx.x; v < total; v += BIN_NT) {
      int r;
      int64_t slot;
      int left;
      resolve(R, v, r, slot, left);
      const double* sf = (r == 7) ? pool.f : bins.f;
      const int64_t sst = (r == 7) ? pool.stride : bins.stride;
      const uint8_t* sm = (r == 7) ? pool.meta : bins.meta;
      const int64_t* sid = (r == 7) ? pool.id : bins.id;
#pragma unroll
      for (int a = 0; a < 6; ++a) dst.f[a * dst.stride + d0 + v] = sf[a * sst + slot];
      dst.meta[d0 + v] = sm[slot];
      dst.elem[d0 + v] = offsetElem + e + 1;
      if (dst.id) dst.id[d0 + v] = sid[slot];
    }
  }
}

// ---- cell_volweight_mean particle loop on the bins (DepositionMethod_CVWM, pic_depo_method.f90:471-544) -----------------------------
// Same arithmetic and the same fixed-order block reduction as k_deposit_cvwm (kernels.cuh); the particles of the element are the
// eight ranges in order.  No reference position is stored: the interpolation of the same step recomputes it (closed form on
// affine elements, the reference's second Newton call otherwise, pic_interpolation_tools.f90:241).
// general path of the deposition for one slot of the bins / pool arrays (no cached reference position)
__device__ __noinline__ void deposit_slot_cold(const double* __restrict__ f, int64_t stride, uint8_t meta, int64_t p, const GeoElem* sg,
                                               const double (*corner)[3], DepAcc* sAcc, int tid) {
  PartBuf pb;
  pb.f = const_cast<double*>(f);
  pb.xif = nullptr;
  pb.stride = stride;
  pb.meta = &meta - p;   // deposit_particle_general reads pb.meta[p] only (no cache to update without xif)
  deposit_particle_general(pb, p, sg, corner, *sAcc, tid);
}

template <bool FAST>
__global__ void __launch_bounds__(BIN_NT, DEP_MINB) k_bin_deposit_cvwm(PartBuf bins, PartBuf pool, BinView bv, int cur, int offsetElem,
                                                                       const GeoElem* __restrict__ geo, const TriaElem* __restrict__ tria,
                                                                       const AffElem* __restrict__ aff, double* __restrict__ elemAcc) {
  __shared__ ElemRanges R;
  __shared__ GeoElem sg;
  __shared__ AffElem sa;
  __shared__ double corner[8][3];
  __shared__ DepAcc sAcc;
  __shared__ double sP[2][6][BIN_NT];
  const int tid = threadIdx.x;
  for (int e = blockIdx.x; e < bv.nElems; e += gridDim.x) {
    __syncthreads();
    load_ranges(R, bv, e, cur, tid);
    if (FAST) stage_words(&sa, aff + (offsetElem + e), sizeof(AffElem));
    __syncthreads();
    if (tid == 0) prefix_ranges(R);
    __syncthreads();
    const int total = R.pre[BIN_NRANGE];
    const bool fastElem = FAST && sa.affine != 0.0;
    if (!fastElem && total > 0) {
      stage_words(&sg, geo + (offsetElem + e), sizeof(GeoElem));
      if (tid < 24) corner[tid / 3][tid % 3] = tria[offsetElem + e].corner[tid / 3][tid % 3];
      __syncthreads();
    }
    double acc[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) acc[a] = 0.;
    if (!fastElem) {
#pragma unroll
      for (int a = 0; a < 32; ++a) sAcc[a][tid] = 0.;
    }
    int nGeneral = 0;
    auto src_of = [&](int v, int& r, int64_t& slot, bool& live) {
      int left;
      resolve(R, v, r, slot, left);
      live = left > 0;
    };
    if (fastElem) {
      int v = tid, stage = 0;
      uint8_t meta = 0, metaNext = 0;
      bool live = false, liveNext = false;
      if (v < total) {
        int r;
        int64_t slot;
        src_of(v, r, slot, live);
        if (live) {
          const double* sf = (r == 7) ? pool.f : bins.f;
          const int64_t sst = (r == 7) ? pool.stride : bins.stride;
#pragma unroll
          for (int a = 0; a < 6; ++a) cp_async8(&sP[0][a][tid], sf + a * sst + slot);
          meta = ((r == 7) ? pool.meta : bins.meta)[slot];
        }
      }
      cp_async_commit();
      for (; v < total; v += BIN_NT, stage ^= 1, meta = metaNext, live = liveNext) {
        const int vn = v + BIN_NT;
        liveNext = false;
        if (vn < total) {
          int r;
          int64_t slot;
          src_of(vn, r, slot, liveNext);
          if (liveNext) {
            const double* sf = (r == 7) ? pool.f : bins.f;
            const int64_t sst = (r == 7) ? pool.stride : bins.stride;
#pragma unroll
            for (int a = 0; a < 6; ++a) cp_async8(&sP[stage ^ 1][a][tid], sf + a * sst + slot);
            metaNext = ((r == 7) ? pool.meta : bins.meta)[slot];
          }
        }
        cp_async_commit();
        cp_async_wait_prev();
        if (!live) continue;
        const double x[3] = {sP[stage][0][tid], sP[stage][1][tid], sP[stage][2][tid]};
        double xi[3];
        if (!affine_xi(&sa, x, xi)) { ++nGeneral; continue; }
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
    }
    if (!fastElem || __syncthreads_or(nGeneral)) {
      if (fastElem) {   // (never observed) particles far outside their affine element: general path in a second sweep
        stage_words(&sg, geo + (offsetElem + e), sizeof(GeoElem));
        if (tid < 24) corner[tid / 3][tid % 3] = tria[offsetElem + e].corner[tid / 3][tid % 3];
        __syncthreads();
      }
      for (int v = tid; v < total; v += BIN_NT) {
        int r;
        int64_t slot;
        bool live;
        src_of(v, r, slot, live);
        if (!live) continue;
        const double* sf = (r == 7) ? pool.f : bins.f;
        const int64_t sst = (r == 7) ? pool.stride : bins.stride;
        if (fastElem) {
          const double x[3] = {sf[slot], sf[sst + slot], sf[2 * sst + slot]};
          double xi[3];
          if (affine_xi(&sa, x, xi)) continue;
        }
        asm volatile("" ::: "memory");
        deposit_slot_cold(sf, sst, ((r == 7) ? pool.meta : bins.meta)[slot], slot, &sg, corner, &sAcc, tid);
      }
    }
    __syncthreads();
    {
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
      for (int a = warp * 8; a < warp * 8 + 8; ++a) {
        double v = ((sAcc[a][lane] + sAcc[a][lane + 32]) + sAcc[a][lane + 64]) + sAcc[a][lane + 96];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) elemAcc[(size_t)e * 32 + a] = v;
      }
    }
  }
}

// ---- interpolate + push + own-element inside test + delivery ------------------------------------------------------------------------
// field tile in shared memory in the host's order: sE[((k*NP + j)*NP + i)*3 + c].  Two particles per thread share every
// shared-memory operand (one LDS feeds two FMAs) and give the FP64 pipe two independent dependency chains per accumulator.
template <int NP>
__device__ __forceinline__ void evaluate_field_fast2(const double xi[2][3], const double* __restrict__ sE, double out[2][3]) {
  double L0[2][NP], L1[2][NP], L2[2][NP];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    lagrange_fast<NP>(xi[q][0], L0[q]);
    lagrange_fast<NP>(xi[q][1], L1[q]);
    lagrange_fast<NP>(xi[q][2], L2[q]);
  }
  double o[2][3] = {{0., 0., 0.}, {0., 0., 0.}};
#pragma unroll 1
  for (int k = 0; k < NP; ++k) {
    double lz[2] = {L2[0][0], L2[1][0]};
#pragma unroll
    for (int m = 1; m < NP; ++m) { lz[0] = (k == m) ? L2[0][m] : lz[0]; lz[1] = (k == m) ? L2[1][m] : lz[1]; }
    double s[2][3] = {{0., 0., 0.}, {0., 0., 0.}};
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double* row = sE + ((k * NP + j) * NP) * 3;
      double t[2][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double u = row[c];
        t[0][c] = u * L0[0][0];
        t[1][c] = u * L0[1][0];
      }
#pragma unroll
      for (int i = 1; i < NP; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double u = row[i * 3 + c];
          t[0][c] = fma(u, L0[0][i], t[0][c]);
          t[1][c] = fma(u, L0[1][i], t[1][c]);
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s[0][c] = fma(t[0][c], L1[0][j], s[0][c]);
        s[1][c] = fma(t[1][c], L1[1][j], s[1][c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[0][c] = fma(s[0][c], lz[0], o[0][c]);
      o[1][c] = fma(s[1][c], lz[1], o[1][c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { out[0][c] = o[0][c]; out[1][c] = o[1][c]; }
}

// one particle through the general path (non-affine element, failed closed form, reference-order arithmetic): field at the
// particle, push; returns the pushed state.  Out of line: keeps the registers of the hot path.
struct PushRes { double x0, x1, x2, v0, v1, v2; };
template <int NP, bool FAST>
__device__ __noinline__ PushRes push_general(double x0, double x1, double x2, double v0, double v1, double v2, int spec, int isNewIn,
                                             const double* __restrict__ sE, const GeoElem* __restrict__ ge, const AffElem* __restrict__ af,
                                             const double* __restrict__ Eg, const double* __restrict__ xgp, double dt) {
  double x[3] = {x0, x1, x2}, v[3] = {v0, v1, v2};
  bool isNew = isNewIn != 0;
  double F[6] = {0., 0., 0., 0., 0., 0.};
  const double q = cst.ChargeIC[spec];
  if (cst.DoInterpolation && fabs(q) > 0.0) {  // isInterpolateParticle
    double xi[3];
    bool suc;
    if (FAST) suc = ref_position_fast(af, ge, x, xi, false);
    else suc = (position_in_ref_elem(ge, x, xi, false, true) & 1) != 0;
    double f3[3];
    if (!suc && cst.DepositionType == PGPU_DEPO_CVWM) field_inverse_distance<NP>(x, Eg, xgp, f3);
    else if (FAST) {
      const double xi2[2][3] = {{xi[0], xi[1], xi[2]}, {xi[0], xi[1], xi[2]}};
      double o2[2][3];
      evaluate_field_fast2<NP>(xi2, sE, o2);
      f3[0] = o2[0][0]; f3[1] = o2[0][1]; f3[2] = o2[0][2];
    } else evaluate_field<NP>(xi, sE, f3);
#pragma unroll
    for (int c = 0; c < 6; ++c) F[c] = cst.externalField[c];
    F[0] = F[0] + f3[0]; F[1] = F[1] + f3[1]; F[2] = F[2] + f3[2];
    F[3] = F[3] + 0.; F[4] = F[4] + 0.; F[5] = F[5] + 0.;
  }
  if (FAST) push_particle_fast(x, v, F, spec, isNew, dt);
  else push_particle(x, v, F, spec, isNew, dt);
  PushRes r;
  r.x0 = x[0]; r.x1 = x[1]; r.x2 = x[2]; r.v0 = v[0]; r.v1 = v[1]; r.v2 = v[2];
  return r;
}

// inside test of the own element for the general path: 0 = inside, 1 = left (far list)
template <bool FAST>
__device__ __noinline__ int inside_general(const PlaneElem* __restrict__ pl, const TriaElem* __restrict__ te, double x0, double x1, double x2) {
  const double x[3] = {x0, x1, x2};
  uint32_t mask;
  const bool in = FAST ? inside_fast<true>(pl, te, x, mask) : inside_quad3d_mask<true>(te, x, mask);
  return in ? 0 : 1;
}

constexpr int CAT_STAY = 0, CAT_FAR = 7, CAT_NONE = 8;

template <int NP, bool FAST>
__global__ void __launch_bounds__(BIN_NT, KB_MINB) k_bin_push(PartBuf bins, PartBuf pool, BinView bv, int cur, FarBuf far,
                                                              const PushElem* __restrict__ pushElem, const double* __restrict__ E,
                                                              const GeoElem* __restrict__ geo, const TriaElem* __restrict__ tria,
                                                              const PlaneElem* __restrict__ planes, const AffElem* __restrict__ aff,
                                                              const double* __restrict__ Elem_xGP, int offsetElem, double dt,
                                                              int* __restrict__ counters /*[2] far records, [4] side movers, [5] main full, [6] inbox full*/) {
  constexpr int ND = NP * NP * NP;
  // field tile: double buffered and prefetched with a bulk copy when it is small and a multiple of 16 bytes (N = 1, 3, 5);
  // otherwise one buffer, staged with plain loads after the element's barrier
  constexpr uint32_t E_BYTES = ND * 24;
  constexpr bool TMA_E = (E_BYTES % 16 == 0) && (E_BYTES <= 4096);
  constexpr int EBUF = TMA_E ? 2 : 1;
  __shared__ __align__(16) PushElem sPE[2];
  __shared__ __align__(16) double sEb[EBUF][ND * 3];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ __align__(16) double sP[2][6][BIN_CHUNK];   // x, v of the current / next pair of every thread
  __shared__ ElemRanges R;
  __shared__ int sCnt[2][BIN_NT / 32][8];
  __shared__ int sRun[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int next = cur ^ 1;
  const int nElems = bv.nElems;
  double* __restrict__ const BF = bins.f;
  const int64_t BS = bins.stride;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto prefetch_elem = [&](int e, int b) {   // one thread: records of element e -> buffer b
    mbar_expect_tx(&mbar[b], (uint32_t)sizeof(PushElem) + (TMA_E ? E_BYTES : 0u));
    bulk_g2s(&sPE[b], pushElem + e, (uint32_t)sizeof(PushElem), &mbar[b]);
    if (TMA_E) bulk_g2s(&sEb[TMA_E ? b : 0][0], E + (size_t)e * ND * 3, E_BYTES, &mbar[b]);
  };
  if (tid == 0 && (int)blockIdx.x < nElems) prefetch_elem(blockIdx.x, 0);
  int it = 0;
  for (int e = blockIdx.x; e < nElems; e += gridDim.x, ++it) {
    const int b = it & 1;
    const int gElem = offsetElem + e + 1;
    // the barrier that ended the previous element also freed buffer b^1: prefetch the records of the next element into it
    if (tid == 0 && e + (int)gridDim.x < nElems) prefetch_elem(e + gridDim.x, b ^ 1);
    load_ranges(R, bv, e, cur, tid);
    if (tid < 8) sRun[0][tid] = 0;
    __syncthreads();
    if (tid == 0) prefix_ranges(R);
    mbar_wait(&mbar[b], (uint32_t)((it >> 1) & 1));
    if (!TMA_E) {
      for (int t = tid; t < ND * 3; t += BIN_NT) sEb[0][t] = __ldg(E + (size_t)e * ND * 3 + t);
    }
    __syncthreads();
    const PushElem& pe = sPE[b];
    const double* sE = sEb[TMA_E ? b : 0];
    const int total = R.pre[BIN_NRANGE];
    const bool fastElem = FAST && pe.affine != 0u;
    const bool planarElem = FAST && pe.planar != 0u;
    const int nChunks = (total + BIN_CHUNK - 1) / BIN_CHUNK;
    const int64_t base_e = R.start[0];
    const int capMain = bv.capMain[e];
    // pair of this thread in chunk c: virtual indices c*BIN_CHUNK + 2*tid, +1.  slot < 0: no live particle
    auto fetch = [&](int c, int stg, int& nLive, int64_t& slot, bool& fromPool, uint32_t& meta2) {
      const int v = c * BIN_CHUNK + 2 * tid;
      nLive = 0;
      slot = -1;
      fromPool = false;
      meta2 = 0;
      if (v < total) {
        int r, left;
        resolve(R, v, r, slot, left);
        fromPool = r == 7;
        const double* sf = fromPool ? pool.f : BF;
        const int64_t sst = fromPool ? pool.stride : BS;
        const uint8_t* sm = fromPool ? pool.meta : bins.meta;
        nLive = left >= 2 ? 2 : (left > 0 ? 1 : 0);
        if (nLive > 0) {
          if ((slot & 1) == 0) {
#pragma unroll
            for (int a = 0; a < 6; ++a) cp_async16(&sP[stg][a][2 * tid], sf + a * sst + slot);
          } else {   // pool ranges start at any slot
#pragma unroll
            for (int a = 0; a < 6; ++a) {
              cp_async8(&sP[stg][a][2 * tid], sf + a * sst + slot);
              cp_async8(&sP[stg][a][2 * tid + 1], sf + a * sst + slot + 1);
            }
          }
          meta2 = (uint32_t)sm[slot] | (nLive > 1 ? ((uint32_t)sm[slot + 1] << 8) : ((uint32_t)sm[slot] << 8));
        }
      }
    };
    int nLive = 0, nLiveNext = 0;
    int64_t slotCur = -1, slotNext = -1;
    bool poolCur = false, poolNext = false;
    uint32_t meta2 = 0, meta2Next = 0;
    if (nChunks > 0) fetch(0, 0, nLive, slotCur, poolCur, meta2);
    cp_async_commit();
    for (int c = 0; c < nChunks; ++c) {
      const int stg = c & 1, cb = c & 1;
      nLiveNext = 0;
      if (c + 1 < nChunks) fetch(c + 1, stg ^ 1, nLiveNext, slotNext, poolNext, meta2Next);
      cp_async_commit();
      cp_async_wait_prev();
      // ---- per-particle work ----------------------------------------------------------------------------------------------
      const int pq = (nLive == 1) ? 0 : 1;   // the idle half of a pair computes on a copy of the live one (finite inputs, discarded)
      const int metaQ[2] = {(int)(meta2 & 0xffu), (int)((meta2 >> 8) & 0xffu)};
      double xn[2][3], vn[2][3];
      int cat[2] = {CAT_NONE, CAT_NONE};
      if (nLive > 0) {
        bool simple = fastElem;
        double f2[2][3];
        if (fastElem) {
          double xi[2][3];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int qs = q ? pq : 0;
            const double r0 = sP[stg][0][2 * tid + qs] - pe.x0[0], r1 = sP[stg][1][2 * tid + qs] - pe.x0[1], r2 = sP[stg][2][2 * tid + qs] - pe.x0[2];
#pragma unroll
            for (int d = 0; d < 3; ++d) xi[q][d] = fma(pe.A[d][0], r0, fma(pe.A[d][1], r1, pe.A[d][2] * r2)) - 1.0;
            simple = simple && fabs(xi[q][0]) <= 1.5 && fabs(xi[q][1]) <= 1.5 && fabs(xi[q][2]) <= 1.5;
          }
          if (simple) evaluate_field_fast2<NP>(xi, sE, f2);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int qs = q ? pq : 0;
#pragma unroll
          for (int d = 0; d < 3; ++d) { xn[q][d] = sP[stg][d][2 * tid + qs]; vn[q][d] = sP[stg][3 + d][2 * tid + qs]; }
        }
        if (simple) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int spec = metaQ[q] & META_SPEC_MASK;
            bool isNew = (metaQ[q] & META_ISNEW) != 0;
            double F[6] = {0., 0., 0., 0., 0., 0.};
            const double qc = cst.ChargeIC[spec];
            if (cst.DoInterpolation && fabs(qc) > 0.0) {
#pragma unroll
              for (int d = 0; d < 6; ++d) F[d] = cst.externalField[d];
              F[0] = F[0] + f2[q][0]; F[1] = F[1] + f2[q][1]; F[2] = F[2] + f2[q][2];
            }
            push_particle_fast(xn[q], vn[q], F, spec, isNew, dt);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const PushRes r = push_general<NP, FAST>(xn[q][0], xn[q][1], xn[q][2], vn[q][0], vn[q][1], vn[q][2], metaQ[q] & META_SPEC_MASK,
                                                     (metaQ[q] & META_ISNEW) ? 1 : 0, sE, geo + (gElem - 1), FAST ? aff + (gElem - 1) : nullptr,
                                                     E + (size_t)e * ND * 3, Elem_xGP + (size_t)(gElem - 1) * ND * 3, dt);
            xn[q][0] = r.x0; xn[q][1] = r.x1; xn[q][2] = r.x2; vn[q][0] = r.v0; vn[q][1] = r.v1; vn[q][2] = r.v2;
          }
        }
        // ---- own-element inside test (first iteration of SingleParticleTriaTracking3D, particle_triatracking.f90:203-218) and,
        //      for planar convex elements, the crossing of exactly one side plane into a face neighbour ------------------------------
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q >= nLive) { cat[q] = CAT_NONE; continue; }
          if (!planarElem) {
            cat[q] = inside_general<FAST>(FAST ? planes + (gElem - 1) : nullptr, tria + (gElem - 1), xn[q][0], xn[q][1], xn[q][2]) ? CAT_FAR : CAT_STAY;
            continue;
          }
          double dx[6];
          uint32_t neg = 0;
          bool ambiguous = false;
          const double tol = pe.tol;
#pragma unroll
          for (int s = 0; s < 6; ++s) {
            dx[s] = fma(pe.pl[s][0], xn[q][0], fma(pe.pl[s][1], xn[q][1], fma(pe.pl[s][2], xn[q][2], -pe.pl[s][3])));
            ambiguous |= fabs(dx[s]) <= tol;
            neg |= (dx[s] < 0.) ? (1u << s) : 0u;
          }
          if (ambiguous) {   // within tol of a side plane: the determinants decide (ParticleInsideQuad3D), leavers take the exact walk
            const uint32_t r = inside_exact_cold<true>(tria + (gElem - 1), xn[q][0], xn[q][1], xn[q][2]);
            cat[q] = (r >> 31) ? CAT_STAY : CAT_FAR;
            continue;
          }
          if (neg == 0u) { cat[q] = CAT_STAY; continue; }
          cat[q] = CAT_FAR;
          if (__popc(neg) != 1) continue;
          const int s = __ffs(neg) - 1;
          if (pe.nbLocal[s] < 0) continue;
          // flight LastPartPos -> x crosses side s at lp + alpha (x - lp); every decision with a margin of tol, otherwise the
          // determinant tests of the exact walk decide
          const double lp0 = sP[stg][0][2 * tid + q], lp1 = sP[stg][1][2 * tid + q], lp2 = sP[stg][2][2 * tid + q];
          const double dl = fma(pe.pl[s][0], lp0, fma(pe.pl[s][1], lp1, fma(pe.pl[s][2], lp2, -pe.pl[s][3])));
          if (!(dl > tol)) continue;
          const double alpha = dl / (dl - dx[s]);
          bool ok = true;
#pragma unroll
          for (int o = 0; o < 6; ++o) {
            const double ol = fma(pe.pl[o][0], lp0, fma(pe.pl[o][1], lp1, fma(pe.pl[o][2], lp2, -pe.pl[o][3])));
            const double oc = fma(alpha, dx[o] - ol, ol);
            if (o != s && !(oc > tol)) ok = false;
          }
          {
            const double gl = fma(pe.dg[s][0], lp0, fma(pe.dg[s][1], lp1, fma(pe.dg[s][2], lp2, -pe.dg[s][3])));
            const double gx = fma(pe.dg[s][0], xn[q][0], fma(pe.dg[s][1], xn[q][1], fma(pe.dg[s][2], xn[q][2], -pe.dg[s][3])));
            const double gc = fma(alpha, gx - gl, gl);
            if (!(fabs(gc) > tol)) ok = false;   // crossing point on the triangle diagonal
          }
          // clearly inside the neighbour: ParticleInsideQuad3D there succeeds, the walk ends (:215-218)
          const double ntol = pe.nbtol[s];
#pragma unroll
          for (int o = 0; o < 6; ++o) {
            const double dn = fma(pe.nbpl[s][o][0], xn[q][0], fma(pe.nbpl[s][o][1], xn[q][1], fma(pe.nbpl[s][o][2], xn[q][2], -pe.nbpl[s][o][3])));
            if (!(dn > ntol)) ok = false;
          }
          if (ok) cat[q] = 1 + s;
        }
      }
      // particle ids (tests) have to be read before the compaction below overwrites the slots of this chunk
      int64_t id[2] = {0, 0};
      if (bins.id && nLive > 0) {
        const int64_t* sid = poolCur ? pool.id : bins.id;
        id[0] = sid[slotCur];
        if (nLive > 1) id[1] = sid[slotCur + 1];
      }
      // ---- stable compaction: rank of every particle within its category in the order of the element's particles (virtual
      //      index 2 * tid + q), chunk by chunk: the order inside an element never changes except by departures and arrivals ----------
      int rank[2] = {0, 0};
      {
        int myCount = 0;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const unsigned bal0 = __ballot_sync(0xffffffffu, cat[0] == k);
          const unsigned bal1 = __ballot_sync(0xffffffffu, cat[1] == k);
          const int below = __popc(bal0 & lt) + __popc(bal1 & lt);
          if (cat[0] == k) rank[0] = below;
          if (cat[1] == k) rank[1] = below + (cat[0] == k ? 1 : 0);
          if (lane == k) myCount = __popc(bal0) + __popc(bal1);
        }
        if (lane < 8) sCnt[cb][warp][lane] = myCount;
      }
      __syncthreads();
      int off[2] = {0, 0};
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (cat[q] >= 8) continue;
        int o = sRun[cb][cat[q]] + rank[q];
#pragma unroll
        for (int w = 0; w < BIN_NT / 32; ++w)
          if (w < warp) o += sCnt[cb][w][cat[q]];
        off[q] = o;
      }
      if (tid < 8) {
        int t = sRun[cb][tid];
#pragma unroll
        for (int w = 0; w < BIN_NT / 32; ++w) t += sCnt[cb][w][tid];
        sRun[cb ^ 1][tid] = t;
      }
      // ---- delivery ----------------------------------------------------------------------------------------------------------------
      const uint32_t srcTag = (uint32_t)((poolCur ? BS : 0) + slotCur);   // origin: slot in the bins arrays, pool slots behind them
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        bool toFar = cat[q] == CAT_FAR;
        int fullKind = 0;
        const uint8_t nmeta = (uint8_t)(metaQ[q] & META_SPEC_MASK);   // IsNewPart is consumed by the push
        int64_t slot = -1;
        if (cat[q] == CAT_STAY) {
          if (off[q] < capMain) slot = base_e + off[q];
          else { toFar = true; fullKind = 5; }   // main full: through the far list back into this element
        } else if (cat[q] >= 1 && cat[q] <= 6) {
          const int s = cat[q] - 1;
          const int nb = pe.nbLocal[s];
          if (off[q] < bv.capIn[nb]) slot = bin_inbox_base(bv, nb, next, pe.nbBox[s]) + off[q];
          else { toFar = true; fullKind = 6; }   // inbox full
        }
        if (slot >= 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) { BF[d * BS + slot] = xn[q][d]; BF[(3 + d) * BS + slot] = vn[q][d]; }
          bins.meta[slot] = nmeta;
          if (bins.id) bins.id[slot] = id[q];
        }
        if (fullKind) atomicAdd(&counters[fullKind], 1);
        const unsigned fm = __ballot_sync(0xffffffffu, toFar);
        if (fm) {
          int slot0 = 0;
          const int leader = __ffs(fm) - 1;
          if (lane == leader) slot0 = atomicAdd(&counters[2], __popc(fm));
          slot0 = __shfl_sync(0xffffffffu, slot0, leader);
          if (toFar) {
            const int f = slot0 + __popc(fm & ((1u << lane) - 1u));
#pragma unroll
            for (int d = 0; d < 3; ++d) { far.x[d][f] = xn[q][d]; far.lp[d][f] = sP[stg][d][2 * tid + q]; far.v[d][f] = vn[q][d]; }
            far.elem[f] = gElem;
            far.meta[f] = nmeta;
            if (far.id) far.id[f] = id[q];
            far.src[f] = srcTag + (uint32_t)q;
          }
        }
      }
      nLive = nLiveNext;
      slotCur = slotNext;
      poolCur = poolNext;
      meta2 = meta2Next;
    }
    __syncthreads();
    // populations after the step: this element's main, and the inboxes this element fills at its face neighbours
    {
      const int fb = nChunks & 1;
      if (tid == 0) {
        const int n = sRun[fb][0];
        bv.nMain[e] = n < capMain ? n : capMain;
      } else if (tid >= 1 && tid <= 6) {
        const int s = tid - 1, nb = pe.nbLocal[s];
        if (nb >= 0) {
          const int n = sRun[fb][tid], ci = bv.capIn[nb];
          const int m = n < ci ? n : ci;
          bv.nIn[((size_t)next * nElems + nb) * 8 + pe.nbBox[s]] = m;
          if (m > 0) atomicAdd(&counters[4], m);
        }
      }
    }
    __syncthreads();
  }
}

// ---- SingleParticleTriaTracking3D for the far list (particle_triatracking.f90:137-484), from the start -----------------------------
// One thread per record, persistent warps with per-lane refill (as k_track_leavers).  key: local element, nElems + rank
// (emigrant) or nElems + nRanks (removed).
template <bool FAST>
__global__ void __launch_bounds__(LV_NT, LV_MINB) k_far_walk(FarBuf far, int nFar, const TriaElem* __restrict__ tria,
                                                             const PlaneElem* __restrict__ planes, const int32_t* __restrict__ elemRank,
                                                             uint32_t* __restrict__ keys, int nElems, int offsetElem, int* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  bool active = false;
  int p = 0, ElemID = 0, guard = 0;
  uint32_t mask = 0;
  double x[3] = {0., 0., 0.}, lp[3] = {0., 0., 0.};
  HopHist h;
  h.clear();
  while (true) {
    int status = -1;
    {
      const unsigned need = __ballot_sync(0xffffffffu, !active);
      if (need) {
        int first = 0;
        if (lane == __ffs(need) - 1) first = atomicAdd(&counters[3], __popc(need));
        first = __shfl_sync(0xffffffffu, first, __ffs(need) - 1);
        const int mine = first + __popc(need & ((1u << lane) - 1u));
        if (!active && mine < nFar) {
          active = true;
          p = mine;
          x[0] = far.x[0][p]; x[1] = far.x[1][p]; x[2] = far.x[2][p];
          lp[0] = far.lp[0][p]; lp[1] = far.lp[1][p]; lp[2] = far.lp[2][p];
          ElemID = far.elem[p];
          guard = 0;
          h.clear();
          // 2a) of the first loop iteration: is the particle in the element it started in? (:203-218)
          const bool in = FAST ? inside_fast<true>(planes + (ElemID - 1), tria + (ElemID - 1), x, mask)
                               : inside_quad3d_mask<true>(tria + (ElemID - 1), x, mask);
          if (in) status = TRK_OK;
        }
      }
      if (__ballot_sync(0xffffffffu, active) == 0) break;
    }
    if (active) {
      if (status == -1) {
        status = tria_hop<FAST, true, FAST ? 2 : 0>(
            tria + (ElemID - 1), tria, FAST ? planes + (ElemID - 1) : nullptr, [&](int, int ne) { return planes + (ne - 1); },
            [&](const double n[3]) {
              const double v0 = far.v[0][p], v1 = far.v[1][p], v2 = far.v[2][p];
              const double vn = (v0 * n[0] + v1 * n[1]) + v2 * n[2];
              far.v[0][p] = v0 - 2. * vn * n[0]; far.v[1][p] = v1 - 2. * vn * n[1]; far.v[2][p] = v2 - 2. * vn * n[2];
            },
            x, lp, ElemID, mask, h);
        if (status == -1 && ++guard > 100000) status = TRK_ERR_LOOP;
      }
      if (status != -1) {
        uint32_t key;
        int newElem = ElemID;
        if (status == TRK_OK) {
          const int rk = (cst.nRanks == 1) ? cst.myRank : elemRank[newElem - 1];
          key = (rk == cst.myRank) ? (uint32_t)(newElem - 1 - offsetElem) : (uint32_t)(nElems + rk);
        } else {
          key = (uint32_t)(nElems + cst.nRanks);
          newElem = 0;
          if (status == TRK_LOST) atomicAdd(&counters[0], 1);
          else if (status != TRK_REMOVED) atomicMax(&counters[1], status);
        }
        far.x[0][p] = x[0]; far.x[1][p] = x[1]; far.x[2][p] = x[2];
        far.elem[p] = newElem;
        keys[p] = key;
        active = false;
      }
    }
  }
}

// sorted far list -> pool: pool slot i takes record perm[i] (i < number of records that stay on this rank)
__global__ void k_far_to_pool(FarBuf far, const uint32_t* __restrict__ perm, int64_t n, PartBuf pool) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = perm[i];
#pragma unroll
  for (int d = 0; d < 3; ++d) { pool.f[d * pool.stride + i] = far.x[d][s]; pool.f[(3 + d) * pool.stride + i] = far.v[d][s]; }
  pool.meta[i] = far.meta[s];
  if (pool.id) pool.id[i] = far.id[s];
}

// origin tags of received particles: behind every local slot, in arrival order
__global__ void k_far_tag_immigrants(uint32_t* __restrict__ src, int64_t n0, int64_t n, uint32_t first) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) src[n0 + i] = first + (uint32_t)i;
}

// out[i] = in[perm[i]] for 32-bit keys (second, stable pass of the far sort works on the keys in origin order)
__global__ void k_gather_u32(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm, int64_t n, uint32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[perm[i]];
}
// perm2[i] = perm1[perm2[i]]
__global__ void k_compose_perm(const uint32_t* __restrict__ perm1, uint32_t* __restrict__ perm2, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) perm2[i] = perm1[perm2[i]];
}
