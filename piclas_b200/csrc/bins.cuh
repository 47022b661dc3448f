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
#include "hint.cuh"

constexpr int BIN_NT = 128;          // threads per CTA of the per-element kernels
constexpr int BIN_PPT = 2;           // particles per thread and sweep (one 128-bit copy per array and thread)
constexpr int BIN_CHUNK = BIN_NT * BIN_PPT;
constexpr int BIN_NRANGE = 8;        // main, six side inboxes, pool
#ifndef MONO_UNROLL_J
#define MONO_UNROLL_J 4      // rows of the field tile in flight per thread in the Horner evaluation
#endif
constexpr int MONO_UJ = MONO_UNROLL_J;
// k_bin_push: threads per CTA and resident CTAs per SM (register cap 65536 / (KB_NT * KB_MINB) = 168).  One warp per CTA: the
// stable compaction of every sweep is a warp scan, no block barrier, no shared-memory exchange of counts; the per-element
// prologue (ranges, records, inbox bases) is executed by one warp instead of four; 64-particle sweeps waste less of the last one
#ifndef KB_NT
#define KB_NT 32
#endif
#ifndef KB_UNROLL_Q
#define KB_UNROLL_Q 1     // 1: push + inside test of the thread's two particles interleaved
#endif
#ifndef KB_MINB
#define KB_MINB (384 / KB_NT)
#endif
constexpr int KB_CHUNK = KB_NT * BIN_PPT;
#if KB_NT == 32
#define KB_SYNC() __syncwarp()
#else
#define KB_SYNC() __syncthreads()
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

// far list: particles that are not delivered by the push kernel itself.  Aliases the (idle) sorted buffers.  Element e owns the
// slots [farBase[e], farBase[e+1]) (farBase = prefix sum of the element populations before the step: a region can never
// overflow): its far records in the order of its particles from the bottom (no atomics, deterministic), the rare particles
// diverted by a full main / inbox region from the top.  k_far_index then lists the used slots densely.
struct FarBuf {
  double* x[3];    // pushed position (in), final position (out: periodic shifts, reflections)
  double* lp[3];   // LastPartPos
  double* v[3];
  int32_t* elem;   // in: global element the walk starts in; out: final global element (0 = removed)
  uint8_t* meta;
  int64_t* id;     // optional
};


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

// ranges of element e by lanes 0..7 of one warp: loads, padded prefix by shuffles, result into shared memory
struct RangeRegs { int64_t start; int cnt; };
__device__ __forceinline__ RangeRegs range_load(const BinView& bv, int e, int cur, int lane) {
  RangeRegs r;
  r.start = 0;
  r.cnt = 0;
  if (lane < BIN_NRANGE) {
    if (lane == 0) { r.start = bv.base[e]; r.cnt = bv.nMain[e]; }
    else if (lane < 7) { r.start = bin_inbox_base(bv, e, cur, lane - 1); r.cnt = bv.nIn[((size_t)cur * bv.nElems + e) * 8 + (lane - 1)]; }
    else { const int64_t* po = cur ? bv.poolOff[1] : bv.poolOff[0]; r.start = po[e]; r.cnt = (int)(po[e + 1] - r.start); }
  }
  return r;
}
__device__ __forceinline__ void range_store(ElemRanges& R, const RangeRegs& r, int lane) {   // whole warp calls
  const int padded = (r.cnt + 1) & ~1;
  int incl = padded;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane < BIN_NRANGE) {
    R.start[lane] = r.start;
    R.cnt[lane] = r.cnt;
    R.pre[lane] = incl - padded;
    if (lane == BIN_NRANGE - 1) R.pre[BIN_NRANGE] = incl;
  }
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

// exclusive scan of up to 2^20 int64 values in three launches (tile sums, scan of the sums, tile scans); out[n] = total
constexpr int SCAN_TILE = 1024;
__global__ void __launch_bounds__(256) k_scan_tile_sums(const int64_t* __restrict__ in, int n, int64_t* __restrict__ sums) {
  __shared__ int64_t ws[8];
  const int base = blockIdx.x * SCAN_TILE;
  int64_t s = 0;
  for (int i = threadIdx.x; i < SCAN_TILE; i += 256) s += (base + i < n) ? in[base + i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    sums[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_scan_tiles(const int64_t* __restrict__ in, int n, const int64_t* __restrict__ tileOff /*scanned sums*/,
                                                    int64_t* __restrict__ out) {
  __shared__ int64_t ws[8];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
  int64_t v[4], t = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; t += v[k]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t incl = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  int64_t off = tileOff[blockIdx.x] + incl - t;
  for (int w = 0; w < warp; ++w) off += ws[w];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n) out[base + k] = off;
    off += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) out[n] = tileOff[gridDim.x];
}

// dense list of the used far slots: element by element, bottom part (far records) then top part (diverted particles)
__global__ void k_far_index(const int64_t* __restrict__ farBase, const int32_t* __restrict__ nFarE /*[nElems][2]*/, const int64_t* __restrict__ dOff,
                            int nElems, uint32_t* __restrict__ idx) {
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int nb = nFarE[2 * e], nt = nFarE[2 * e + 1];
    const int64_t b0 = farBase[e], b1 = farBase[e + 1], d0 = dOff[e];
    for (int i = threadIdx.x; i < nb + nt; i += blockDim.x) idx[d0 + i] = (uint32_t)(i < nb ? b0 + i : b1 - 1 - (i - nb));
  }
}
__global__ void k_far_total(const int32_t* __restrict__ nFarE, int nElems, int64_t* __restrict__ tot) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nElems) tot[e] = (int64_t)nFarE[2 * e] + nFarE[2 * e + 1];
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
constexpr int DB_NT = 32;   // one warp per CTA: no block barriers, the 32 sums of an element leave the warp by a shuffle reduce-scatter
typedef double DepAccW[32][DB_NT];
__device__ __noinline__ void deposit_slot_cold(const double* __restrict__ f, int64_t stride, uint8_t meta, int64_t p, const GeoElem* sg,
                                               const double (*corner)[3], DepAccW* sAcc, int tid) {
  PartBuf pb;
  pb.f = const_cast<double*>(f);
  pb.xif = nullptr;
  pb.stride = stride;
  pb.meta = &meta - p;   // deposit_particle_general reads pb.meta[p] only (no cache to update without xif)
  deposit_particle_general(pb, p, sg, corner, *sAcc, tid);
}

// ---- cell_volweight_mean particle loop on the bins (DepositionMethod_CVWM, pic_depo_method.f90:471-544) -----------------------------
// Same arithmetic as k_deposit_cvwm (kernels.cuh); the particles of the element are the eight ranges in order, two particles per
// thread and sweep on 128-bit copies.  The 32 element sums (8 nodes x 4 components) are reduced over the warp in a fixed
// pattern (deterministic).  No reference position is stored: the interpolation of the same step recomputes it (closed form on
// affine elements, the reference's second Newton call otherwise, pic_interpolation_tools.f90:241).
template <bool FAST>
__global__ void __launch_bounds__(DB_NT, 16) k_bin_deposit_cvwm(PartBuf bins, PartBuf pool, BinView bv, int cur, int offsetElem,
                                                                const GeoElem* __restrict__ geo, const TriaElem* __restrict__ tria,
                                                                const AffElem* __restrict__ aff, double* __restrict__ elemAcc) {
  __shared__ ElemRanges R;
  __shared__ GeoElem sg;
  __shared__ AffElem sa;
  __shared__ double corner[8][3];
  // the accumulators of the general path and the staging slots of the fast path are never live at the same time: one buffer
  typedef double StageBuf[2][6][2 * DB_NT];
  __shared__ __align__(16) unsigned char sBuf[sizeof(DepAccW) > sizeof(StageBuf) ? sizeof(DepAccW) : sizeof(StageBuf)];
  DepAccW& sAcc = *reinterpret_cast<DepAccW*>(sBuf);
  StageBuf& sP = *reinterpret_cast<StageBuf*>(sBuf);
  __shared__ uint32_t sM[2][DB_NT][2];
  const int tid = threadIdx.x, lane = tid;
  for (int e = blockIdx.x; e < bv.nElems; e += gridDim.x) {
    __syncwarp();
    range_store(R, range_load(bv, e, cur, lane), lane);
    if (FAST) stage_words(&sa, aff + (offsetElem + e), sizeof(AffElem));
    __syncwarp();
    const int total = R.pre[BIN_NRANGE];
    const bool fastElem = FAST && sa.affine != 0.0;
    double acc[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) acc[a] = 0.;
    int nGeneral = 0;
    if (fastElem) {
      int stage = 0, nLive = 0, nLiveNext = 0, sh = 0, shNext = 0;
      auto fetch2 = [&](int v, int stg, int& nl, int& shifts) {
        nl = 0;
        shifts = 0;
        if (v < total) {
          int r, left;
          int64_t slot;
          resolve(R, v, r, slot, left);
          nl = left >= 2 ? 2 : (left > 0 ? 1 : 0);
          if (nl > 0) {
            const double* sf = (r == 7) ? pool.f : bins.f;
            const int64_t sst = (r == 7) ? pool.stride : bins.stride;
            const uint8_t* sm = (r == 7) ? pool.meta : bins.meta;
            if ((slot & 1) == 0) {
#pragma unroll
              for (int a = 0; a < 6; ++a) cp_async16(&sP[stg][a][2 * tid], sf + a * sst + slot);
            } else {
#pragma unroll
              for (int a = 0; a < 6; ++a) {
                cp_async8(&sP[stg][a][2 * tid], sf + a * sst + slot);
                cp_async8(&sP[stg][a][2 * tid + 1], sf + a * sst + slot + 1);
              }
            }
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&sM[stg][tid][0])), "l"(sm + (slot & ~(int64_t)3)) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&sM[stg][tid][1])), "l"(sm + ((slot + 1) & ~(int64_t)3)) : "memory");
            shifts = (int)(8 * (slot & 3)) | ((int)(8 * ((slot + 1) & 3)) << 8);
          }
        }
      };
      fetch2(2 * tid, 0, nLive, sh);
      cp_async_commit();
      for (int v = 2 * tid; v < total; v += 2 * DB_NT, stage ^= 1, nLive = nLiveNext, sh = shNext) {
        fetch2(v + 2 * DB_NT, stage ^ 1, nLiveNext, shNext);
        cp_async_commit();
        cp_async_wait_prev();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q >= nLive) continue;
          const double x[3] = {sP[stage][0][2 * tid + q], sP[stage][1][2 * tid + q], sP[stage][2][2 * tid + q]};
          double xi[3];
          if (!affine_xi(&sa, x, xi)) { ++nGeneral; continue; }
          const int spec = (int)((sM[stage][tid][q] >> ((sh >> (8 * q)) & 0xff)) & META_SPEC_MASK);
          const double qc = cst.ChargeIC[spec];
          if (!(fabs(qc) > 0.0)) continue;  // isDepositParticle
          const double Charge = qc * cst.MPF[spec];
          const double T[4] = {sP[stage][3][2 * tid + q] * Charge, sP[stage][4][2 * tid + q] * Charge, sP[stage][5][2 * tid + q] * Charge, Charge};
          const double a1 = 0.5 * (xi[0] + 1.0), a2 = 0.5 * (xi[1] + 1.0), a3 = 0.5 * (xi[2] + 1.0);
          const double b1 = 1 - a1, b2 = 1 - a2, b3 = 1 - a3;
          const double w[8] = {(b1 * b2) * b3, (a1 * b2) * b3, (a1 * a2) * b3, (b1 * a2) * b3,
                               (b1 * b2) * a3, (a1 * b2) * a3, (a1 * a2) * a3, (b1 * a2) * a3};
#pragma unroll
          for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[n * 4 + c] = fma(T[c], w[n], acc[n * 4 + c]);
        }
      }
    }
    if (!fastElem || __any_sync(0xffffffffu, nGeneral != 0)) {
      // general path (Newton; inverse-distance fallback): accumulators in shared memory so that the iteration keeps the registers
      __syncwarp();   // every lane is done with its staging slots before the buffer takes the accumulators
#pragma unroll
      for (int a = 0; a < 32; ++a) sAcc[a][tid] = acc[a];
      if (total > 0) {
        stage_words(&sg, geo + (offsetElem + e), sizeof(GeoElem));
        if (tid < 24) corner[tid / 3][tid % 3] = tria[offsetElem + e].corner[tid / 3][tid % 3];
      }
      __syncwarp();
      for (int v = tid; v < total; v += DB_NT) {
        int r, left;
        int64_t slot;
        resolve(R, v, r, slot, left);
        if (left <= 0) continue;
        const double* sf = (r == 7) ? pool.f : bins.f;
        const int64_t sst = (r == 7) ? pool.stride : bins.stride;
        if (fastElem) {   // (never observed) particles far outside their affine element only
          const double x[3] = {sf[slot], sf[sst + slot], sf[2 * sst + slot]};
          double xi[3];
          if (affine_xi(&sa, x, xi)) continue;
        }
        asm volatile("" ::: "memory");
        deposit_slot_cold(sf, sst, ((r == 7) ? pool.meta : bins.meta)[slot], slot, &sg, corner, &sAcc, tid);
      }
      __syncwarp();
#pragma unroll
      for (int a = 0; a < 32; ++a) acc[a] = sAcc[a][tid];
    }
    // reduce-scatter over the warp: after the step with offset o every lane keeps the half of its values whose index has bit o
    // equal to the lane's; lane L ends with the element's sum number L.  Fixed pattern: the result does not depend on scheduling.
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const double send = up ? acc[i] : acc[i + o];
        const double keep = up ? acc[i + o] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    elemAcc[(size_t)e * 32 + lane] = acc[0];
  }
}

// ---- interpolate + push + own-element inside test + delivery ------------------------------------------------------------------------
// Restructured arithmetic: the field tile of an element is converted once per field update (k_nodal_to_mono, at
// piclas_gpu_set_field) from values at the Gauss points to the coefficients of the same polynomial in the monomial basis,
//   E_c(xi, eta, zeta) = sum_ijk a[k][j][i][c] xi^i eta^j zeta^k ,
// and evaluated by nested Horner recurrences: 3 ((N+1)^3 - 1) multiply-adds per particle (189 at N = 3) instead of the 252 + 54
// of the sum-factorised Lagrange form, no basis values to hold in registers.  Same polynomial, differences O(1e-16) of the
// field's magnitude (the reference's own form, eval_xyz.f90:207-215, is what arithmetic = 0 runs).  Two particles per thread
// share every shared-memory operand (128-bit loads, one LDS feeds four FMAs).
template <int NP>
__device__ __forceinline__ void evaluate_field_mono2(const double xi[2][3], const double* __restrict__ sA, double out[2][3]) {
  static_assert(NP >= 2, "N >= 1");
  double o[2][3] = {{0., 0., 0.}, {0., 0., 0.}};
#pragma unroll 1
  for (int k = NP - 1; k >= 0; --k) {
    double s[2][3];
#pragma unroll MONO_UJ
    for (int j = NP - 1; j >= 0; --j) {
      const double* row = sA + ((k * NP + j) * NP) * 3;   // [i][c], NP * 3 doubles
      double u[NP * 3];
      if ((NP * 3) % 2 == 0) {
        const double2* row2 = reinterpret_cast<const double2*>(row);
#pragma unroll
        for (int m = 0; m < NP * 3 / 2; ++m) { const double2 w = row2[m]; u[2 * m] = w.x; u[2 * m + 1] = w.y; }
      } else {
#pragma unroll
        for (int m = 0; m < NP * 3; ++m) u[m] = row[m];
      }
      // Horner in xi; the first step takes both coefficients straight from the tile (no copy of the leading one)
      double r[2][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        r[0][c] = fma(u[(NP - 1) * 3 + c], xi[0][0], u[(NP - 2) * 3 + c]);
        r[1][c] = fma(u[(NP - 1) * 3 + c], xi[1][0], u[(NP - 2) * 3 + c]);
      }
#pragma unroll
      for (int i = NP - 3; i >= 0; --i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          r[0][c] = fma(r[0][c], xi[0][0], u[i * 3 + c]);
          r[1][c] = fma(r[1][c], xi[1][0], u[i * 3 + c]);
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (j == NP - 1) { s[0][c] = r[0][c]; s[1][c] = r[1][c]; }
        else { s[0][c] = fma(s[0][c], xi[0][1], r[0][c]); s[1][c] = fma(s[1][c], xi[1][1], r[1][c]); }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[0][c] = fma(o[0][c], xi[0][2], s[0][c]);
      o[1][c] = fma(o[1][c], xi[1][2], s[1][c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { out[0][c] = o[0][c]; out[1][c] = o[1][c]; }
}

// values at the Gauss points -> monomial coefficients, one element per CTA: three sweeps (xi, eta, zeta) of the inverse
// Vandermonde matrix cst.n2m over the tile
template <int NP>
__global__ void __launch_bounds__(128) k_nodal_to_mono(const double* __restrict__ E, double* __restrict__ A, int nElems) {
  constexpr int ND = NP * NP * NP;
  __shared__ double t0[ND * 3], t1[ND * 3];
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    __syncthreads();
    for (int t = threadIdx.x; t < ND * 3; t += blockDim.x) t0[t] = E[(size_t)e * ND * 3 + t];
    __syncthreads();
    for (int t = threadIdx.x; t < ND * 3; t += blockDim.x) {   // xi direction
      const int c = t % 3, n = t / 3, p = n % NP, kj = n / NP;
      double a = 0.;
#pragma unroll
      for (int i = 0; i < NP; ++i) a = fma(cst.n2m[p][i], t0[(kj * NP + i) * 3 + c], a);
      t1[t] = a;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ND * 3; t += blockDim.x) {   // eta direction
      const int c = t % 3, n = t / 3, i = n % NP, p = (n / NP) % NP, k = n / (NP * NP);
      double a = 0.;
#pragma unroll
      for (int j = 0; j < NP; ++j) a = fma(cst.n2m[p][j], t1[((k * NP + j) * NP + i) * 3 + c], a);
      t0[t] = a;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ND * 3; t += blockDim.x) {   // zeta direction
      const int c = t % 3, n = t / 3, ji = n % (NP * NP), p = n / (NP * NP);
      double a = 0.;
#pragma unroll
      for (int k = 0; k < NP; ++k) a = fma(cst.n2m[p][k], t0[(k * NP * NP + ji) * 3 + c], a);
      A[(size_t)e * ND * 3 + t] = a;
    }
  }
}

// One particle through the general path — non-affine or non-planar element, closed form far outside the element, position within
// tol of a side plane, B != 0, reference-order arithmetic: reference position (Newton), field, push, ParticleInsideQuad3D of the
// own element.  Out of line and self-contained, so that the hot path of k_bin_push holds no state across a call.  left != 0:
// the particle is not in its element any more (far list: exact walk).
struct ColdRes { double x0, x1, x2, v0, v1, v2; int left; };
template <int NP, bool FAST>
__device__ __noinline__ ColdRes cold_particle(double x0, double x1, double x2, double v0, double v1, double v2, int meta,
                                              const double* __restrict__ sE, const GeoElem* __restrict__ ge, const AffElem* __restrict__ af,
                                              const PlaneElem* __restrict__ pl, const TriaElem* __restrict__ te, const double* __restrict__ Eg,
                                              const double* __restrict__ xgp, double dt) {
  double x[3] = {x0, x1, x2}, v[3] = {v0, v1, v2};
  const int spec = meta & META_SPEC_MASK;
  bool isNew = (meta & META_ISNEW) != 0;
  double F[6] = {0., 0., 0., 0., 0., 0.};
  const double q = cst.ChargeIC[spec];
  if (cst.DoInterpolation && fabs(q) > 0.0) {  // isInterpolateParticle
    double xi[3];
    bool suc;
    if (FAST) suc = ref_position_fast(af, ge, x, xi, false);
    else suc = (position_in_ref_elem(ge, x, xi, false, true) & 1) != 0;
    double f3[3];
    if (!suc && cst.DepositionType == PGPU_DEPO_CVWM) field_inverse_distance<NP>(x, Eg, xgp, f3);
    else if (FAST) {   // sE: monomial coefficients
      const double xi2[2][3] = {{xi[0], xi[1], xi[2]}, {xi[0], xi[1], xi[2]}};
      double o2[2][3];
      evaluate_field_mono2<NP>(xi2, sE, o2);
      f3[0] = o2[0][0]; f3[1] = o2[0][1]; f3[2] = o2[0][2];
    } else evaluate_field<NP>(xi, sE, f3);   // sE: values at the Gauss points
#pragma unroll
    for (int c = 0; c < 6; ++c) F[c] = cst.externalField[c];
    F[0] = F[0] + f3[0]; F[1] = F[1] + f3[1]; F[2] = F[2] + f3[2];
    F[3] = F[3] + 0.; F[4] = F[4] + 0.; F[5] = F[5] + 0.;
  }
  if (FAST) push_particle_fast(x, v, F, spec, isNew, dt);
  else push_particle(x, v, F, spec, isNew, dt);
  uint32_t mask;
  const bool in = FAST ? inside_fast<true>(pl, te, x, mask) : inside_quad3d_mask<true>(te, x, mask);
  ColdRes r;
  r.x0 = x[0]; r.x1 = x[1]; r.x2 = x[2]; r.v0 = v[0]; r.v1 = v[1]; r.v2 = v[2];
  r.left = in ? 0 : 1;
  return r;
}

// hot-path push with B = 0 (both time discretisations), fused multiply-adds; same formulas as push_particle_fast (fastmath.cuh)
__device__ __forceinline__ void push_inline_b0(double x[3], double v[3], const double E3[3], int spec, bool isNew, bool boris, double dt) {
  const double q = cst.ChargeIC[spec], mass = cst.MassIC[spec];
  const bool isPush = fabs(q) > 0.0;
  if (boris) {
    if (isPush && cst.DoInterpolation) {
      if (isNew) {
        const double h = -0.5 * dt * (q / mass);
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = fma(E3[d], h, v[d]);
      }
      const double c_1 = (q * dt) / (mass * 2.);
      const double c2_inv = cst.c2_inv;
      const double gamma = rsqrt(fma(-fma(v[0], v[0], fma(v[1], v[1], v[2] * v[2])), c2_inv, 1.0));
      double vn[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) vn[d] = fma(2. * c_1, E3[d], v[d] * gamma);
      const double s = rsqrt(fma(fma(vn[0], vn[0], fma(vn[1], vn[1], vn[2] * vn[2])), c2_inv, 1.0));
#pragma unroll
      for (int d = 0; d < 3; ++d) v[d] = vn[d] * s;
    }
  } else {   // Leapfrog (timedisc_TimeStepPoisson.f90:124-181), operation order of push_particle
    double Pt[3] = {0., 0., 0.};
    if (cst.DoInterpolation && isPush) {
      const double qmt = q / mass;
#pragma unroll
      for (int d = 0; d < 3; ++d) Pt[d] = E3[d] * qmt;
    }
    if (isPush) {
      if (isNew) {
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = v[d] - (Pt[d] * dt) * 0.5;
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) v[d] = v[d] + Pt[d] * dt;
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) x[d] = fma(v[d], dt, x[d]);
}

constexpr int UNPUSHED_LIST = 4096;   // far slots of unpushed records kept behind the eight counters; more: scan of the whole far list
constexpr int CAT_STAY = 0, CAT_FAR = 7, CAT_NONE = 8;

// HOTONLY: every local element is affine with planar sides, B = 0, restructured arithmetic.  The kernel then contains no call at
// all (a call anywhere in the loop makes the compiler keep the loop's state in local memory: 340 bytes of spills per thread); the
// rare particle that needs the general path (closed-form reference position far outside the element, pushed position within tol
// of a side plane) is handed UNPUSHED to the far list, where k_far_unpushed runs the general path for it before the walk.
template <int NP, bool FAST, bool HOTONLY>
__global__ void __launch_bounds__(KB_NT, KB_MINB) k_bin_push(PartBuf bins, PartBuf pool, BinView bv, int cur, FarBuf far,
                                                              const int64_t* __restrict__ farBase, int32_t* __restrict__ nFarE,
                                                              const PushElem* __restrict__ pushElem, const double* __restrict__ Et /*tile source: monomial
                                                              coefficients (FAST) or Gauss-point values*/, const double* __restrict__ E /*Gauss-point values*/,
                                                              const GeoElem* __restrict__ geo, const TriaElem* __restrict__ tria,
                                                              const PlaneElem* __restrict__ planes, const AffElem* __restrict__ aff,
                                                              const double* __restrict__ Elem_xGP, int offsetElem, double dt,
                                                              int* __restrict__ counters /*[4] side movers, [5] main full, [6] inbox full*/) {
  constexpr int ND = NP * NP * NP;
  // field tile: double buffered and prefetched with a bulk copy when it is small and a multiple of 16 bytes (N = 1, 3);
  // otherwise one buffer, staged with plain loads after the element's barrier
  constexpr uint32_t E_BYTES = ND * 24;
  constexpr bool TMA_E = (E_BYTES % 16 == 0) && (E_BYTES <= 4096);
  constexpr int EBUF = TMA_E ? 2 : 1;
  __shared__ __align__(16) PushElem sPE[2];
  __shared__ __align__(16) double sEb[EBUF][ND * 3];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ __align__(16) double sP[2][6][KB_CHUNK];   // x, v of the current / next pair of every thread (cp.async)
  __shared__ uint32_t sM[2][KB_NT][2];                  // the aligned 32-bit words that hold the pair's two meta bytes
  __shared__ __align__(16) double sN[6][KB_CHUNK];       // pushed x, v of the sweep's particles
  __shared__ ElemRanges R2[2];
  __shared__ int sCnt[2][KB_NT / 32][8];
  __shared__ int sRun[2][8];
  __shared__ int64_t sNbBase[6];                         // first slot of the inbox this element fills at the neighbour behind side s
  __shared__ int sNbCap[6];
  __shared__ int sDiverted;                              // particles of this element sent to the far list by a full region
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
  KB_SYNC();
  auto prefetch_elem = [&](int e, int b) {   // one thread: records of element e -> buffer b
    mbar_expect_tx(&mbar[b], (uint32_t)sizeof(PushElem) + (TMA_E ? E_BYTES : 0u));
    bulk_g2s(&sPE[b], pushElem + e, (uint32_t)sizeof(PushElem), &mbar[b]);
    if (TMA_E) bulk_g2s(&sEb[TMA_E ? b : 0][0], Et + (size_t)e * ND * 3, E_BYTES, &mbar[b]);
  };
  // pair of this thread in chunk c of the element with ranges R: virtual indices c*KB_CHUNK + 2*tid, +1.  Copies only: nothing
  // here waits for memory.  Returns the slot of the first particle (-1: none), the number of live particles and the source.
  auto fetch = [&](const ElemRanges& R, int c, int stg, int& nLive, int64_t& slot, bool& fromPool) {
    const int v = c * KB_CHUNK + 2 * tid;
    nLive = 0;
    slot = -1;
    fromPool = false;
    if (v < R.pre[BIN_NRANGE]) {
      int r, left;
      resolve(R, v, r, slot, left);
      fromPool = r == 7;
      nLive = left >= 2 ? 2 : (left > 0 ? 1 : 0);
      if (nLive > 0) {
        const double* sf = fromPool ? pool.f : BF;
        const int64_t sst = fromPool ? pool.stride : BS;
        const uint8_t* sm = fromPool ? pool.meta : bins.meta;
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
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&sM[stg][tid][0])), "l"(sm + (slot & ~(int64_t)3)) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&sM[stg][tid][1])), "l"(sm + ((slot + 1) & ~(int64_t)3)) : "memory");
      } else slot = -1;
    }
  };
  int nLive = 0, nLiveNext = 0;
  int64_t slotCur = -1, slotNext = -1;
  bool poolCur = false, poolNext = false;
  int stg = 0;
  if ((int)blockIdx.x < nElems) {
    if (tid == 0) prefetch_elem(blockIdx.x, 0);
    if (warp == 0) range_store(R2[0], range_load(bv, blockIdx.x, cur, lane), lane);
    KB_SYNC();
    fetch(R2[0], 0, 0, nLive, slotCur, poolCur);
  }
  cp_async_commit();
  int it = 0;
  for (int e = blockIdx.x; e < nElems; e += gridDim.x, ++it) {
    // invariant: ranges of e in R2[it & 1]; its chunk 0 in flight in stage stg (latest cp.async group); records of e in flight in
    // buffer it & 1; every thread has passed the barrier that ended the previous element
    const int b = it & 1;
    const int gElem = offsetElem + e + 1;
    const int e2 = e + gridDim.x;
    const bool haveNext = e2 < nElems;
    const ElemRanges& R = R2[b];
    if (tid == 0 && haveNext) prefetch_elem(e2, b ^ 1);
    RangeRegs rNext;
    rNext.start = 0; rNext.cnt = 0;
    if (warp == 0 && haveNext) rNext = range_load(bv, e2, cur, lane);   // consumed after the work of chunk 0
    if (tid < 8) sRun[0][tid] = 0;
    if (tid == 8) sDiverted = 0;
    const int64_t farLo = farBase[e], farHi = farBase[e + 1];
    mbar_wait(&mbar[b], (uint32_t)((it >> 1) & 1));
    if (!TMA_E) {
      for (int t = tid; t < ND * 3; t += KB_NT) sEb[0][t] = __ldg(Et + (size_t)e * ND * 3 + t);
      KB_SYNC();
    }
    const PushElem& pe = sPE[b];
    const double* sE = sEb[TMA_E ? b : 0];
    int64_t nbBase = 0;
    int nbCap = 0;
    if (warp == 0 && lane >= 8 && lane < 14) {   // inboxes this element fills: consumed after the work of chunk 0
      const int nb = pe.nbLocal[lane - 8];
      if (nb >= 0) { nbBase = bin_inbox_base(bv, nb, next, pe.nbBox[lane - 8]); nbCap = bv.capIn[nb]; }
    }
    const int total = R.pre[BIN_NRANGE];
    const bool boris = cst.TimeDiscMethod == PGPU_TIMEDISC_BORIS_LEAPFROG;
    const bool noB = cst.externalField[3] == 0. && cst.externalField[4] == 0. && cst.externalField[5] == 0.;
    const bool hotElem = HOTONLY || (FAST && pe.affine != 0u && pe.planar != 0u && noB);
    const int nChunks = (total + KB_CHUNK - 1) / KB_CHUNK;
    const int64_t base_e = R.start[0];
    const int capMain = bv.capMain[e];
    auto publish_next = [&]() {   // warp 0: ranges of the next element, inboxes of this one
      if (haveNext) range_store(R2[b ^ 1], rNext, lane);
      if (lane >= 8 && lane < 14) { sNbBase[lane - 8] = nbBase; sNbCap[lane - 8] = nbCap; }
    };
    if (nChunks <= 1) {   // too short to hide the loads behind chunk 0: publish now
      if (warp == 0) publish_next();
      KB_SYNC();
      if (nChunks == 0) {   // empty element: only the chunk 0 of the next element has to be put in flight
        if (haveNext) fetch(R2[b ^ 1], 0, stg, nLive, slotCur, poolCur);
        cp_async_commit();
      }
    }
    for (int c = 0; c < nChunks; ++c) {
      const int cb = c & 1;
      nLiveNext = 0;
      slotNext = -1;
      poolNext = false;
      if (c + 1 < nChunks) fetch(R, c + 1, stg ^ 1, nLiveNext, slotNext, poolNext);
      else if (haveNext) fetch(R2[b ^ 1], 0, stg ^ 1, nLiveNext, slotNext, poolNext);   // chunk 0 of the next element
      cp_async_commit();
      cp_async_wait_prev();
      // ---- per-particle work ----------------------------------------------------------------------------------------------
      // hot path (no calls): affine + planar element, closed-form reference position, B = 0, position clear of every side plane;
      // everything else goes through cold_particle.  The field of the thread's two particles is evaluated together (shared
      // operands); push, inside test and delivery then run particle by particle in rolled loops, the pushed state parked in
      // shared memory: half the code and far fewer live registers than two interleaved copies.
      const int pq = (nLive == 1) ? 0 : 1;   // the idle half of a pair computes on a copy of the live one (finite inputs, discarded)
      const double* sPs = &sP[stg][0][2 * tid];   // [a * KB_CHUNK + q]: x, v as loaded (x = LastPartPos)
      double* sNs = &sN[0][2 * tid];              // [a * KB_CHUNK + q]: pushed x, v
      int cats = CAT_NONE | (CAT_NONE << 4);
      int unpushed = 0;
      if (nLive > 0) {
        double f2[2][3];
        bool farOut[2] = {false, false};   // closed-form reference position far outside the element
        if (hotElem) {
          double xi[2][3];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int qs = q ? pq : 0;
            const double r0 = sPs[qs] - pe.x0[0], r1 = sPs[KB_CHUNK + qs] - pe.x0[1], r2 = sPs[2 * KB_CHUNK + qs] - pe.x0[2];
#pragma unroll
            for (int d = 0; d < 3; ++d) xi[q][d] = fma(pe.A[d][0], r0, fma(pe.A[d][1], r1, pe.A[d][2] * r2)) - 1.0;
            farOut[q] = !(fabs(xi[q][0]) <= 1.5 && fabs(xi[q][1]) <= 1.5 && fabs(xi[q][2]) <= 1.5);
            if (farOut[q]) { xi[q][0] = 0.; xi[q][1] = 0.; xi[q][2] = 0.; }
          }
          evaluate_field_mono2<NP>(xi, sE, f2);
        }
#if KB_UNROLL_Q
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int q = 0; q < 2; ++q) {
          if (q >= nLive) continue;
          const int metaQ = (int)((sM[stg][tid][q] >> (8 * (int)((slotCur + q) & 3))) & 0xffu);
          double xq[3] = {sPs[q], sPs[KB_CHUNK + q], sPs[2 * KB_CHUNK + q]};
          double vq[3] = {sPs[3 * KB_CHUNK + q], sPs[4 * KB_CHUNK + q], sPs[5 * KB_CHUNK + q]};
          int cat = CAT_FAR;
          bool cold = !hotElem || (q ? farOut[1] : farOut[0]);
          if (!cold) {
            const int spec = metaQ & META_SPEC_MASK;
            double E3[3] = {0., 0., 0.};
            if (cst.DoInterpolation && fabs(cst.ChargeIC[spec]) > 0.0) {
#pragma unroll
              for (int d = 0; d < 3; ++d) E3[d] = cst.externalField[d] + (q ? f2[1][d] : f2[0][d]);
            }
            const double lp0 = xq[0], lp1 = xq[1], lp2 = xq[2];
            push_inline_b0(xq, vq, E3, spec, (metaQ & META_ISNEW) != 0, boris, dt);
            // ---- own-element inside test (first iteration of SingleParticleTriaTracking3D, particle_triatracking.f90:203-218)
            //      and the crossing of exactly one side plane into a face neighbour ------------------------------------------------
            double dx[6];
            uint32_t neg = 0;
            bool ambiguous = false;
            const double tol = pe.tol;
#pragma unroll
            for (int s = 0; s < 6; ++s) {
              dx[s] = fma(pe.pl[s][0], xq[0], fma(pe.pl[s][1], xq[1], fma(pe.pl[s][2], xq[2], -pe.pl[s][3])));
              ambiguous |= fabs(dx[s]) <= tol;
              neg |= ((uint32_t)__double2hiint(dx[s]) >> 31) << s;   // sign bit (a zero of either sign is ambiguous anyway)
            }
            if (ambiguous) cold = true;   // within tol of a side plane: the determinants decide (ParticleInsideQuad3D)
            else if (neg == 0u) cat = CAT_STAY;
            else if (__popc(neg) == 1 && pe.nbLocal[__ffs(neg) - 1] >= 0) {
              const int s = __ffs(neg) - 1;
              // flight LastPartPos -> x crosses side s at lp + alpha (x - lp); every decision with a margin of tol, otherwise
              // the determinant tests of the exact walk decide
              const double dl = fma(pe.pl[s][0], lp0, fma(pe.pl[s][1], lp1, fma(pe.pl[s][2], lp2, -pe.pl[s][3])));
              const double dxs = fma(pe.pl[s][0], xq[0], fma(pe.pl[s][1], xq[1], fma(pe.pl[s][2], xq[2], -pe.pl[s][3])));   // == dx[s], no dynamic index
              const double alpha = dl / (dl - dxs);
              bool ok = dl > tol;
#pragma unroll
              for (int o = 0; o < 6; ++o) {
                const double ol = fma(pe.pl[o][0], lp0, fma(pe.pl[o][1], lp1, fma(pe.pl[o][2], lp2, -pe.pl[o][3])));
                const double oc = fma(alpha, dx[o] - ol, ol);
                if (o != s && !(oc > tol)) ok = false;
              }
              {
                const double gl = fma(pe.dg[s][0], lp0, fma(pe.dg[s][1], lp1, fma(pe.dg[s][2], lp2, -pe.dg[s][3])));
                const double gx = fma(pe.dg[s][0], xq[0], fma(pe.dg[s][1], xq[1], fma(pe.dg[s][2], xq[2], -pe.dg[s][3])));
                const double gc = fma(alpha, gx - gl, gl);
                if (!(fabs(gc) > tol)) ok = false;   // crossing point on the triangle diagonal
              }
              // periodic side: the particle goes on behind the partner side (PeriodicBoundary, particle_boundary_condition.f90:224-284:
              // LastPartPos = crossing point + vector, PartState = LastPartPos + rest of the flight; here the vector is added
              // to the pushed position, the same point to rounding)
              double xs[3] = {xq[0], xq[1], xq[2]};
              const uint32_t sc = (pe.shiftCode >> (5 * s)) & 31u;
              if (sc) {
                const int pv = (int)(sc & 15u) - 1;
#pragma unroll
                for (int d = 0; d < 3; ++d) xs[d] = (sc & 16u) ? xq[d] - cst.PeriodicVectors[pv][d] : xq[d] + cst.PeriodicVectors[pv][d];
              }
              // clearly inside the neighbour: ParticleInsideQuad3D there succeeds, the walk ends (:215-218)
              const double ntol = pe.nbtol[s];
#pragma unroll
              for (int o = 0; o < 6; ++o) {
                const double dn = fma(pe.nbpl[s][o][0], xs[0], fma(pe.nbpl[s][o][1], xs[1], fma(pe.nbpl[s][o][2], xs[2], -pe.nbpl[s][o][3])));
                if (!(dn > ntol)) ok = false;
              }
              if (ok) {
                cat = 1 + s;
                xq[0] = xs[0]; xq[1] = xs[1]; xq[2] = xs[2];
              }
            }
          }
          if (cold) {
            if (HOTONLY) {   // unpushed to the far list: state as loaded
#pragma unroll
              for (int d = 0; d < 3; ++d) { xq[d] = sPs[d * KB_CHUNK + q]; vq[d] = sPs[(3 + d) * KB_CHUNK + q]; }
              cat = CAT_FAR;
              unpushed |= 1 << q;
            } else {
              const ColdRes r = cold_particle<NP, FAST>(sPs[q], sPs[KB_CHUNK + q], sPs[2 * KB_CHUNK + q], sPs[3 * KB_CHUNK + q], sPs[4 * KB_CHUNK + q],
                                                        sPs[5 * KB_CHUNK + q], metaQ, sE, geo + (gElem - 1), FAST ? aff + (gElem - 1) : nullptr,
                                                        FAST ? planes + (gElem - 1) : nullptr, tria + (gElem - 1), E + (size_t)e * ND * 3,
                                                        Elem_xGP + (size_t)(gElem - 1) * ND * 3, dt);
              xq[0] = r.x0; xq[1] = r.x1; xq[2] = r.x2; vq[0] = r.v0; vq[1] = r.v1; vq[2] = r.v2;
              cat = r.left ? CAT_FAR : CAT_STAY;
            }
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) { sNs[d * KB_CHUNK + q] = xq[d]; sNs[(3 + d) * KB_CHUNK + q] = vq[d]; }
          cats = q ? ((cats & 0xf) | (cat << 4)) : ((cats & 0xf0) | cat);
        }
      }
      if (c == 0 && nChunks > 1 && warp == 0) publish_next();   // the loads issued at the element's start have long arrived
      // particle ids (tests) have to be read before the delivery below overwrites slots of this sweep
      int64_t id0 = 0, id1 = 0;
      if (bins.id && nLive > 0) {
        const int64_t* sid = poolCur ? pool.id : bins.id;
        id0 = sid[slotCur];
        if (nLive > 1) id1 = sid[slotCur + 1];
      }
      // ---- stable compaction: rank of every particle within its category in the order of the element's particles (virtual
      //      index 2 * tid + q), chunk by chunk: the order inside an element never changes except by departures and arrivals ----------
      // eight 8-bit counters (one per category, <= 64 particles per warp and sweep) packed in a 64-bit word: one inclusive warp
      // scan yields every particle's rank within its category and the warp's totals
      const int cat0 = cats & 0xf, cat1 = cats >> 4;
      int rank0, rank1;
      {
        const unsigned long long w0 = cat0 < 8 ? (1ull << (8 * cat0)) : 0ull, w1 = cat1 < 8 ? (1ull << (8 * cat1)) : 0ull;
        unsigned long long incl = w0 + w1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const unsigned long long excl = incl - w0 - w1;
        rank0 = (int)((excl >> (8 * (cat0 & 7))) & 0xffull);
        rank1 = (int)((excl >> (8 * (cat1 & 7))) & 0xffull) + (cat0 == cat1 ? 1 : 0);
        const unsigned long long tot = __shfl_sync(0xffffffffu, incl, 31);
        if (lane < 8) sCnt[cb][warp][lane] = (int)((tot >> (8 * lane)) & 0xffull);
      }
      KB_SYNC();
      if (tid < 8) {
        int t = sRun[cb][tid];
#pragma unroll
        for (int w = 0; w < KB_NT / 32; ++w) t += sCnt[cb][w][tid];
        sRun[cb ^ 1][tid] = t;
      }
      // ---- delivery ----------------------------------------------------------------------------------------------------------------
      // stayers -> main, side movers -> the neighbour's inbox, far records -> this element's far region from the bottom (slot =
      // rank among the element's far particles: no atomics); a particle whose region is full -> far region from the top
#pragma unroll 1
      for (int q = 0; q < nLive; ++q) {
        const int cat = q ? cat1 : cat0;
        int off = sRun[cb][cat] + (q ? rank1 : rank0);
#pragma unroll
        for (int w = 0; w < KB_NT / 32; ++w)
          if (w < warp) off += sCnt[cb][w][cat];
        int64_t dst = -1, fdst = -1;
        if (cat == CAT_STAY) {
          if (off < capMain) dst = base_e + off;
          else { fdst = farHi - 1 - atomicAdd(&sDiverted, 1); atomicAdd(&counters[5], 1); }   // main full: through the far list back in
        } else if (cat <= 6) {
          if (off < sNbCap[cat - 1]) dst = sNbBase[cat - 1] + off;
          else { fdst = farHi - 1 - atomicAdd(&sDiverted, 1); atomicAdd(&counters[6], 1); }   // inbox full
        } else fdst = farLo + off;
        const int metaQ = (int)((sM[stg][tid][q] >> (8 * (int)((slotCur + q) & 3))) & 0xffu);
        // IsNewPart is consumed by the push; an unpushed far record keeps it
        const uint8_t nmeta = (HOTONLY && ((unpushed >> q) & 1)) ? (uint8_t)((metaQ & (META_SPEC_MASK | META_ISNEW)) | META_UNPUSHED)
                                                                 : (uint8_t)(metaQ & META_SPEC_MASK);
        if (HOTONLY && ((unpushed >> q) & 1)) {   // the few unpushed records are listed behind the counters (k_far_unpushed)
          const int pos = atomicAdd(&counters[7], 1);
          if (pos < UNPUSHED_LIST) counters[8 + pos] = (int)fdst;
        }
        if (dst >= 0) {
#pragma unroll
          for (int a = 0; a < 6; ++a) BF[a * BS + dst] = sNs[a * KB_CHUNK + q];
          bins.meta[dst] = nmeta;
          if (bins.id) bins.id[dst] = q ? id1 : id0;
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) { far.x[d][fdst] = sNs[d * KB_CHUNK + q]; far.lp[d][fdst] = sPs[d * KB_CHUNK + q]; far.v[d][fdst] = sNs[(3 + d) * KB_CHUNK + q]; }
          far.elem[fdst] = gElem;
          far.meta[fdst] = nmeta;
          if (far.id) far.id[fdst] = q ? id1 : id0;
        }
      }
      nLive = nLiveNext;
      slotCur = slotNext;
      poolCur = poolNext;
      stg ^= 1;
    }
    KB_SYNC();
    // populations after the step: this element's main, and the inboxes this element fills at its face neighbours
    {
      const int fb = nChunks & 1;
      if (tid == 0) {
        const int n = sRun[fb][0];
        bv.nMain[e] = n < capMain ? n : capMain;
      } else if (tid == 7) {
        nFarE[2 * e] = sRun[fb][7];
        nFarE[2 * e + 1] = sDiverted;
      } else if (tid >= 1 && tid <= 6) {
        const int s = tid - 1, nb = pe.nbLocal[s];
        if (nb >= 0) {
          const int n = sRun[fb][tid], ci = sNbCap[s];
          const int m = n < ci ? n : ci;
          bv.nIn[((size_t)next * nElems + nb) * 8 + pe.nbBox[s]] = m;
          if (m > 0) atomicAdd(&counters[4], m);
        }
      }
    }
    KB_SYNC();
  }
}

// general path for the far records that k_bin_push<.., HOTONLY> handed over unpushed: field at the particle, push (x = pushed
// position, lp = LastPartPos stays); the walk that follows does the inside test of the own element
template <int NP>
__global__ void k_far_unpushed(FarBuf far, const uint32_t* __restrict__ idx, const int* __restrict__ list, int nFar, const double* __restrict__ Emono, const double* __restrict__ E,
                               const GeoElem* __restrict__ geo, const AffElem* __restrict__ aff, const PlaneElem* __restrict__ planes,
                               const TriaElem* __restrict__ tria, const double* __restrict__ Elem_xGP, int offsetElem, double dt) {
  constexpr int ND = NP * NP * NP;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nFar) return;
  const int64_t p = idx ? (int64_t)idx[i] : (int64_t)list[i];   // list: the slots k_bin_push noted; idx: every far slot
  const int meta = far.meta[p];
  if (!(meta & META_UNPUSHED)) return;
  const int g = far.elem[p], e = g - 1 - offsetElem;
  const ColdRes r = cold_particle<NP, true>(far.x[0][p], far.x[1][p], far.x[2][p], far.v[0][p], far.v[1][p], far.v[2][p], meta & ~META_UNPUSHED,
                                            Emono + (size_t)e * ND * 3, geo + (g - 1), aff + (g - 1), planes + (g - 1), tria + (g - 1),
                                            E + (size_t)e * ND * 3, Elem_xGP + (size_t)(g - 1) * ND * 3, dt);
  far.x[0][p] = r.x0; far.x[1][p] = r.x1; far.x[2][p] = r.x2;
  far.v[0][p] = r.v0; far.v[1][p] = r.v1; far.v[2][p] = r.v2;
  far.meta[p] = (uint8_t)(meta & META_SPEC_MASK);
}

// ---- far records that come to rest within two side crossings --------------------------------------------------------------------
// Most far records left their element through an edge region: beyond two side planes, at rest in the neighbour's neighbour.
// One warp per element stages the element's PushElem (own planes and the planes of the six neighbours, one coalesced copy for
// all its records) and takes up to two crossings with the side planes exactly as exit_side_planar does (every decision with a
// margin of tol, the flight taken from LastPartPos as SingleParticleTriaTracking3D does, particle_triatracking.f90:171-173);
// only the planes of the last element are gathered from the element table.  A record is settled when it is clearly inside
// the element reached (ParticleInsideQuad3D there succeeds, :215-218); everything else - boundary sides, non-planar
// elements, a third crossing, any decision within tol - is appended to the pending list, which k_far_walk takes from the start
// with the determinant tests.  Also writes the dense index of the element's far slots (k_far_index).
#ifndef FH_MINB
#define FH_MINB 5
#endif
constexpr int FH_WARPS = 4, FH_PEND = 512;
__global__ void __launch_bounds__(FH_WARPS * 32, FH_MINB) k_far_hint(FarBuf far, const int64_t* __restrict__ farBase, const int32_t* __restrict__ nFarE,
                                                            const int64_t* __restrict__ dOff, int nElems, int offsetElem,
                                                            const PushElem* __restrict__ pushElems, const HintNb* __restrict__ hintNb,
                                                            const TriaElem* __restrict__ tria,
                                                            const PlaneElem* __restrict__ planes, const int32_t* __restrict__ elemRank,
                                                            uint32_t* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t* __restrict__ pend,
                                                            int* __restrict__ counters, int* __restrict__ emigCnt) {
  __shared__ PushElem sPe[FH_WARPS];
  __shared__ HintNb sHn[FH_WARPS];
  __shared__ uint32_t sPend[FH_WARPS][FH_PEND];
  int nPend = 0;   // warp-uniform
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const PushElem& pe = sPe[w];
  const HintNb& hn = sHn[w];
  for (int e = blockIdx.x * FH_WARPS + w; e < nElems; e += gridDim.x * FH_WARPS) {
    const int nb = nFarE[2 * e], n = nb + nFarE[2 * e + 1];
    if (n == 0) continue;
    const int64_t b0 = farBase[e], b1 = farBase[e + 1], d0 = dOff[e];
    __syncwarp();
    {
      const int4* src = reinterpret_cast<const int4*>(pushElems + e);
      int4* dst = reinterpret_cast<int4*>(&sPe[w]);
      for (int i = lane; i < (int)(sizeof(PushElem) / 16); i += 32) dst[i] = __ldg(src + i);
      if (lane < (int)(sizeof(HintNb) / 16)) reinterpret_cast<int4*>(&sHn[w])[lane] = __ldg(reinterpret_cast<const int4*>(hintNb + e) + lane);
    }
    __syncwarp();
    const int ge = offsetElem + e + 1;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < n;
      int64_t slot = 0;
      int fin = 0;   // element the record comes to rest in, 0: undecided
      if (valid) {
        slot = i < nb ? b0 + i : b1 - 1 - (i - nb);
        idx[d0 + i] = (uint32_t)slot;
        if (pe.planar) {
          double x[3] = {far.x[0][slot], far.x[1][slot], far.x[2][slot]};
          double lp[3] = {far.lp[0][slot], far.lp[1][slot], far.lp[2][slot]};
          bool moved = false;
          fin = far_hint_record<true>(pe, hn, tria, planes, ge, x, lp, moved);
          if (fin > 0 && moved) { far.x[0][slot] = x[0]; far.x[1][slot] = x[1]; far.x[2][slot] = x[2]; }
        }
        if (fin > 0) {
          const int rk = (cst.nRanks == 1) ? cst.myRank : elemRank[fin - 1];
          keys[d0 + i] = (rk == cst.myRank) ? (uint32_t)(fin - 1 - offsetElem) : (uint32_t)(nElems + rk);
          if (rk != cst.myRank && emigCnt) atomicAdd(&emigCnt[rk], 1);
          if (fin != ge) far.elem[slot] = fin;
        }
      }
      // undecided records: collected per warp, appended to the pending list a few hundred at a time (one atomic on the list's
      // counter per flush; one per 32 records made the kernel wait on that one address for most of its run time)
      const unsigned open = __ballot_sync(0xffffffffu, valid && fin == 0);
      if (open) {
        if (valid && fin == 0) sPend[w][nPend + __popc(open & ((1u << lane) - 1u))] = (uint32_t)(d0 + i);
        nPend += __popc(open);
        if (nPend > FH_PEND - 32) {
          __syncwarp();
          int first = 0;
          if (lane == 0) first = atomicAdd(&counters[2], nPend);
          first = __shfl_sync(0xffffffffu, first, 0);
          for (int j = lane; j < nPend; j += 32) pend[first + j] = sPend[w][j];
          nPend = 0;
          __syncwarp();
        }
      }
    }
  }
  if (nPend > 0) {
    __syncwarp();
    int first = 0;
    if (lane == 0) first = atomicAdd(&counters[2], nPend);
    first = __shfl_sync(0xffffffffu, first, 0);
    for (int j = lane; j < nPend; j += 32) pend[first + j] = sPend[w][j];
  }
}

// ---- SingleParticleTriaTracking3D for the far list (particle_triatracking.f90:137-484), from the start -----------------------------
// One thread per record, persistent warps with per-lane refill (as k_track_leavers).  key: local element, nElems + rank
// (emigrant) or nElems + nRanks (removed).
template <bool FAST>
__global__ void __launch_bounds__(LV_NT, LV_MINB) k_far_walk(FarBuf far, const uint32_t* __restrict__ idx, int nFar, const TriaElem* __restrict__ tria,
                                                             const PlaneElem* __restrict__ planes, const int32_t* __restrict__ elemRank,
                                                             uint32_t* __restrict__ keys, int nElems, int offsetElem, int* __restrict__ counters,
                                                             int* __restrict__ emigCnt /*[nRanks] emigrants per destination rank, or null*/,
                                                             const uint32_t* __restrict__ pend /*dense indices left by k_far_hint (count in counters[2]), or null: all*/) {
  const int lane = threadIdx.x & 31;
  if (pend) nFar = counters[2];
  // records per grab: the list spread over all warps of the grid (a short pending list in 256-record chunks would be walked by a
  // handful of warps, eight rounds each), between 32 and 256
  int FW_CHUNK = (nFar / (int)(gridDim.x * (blockDim.x >> 5)) + 31) & ~31;
  FW_CHUNK = FW_CHUNK < 32 ? 32 : (FW_CHUNK > 256 ? 256 : FW_CHUNK);
  int chunkNext = 0, chunkEnd = 0;   // warp-uniform
  bool drained = false;
  bool active = false;
  int p = 0, dense = 0, ElemID = 0, guard = 0;
  uint32_t mask = 0;
  double x[3] = {0., 0., 0.}, lp[3] = {0., 0., 0.};
  HopHist h;
  h.clear();
  while (true) {
    int status = -1;
    {
      // idle lanes take the next records of the warp's chunk of the list; a new chunk (one atomic on the list's counter per
      // chunk, not one per refill) when it runs out
      unsigned need = __ballot_sync(0xffffffffu, !active);
      while (need) {
        if (chunkNext >= chunkEnd) {
          if (drained) break;
          int first = 0;
          if (lane == 0) first = atomicAdd(&counters[3], FW_CHUNK);
          first = __shfl_sync(0xffffffffu, first, 0);
          if (first >= nFar) { drained = true; break; }
          chunkNext = first;
          chunkEnd = first + FW_CHUNK < nFar ? first + FW_CHUNK : nFar;
        }
        const int avail = chunkEnd - chunkNext, want = __popc(need);
        const int rank = __popc(need & ((1u << lane) - 1u));
        if (!active && rank < avail) {
          const int mine = chunkNext + rank;
          active = true;
          dense = pend ? (int)pend[mine] : mine;
          p = (int)idx[dense];
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
        chunkNext += want < avail ? want : avail;
        if (want <= avail) break;
        need = __ballot_sync(0xffffffffu, !active);
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
          if (rk != cst.myRank && emigCnt) atomicAdd(&emigCnt[rk], 1);
        } else {
          key = (uint32_t)(nElems + cst.nRanks);
          newElem = 0;
          if (status == TRK_LOST) atomicAdd(&counters[0], 1);
          else if (status != TRK_REMOVED) atomicMax(&counters[1], status);
        }
        far.x[0][p] = x[0]; far.x[1][p] = x[1]; far.x[2][p] = x[2];
        far.elem[p] = newElem;
        keys[dense] = key;
        active = false;
      }
    }
  }
}

// sorted far list -> pool: pool slot i takes record perm[i] (i < number of records that stay on this rank)
__global__ void k_far_to_pool(FarBuf far, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ perm, int64_t n, PartBuf pool) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = idx[perm[i]];
#pragma unroll
  for (int d = 0; d < 3; ++d) { pool.f[d * pool.stride + i] = far.x[d][s]; pool.f[(3 + d) * pool.stride + i] = far.v[d][s]; }
  pool.meta[i] = far.meta[s];
  if (pool.id) pool.id[i] = far.id[s];
}

// received particles lie behind every element's far region: their index entries follow the local ones in arrival order
__global__ void k_far_index_immigrants(uint32_t* __restrict__ idx, int64_t n0, int64_t n, uint32_t firstSlot) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) idx[n0 + i] = firstSlot + (uint32_t)i;
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
