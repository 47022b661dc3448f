// sort.cu — stable LSD radix sort (8- or 10-bit digits), permutation gathers, segment offsets.
#include "sort.cuh"

namespace {

constexpr int SORT_NT = 256;              // threads per CTA
constexpr int SORT_IPT = 16;              // keys per thread
constexpr int SORT_TILE = SORT_NT * SORT_IPT;
constexpr int SORT_WARPS = SORT_NT / 32;

// per-tile digit histogram -> blockHist[block][digit] (block-major: every table access below is a coalesced row)
template <int RADIX>
__global__ void __launch_bounds__(SORT_NT) k_hist(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t mask,
                                                  uint32_t* __restrict__ blockHist) {
  __shared__ uint32_t h[RADIX];
  for (int i = threadIdx.x; i < RADIX; i += SORT_NT) h[i] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
  const int lane = threadIdx.x & 31;
#pragma unroll 4
  for (int r = 0; r < SORT_IPT; ++r) {
    const size_t i = base + (size_t)r * SORT_NT + threadIdx.x;
    // the particles arrive almost sorted by element: neighbouring keys share the digit, so one atomic per group of equal
    // digits in the warp instead of 32 serialised atomics on one shared-memory word
    const uint32_t digit = (i < n) ? ((keys[i] >> shift) & mask) : (uint32_t)RADIX;
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    if (i < n && lane == __ffs(peers) - 1) atomicAdd(&h[digit], (uint32_t)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RADIX; i += SORT_NT) blockHist[(size_t)blockIdx.x * RADIX + i] = h[i];
}

// Column scan of the table: entry (block b, digit d) becomes the first output index of the keys with digit d in tile b,
//   sum_{d' < d} total[d'] + sum_{b' < b} hist[b'][d]     (stable: tiles in order within a digit)
// in three sweeps over groups of COL_GB tiles; thread <-> digit, so every access is a coalesced row.
constexpr int COL_GB = 256;

template <int RADIX>
__global__ void __launch_bounds__(256) k_col_reduce(const uint32_t* __restrict__ hist, uint32_t nBlocks, uint32_t* __restrict__ groupSum) {
  const uint32_t b0 = blockIdx.x * COL_GB, b1 = min(b0 + COL_GB, nBlocks);
  for (int d = threadIdx.x; d < RADIX; d += 256) {
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; ++b) sum += hist[(size_t)b * RADIX + d];
    groupSum[(size_t)blockIdx.x * RADIX + d] = sum;
  }
}

template <int RADIX>
__global__ void __launch_bounds__(RADIX) k_col_scan(uint32_t* __restrict__ groupSum, uint32_t nGroups) {
  __shared__ uint32_t tot[RADIX];
  const int d = threadIdx.x;
  uint32_t run = 0;
  for (uint32_t g = 0; g < nGroups; ++g) {
    const uint32_t t = groupSum[(size_t)g * RADIX + d];
    groupSum[(size_t)g * RADIX + d] = run;
    run += t;
  }
  tot[d] = run;
  __syncthreads();
  // exclusive scan of the digit totals (Hillis-Steele in shared memory, RADIX <= 1024 threads)
  for (int o = 1; o < RADIX; o <<= 1) {
    const uint32_t t = (d >= o) ? tot[d - o] : 0u;
    __syncthreads();
    tot[d] += t;
    __syncthreads();
  }
  const uint32_t base = tot[d] - run;
  for (uint32_t g = 0; g < nGroups; ++g) groupSum[(size_t)g * RADIX + d] += base;
}

template <int RADIX>
__global__ void __launch_bounds__(256) k_col_apply(uint32_t* __restrict__ hist, uint32_t nBlocks, const uint32_t* __restrict__ groupSum) {
  const uint32_t b0 = blockIdx.x * COL_GB, b1 = min(b0 + COL_GB, nBlocks);
  for (int d = threadIdx.x; d < RADIX; d += 256) {
    uint32_t run = groupSum[(size_t)blockIdx.x * RADIX + d];
    for (uint32_t b = b0; b < b1; ++b) {
      const uint32_t t = hist[(size_t)b * RADIX + d];
      hist[(size_t)b * RADIX + d] = run;
      run += t;
    }
  }
}

// Stable scatter.  A warp owns SORT_IPT * 32 consecutive keys of the tile and keeps them in registers; ranks among equal
// digits come from __match_any_sync, the per-warp digit counters live in shared memory and are private to the warp, so the
// two sweeps over the keys need no block barrier (the first version synchronised the block four times per 256 keys and sat
// at 53 % of the shared-memory pipe).  permIn == nullptr: identity (first pass).
template <int RADIX>
__global__ void __launch_bounds__(SORT_NT) k_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ permIn,
                                                     uint32_t* __restrict__ keysOut, uint32_t* __restrict__ permOut, size_t n,
                                                     int shift, uint32_t mask, const uint32_t* __restrict__ blockOff) {
  __shared__ uint32_t warpCnt[SORT_WARPS][RADIX];  // count per warp and digit, then running output offset
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_NT) (&warpCnt[0][0])[i] = 0;
  __syncthreads();
  const size_t wbase = (size_t)blockIdx.x * SORT_TILE + (size_t)warp * (SORT_IPT * 32);
  uint32_t key[SORT_IPT], src[SORT_IPT], peers[SORT_IPT];
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    const size_t i = wbase + (size_t)r * 32 + lane;
    key[r] = 0xffffffffu;
    src[r] = 0;
    if (i < n) {
      key[r] = keysIn[i];
      src[r] = permIn ? permIn[i] : (uint32_t)i;
    }
  }
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    const bool valid = wbase + (size_t)r * 32 + lane < n;
    const uint32_t digit = valid ? ((key[r] >> shift) & mask) : (uint32_t)RADIX;  // invalid lanes match among themselves only
    peers[r] = __match_any_sync(0xffffffffu, digit);
    if (valid && lane == __ffs(peers[r]) - 1) warpCnt[warp][digit] += __popc(peers[r]);
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over the warps for each digit on top of the tile's first output index
  for (int d = threadIdx.x; d < RADIX; d += SORT_NT) {
    uint32_t run = blockOff[(size_t)blockIdx.x * RADIX + d];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      const uint32_t c = warpCnt[w][d];
      warpCnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_IPT; ++r) {
    const bool valid = wbase + (size_t)r * 32 + lane < n;
    const uint32_t digit = (key[r] >> shift) & mask;
    if (valid) {
      const uint32_t dst = warpCnt[warp][digit] + __popc(peers[r] & ((1u << lane) - 1u));
      keysOut[dst] = key[r];
      permOut[dst] = src[r];
    }
    __syncwarp();
    if (valid && lane == __ffs(peers[r]) - 1) warpCnt[warp][digit] += __popc(peers[r]);
    __syncwarp();
  }
}

template <typename T>
__global__ void k_gather(const T* __restrict__ in, T* __restrict__ out, const uint32_t* __restrict__ perm, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[perm[i]];
}

__global__ void k_segment_offsets(const uint32_t* __restrict__ keys, size_t n, uint32_t nKeys, int64_t* __restrict__ off) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nKeys) return;
  size_t lo = 0, hi = n;  // first i with keys[i] >= k
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1;
    else hi = mid;
  }
  off[k] = (int64_t)lo;
}

}  // namespace

cudaError_t sort_workspace_reserve(SortWorkspace& ws, size_t n) {
  if (n <= ws.capacity && ws.keysA) return cudaSuccess;
  sort_workspace_free(ws);
  if (n == 0) n = 1;
  cudaError_t e;
  const size_t nBlocks = (n + SORT_TILE - 1) / SORT_TILE;
  const size_t nh = nBlocks * 1024;   // up to 10-bit digits
  if ((e = cudaMalloc(&ws.keysA, n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&ws.keysB, n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&ws.permA, n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&ws.permB, n * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&ws.blockHist, nh * 4)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&ws.groupSum, ((nBlocks + COL_GB - 1) / COL_GB) * 1024 * 4)) != cudaSuccess) return e;
  ws.capacity = n;
  ws.histCapacity = nh;
  return cudaSuccess;
}

void sort_workspace_free(SortWorkspace& ws) {
  cudaFree(ws.keysA); cudaFree(ws.keysB); cudaFree(ws.permA); cudaFree(ws.permB);
  cudaFree(ws.blockHist); cudaFree(ws.groupSum);
  ws = SortWorkspace();
}

template <int RADIX_BITS>
static cudaError_t sort_passes(SortWorkspace& ws, const uint32_t* keys, size_t n, int passes, cudaStream_t st, uint32_t** sortedKeys,
                               uint32_t** perm, int* nLaunches) {
  constexpr int RADIX = 1 << RADIX_BITS;
  const uint32_t nBlocks = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
  const uint32_t nGroups = (nBlocks + COL_GB - 1) / COL_GB;
  const uint32_t* kin = keys;
  const uint32_t* pin = nullptr;   // identity
  uint32_t* kout = ws.keysB;
  uint32_t* pout = ws.permB;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * RADIX_BITS;
    k_hist<RADIX><<<nBlocks, SORT_NT, 0, st>>>(kin, n, shift, RADIX - 1, ws.blockHist);
    k_col_reduce<RADIX><<<nGroups, 256, 0, st>>>(ws.blockHist, nBlocks, ws.groupSum);
    k_col_scan<RADIX><<<1, RADIX, 0, st>>>(ws.groupSum, nGroups);
    k_col_apply<RADIX><<<nGroups, 256, 0, st>>>(ws.blockHist, nBlocks, ws.groupSum);
    k_scatter<RADIX><<<nBlocks, SORT_NT, 0, st>>>(kin, pin, kout, pout, n, shift, RADIX - 1, ws.blockHist);
    *nLaunches += 5;
    kin = kout;
    pin = pout;
    pout = (pout == ws.permB) ? ws.permA : ws.permB;
    kout = (kout == ws.keysB) ? ws.keysA : ws.keysB;
  }
  *sortedKeys = const_cast<uint32_t*>(kin);
  *perm = const_cast<uint32_t*>(pin);
  return cudaGetLastError();
}

cudaError_t radix_sort_by_key(SortWorkspace& ws, const uint32_t* keys, size_t n, int bits, cudaStream_t st,
                              uint32_t** sortedKeys, uint32_t** perm, int* nLaunches) {
  if (n > ws.capacity) return cudaErrorInvalidValue;
  if (n == 0) {
    *sortedKeys = ws.keysA;
    *perm = ws.permA;
    return cudaSuccess;
  }
  // digit widths: as few passes as the key width allows with 8- or 10-bit digits (64^3 elements: 19 key bits -> 2 x 10); short
  // lists (the far list of the binned layout) use 8-bit digits, see plan_digits
  if (n < ((size_t)1 << 26) && bits > 10) return sort_passes<8>(ws, keys, n, (bits + 7) / 8, st, sortedKeys, perm, nLaunches);
  if (bits <= 8) return sort_passes<8>(ws, keys, n, 1, st, sortedKeys, perm, nLaunches);
  if (bits <= 10) return sort_passes<10>(ws, keys, n, 1, st, sortedKeys, perm, nLaunches);
  if (bits <= 16) return sort_passes<8>(ws, keys, n, 2, st, sortedKeys, perm, nLaunches);
  if (bits <= 20) return sort_passes<10>(ws, keys, n, 2, st, sortedKeys, perm, nLaunches);
  if (bits <= 24) return sort_passes<8>(ws, keys, n, 3, st, sortedKeys, perm, nLaunches);
  if (bits <= 30) return sort_passes<10>(ws, keys, n, 3, st, sortedKeys, perm, nLaunches);
  return sort_passes<8>(ws, keys, n, 4, st, sortedKeys, perm, nLaunches);
}

// ---- stable sort by (keysHi, keysLo): LSD passes over the digits of keysLo, then over those of keysHi ------------------------------
// Used for the far list of the binned layout: keysHi = destination, keysLo = unique origin tag, so that the order inside a
// destination does not depend on the order in which the CTAs appended their records.
namespace {
template <int RADIX_BITS>
void one_pass(SortWorkspace& ws, const uint32_t* kin, const uint32_t* pin, uint32_t* kout, uint32_t* pout, size_t n, int shift,
              cudaStream_t st, int* nLaunches) {
  constexpr int RADIX = 1 << RADIX_BITS;
  const uint32_t nBlocks = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
  const uint32_t nGroups = (nBlocks + COL_GB - 1) / COL_GB;
  k_hist<RADIX><<<nBlocks, SORT_NT, 0, st>>>(kin, n, shift, RADIX - 1, ws.blockHist);
  k_col_reduce<RADIX><<<nGroups, 256, 0, st>>>(ws.blockHist, nBlocks, ws.groupSum);
  k_col_scan<RADIX><<<1, RADIX, 0, st>>>(ws.groupSum, nGroups);
  k_col_apply<RADIX><<<nGroups, 256, 0, st>>>(ws.blockHist, nBlocks, ws.groupSum);
  k_scatter<RADIX><<<nBlocks, SORT_NT, 0, st>>>(kin, pin, kout, pout, n, shift, RADIX - 1, ws.blockHist);
  *nLaunches += 5;
}
// far lists are a few per cent of the particles: with so few tiles the column kernels of a 10-bit pass (1024-entry rows) run on a
// handful of CTAs and cost 3.5 x an 8-bit pass (measured at 4.4e6 keys: 270 vs 77 us), so short lists use 8-bit digits throughout
void plan_digits(int bits, int* width, int* passes, size_t n) {
  if (n < ((size_t)1 << 26)) { *width = 8; *passes = (bits + 7) / 8; return; }
  if (bits <= 8) { *width = 8; *passes = 1; }
  else if (bits <= 10) { *width = 10; *passes = 1; }
  else if (bits <= 16) { *width = 8; *passes = 2; }
  else if (bits <= 20) { *width = 10; *passes = 2; }
  else if (bits <= 24) { *width = 8; *passes = 3; }
  else if (bits <= 30) { *width = 10; *passes = 3; }
  else { *width = 8; *passes = 4; }
}
}  // namespace

cudaError_t radix_sort_two_keys(SortWorkspace& ws, const uint32_t* keysLo, int bitsLo, const uint32_t* keysHi, int bitsHi, size_t n,
                                cudaStream_t st, uint32_t** sortedHi, uint32_t** perm, int* nLaunches) {
  if (n > ws.capacity) return cudaErrorInvalidValue;
  if (n == 0) {
    *sortedHi = ws.keysA;
    *perm = ws.permA;
    return cudaSuccess;
  }
  const uint32_t* kin = keysLo;
  const uint32_t* pin = nullptr;
  uint32_t* kout = ws.keysB;
  uint32_t* pout = ws.permB;
  auto advance = [&]() {
    kin = kout;
    pin = pout;
    pout = (pout == ws.permB) ? ws.permA : ws.permB;
    kout = (kout == ws.keysB) ? ws.keysA : ws.keysB;
  };
  int w, np;
  plan_digits(bitsLo, &w, &np, n);
  for (int p = 0; p < np; ++p) {
    if (w == 8) one_pass<8>(ws, kin, pin, kout, pout, n, p * 8, st, nLaunches);
    else one_pass<10>(ws, kin, pin, kout, pout, n, p * 10, st, nLaunches);
    advance();
  }
  // destination keys in the order reached so far (kout is free: it holds the keys of the pass before the last one)
  k_gather<uint32_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keysHi, kout, pin, n);
  ++*nLaunches;
  kin = kout;
  kout = (kout == ws.keysB) ? ws.keysA : ws.keysB;
  plan_digits(bitsHi, &w, &np, n);
  for (int p = 0; p < np; ++p) {
    if (w == 8) one_pass<8>(ws, kin, pin, kout, pout, n, p * 8, st, nLaunches);
    else one_pass<10>(ws, kin, pin, kout, pout, n, p * 10, st, nLaunches);
    advance();
  }
  *sortedHi = const_cast<uint32_t*>(kin);
  *perm = const_cast<uint32_t*>(pin);
  return cudaGetLastError();
}

// all particle arrays in one pass: the permutation is read once
__global__ void k_gather_particles(const double* __restrict__ x0, const double* __restrict__ x1, const double* __restrict__ x2,
                                   const double* __restrict__ v0, const double* __restrict__ v1, const double* __restrict__ v2,
                                   const int32_t* __restrict__ elem, const uint8_t* __restrict__ meta, const int64_t* __restrict__ id,
                                   double* __restrict__ y0, double* __restrict__ y1, double* __restrict__ y2, double* __restrict__ w0,
                                   double* __restrict__ w1, double* __restrict__ w2, int32_t* __restrict__ elemOut,
                                   uint8_t* __restrict__ metaOut, int64_t* __restrict__ idOut, const uint32_t* __restrict__ perm, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = perm[i];
  const double a0 = x0[s], a1 = x1[s], a2 = x2[s], b0 = v0[s], b1 = v1[s], b2 = v2[s];
  const uint8_t m = meta[s];
  y0[i] = a0; y1[i] = a1; y2[i] = a2;
  w0[i] = b0; w1[i] = b1; w2[i] = b2;
  if (elem) elemOut[i] = elem[s];   // null: the caller regenerates the element ids from the sorted segments
  metaOut[i] = m;
  if (id) idOut[i] = id[s];
}

cudaError_t gather_particles(const double* const x[3], const double* const v[3], const int32_t* elem, const uint8_t* meta,
                             const int64_t* id, double* const y[3], double* const w[3], int32_t* elemOut, uint8_t* metaOut,
                             int64_t* idOut, const uint32_t* perm, size_t n, cudaStream_t st) {
  if (n)
    k_gather_particles<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x[0], x[1], x[2], v[0], v[1], v[2], elem, meta, id, y[0], y[1],
                                                                     y[2], w[0], w[1], w[2], elemOut, metaOut, idOut, perm, n);
  return cudaGetLastError();
}

cudaError_t gather_f64(const double* in, double* out, const uint32_t* perm, size_t n, cudaStream_t st) {
  if (n) k_gather<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, perm, n);
  return cudaGetLastError();
}
cudaError_t gather_i32(const int32_t* in, int32_t* out, const uint32_t* perm, size_t n, cudaStream_t st) {
  if (n) k_gather<int32_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, perm, n);
  return cudaGetLastError();
}
cudaError_t gather_u8(const uint8_t* in, uint8_t* out, const uint32_t* perm, size_t n, cudaStream_t st) {
  if (n) k_gather<uint8_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, perm, n);
  return cudaGetLastError();
}
cudaError_t gather_i64(const int64_t* in, int64_t* out, const uint32_t* perm, size_t n, cudaStream_t st) {
  if (n) k_gather<int64_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, perm, n);
  return cudaGetLastError();
}

cudaError_t segment_offsets(const uint32_t* sortedKeys, size_t n, uint32_t nKeys, int64_t* off, cudaStream_t st) {
  k_segment_offsets<<<(nKeys + 1 + 255) / 256, 256, 0, st>>>(sortedKeys, n, nKeys, off);
  return cudaGetLastError();
}
