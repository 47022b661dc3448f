// piclas_gpu.cu — C ABI (include/piclas_gpu.h) of the B200 particle step: context, host<->device transfers, step driver.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"
#include "sf.cuh"
#include "ref.cuh"
#include "sort.cuh"
#include "bins.cuh"
#include "emit.cuh"

namespace {

enum { ELEM_FIRSTSIDEIND = 2, ELEM_LASTSIDEIND = 3, ELEM_FIRSTNODEIND = 4, ELEM_LASTNODEIND = 5, ELEM_RANK = 6 };
enum { SIDE_NBELEMID = 2, SIDE_BCID = 4, SIDE_LOCALID = 6 };

struct Ctx {
  bool ready = false;
  int device = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evp[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // phase marks
  double phaseMs[4] = {0., 0., 0., 0.};  // deposit particle kernel, deposit node/DOF kernels, push+track kernel, sort+permute
  pgpu_params_t prm;
  int nGlobalElems = 0, nElems = 0, offsetElem = 0, N = 0, NP = 0, ND = 0, nNodes = 0, nRanks = 1, myRank = 0;
  int nSMs = 148;
  // mesh records
  TriaElem* dTria = nullptr;
  GeoElem* dGeo = nullptr;
  PlaneElem* dPlanes = nullptr;   // arithmetic == 1 only
  AffElem* dAff = nullptr;        // arithmetic == 1 only
  bool fast = false;
  int32_t* dElemRank = nullptr;
  double* dElemXGP = nullptr;   // Elem_xGP (global) — inverse-distance fallback only
  // CVWM
  int32_t *dAdjOff = nullptr, *dAdj = nullptr, *dElemNodeU = nullptr;
  int32_t *dPerN = nullptr, *dPerOff = nullptr, *dPerNodes = nullptr;
  double *dNodeVolume = nullptr, *dElemAcc = nullptr, *dS = nullptr, *dNodeSource = nullptr, *dPartSource = nullptr;
  double* dCharge = nullptr;   // PartSource(4,:) of the local elements, compact (piclas_gpu_get_charge)
  // RefMapping
  bool ref = false;
  RefTables refT;
  double* dXiB[3] = {nullptr, nullptr, nullptr};   // second PartPosRef buffer (RefMapping: PartPosRef is particle state)
  int32_t *dElemToBCSides = nullptr, *dSideInfo = nullptr;
  double *dSideBCMetrics = nullptr, *dSideNormVec = nullptr, *dSideDistance = nullptr, *dBV0 = nullptr, *dBV1 = nullptr, *dBV2 = nullptr, *dBV3 = nullptr;
  int32_t* dSideType = nullptr;
  double *dElemRadius2 = nullptr, *dElemEpsOneCell = nullptr;
  // shape function
  bool sfActive = false;
  SFTables sfT;
  int32_t *dFibN = nullptr, *dFibOff = nullptr, *dFibElem = nullptr, *dElemToBGM = nullptr, *dCandOff = nullptr, *dCandSrc = nullptr;
  uint8_t* dCandCase = nullptr;
  double *dElemBary = nullptr, *dElemRadius = nullptr, *dElemsJ = nullptr, *dSFElemr2 = nullptr;
  double* dSfFac[4] = {nullptr, nullptr, nullptr, nullptr};
  int nSfTargets = 0, nSfHalo = 0;               // gather targets: nElems local + nSfHalo elements of other ranks
  int32_t *dSfTarget = nullptr, *dSfRecvElem = nullptr, *dSfRecvOff = nullptr, *dSfRecvIdx = nullptr;
  int nSfRecvElems = 0;                            // distinct local elements that receive halo contributions
  double* dSfRecv = nullptr;                       // received halo contributions [nRecvElems][ND][4]
  std::vector<int64_t> sfSendCount, sfRecvCount;   // elements per rank
  int64_t sfRecvTotal = 0;
  int64_t sfFacCap = 0;
  // field
  double* dE = nullptr;
  double* dEmono = nullptr;   // the same polynomials in the monomial basis (restructured arithmetic on the bins, bins.cuh)
  bool haveField = false;
  // particles
  PartBuf buf[2];
  double* dXi[3] = {nullptr, nullptr, nullptr};  // cached reference positions (never permuted: recomputed after a sort)
  int cur = 0;
  int64_t cap = 0, nPart = 0, nTotalSorted = 0;
  bool carryIDs = false;
  bool xiValid = false;
  int64_t* dElemOff = nullptr;        // [nElems + nRanks + 2]
  std::vector<int64_t> hTailOff;      // host copy of elemOff[nElems .. nElems+nRanks+1]
  uint32_t* dKeys = nullptr;          // [cap]
  int* dCounters = nullptr;           // [8]: lost, error, leavers handed to the walk kernel, work counter, queue A, queue B, -, -
  SortWorkspace sortws;
  int keyBits = 1;
  // staging for AoS transfers
  double* dStage = nullptr;
  int32_t* dStageI = nullptr;
  int64_t* dStageL = nullptr;
  int64_t stageCap = 0;
  // particle exchange between ranks (AoS messages as particle_mpi.f90:472-502)
  double *dCommSend = nullptr, *dCommRecv = nullptr;
  // multi-rank: emigrants are extracted from the unsorted arrays after tracking, the one sort of the step follows the exchange
  bool exchangePending = false;
  int64_t nUnsorted = 0, nEmig = 0;
  uint32_t *dTileCnt = nullptr, *dEmigIdx = nullptr, *dEmigKey = nullptr;
  int64_t tileCntCap = 0, emigCap = 0;
  int64_t* dEmigOff = nullptr;   // [nRanks+1]
  int64_t commSendCap = 0, commRecvCap = 0;
  int commSize = 8;
  // compact node halo of cell_volweight_mean (nRanks > 1): the unique nodes whose sums need contributions of several ranks
  int32_t* dHaloNodes = nullptr;
  int nHaloNodes = 0;
  double *dHaloSend = nullptr, *dHaloRecv = nullptr;   // [nHaloNodes][4], [nRanks][nHaloNodes][4]
  bool haloPacked = false;
  bool ownStream = true;
  int64_t* hPin = nullptr;          // pinned host words for the small device->host readbacks of a step (counters, offsets)
  int64_t* dSendCounts = nullptr;   // [nRanks] emigrant counts of the open step on the device (for a device-side count exchange)
  int* dEmigCnt = nullptr;          // [nRanks] counted by k_far_walk
  std::vector<int64_t> hEmigCnt;    // host copy (valid while the exchange of a binned step is pending)
  // binned layout (bins.cuh): TriaTracking + cell_volweight_mean (or no deposition)
  bool binEligible = false;      // this configuration steps on the bins
  bool binned = false;           // the particles live in the bins (else: sorted arrays buf[cur] + dElemOff)
  bool sortedViewValid = false;  // binned, and buf[0] + dElemOff hold a current copy in sorted order (download, analysis)
  bool allHot = false;           // every local element affine + planar, B = 0, restructured arithmetic: call-free push kernel
  bool wantRebin = false;        // regions overflowed in the last step: re-plan the capacities before the next one
  PartBuf bins;
  int64_t binSlotsCap = 0;
  PartBuf pool[2];
  int64_t poolCap[2] = {0, 0};
  int64_t* dPoolOff[2] = {nullptr, nullptr};   // [nElems + nRanks + 2] each
  int32_t *dCapMain = nullptr, *dCapIn = nullptr, *dNMain = nullptr, *dNIn = nullptr;
  int64_t *dBinBase = nullptr, *dBinTmp = nullptr;
  int64_t *dFarBase = nullptr, *dFarDOff = nullptr, *dScanSums = nullptr, *dScanOff = nullptr;   // far regions, dense offsets, scan scratch
  int32_t* dNFarE = nullptr;     // [nElems][2] far records / diverted particles of every element in the last step
  uint32_t* dFarIdx = nullptr;   // [cap] used far slots, densely listed
  int64_t farIdxCap = 0;
  PushElem* dPushElem = nullptr;
  RefTables locT{};                   // FIBGM + barycentres + radii for SinglePointToElement (emission); zero when the mesh has no FIBGM
  double globLo[3] = {0., 0., 0.}, globHi[3] = {0., 0., 0.};   // GEO%xminglob .. zmaxglob
  cudaStream_t stCopy = nullptr;      // piclas_gpu_get_partsource_async: device -> host copy beside the step's own stream
  cudaEvent_t evDepoDone = nullptr, evCopyDone = nullptr;
  bool psCopyPending = false;
  HintNb* dHintNb = nullptr;
  uint32_t* dFarPend = nullptr;   // dense indices of the far records k_far_hint left to the exact walk
  int64_t farPendCap = 0;
  int binParity = 0;
  int64_t nFar = 0;              // records of the far list of the open step
  double mainSlack = 0.10, inFrac = 0.09;
  bool inFracCapped = false;     // a larger inbox share did not fit into the memory: no re-plan for full inboxes any more
  double detourMs = 0.;          // time the diverted particles have cost since the bins were planned (estimate)
  int64_t farStats[4] = {0, 0, 0, 0};   // last step: far records, side movers, overflowed main, overflowed inbox
  // timing
  double lastMs = 0.;
  int lastLaunches = 0;
};

Ctx g;
std::string g_err;

// NVTX ranges named after the reference's load-balance sections (LB_* of piclas.h:298-321, loadbalance_timers.f90:70-266), so that
// an nsys / ncu timeline of the host shows the same sections the Fortran timers report (header-only NVTX3: no library to link)
struct LbRange {
  explicit LbRange(const char* name) { nvtxRangePushA(name); }
  ~LbRange() { nvtxRangePop(); }
};

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
                                       cudaGetErrorString(e_));                                        \
  } while (0)

template <typename T>
int upload(T** d, const T* h, size_t n) {
  CK(cudaMalloc((void**)d, (n ? n : 1) * sizeof(T)));
  if (n) CK(cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int alloc_partbuf(PartBuf& b, int64_t cap, bool ids) {
  const int64_t stride = (cap + 31) & ~(int64_t)31;   // keeps every component 256-byte aligned
  b.f = nullptr;
  CK(cudaMalloc((void**)&b.f, (size_t)stride * 6 * 8));
  b.stride = stride;
  b.xif = nullptr;
  for (int d = 0; d < 3; ++d) {
    b.x[d] = b.f + (size_t)d * stride;
    b.v[d] = b.f + (size_t)(3 + d) * stride;
    b.xi[d] = nullptr;
  }
  CK(cudaMalloc((void**)&b.elem, cap * 4));
  CK(cudaMalloc((void**)&b.meta, cap));
  b.id = nullptr;
  if (ids) CK(cudaMalloc((void**)&b.id, cap * 8));
  return 0;
}
void free_partbuf(PartBuf& b) {
  cudaFree(b.f);
  b.f = nullptr;
  for (int d = 0; d < 3; ++d) { b.x[d] = b.v[d] = b.xi[d] = nullptr; }
  cudaFree(b.elem); cudaFree(b.meta); cudaFree(b.id);
  b.elem = nullptr; b.meta = nullptr; b.id = nullptr;
}

int reserve_particles(int64_t need) {
  if (need <= g.cap) return 0;
  int64_t ncap = need + need / 8 + 1024;
  if (g.prm.maxParticleNumber > ncap) ncap = g.prm.maxParticleNumber;
  PartBuf nb[2];
  double* oldXi[3] = {nullptr, nullptr, nullptr};
  if (g.cap > 0) for (int d = 0; d < 3; ++d) oldXi[d] = g.buf[g.cur].xi[d];
  if (alloc_partbuf(nb[0], ncap, g.carryIDs)) return 1;
  if (alloc_partbuf(nb[1], ncap, g.carryIDs)) return 1;
  if (g.cap > 0 && g.nPart > 0) {
    const PartBuf& o = g.buf[g.cur];
    for (int d = 0; d < 3; ++d) {
      CK(cudaMemcpy(nb[0].x[d], o.x[d], g.nPart * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(nb[0].v[d], o.v[d], g.nPart * 8, cudaMemcpyDeviceToDevice));
    }
    CK(cudaMemcpy(nb[0].elem, o.elem, g.nPart * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(nb[0].meta, o.meta, g.nPart, cudaMemcpyDeviceToDevice));
    if (g.carryIDs) CK(cudaMemcpy(nb[0].id, o.id, g.nPart * 8, cudaMemcpyDeviceToDevice));
  }
  if (g.cap > 0) {
    free_partbuf(g.buf[0]);
    free_partbuf(g.buf[1]);
  }
  if (!g.binEligible) {
    const int64_t stride = nb[0].stride;
    double* nx = nullptr;
    CK(cudaMalloc((void**)&nx, (size_t)stride * 3 * 8));
    for (int d = 0; d < 3; ++d)
      if (g.ref && g.cap > 0 && g.nPart > 0 && oldXi[d]) CK(cudaMemcpy(nx + (size_t)d * stride, oldXi[d], g.nPart * 8, cudaMemcpyDeviceToDevice));
    cudaFree(g.dXi[0]);
    for (int d = 0; d < 3; ++d) {
      g.dXi[d] = nx + (size_t)d * stride;
      nb[0].xi[d] = nb[1].xi[d] = g.dXi[d];
    }
    nb[0].xif = nb[1].xif = nx;
    if (g.ref) {
      cudaFree(g.dXiB[0]);
      double* nb2 = nullptr;
      CK(cudaMalloc((void**)&nb2, (size_t)stride * 3 * 8));
      for (int d = 0; d < 3; ++d) { g.dXiB[d] = nb2 + (size_t)d * stride; nb[1].xi[d] = g.dXiB[d]; }
      nb[1].xif = nb2;
    }
  }
  g.buf[0] = nb[0];
  g.buf[1] = nb[1];
  g.cur = 0;
  g.cap = ncap;
  {
    uint32_t* nk = nullptr;
    CK(cudaMalloc((void**)&nk, ncap * 4));
    if (g.exchangePending && g.dKeys) CK(cudaMemcpy(nk, g.dKeys, g.nUnsorted * 4, cudaMemcpyDeviceToDevice));   // keys of the open step
    cudaFree(g.dKeys);
    g.dKeys = nk;
  }
  CK(sort_workspace_reserve(g.sortws, (size_t)ncap));
  g.xiValid = false;
  return 0;
}

int reserve_stage(int64_t n) {
  if (n <= g.stageCap) return 0;
  cudaFree(g.dStage); cudaFree(g.dStageI); cudaFree(g.dStageL);
  CK(cudaMalloc((void**)&g.dStage, n * 9 * 8));
  CK(cudaMalloc((void**)&g.dStageI, n * 4 * 4));
  CK(cudaMalloc((void**)&g.dStageL, n * 8));
  g.stageCap = n;
  return 0;
}

__global__ void k_aos_to_soa(PartBuf pb, int64_t dst0, int64_t n, const double* __restrict__ ps, const int32_t* __restrict__ spec,
                             const int32_t* __restrict__ elem, const int32_t* __restrict__ inside, const int32_t* __restrict__ isnew,
                             const int64_t* __restrict__ ids, int64_t idBase, const double* __restrict__ ref, int nGlobalElems,
                             int nSpecies, int* __restrict__ badInput) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = dst0 + i;
  {  // host input indexes device tables: 1 <= GlobalElemID <= nGlobalElems (live particles), 1 <= PartSpecies <= nSpecies
    const bool live = inside ? (inside[i] != 0) : true;
    if (live && (elem[i] < 1 || elem[i] > nGlobalElems)) atomicOr(badInput, 1);
    if (live && (spec[i] < 1 || spec[i] > nSpecies)) atomicOr(badInput, 2);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    pb.x[d][p] = ps[i * 6 + d];
    pb.v[d][p] = ps[i * 6 + 3 + d];
  }
  const bool in = (inside ? (inside[i] != 0) : true) && elem[i] >= 1 && elem[i] <= nGlobalElems;
  pb.elem[p] = in ? elem[i] : 0;
  pb.meta[p] = (uint8_t)(((spec[i] - 1) & META_SPEC_MASK) | ((isnew && isnew[i]) ? META_ISNEW : 0));
  if (pb.id) pb.id[p] = ids ? ids[i] : (idBase + i);
  if (ref) {
#pragma unroll
    for (int d = 0; d < 3; ++d) pb.xi[d][p] = ref[i * 3 + d];
  }
}

__global__ void k_soa_to_aos(PartBuf pb, int64_t src0, int64_t n, double* __restrict__ ps, int32_t* __restrict__ spec,
                             int32_t* __restrict__ elem, double* __restrict__ xi, int64_t* __restrict__ ids) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = src0 + i;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    ps[i * 6 + d] = pb.x[d][p];
    ps[i * 6 + 3 + d] = pb.v[d][p];
    if (xi) xi[i * 3 + d] = pb.xi[d][p];
  }
  spec[i] = (pb.meta[p] & META_SPEC_MASK) + 1;
  elem[i] = pb.elem[p];
  if (ids) ids[i] = pb.id ? pb.id[p] : -1;
}

// PartPosRef(1:3) for a download under TriaTracking: GetPositionInRefElem (ForceMode, as DepositionMethod_CVWM calls it) at the
// particle's current position in its element; closed form on affine elements with the restructured arithmetic
__global__ void k_posref_download(PartBuf pb, int64_t src0, int64_t n, const GeoElem* __restrict__ geo, const AffElem* __restrict__ aff,
                                  double* __restrict__ xiOut) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = src0 + i;
  const int e = pb.elem[p];
  double xi[3] = {0., 0., 0.};
  if (e >= 1) {
    const double x[3] = {pb.x[0][p], pb.x[1][p], pb.x[2][p]};
    if (aff) ref_position_fast(aff + (e - 1), geo + (e - 1), x, xi, true);
    else position_in_ref_elem(geo + (e - 1), x, xi, true, true);
  }
  xiOut[i * 3 + 0] = xi[0]; xiOut[i * 3 + 1] = xi[1]; xiOut[i * 3 + 2] = xi[2];
}

// MPIParticleRecv unpack (particle_mpi.f90:831-989): received particles are appended; IsNewPart = F
__global__ void k_unpack_immigrants(PartBuf pb, int64_t dst0, int64_t n, int cs, int withRef, const double* __restrict__ buf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = dst0 + i;
  const double* b = buf + i * cs;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    pb.x[d][p] = b[d];
    pb.v[d][p] = b[3 + d];
  }
  int o = 6;
  if (withRef) { pb.xi[0][p] = b[6]; pb.xi[1][p] = b[7]; pb.xi[2][p] = b[8]; o = 9; }
  pb.meta[p] = (uint8_t)(((int)b[o] - 1) & META_SPEC_MASK);
  pb.elem[p] = (int)b[o + 1];
  if (pb.id) pb.id[p] = (cs > o + 2) ? __double_as_longlong(b[o + 2]) : -1;
}

// ---- emigrant extraction (multi-rank): stable compaction of the particles whose key addresses another rank -----------------------
constexpr int EM_NT = 256, EM_IPT = 8, EM_TILE = EM_NT * EM_IPT;

__global__ void __launch_bounds__(EM_NT) k_emig_count(const uint32_t* __restrict__ keys, int64_t n, uint32_t lo, uint32_t hi,
                                                      uint32_t* __restrict__ tileCnt) {
  __shared__ uint32_t ws[EM_NT / 32];
  const int64_t base = (int64_t)blockIdx.x * EM_TILE;
  uint32_t c = 0;
#pragma unroll
  for (int r = 0; r < EM_IPT; ++r) {
    const int64_t i = base + (int64_t)r * EM_NT + threadIdx.x;
    if (i < n) {
      const uint32_t k = keys[i];
      c += (k >= lo && k < hi) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < EM_NT / 32; ++w) t += ws[w];
    tileCnt[blockIdx.x] = t;
  }
}

// exclusive scan of the tile counts by one CTA; total -> *total
__global__ void __launch_bounds__(1024) k_emig_scan(uint32_t* __restrict__ tileCnt, uint32_t nTiles, int64_t* __restrict__ total) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nTiles; b0 += 1024) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t v = i < nTiles ? tileCnt[i] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nTiles) tileCnt[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// emigIdx / emigKey in particle order (thread t owns EM_IPT consecutive keys of the tile)
__global__ void __launch_bounds__(EM_NT) k_emig_compact(const uint32_t* __restrict__ keys, int64_t n, uint32_t lo, uint32_t hi,
                                                        const uint32_t* __restrict__ tileOff, uint32_t* __restrict__ emigIdx,
                                                        uint32_t* __restrict__ emigKey) {
  __shared__ uint32_t ws[EM_NT / 32];
  const int64_t base = (int64_t)blockIdx.x * EM_TILE + (int64_t)threadIdx.x * EM_IPT;
  uint32_t k[EM_IPT];
  uint32_t c = 0;
#pragma unroll
  for (int r = 0; r < EM_IPT; ++r) {
    k[r] = (base + r < n) ? keys[base + r] : 0xffffffffu;
    c += (k[r] >= lo && k[r] < hi) ? 1u : 0u;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  uint32_t off = tileOff[blockIdx.x] + inc - c;
  for (int w = 0; w < warp; ++w) off += ws[w];
#pragma unroll
  for (int r = 0; r < EM_IPT; ++r)
    if (k[r] >= lo && k[r] < hi) {
      emigIdx[off] = (uint32_t)(base + r);
      emigKey[off] = k[r] - lo;   // destination rank
      ++off;
    }
}

// message layout of one migrating particle (particle_mpi.f90:472-502, no LSERK/vMPF/DSMC): PartState(1:6) [, PartPosRef(1:3) with
// RefMapping], REAL(PartSpecies), REAL(PEM%GlobalElemID) [, particle id bits when ids are carried].
// message i <- particle emigIdx[perm[i]] (grouped by destination rank); the particle's key becomes "removed"
__global__ void k_pack_emigrants_idx(PartBuf pb, const uint32_t* __restrict__ emigIdx, const uint32_t* __restrict__ perm, int64_t n, int cs,
                                     int withRef, double* __restrict__ buf, uint32_t* __restrict__ keys, uint32_t removedKey,
                                     const uint32_t* __restrict__ slotOf /*bins: key index -> slot of the far list; else null*/) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t kp = emigIdx[perm[i]];
  const int64_t p = slotOf ? (int64_t)slotOf[kp] : kp;
  double* b = buf + i * cs;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    b[d] = pb.x[d][p];
    b[3 + d] = pb.v[d][p];
  }
  int o = 6;
  if (withRef) { b[6] = pb.xi[0][p]; b[7] = pb.xi[1][p]; b[8] = pb.xi[2][p]; o = 9; }   // PartPosRef (RefMapping)
  b[o] = (double)((pb.meta[p] & META_SPEC_MASK) + 1);
  b[o + 1] = (double)pb.elem[p];
  if (cs > o + 2) b[o + 2] = __longlong_as_double(pb.id ? pb.id[p] : -1);
  keys[kp] = removedKey;
}

// PEM%GlobalElemID of the sorted particles follows from their segment: cheaper to write than to gather (a gathered mover costs a
// 32-byte sector for 4 bytes)
__global__ void k_fill_elem(const int64_t* __restrict__ elemOff, int nElems, int offsetElem, int32_t* __restrict__ elem) {
  for (int e = blockIdx.x; e < nElems; e += gridDim.x) {
    const int64_t p0 = elemOff[e], p1 = elemOff[e + 1];
    for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) elem[p] = offsetElem + e + 1;
  }
}

// sort the first nIn particles of the current buffer by key (keys already in g.dKeys), gather into the other buffer
int sort_and_permute(int64_t nIn) {
  uint32_t *sk = nullptr, *perm = nullptr;
  CK(radix_sort_by_key(g.sortws, g.dKeys, (size_t)nIn, g.keyBits, g.st, &sk, &perm, &g.lastLaunches));
  const PartBuf& a = g.buf[g.cur];
  PartBuf& b = g.buf[g.cur ^ 1];
  CK(gather_particles(a.x, a.v, nullptr, a.meta, g.carryIDs ? a.id : nullptr, b.x, b.v, b.elem, b.meta, b.id, perm, (size_t)nIn, g.st));
  ++g.lastLaunches;
  if (g.ref) {
    for (int d = 0; d < 3; ++d) CK(gather_f64(a.xi[d], b.xi[d], perm, (size_t)nIn, g.st));
    g.lastLaunches += 3;
  }
  const uint32_t nKeys = (uint32_t)(g.nElems + g.nRanks + 1);
  CK(segment_offsets(sk, (size_t)nIn, nKeys, g.dElemOff, g.st));
  if (g.nElems > 0) k_fill_elem<<<g.nElems < g.nSMs * 16 ? g.nElems : g.nSMs * 16, 128, 0, g.st>>>(g.dElemOff, g.nElems, g.offsetElem, b.elem);
  g.lastLaunches += 2;
  g.hTailOff.assign(g.nRanks + 2, 0);
  CK(cudaMemcpyAsync(g.hTailOff.data(), g.dElemOff + g.nElems, (g.nRanks + 2) * sizeof(int64_t), cudaMemcpyDeviceToHost, g.st));
  CK(cudaStreamSynchronize(g.st));
  g.cur ^= 1;
  g.nPart = g.hTailOff[0];                  // particles owned by this rank
  g.nTotalSorted = g.hTailOff[g.nRanks];    // + emigrants grouped by destination rank behind them
  g.xiValid = false;
  return 0;
}

template <int NP, bool FAST>
void launch_push_track_t(double dt) {
  const int grid = g.nElems < g.nSMs * 8 ? g.nElems : g.nSMs * 8;
  const PartBuf& o = g.buf[g.cur ^ 1];
  uint32_t* leaverIdx = g.sortws.permB;  // idle until the sort that follows
  uint32_t *leaverHistE = g.sortws.keysB, *leaverHistS = g.sortws.permA;
  if (g.ref) {
    k_interp_push<NP, FAST, true><<<grid, STEP_NT, 0, g.st>>>(g.buf[g.cur], o.x[0], o.x[1], o.x[2], g.dElemOff, g.nElems, g.offsetElem,
                                                             g.dGeo, g.dTria, g.dPlanes, g.dAff, g.dE, g.dElemXGP, g.dKeys, leaverIdx,
                                                             leaverHistE, leaverHistS, g.dElemRank, dt, 1, g.dCounters);
    k_track_ref<<<g.nSMs * 8, 128, 0, g.st>>>(g.buf[g.cur], o.x[0], o.x[1], o.x[2], g.nPart, g.refT, g.dElemRank, g.dKeys, g.nElems,
                                              g.offsetElem, g.dCounters);
  } else {
    k_interp_push<NP, FAST, false><<<grid, STEP_NT, 0, g.st>>>(g.buf[g.cur], o.x[0], o.x[1], o.x[2], g.dElemOff, g.nElems, g.offsetElem,
                                                              g.dGeo, g.dTria, g.dPlanes, g.dAff, g.dE, g.dElemXGP, g.dKeys, leaverIdx,
                                                              leaverHistE, leaverHistS, g.dElemRank, dt, g.xiValid ? 1 : 0, g.dCounters);
    k_track_leavers<FAST><<<g.nSMs * 8, LV_NT, 0, g.st>>>(g.buf[g.cur], o.x[0], o.x[1], o.x[2], leaverIdx, leaverHistE, leaverHistS, g.dTria, g.dPlanes,
                                                        g.dElemRank, g.dKeys, g.nElems, g.offsetElem, g.dCounters);
  }
  g.lastLaunches += 2;
}
template <int NP>
void launch_push_track(double dt) {
  if (g.fast) launch_push_track_t<NP, true>(dt);
  else launch_push_track_t<NP, false>(dt);
}
template <int NP>
void launch_dofs() {
  const size_t total = (size_t)g.nElems * NP * NP * NP * 4;
  k_nodes_to_dofs<NP><<<(unsigned)((total + 255) / 256), 256, 0, g.st>>>(g.dNodeSource, g.dElemNodeU, g.dPartSource, g.nElems);
}

// grows the sorted buffers (and what is sized with them) while the far list of an open step lives in them
int reserve_far(int64_t need, int64_t keep) {
  if (need <= g.cap) return 0;
  int64_t ncap = need + need / 8 + 1024;
  PartBuf nb[2];
  if (alloc_partbuf(nb[0], ncap, g.carryIDs)) return 1;
  if (alloc_partbuf(nb[1], ncap, g.carryIDs)) return 1;
  uint32_t* nk = nullptr;
  CK(cudaMalloc((void**)&nk, ncap * 4));
  if (keep > 0 && g.cap > 0) {
    for (int d = 0; d < 3; ++d) {
      CK(cudaMemcpy(nb[0].x[d], g.buf[0].x[d], keep * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(nb[0].v[d], g.buf[0].v[d], keep * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(nb[1].x[d], g.buf[1].x[d], keep * 8, cudaMemcpyDeviceToDevice));
    }
    CK(cudaMemcpy(nb[0].elem, g.buf[0].elem, keep * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(nb[1].elem, g.buf[1].elem, keep * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(nb[0].meta, g.buf[0].meta, keep, cudaMemcpyDeviceToDevice));
    if (g.carryIDs) CK(cudaMemcpy(nb[0].id, g.buf[0].id, keep * 8, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(nk, g.dKeys, keep * 4, cudaMemcpyDeviceToDevice));
  }
  {
    uint32_t* ni = nullptr;
    CK(cudaMalloc((void**)&ni, ncap * 4));
    if (keep > 0 && g.dFarIdx) CK(cudaMemcpy(ni, g.dFarIdx, (keep < g.farIdxCap ? keep : g.farIdxCap) * 4, cudaMemcpyDeviceToDevice));
    cudaFree(g.dFarIdx);
    g.dFarIdx = ni;
    g.farIdxCap = ncap;
  }
  if (g.cap > 0) { free_partbuf(g.buf[0]); free_partbuf(g.buf[1]); }
  cudaFree(g.dKeys);
  g.dKeys = nk;
  g.buf[0] = nb[0];
  g.buf[1] = nb[1];
  g.cap = ncap;
  g.cur = 0;
  g.sortedViewValid = false;
  CK(sort_workspace_reserve(g.sortws, (size_t)ncap));
  return 0;
}

#include "bins_host.inc"

void begin_timing() {
  g.lastLaunches = 0;
  cudaEventRecord(g.ev0, g.st);
}
void end_timing() {
  cudaEventRecord(g.ev1, g.st);
  cudaEventSynchronize(g.ev1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, g.ev0, g.ev1);
  g.lastMs = ms;
}

}  // namespace

extern "C" {

const char* piclas_gpu_last_error(void) { return g_err.c_str(); }

int piclas_gpu_finalize(void) {
  if (!g.ready && !g.st) return 0;
  cudaSetDevice(g.device);
  cudaDeviceSynchronize();
  cudaFree(g.dTria); cudaFree(g.dGeo); cudaFree(g.dPlanes); cudaFree(g.dAff); cudaFree(g.dElemRank); cudaFree(g.dElemXGP);
  cudaFree(g.dAdjOff); cudaFree(g.dAdj); cudaFree(g.dElemNodeU); cudaFree(g.dPerN); cudaFree(g.dPerOff); cudaFree(g.dPerNodes);
  cudaFree(g.dNodeVolume); cudaFree(g.dElemAcc); cudaFree(g.dS); cudaFree(g.dNodeSource); cudaFree(g.dPartSource); cudaFree(g.dCharge);
  cudaFree(g.dE); cudaFree(g.dEmono); cudaFree(g.dElemOff); cudaFree(g.dKeys); cudaFree(g.dCounters);
  cudaFree(g.dStage); cudaFree(g.dStageI); cudaFree(g.dStageL);
  cudaFree(g.dCommSend); cudaFree(g.dCommRecv);
  cudaFree(g.dHaloNodes); cudaFree(g.dHaloSend); cudaFree(g.dHaloRecv);
  cudaFree(g.dTileCnt); cudaFree(g.dEmigIdx); cudaFree(g.dEmigKey); cudaFree(g.dEmigOff);
  cudaFree(g.dFibN); cudaFree(g.dFibOff); cudaFree(g.dFibElem); cudaFree(g.dElemToBGM); cudaFree(g.dCandOff); cudaFree(g.dCandSrc);
  cudaFree(g.dCandCase); cudaFree(g.dElemBary); cudaFree(g.dElemRadius); cudaFree(g.dElemsJ); cudaFree(g.dSFElemr2);
  for (int c = 0; c < 4; ++c) cudaFree(g.dSfFac[c]);
  cudaFree(g.dSfTarget); cudaFree(g.dSfRecvElem); cudaFree(g.dSfRecvOff); cudaFree(g.dSfRecvIdx); cudaFree(g.dSfRecv);
  cudaFree(g.dXiB[0]);
  cudaFree(g.dElemToBCSides); cudaFree(g.dSideInfo); cudaFree(g.dSideBCMetrics); cudaFree(g.dSideNormVec); cudaFree(g.dSideDistance);
  cudaFree(g.dBV0); cudaFree(g.dBV1); cudaFree(g.dBV2); cudaFree(g.dBV3); cudaFree(g.dSideType); cudaFree(g.dElemRadius2); cudaFree(g.dElemEpsOneCell);
  if (g.cap > 0) {
    free_partbuf(g.buf[0]);
    free_partbuf(g.buf[1]);
  }
  cudaFree(g.dXi[0]);
  sort_workspace_free(g.sortws);
  bins_free();
  for (int i = 0; i < 10; ++i) if (g.evp[i]) cudaEventDestroy(g.evp[i]);
  if (g.ev0) cudaEventDestroy(g.ev0);
  if (g.ev1) cudaEventDestroy(g.ev1);
  if (g.stCopy) { cudaStreamSynchronize(g.stCopy); cudaStreamDestroy(g.stCopy); cudaEventDestroy(g.evDepoDone); cudaEventDestroy(g.evCopyDone); }
  if (g.st && g.ownStream) cudaStreamDestroy(g.st);
  if (g.hPin) cudaFreeHost(g.hPin);
  cudaFree(g.dSendCounts); cudaFree(g.dEmigCnt);
  g = Ctx();
  return 0;
}

int piclas_gpu_init(const pgpu_mesh_t* m, const pgpu_params_t* p) {
  if (g.ready) piclas_gpu_finalize();
  g_err.clear();
  if (!m || !p) return fail("piclas_gpu_init: null argument");
  // ---- what this build supports; everything else aborts loudly ----------------------
  if (m->NGeo != 1) return fail("piclas_gpu_init: NGeo=%d not supported (straight-sided NGeo=1 meshes only)", m->NGeo);
  if (m->N < 1 || m->N > PGPU_MAX_N) return fail("piclas_gpu_init: N=%d outside 1..%d", m->N, PGPU_MAX_N);
  if (p->TrackingMethod != PGPU_TRIATRACKING && p->TrackingMethod != PGPU_REFMAPPING)
    return fail("piclas_gpu_init: TrackingMethod=%d not supported (triatracking, refmapping)", p->TrackingMethod);
  const bool isRef = p->TrackingMethod == PGPU_REFMAPPING;
  if (isRef) {
    if (p->DoDeposition && p->DepositionType == PGPU_DEPO_CVWM)
      return fail("piclas_gpu_init: cell_volweight_mean requires TrackingMethod=triatracking (pic_depo_method.f90:127-129)");
    if (!m->ElemToBCSides || !m->SideBCMetrics || !m->SideType || !m->SideNormVec || !m->SideDistance || !m->BaseVectors0 ||
        !m->BaseVectors1 || !m->BaseVectors2 || !m->ElemEpsOneCell || !m->ElemRadiusNGeo || !m->FIBGM_nElems || !m->FIBGM_offsetElem ||
        !m->FIBGM_Element)
      return fail("piclas_gpu_init: refmapping needs ElemToBCSides, SideBCMetrics, SideType, SideNormVec, SideDistance, BaseVectors0-2, "
                  "ElemEpsOneCell, ElemRadiusNGeo and the FIBGM tables");
    for (int b = 0; b < m->nBCSidesTotal; ++b) {
      const int sid = (int)m->SideBCMetrics[(size_t)b * 7];
      if (sid < 1 || sid > m->nSides) return fail("piclas_gpu_init: SideBCMetrics holds an invalid side id");
      if (m->SideType[sid - 1] < 0 || m->SideType[sid - 1] > 2)
        return fail("piclas_gpu_init: BC side %d is curved (only PLANAR_RECT, PLANAR_NONRECT and BILINEAR sides are supported)", sid);
      if (m->SideType[sid - 1] != 0 && !m->BaseVectors3)
        return fail("piclas_gpu_init: BC side %d is not PLANAR_RECT: BaseVectors3 is needed (ComputeBiLinearIntersection)", sid);
    }
  }
  const bool isSF = p->DepositionType == PGPU_DEPO_SF || p->DepositionType == PGPU_DEPO_SF_CC || p->DepositionType == PGPU_DEPO_SF_ADAPTIVE;
  if (p->DoDeposition && p->DepositionType != PGPU_DEPO_CVWM && !isSF)
    return fail("piclas_gpu_init: PIC-Deposition-Type %d not supported (cell_volweight_mean, shape_function, shape_function_cc, shape_function_adaptive)", p->DepositionType);
  if (p->DoDeposition && isSF) {
    if (!m->FIBGM_nElems || !m->FIBGM_offsetElem || !m->FIBGM_Element || !m->ElemToBGM || !m->ElemRadiusNGeo)
      return fail("piclas_gpu_init: shape-function deposition needs FIBGM_nElems/offsetElem/Element, ElemToBGM and ElemRadiusNGeo");
    if (p->dim_sf < 1 || p->dim_sf > 3 || p->dim_sf_dir < 1 || p->dim_sf_dir > 3) return fail("piclas_gpu_init: bad shape-function dimension/direction");
    if (p->DepositionType == PGPU_DEPO_SF_ADAPTIVE && !m->SFElemr2) return fail("piclas_gpu_init: shape_function_adaptive needs SFElemr2");
    if (p->DepositionType != PGPU_DEPO_SF_ADAPTIVE && !(p->r_sf > 0.)) return fail("piclas_gpu_init: PIC-shapefunction-radius must be > 0");
    if (p->alpha_sf < 1) return fail("piclas_gpu_init: PIC-shapefunction-alpha must be >= 1");
  }
  if (p->TimeDiscMethod != PGPU_TIMEDISC_BORIS_LEAPFROG && p->TimeDiscMethod != PGPU_TIMEDISC_LEAPFROG)
    return fail("piclas_gpu_init: TimeDiscMethod=%d not supported (508 Boris-Leapfrog, 509 Leapfrog)", p->TimeDiscMethod);
  if (p->CartesianPeriodic) return fail("piclas_gpu_init: CartesianPeriodic=T not supported");
  if (p->PartLorentzType != 0) return fail("piclas_gpu_init: Part-LorentzType=%d; only non-relativistic (0) is implemented for the new-particle half step", p->PartLorentzType);
  if (p->NoDirichletDeposition) return fail("piclas_gpu_init: PIC-DoDirichletDeposition=F (NullifyNodeSourceDirichletSides) is not implemented");
  if (p->DoDielectricSurfaceCharge) return fail("piclas_gpu_init: DoDielectricSurfaceCharge=T (NodeSourceExt) is not implemented");
  if (p->RefMappingGuess == 2) return fail("piclas_gpu_init: RefMappingGuess=2 not supported (use 1, 3 or 4)");
  if (p->RefMappingGuess < 1 || p->RefMappingGuess > 4) return fail("piclas_gpu_init: RefMappingGuess=%d invalid", p->RefMappingGuess);
  if (p->nSpecies < 1 || p->nSpecies > 32) return fail("piclas_gpu_init: nSpecies=%d outside 1..32", p->nSpecies);
  if (m->nBCs > 255) return fail("piclas_gpu_init: more than 255 boundary conditions");
  if (m->nPeriodicVectors > 8) return fail("piclas_gpu_init: more than 8 periodic vectors");
  if (m->elemInfoSize < 7 || m->sideInfoSize < 8) return fail("piclas_gpu_init: ElemInfo/SideInfo leading dimension too small");
  if (p->nRanks < 1 || p->myRank < 0 || p->myRank >= p->nRanks) return fail("piclas_gpu_init: bad rank layout");
  if (m->nGlobalElems < 1 || m->offsetElem < 0 || m->nElems < 0 || m->offsetElem + m->nElems > m->nGlobalElems)
    return fail("piclas_gpu_init: offsetElem=%d / nElems=%d outside the %d global elements", m->offsetElem, m->nElems, m->nGlobalElems);
  for (int b = 0; b < m->nBCs; ++b) {
    if (m->bc_kind[b] != PGPU_BC_OPEN && m->bc_kind[b] != PGPU_BC_PERIODIC && m->bc_kind[b] != PGPU_BC_REFLECTIVE)
      return fail("piclas_gpu_init: boundary %d has TargetBoundCond=%d; only open (1), reflective (2, specular) and periodic (3) are supported", b + 1, m->bc_kind[b]);
  }

  g.device = p->device;
  CK(cudaSetDevice(g.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, g.device));
  g.nSMs = prop.multiProcessorCount;
  CK(cudaStreamCreate(&g.st));
  CK(cudaMallocHost((void**)&g.hPin, 256 * sizeof(int64_t)));
  CK(cudaMalloc((void**)&g.dSendCounts, (size_t)(p->nRanks > 0 ? p->nRanks : 1) * sizeof(int64_t)));
  CK(cudaMemset(g.dSendCounts, 0, (size_t)(p->nRanks > 0 ? p->nRanks : 1) * sizeof(int64_t)));
  CK(cudaMalloc((void**)&g.dEmigCnt, (size_t)(p->nRanks > 0 ? p->nRanks : 1) * sizeof(int)));
  CK(cudaEventCreate(&g.ev0));
  CK(cudaEventCreate(&g.ev1));
  for (int i = 0; i < 10; ++i) CK(cudaEventCreate(&g.evp[i]));
  g.prm = *p;
  g.prm.ChargeIC = g.prm.MassIC = g.prm.MacroParticleFactor = nullptr;
  g.nGlobalElems = m->nGlobalElems;
  g.nElems = m->nElems;
  g.offsetElem = m->offsetElem;
  g.N = m->N;
  g.NP = m->N + 1;
  g.ND = g.NP * g.NP * g.NP;
  g.nNodes = m->nUniqueGlobalNodes;
  g.nRanks = p->nRanks;
  g.myRank = p->myRank;
  g.carryIDs = p->carryParticleIDs != 0;
  g.commSize = 8 + (p->TrackingMethod == PGPU_REFMAPPING ? 3 : 0) + (g.carryIDs ? 1 : 0);  // particle_mpi.f90:158-183
  const int nG = m->nGlobalElems;

  // ---- per-element records ------------------------------------------------------------------------------------
  std::vector<TriaElem> tria(nG);
  std::vector<GeoElem> geo(nG);
  std::vector<int32_t> rank(nG);
  for (int e = 0; e < nG; ++e) {
    const int32_t* ei = m->ElemInfo + (size_t)e * m->elemInfoSize;
    const int firstSide = ei[ELEM_FIRSTSIDEIND], lastSide = ei[ELEM_LASTSIDEIND];
    const int firstNode = ei[ELEM_FIRSTNODEIND], lastNode = ei[ELEM_LASTNODEIND];
    if (lastSide - firstSide != 6) return fail("piclas_gpu_init: element %d has %d sides (mortar meshes are not supported)", e + 1, lastSide - firstSide);
    if (lastNode - firstNode != 8) return fail("piclas_gpu_init: element %d has %d nodes (NGeo=1 hexahedra only)", e + 1, lastNode - firstNode);
    rank[e] = ei[ELEM_RANK];
    if (rank[e] < 0 || rank[e] >= g.nRanks) return fail("piclas_gpu_init: ELEM_RANK of element %d outside 0..nRanks-1", e + 1);
    // the sort keys and the halo lists rely on the reference's partition: contiguous element ranges in rank order
    // (loadbalance/loaddistribution.f90:362-369), this rank owning exactly offsetElem+1 .. offsetElem+nElems
    if (e > 0 && rank[e] < rank[e - 1]) return fail("piclas_gpu_init: ELEM_RANK is not ascending at element %d (ranks must own contiguous element ranges)", e + 1);
    if ((rank[e] == g.myRank) != (e >= g.offsetElem && e < g.offsetElem + g.nElems))
      return fail("piclas_gpu_init: ELEM_RANK of element %d disagrees with offsetElem=%d / nElems=%d of rank %d", e + 1, g.offsetElem, g.nElems, g.myRank);
    TriaElem& t = tria[e];
    memset(&t, 0, sizeof(t));
    for (int n = 0; n < 8; ++n)
      for (int d = 0; d < 3; ++d) t.corner[n][d] = m->NodeCoords[(size_t)(firstNode + n) * 3 + d];
    for (int s = 0; s < 6; ++s) {
      const int sid = firstSide + s + 1;  // 1-based SideInfo index
      const int32_t* si = m->SideInfo + (size_t)(sid - 1) * m->sideInfoSize;
      if (si[SIDE_LOCALID] != s + 1) return fail("piclas_gpu_init: element %d: SIDE_LOCALID out of order (mortar meshes are not supported)", e + 1);
      if (si[SIDE_NBELEMID] < 0) return fail("piclas_gpu_init: element %d has a mortar side", e + 1);
      if (si[SIDE_BCID] < 0 || si[SIDE_BCID] > m->nBCs) return fail("piclas_gpu_init: element %d: SIDE_BCID out of range", e + 1);
      t.nbElem[s] = si[SIDE_NBELEMID];
      t.sideID[s] = sid;
      t.bcid[s] = (uint8_t)si[SIDE_BCID];
      if (m->ConcaveElemSide[(size_t)e * 6 + s]) t.concave |= (uint8_t)(1u << s);
      for (int n = 0; n < 4; ++n) {
        const int loc = m->ElemSideNodeID[((size_t)e * 6 + s) * 4 + n] - firstNode;
        if (loc < 0 || loc > 7) return fail("piclas_gpu_init: ElemSideNodeID of element %d points outside the element", e + 1);
        t.sideNode[s][n] = (uint8_t)loc;
      }
    }
    GeoElem& ge = geo[e];
    memset(&ge, 0, sizeof(ge));
    memcpy(ge.XCL, m->XCL_NGeo + (size_t)e * 24, 24 * 8);
    memcpy(ge.dXCL, m->dXCL_NGeo + (size_t)e * 72, 72 * 8);
    memcpy(ge.bary, m->ElemBaryNGeo + (size_t)e * 3, 3 * 8);
    memcpy(ge.xez, m->XiEtaZetaBasis + (size_t)e * 18, 18 * 8);
    memcpy(ge.slen, m->slenXiEtaZetaBasis + (size_t)e * 6, 6 * 8);
  }
  g.fast = p->arithmetic != 0;
  std::vector<PlaneElem> planes;
  std::vector<AffElem> affs;
  if (g.fast) {
    planes.resize(nG);
    affs.resize(nG);
    for (int e = 0; e < nG; ++e) {
      const TriaElem& t = tria[e];
      PlaneElem& pl = planes[e];
      memset(&pl, 0, sizeof(pl));
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int n = 0; n < 8; ++n)
        for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], t.corner[n][d]); hi[d] = std::max(hi[d], t.corner[n][d]); }
      const double diam = sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
      pl.tol = 1e-8 * diam;
      for (int s = 0; s < 6; ++s)
        for (int tr = 0; tr < 2; ++tr) {
          // triangle (node1, node_{tr+2}, node_{tr+3}); det = -(x - P1) . N with N = (Pb - P1) x (Pc - P1)
          const double* P1 = t.corner[t.sideNode[s][0]];
          const double* Pb = t.corner[t.sideNode[s][tr + 1]];
          const double* Pc = t.corner[t.sideNode[s][tr + 2]];
          const double u[3] = {Pb[0] - P1[0], Pb[1] - P1[1], Pb[2] - P1[2]}, w[3] = {Pc[0] - P1[0], Pc[1] - P1[1], Pc[2] - P1[2]};
          double N[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
          const double len = sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
          if (!(len > 0.)) return fail("piclas_gpu_init: element %d has a degenerate side triangle", e + 1);
          const int k = 2 * s + tr;
          for (int d = 0; d < 3; ++d) pl.pl[k][d] = -N[d] / len;
          pl.pl[k][3] = pl.pl[k][0] * P1[0] + pl.pl[k][1] * P1[1] + pl.pl[k][2] * P1[2];
        }
      for (int s = 0; s < 6; ++s) pl.concave2 |= (uint32_t)((t.concave >> s) & 1u) << (2 * s);
      // planar sides (ConcaveElemSide is then irrelevant: both triangles lie in one plane) and all corners inside all
      // side planes: convex polyhedron
      bool planar = true;
      for (int k = 0; k < 12; ++k)
        for (int n = 0; n < 8; ++n)
          if (pl.pl[k][0] * t.corner[n][0] + pl.pl[k][1] * t.corner[n][1] + pl.pl[k][2] * t.corner[n][2] - pl.pl[k][3] < -1e-12 * diam)
            planar = false;
      for (int s = 0; s < 6; ++s) {
        const double *a = pl.pl[2 * s], *b = pl.pl[2 * s + 1];
        if (fabs(a[0] - b[0]) > 1e-12 || fabs(a[1] - b[1]) > 1e-12 || fabs(a[2] - b[2]) > 1e-12 || fabs(a[3] - b[3]) > 1e-12 * diam)
          planar = false;
        // diagonal plane: normal = n_side x (P3 - P1), oriented towards node 2
        const double* P1 = t.corner[t.sideNode[s][0]];
        const double* P2 = t.corner[t.sideNode[s][1]];
        const double* P3 = t.corner[t.sideNode[s][2]];
        const double dgn[3] = {P3[0] - P1[0], P3[1] - P1[1], P3[2] - P1[2]};
        double mm[3] = {a[1] * dgn[2] - a[2] * dgn[1], a[2] * dgn[0] - a[0] * dgn[2], a[0] * dgn[1] - a[1] * dgn[0]};
        const double ml = sqrt(mm[0] * mm[0] + mm[1] * mm[1] + mm[2] * mm[2]);
        if (!(ml > 0.)) { planar = false; continue; }
        double sgn = (mm[0] * (P2[0] - P1[0]) + mm[1] * (P2[1] - P1[1]) + mm[2] * (P2[2] - P1[2])) >= 0. ? 1. : -1.;
        for (int d = 0; d < 3; ++d) pl.dg[s][d] = sgn * mm[d] / ml;
        pl.dg[s][3] = pl.dg[s][0] * P1[0] + pl.dg[s][1] * P1[1] + pl.dg[s][2] * P1[2];
      }
      pl.planar = planar ? 1u : 0u;
      // affine test on the trilinear map X(i,j,k): all mixed differences vanish
      AffElem& a = affs[e];
      memset(&a, 0, sizeof(a));
      const double(*X)[3] = geo[e].XCL;  // node i + 2j + 4k
      double ea[3], eb[3], ec[3], dev = 0.;
      for (int d = 0; d < 3; ++d) {
        ea[d] = X[1][d] - X[0][d]; eb[d] = X[2][d] - X[0][d]; ec[d] = X[4][d] - X[0][d];
        dev = std::max(dev, fabs(X[3][d] - (X[0][d] + ea[d] + eb[d])));
        dev = std::max(dev, fabs(X[5][d] - (X[0][d] + ea[d] + ec[d])));
        dev = std::max(dev, fabs(X[6][d] - (X[0][d] + eb[d] + ec[d])));
        dev = std::max(dev, fabs(X[7][d] - (X[0][d] + ea[d] + eb[d] + ec[d])));
      }
      if (dev <= 1e-14 * diam) {
        // x = X0 + M (xi + 1)/2 with M = [ea eb ec]  =>  xi = 2 M^-1 (x - X0) - 1
        const double M[3][3] = {{ea[0], eb[0], ec[0]}, {ea[1], eb[1], ec[1]}, {ea[2], eb[2], ec[2]}};
        const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
        if (det > 0.) {
          const double id = 2.0 / det;
          a.A[0][0] = (M[1][1] * M[2][2] - M[1][2] * M[2][1]) * id; a.A[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id;
          a.A[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id; a.A[1][0] = (M[1][2] * M[2][0] - M[1][0] * M[2][2]) * id;
          a.A[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id; a.A[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
          a.A[2][0] = (M[1][0] * M[2][1] - M[1][1] * M[2][0]) * id; a.A[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id;
          a.A[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
          for (int d = 0; d < 3; ++d) a.x0[d] = X[0][d];
          a.affine = 1.0;
        }
      }
    }
    if (upload(&g.dPlanes, planes.data(), (size_t)nG)) return 1;
    if (upload(&g.dAff, affs.data(), (size_t)nG)) return 1;
  }
  if (upload(&g.dTria, tria.data(), (size_t)nG)) return 1;
  if (upload(&g.dGeo, geo.data(), (size_t)nG)) return 1;
  if (upload(&g.dElemRank, rank.data(), (size_t)nG)) return 1;
  if (upload(&g.dElemXGP, m->Elem_xGP, (size_t)nG * g.ND * 3)) return 1;

  // ---- particle layout: bins (bins.cuh) for TriaTracking with cell_volweight_mean or without deposition; the sorted arrays of
  //      round 1 for RefMapping and the shape functions (PICLAS_GPU_LAYOUT=sorted forces them everywhere: A/B measurements)
  {
    const char* lay = getenv("PICLAS_GPU_LAYOUT");
    const bool forceSorted = lay && strcmp(lay, "sorted") == 0;
    g.binEligible = !isRef && !(p->DoDeposition && isSF) && !forceSorted;
    if (const char* v = getenv("PICLAS_GPU_BIN_MAIN_SLACK")) g.mainSlack = atof(v);
    if (const char* v = getenv("PICLAS_GPU_BIN_INBOX_FRAC")) g.inFrac = atof(v);
    if (g.binEligible) {
      if (bins_alloc_tables()) return 1;
      if (build_push_elems(m, tria, planes, affs, rank)) return 1;
    }
  }

  // ---- cell_volweight_mean tables ---------------------------------------------------------------------------------
  // The reference builds ElemNodeID for TriaTracking only (particle_mesh.f90:411-450) and NodeVolume / Periodic_* for
  // cell_volweight_mean only: the pointers are read in that mode alone (a host may pass NULL otherwise).
  CK(cudaMalloc((void**)&g.dPartSource, (size_t)(g.nElems ? g.nElems : 1) * g.ND * 4 * 8));
  if (p->DoDeposition && p->DepositionType == PGPU_DEPO_CVWM) {
    if (!m->ElemNodeID || !m->NodeInfo || !m->Periodic_nNodes || !m->Periodic_offsetNode || !m->NodeVolume || (m->nPeriodicNodesTotal > 0 && !m->Periodic_Nodes))
      return fail("piclas_gpu_init: cell_volweight_mean needs ElemNodeID, NodeInfo, Periodic_nNodes / offsetNode / Nodes and NodeVolume");
    std::vector<int32_t> elemNodeU((size_t)g.nElems * 8), cnt(g.nNodes + 1, 0);
    for (int e = 0; e < g.nElems; ++e)
      for (int c = 0; c < 8; ++c) {
        const int nu = m->NodeInfo[m->ElemNodeID[(size_t)(g.offsetElem + e) * 8 + c] - 1] - 1;
        if (nu < 0 || nu >= g.nNodes) return fail("piclas_gpu_init: NodeInfo out of range");
        elemNodeU[(size_t)e * 8 + c] = nu;
        cnt[nu + 1]++;
      }
    for (int n = 0; n < g.nNodes; ++n) cnt[n + 1] += cnt[n];
    std::vector<int32_t> adj((size_t)g.nElems * 8), fill(cnt.begin(), cnt.end() - 1);
    for (int e = 0; e < g.nElems; ++e)
      for (int c = 0; c < 8; ++c) adj[fill[elemNodeU[(size_t)e * 8 + c]]++] = e * 8 + c;
    if (upload(&g.dAdjOff, cnt.data(), cnt.size())) return 1;
    if (upload(&g.dAdj, adj.data(), adj.size())) return 1;
    if (upload(&g.dElemNodeU, elemNodeU.data(), elemNodeU.size())) return 1;
    if (upload(&g.dPerN, m->Periodic_nNodes, (size_t)g.nNodes)) return 1;
    if (upload(&g.dPerOff, m->Periodic_offsetNode, (size_t)g.nNodes)) return 1;
    if (upload(&g.dPerNodes, m->Periodic_Nodes, (size_t)m->nPeriodicNodesTotal)) return 1;
    if (upload(&g.dNodeVolume, m->NodeVolume, (size_t)g.nNodes)) return 1;
    CK(cudaMalloc((void**)&g.dElemAcc, (size_t)(g.nElems ? g.nElems : 1) * 32 * 8));
    CK(cudaMalloc((void**)&g.dS, (size_t)g.nNodes * 4 * 8));
    CK(cudaMalloc((void**)&g.dNodeSource, (size_t)g.nNodes * 4 * 8));
  }

  if (g.nRanks > 1 && p->DoDeposition && p->DepositionType == PGPU_DEPO_CVWM) {
    // nodes whose NodeSource needs the sums of several ranks: touched by elements of two ranks, directly or through a periodic
    // partner (pic_depo.f90:298-571 builds the same lists per neighbour rank)
    std::vector<int32_t> halo;
    if (g.nRanks <= 64) {
      std::vector<uint64_t> touch((size_t)g.nNodes, 0);
      for (int e = 0; e < nG; ++e)
        for (int c = 0; c < 8; ++c) touch[m->NodeInfo[m->ElemNodeID[(size_t)e * 8 + c] - 1] - 1] |= 1ull << rank[e];
      for (int n = 0; n < g.nNodes; ++n) {
        uint64_t need = touch[n];
        for (int j = 0; j < m->Periodic_nNodes[n]; ++j) need |= touch[m->Periodic_Nodes[m->Periodic_offsetNode[n] + j] - 1];
        if (touch[n] != 0 && (need & (need - 1)) != 0) halo.push_back(n);
      }
    } else {
      for (int n = 0; n < g.nNodes; ++n) halo.push_back(n);
    }
    g.nHaloNodes = (int)halo.size();
    if (upload(&g.dHaloNodes, halo.data(), halo.size())) return 1;
    CK(cudaMalloc((void**)&g.dHaloSend, (size_t)(g.nHaloNodes ? g.nHaloNodes : 1) * 4 * 8));
    CK(cudaMalloc((void**)&g.dHaloRecv, (size_t)(g.nHaloNodes ? g.nHaloNodes : 1) * 4 * 8 * g.nRanks));
  }

  // ---- tables shared by RefMapping and the shape functions -----------------------------------------------------------
  g.sfActive = p->DoDeposition && isSF;
  g.ref = isRef;
  g.locT = RefTables{};
  for (int d = 0; d < 3; ++d) { g.globLo[d] = m->xyzminglob[d]; g.globHi[d] = m->xyzmaxglob[d]; }
  if (g.sfActive || g.ref || (m->FIBGM_nElems && m->FIBGM_offsetElem && m->FIBGM_Element && m->ElemBaryNGeo && m->ElemRadius2NGeo)) {
    const int ni0 = m->FIBGMmax[0] - m->FIBGMmin[0] + 1, nj0 = m->FIBGMmax[1] - m->FIBGMmin[1] + 1, nk0 = m->FIBGMmax[2] - m->FIBGMmin[2] + 1;
    const size_t nCells0 = (size_t)ni0 * nj0 * nk0;
    if (upload(&g.dFibN, m->FIBGM_nElems, nCells0)) return 1;
    if (upload(&g.dFibOff, m->FIBGM_offsetElem, nCells0)) return 1;
    if (upload(&g.dFibElem, m->FIBGM_Element, (size_t)m->nFIBGMElemsTotal)) return 1;
    if (upload(&g.dElemBary, m->ElemBaryNGeo, (size_t)nG * 3)) return 1;
    if (m->ElemRadiusNGeo && upload(&g.dElemRadius, m->ElemRadiusNGeo, (size_t)nG)) return 1;
    if (m->ElemRadius2NGeo) {   // SinglePointToElement of the emission (any tracking method)
      if (upload(&g.dElemRadius2, m->ElemRadius2NGeo, (size_t)nG)) return 1;
      g.locT.geo = g.dGeo; g.locT.ElemBary = g.dElemBary; g.locT.ElemRadius2 = g.dElemRadius2;
      g.locT.FIBGM_nElems = g.dFibN; g.locT.FIBGM_offsetElem = g.dFibOff; g.locT.FIBGM_Element = g.dFibElem;
    }
  }
  if (g.ref) {
    if (upload(&g.dElemToBCSides, m->ElemToBCSides, (size_t)nG * 2)) return 1;
    if (upload(&g.dSideBCMetrics, m->SideBCMetrics, (size_t)m->nBCSidesTotal * 7)) return 1;
    if (upload(&g.dSideInfo, m->SideInfo, (size_t)m->nSides * m->sideInfoSize)) return 1;
    if (upload(&g.dSideNormVec, m->SideNormVec, (size_t)m->nSides * 3)) return 1;
    if (upload(&g.dSideDistance, m->SideDistance, (size_t)m->nSides)) return 1;
    if (upload(&g.dBV0, m->BaseVectors0, (size_t)m->nSides * 3)) return 1;
    if (upload(&g.dBV1, m->BaseVectors1, (size_t)m->nSides * 3)) return 1;
    if (upload(&g.dBV2, m->BaseVectors2, (size_t)m->nSides * 3)) return 1;
    if (m->BaseVectors3 && upload(&g.dBV3, m->BaseVectors3, (size_t)m->nSides * 3)) return 1;
    if (upload(&g.dSideType, m->SideType, (size_t)m->nSides)) return 1;
    if (!g.dElemRadius2 && upload(&g.dElemRadius2, m->ElemRadius2NGeo, (size_t)nG)) return 1;
    if (upload(&g.dElemEpsOneCell, m->ElemEpsOneCell, (size_t)nG)) return 1;
    g.locT.ElemEpsOneCell = g.dElemEpsOneCell;
    g.refT.geo = g.dGeo; g.refT.ElemToBCSides = g.dElemToBCSides; g.refT.SideBCMetrics = g.dSideBCMetrics; g.refT.SideInfo = g.dSideInfo;
    g.refT.sideInfoSize = m->sideInfoSize; g.refT.SideNormVec = g.dSideNormVec; g.refT.SideDistance = g.dSideDistance;
    g.refT.BaseVectors0 = g.dBV0; g.refT.BaseVectors1 = g.dBV1; g.refT.BaseVectors2 = g.dBV2; g.refT.BaseVectors3 = g.dBV3;
    g.refT.SideType = g.dSideType; g.refT.ElemBary = g.dElemBary;
    g.refT.ElemRadius = g.dElemRadius; g.refT.ElemRadius2 = g.dElemRadius2; g.refT.ElemEpsOneCell = g.dElemEpsOneCell;
    g.refT.FIBGM_nElems = g.dFibN; g.refT.FIBGM_offsetElem = g.dFibOff; g.refT.FIBGM_Element = g.dFibElem;
  }
  // ---- shape-function tables and gather candidates ------------------------------------------------------------------
  int sfCaseM[27][3];
  memset(sfCaseM, 0, sizeof(sfCaseM));
  int nSFCases = 0, sfDir1 = (p->dim_sf_dir == 2) ? 1 : 2, sfDir2 = (p->dim_sf_dir == 3) ? 1 : 3, pvec1 = 0, pvec2 = 0;
  if (g.sfActive) {
    const int dim_sf = p->dim_sf;
    if (m->nPeriodicVectors > 0) {  // InitPeriodicSFCaseMatrix, pic_depo.f90:828-910
      nSFCases = 1;
      for (int d = 0; d < dim_sf; ++d) nSFCases *= 3;
      auto M = [&](int i1, int c1) -> int& { return sfCaseM[i1 - 1][c1 - 1]; };
      if (dim_sf == 1) { M(1, 1) = 1; M(3, 1) = -1; }
      if (dim_sf == 2) {
        for (int i = 1; i <= 3; ++i) M(i, 1) = 1;
        for (int i = 7; i <= 9; ++i) M(i, 1) = -1;
        for (int I = 1; I <= 3; ++I) { M(I * 3 - 2, 2) = 1; M(I * 3, 2) = -1; }
        if (m->nPeriodicVectors == 1) { pvec1 = 1; pvec2 = 0; }
        else if (m->nPeriodicVectors == 2) { pvec1 = 1; pvec2 = 2; }
        else { pvec1 = sfDir1; pvec2 = sfDir2; }
      }
      if (dim_sf == 3) {
        for (int i = 1; i <= 9; ++i) M(i, 1) = 1;
        for (int i = 19; i <= 27; ++i) M(i, 1) = -1;
        for (int I = 1; I <= 3; ++I) {
          for (int i = I * 9 - 8; i <= I * 9 - 6; ++i) M(i, 2) = 1;
          for (int i = I * 9 - 2; i <= I * 9; ++i) M(i, 2) = -1;
          for (int J = 1; J <= 3; ++J) { M((J * 3 - 2) + (I - 1) * 9, 3) = 1; M((J * 3) + (I - 1) * 9, 3) = -1; }
        }
        if (m->nPeriodicVectors < 3) return fail("piclas_gpu_init: 3-D shape function with periodic sides needs three periodic vectors");
      }
    }
    const int ni = m->FIBGMmax[0] - m->FIBGMmin[0] + 1, nj = m->FIBGMmax[1] - m->FIBGMmin[1] + 1, nk = m->FIBGMmax[2] - m->FIBGMmin[2] + 1;
    const size_t nCells = (size_t)ni * nj * nk;
    (void)nCells;
    if (upload(&g.dElemToBGM, m->ElemToBGM, (size_t)nG * 6)) return 1;
    if (upload(&g.dElemsJ, m->ElemsJ, (size_t)nG * g.ND)) return 1;
    if (m->SFElemr2 && upload(&g.dSFElemr2, m->SFElemr2, (size_t)nG * 2)) return 1;
    // candidates: (source element s, case c) can reach target e iff SFNorm(bary_e - (bary_s + shift_c)) <= r + R_e + R_s
    double rmax = p->r_sf;
    if (p->DepositionType == PGPU_DEPO_SF_ADAPTIVE) {
      rmax = 0.;
      for (int e = 0; e < nG; ++e) rmax = std::max(rmax, m->SFElemr2[(size_t)e * 2]);
    }
    auto sfnorm = [&](const double* v) {
      if (dim_sf == 1) return fabs(v[p->dim_sf_dir - 1]);
      if (dim_sf == 2) return sqrt(v[sfDir1 - 1] * v[sfDir1 - 1] + v[sfDir2 - 1] * v[sfDir2 - 1]);
      return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    };
    auto PVc = [&](int I, int iVec) { return m->PeriodicVectors[(size_t)(iVec - 1) * 3 + (I - 1)]; };
    const int nCaseLoop = nSFCases > 1 ? nSFCases : 1;
    std::vector<std::array<double, 3>> shift(nCaseLoop);
    for (int c = 0; c < nCaseLoop; ++c) {
      shift[c] = {0., 0., 0.};
      if (nSFCases > 1) {
        const int* cm = sfCaseM[c];
        if (dim_sf == 1) shift[c][p->dim_sf_dir - 1] = cm[0] * PVc(p->dim_sf_dir, p->dim_sf_dir);
        else if (dim_sf == 2) {
          shift[c][sfDir1 - 1] = cm[0] * PVc(sfDir1, pvec1) + (pvec2 > 0 ? cm[1] * PVc(sfDir1, pvec2) : 0.);
          shift[c][sfDir2 - 1] = cm[0] * PVc(sfDir2, pvec1) + (pvec2 > 0 ? cm[1] * PVc(sfDir2, pvec2) : 0.);
        } else for (int I = 1; I <= 3; ++I) shift[c][I - 1] = cm[0] * PVc(I, 1) + cm[1] * PVc(I, 2) + cm[2] * PVc(I, 3);
      }
    }
    // gather targets: the local elements plus, for multi-rank runs, every element of another rank that a local particle
    // can reach (sorted by owner rank, then element id, so that each rank's part of the halo buffer is contiguous)
    std::vector<int32_t> targets(g.nElems);
    for (int e = 0; e < g.nElems; ++e) targets[e] = g.offsetElem + e + 1;
    std::vector<int32_t> candOff, candSrc;
    std::vector<uint8_t> candCase;
    std::vector<int> stamp(nG, -1);
    // corner boxes of the elements: a target's Gauss points lie inside its box, a source's particles inside its box up to the
    // localisation tolerance of RefMapping (ElemEpsOneCell, a few per cent of the element) -> source boxes are inflated by 5 %
    std::vector<std::array<double, 6>> ebox(nG);
    for (int e2 = 0; e2 < nG; ++e2) {
      for (int d = 0; d < 3; ++d) { ebox[e2][d] = 1e300; ebox[e2][3 + d] = -1e300; }
      for (int n = 0; n < 8; ++n)
        for (int d = 0; d < 3; ++d) {
          ebox[e2][d] = std::min(ebox[e2][d], tria[e2].corner[n][d]);
          ebox[e2][3 + d] = std::max(ebox[e2][3 + d], tria[e2].corner[n][d]);
        }
    }
    double RsMaxAll = 0.;
    for (int s2 = 0; s2 < nG; ++s2) RsMaxAll = std::max(RsMaxAll, m->ElemRadiusNGeo[s2]);
    int tagCounter = 0;
    // sources (restricted to [srcLo, srcHi) global 0-based) that reach global element ge; appended to `out` as (src, case)
    auto reach = [&](int ge, int srcLo, int srcHi, std::vector<std::pair<int, int>>& out) {
      const double* be = m->ElemBaryNGeo + (size_t)ge * 3;
      const double Re = m->ElemRadiusNGeo[ge];
      for (int c = 0; c < nCaseLoop; ++c) {
        const double q[3] = {be[0] - shift[c][0], be[1] - shift[c][1], be[2] - shift[c][2]};
        const double Rq = rmax + Re;
        int lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
          lo[d] = (int)floor((q[d] - (Rq + RsMaxAll) - m->xyzminglob[d]) / m->FIBGMdeltas[d]) + 1;
          hi[d] = (int)floor((q[d] + (Rq + RsMaxAll) - m->xyzminglob[d]) / m->FIBGMdeltas[d]) + 1;
          const bool used = (dim_sf == 3) || (dim_sf == 1 && d == p->dim_sf_dir - 1) || (dim_sf == 2 && d != p->dim_sf_dir - 1);
          if (!used) { lo[d] = m->FIBGMmin[d]; hi[d] = m->FIBGMmax[d]; }
          lo[d] = std::max(lo[d], m->FIBGMmin[d]);
          hi[d] = std::min(hi[d], m->FIBGMmax[d]);
        }
        const int tag = tagCounter++;
        std::vector<int> found;
        for (int kk = lo[0]; kk <= hi[0]; ++kk) for (int ll = lo[1]; ll <= hi[1]; ++ll) for (int mm = lo[2]; mm <= hi[2]; ++mm) {
          const size_t cell = (size_t)(kk - m->FIBGMmin[0]) + (size_t)ni * ((size_t)(ll - m->FIBGMmin[1]) + (size_t)nj * (size_t)(mm - m->FIBGMmin[2]));
          for (int k2 = 0; k2 < m->FIBGM_nElems[cell]; ++k2) {
            const int gs = m->FIBGM_Element[m->FIBGM_offsetElem[cell] + k2] - 1;
            if (stamp[gs] == tag) continue;
            stamp[gs] = tag;
            if (gs < srcLo || gs >= srcHi) continue;
            const double* bs = m->ElemBaryNGeo + (size_t)gs * 3;
            const double dv[3] = {q[0] - bs[0], q[1] - bs[1], q[2] - bs[2]};
            if (!(sfnorm(dv) <= 1.0000001 * (rmax + Re + m->ElemRadiusNGeo[gs]))) continue;
            // sharper: distance between the target's box (shifted as the particle images are) and the inflated source box
            double gap[3];
            for (int d = 0; d < 3; ++d) {
              const double ext = 0.05 * (ebox[gs][3 + d] - ebox[gs][d]) + 1e-12;
              const double tlo = ebox[ge][d] - shift[c][d], thi = ebox[ge][3 + d] - shift[c][d];
              gap[d] = std::max(std::max(tlo - (ebox[gs][3 + d] + ext), (ebox[gs][d] - ext) - thi), 0.0);
            }
            if (sfnorm(gap) <= 1.0000001 * rmax) found.push_back(gs);
          }
        }
        std::sort(found.begin(), found.end());
        for (int gs : found) out.emplace_back(gs, nSFCases > 1 ? c + 1 : 0);
      }
    };
    const int myLo = g.offsetElem, myHi = g.offsetElem + g.nElems;
    // halo targets and the mirror-image receive lists (both follow from the geometry, no communication needed)
    g.sfSendCount.assign(g.nRanks, 0);
    g.sfRecvCount.assign(g.nRanks, 0);
    std::vector<int32_t> recvLocal;
    if (g.nRanks > 1) {
      std::vector<int> rlo(g.nRanks, nG), rhi(g.nRanks, 0);
      for (int e = 0; e < nG; ++e) { rlo[rank[e]] = std::min(rlo[rank[e]], e); rhi[rank[e]] = std::max(rhi[rank[e]], e + 1); }
      std::vector<std::pair<int, int>> tmp;
      for (int r = 0; r < g.nRanks; ++r) {
        if (r == g.myRank) continue;
        for (int ge = rlo[r]; ge < rhi[r]; ++ge) {          // elements of rank r reached by my particles -> I send
          tmp.clear();
          reach(ge, myLo, myHi, tmp);
          if (!tmp.empty()) { targets.push_back(ge + 1); g.sfSendCount[r]++; }
        }
        for (int ge = myLo; ge < myHi; ++ge) {              // my elements reached by particles of rank r -> I receive
          tmp.clear();
          reach(ge, rlo[r], rhi[r], tmp);
          if (!tmp.empty()) { recvLocal.push_back(ge - g.offsetElem); g.sfRecvCount[r]++; }
        }
      }
    }
    g.nSfTargets = (int)targets.size();
    g.nSfHalo = g.nSfTargets - g.nElems;
    g.sfRecvTotal = (int64_t)recvLocal.size();
    candOff.assign(1, 0);
    {
      std::vector<std::pair<int, int>> tmp;
      for (int t = 0; t < g.nSfTargets; ++t) {
        tmp.clear();
        reach(targets[t] - 1, myLo, myHi, tmp);
        for (auto& sc : tmp) { candSrc.push_back(sc.first - g.offsetElem); candCase.push_back((uint8_t)sc.second); }
        candOff.push_back((int32_t)candSrc.size());
      }
    }
    if (upload(&g.dSfTarget, targets.data(), targets.size())) return 1;
    {  // receive list grouped by local element (CSR), entries of one element in rank order
      std::vector<int32_t> cnt(g.nElems + 1, 0), rElem, rOff(1, 0), rIdx(recvLocal.size());
      for (int32_t e : recvLocal) cnt[e + 1]++;
      for (int e = 0; e < g.nElems; ++e) cnt[e + 1] += cnt[e];
      std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
      for (size_t i = 0; i < recvLocal.size(); ++i) rIdx[fill[recvLocal[i]]++] = (int32_t)i;
      for (int e = 0; e < g.nElems; ++e)
        if (cnt[e + 1] > cnt[e]) { rElem.push_back(e); rOff.push_back(cnt[e + 1]); }
      g.nSfRecvElems = (int)rElem.size();
      if (upload(&g.dSfRecvElem, rElem.data(), rElem.size())) return 1;
      if (upload(&g.dSfRecvOff, rOff.data(), rOff.size())) return 1;
      if (upload(&g.dSfRecvIdx, rIdx.data(), rIdx.size())) return 1;
    }
    CK(cudaMalloc((void**)&g.dSfRecv, (size_t)(recvLocal.size() ? recvLocal.size() : 1) * g.ND * 4 * 8));
    cudaFree(g.dPartSource);   // local elements followed by the halo targets
    CK(cudaMalloc((void**)&g.dPartSource, (size_t)(g.nSfTargets ? g.nSfTargets : 1) * g.ND * 4 * 8));
    if (upload(&g.dCandOff, candOff.data(), candOff.size())) return 1;
    if (upload(&g.dCandSrc, candSrc.data(), candSrc.size())) return 1;
    if (upload(&g.dCandCase, candCase.data(), candCase.size())) return 1;
    g.sfT.FIBGM_nElems = g.dFibN; g.sfT.FIBGM_offsetElem = g.dFibOff; g.sfT.FIBGM_Element = g.dFibElem; g.sfT.ElemToBGM = g.dElemToBGM;
    g.sfT.ElemBary = g.dElemBary; g.sfT.ElemRadius = g.dElemRadius; g.sfT.Elem_xGP = g.dElemXGP; g.sfT.ElemsJ = g.dElemsJ;
    g.sfT.target = g.dSfTarget;
    g.sfT.SFElemr2 = g.dSFElemr2; g.sfT.candOff = g.dCandOff; g.sfT.candSrc = g.dCandSrc; g.sfT.candCase = g.dCandCase;
  }
  CK(cudaMalloc((void**)&g.dE, (size_t)(g.nElems ? g.nElems : 1) * g.ND * 3 * 8));
  CK(cudaMemset(g.dE, 0, (size_t)(g.nElems ? g.nElems : 1) * g.ND * 3 * 8));
  if (g.binEligible && g.fast) {
    CK(cudaMalloc((void**)&g.dEmono, (size_t)(g.nElems ? g.nElems : 1) * g.ND * 3 * 8));
    CK(cudaMemset(g.dEmono, 0, (size_t)(g.nElems ? g.nElems : 1) * g.ND * 3 * 8));
  }
  CK(cudaMalloc((void**)&g.dElemOff, (size_t)(g.nElems + g.nRanks + 2) * sizeof(int64_t)));
  CK(cudaMemset(g.dElemOff, 0, (size_t)(g.nElems + g.nRanks + 2) * sizeof(int64_t)));
  CK(cudaMalloc((void**)&g.dCounters, (8 + UNPUSHED_LIST) * sizeof(int)));   // counters, then the far slots of unpushed records
  g.keyBits = 1;
  while ((1u << g.keyBits) < (uint32_t)(g.nElems + g.nRanks + 1)) ++g.keyBits;

  // ---- constant tables ------------------------------------------------------------------------------------------------
  static ConstTables h;
  memset(&h, 0, sizeof(h));
  for (int i = 0; i <= m->N; ++i) {
    h.xGP[i] = m->xGP[i]; h.wGP[i] = m->wGP[i]; h.wBary[i] = m->wBary[i];
    h.cvwFac[i] = (m->xGP[i] + 1.0) / 2.0;  // pic_depo.f90:271-275
  }
  for (int i = 0; i < 2; ++i) { h.XiCL[i] = m->XiCL_NGeo[i]; h.wBaryCL[i] = m->wBaryCL_NGeo[i]; }
  for (int i = 0; i < 6; ++i) h.externalField[i] = p->externalField[i];
  h.c2_inv = p->c2_inv;
  h.RefMappingEps = p->RefMappingEps;
  h.RefMappingGuess = p->RefMappingGuess;
  h.TrackingMethod = p->TrackingMethod;
  h.TimeDiscMethod = p->TimeDiscMethod;
  h.DoInterpolation = p->DoInterpolation;
  h.DepositionType = p->DepositionType;
  h.nSpecies = p->nSpecies;
  for (int s = 0; s < p->nSpecies; ++s) { h.ChargeIC[s] = p->ChargeIC[s]; h.MassIC[s] = p->MassIC[s]; h.MPF[s] = p->MacroParticleFactor[s]; }
  h.nBCs = m->nBCs;
  for (int b = 0; b < m->nBCs; ++b) { h.bc_kind[b] = m->bc_kind[b]; h.bc_alpha[b] = m->bc_alpha[b]; }
  h.nPeriodicVectors = m->nPeriodicVectors;
  for (int v = 0; v < m->nPeriodicVectors; ++v)
    for (int d = 0; d < 3; ++d) h.PeriodicVectors[v][d] = m->PeriodicVectors[v * 3 + d];
  h.arithmetic = p->arithmetic;
  h.dim_sf = p->dim_sf; h.dim_sf_dir = p->dim_sf_dir; h.dim_sf_dir1 = sfDir1; h.dim_sf_dir2 = sfDir2;
  h.dim_periodic_vec1 = pvec1; h.dim_periodic_vec2 = pvec2; h.nSFCases = nSFCases; h.alpha_sf = p->alpha_sf; h.sfDepo3D = p->sfDepo3D;
  memcpy(h.sfCase, sfCaseM, sizeof(sfCaseM));
  h.r_sf = p->r_sf; h.r2_sf = p->r_sf * p->r_sf; h.r2_sf_inv = (p->r_sf > 0.) ? 1. / (p->r_sf * p->r_sf) : 0.;
  h.w_sf = p->w_sf; h.dimFactorSF = p->dimFactorSF;
  for (int d = 0; d < 3; ++d) { h.FIBGMdeltas[d] = m->FIBGMdeltas[d]; h.xyzminglob[d] = m->xyzminglob[d]; h.FIBGMmin[d] = m->FIBGMmin[d]; h.FIBGMmax[d] = m->FIBGMmax[d]; }
  {
    // inverse Vandermonde matrix of the Gauss points: column i = monomial coefficients of the i-th Lagrange polynomial
    // (Gauss-Jordan with partial pivoting in long double; (N+1) <= 8)
    const int n = m->N + 1;
    long double V[8][16];
    for (int r = 0; r < n; ++r) {
      long double pw = 1.0L;
      for (int c = 0; c < n; ++c) { V[r][c] = pw; pw *= (long double)m->xGP[r]; }
      for (int c = 0; c < n; ++c) V[r][n + c] = (r == c) ? 1.0L : 0.0L;
    }
    for (int c = 0; c < n; ++c) {
      int piv = c;
      for (int r = c + 1; r < n; ++r) if (fabsl(V[r][c]) > fabsl(V[piv][c])) piv = r;
      for (int k = 0; k < 2 * n; ++k) std::swap(V[c][k], V[piv][k]);
      const long double d = V[c][c];
      for (int k = 0; k < 2 * n; ++k) V[c][k] /= d;
      for (int r = 0; r < n; ++r) {
        if (r == c) continue;
        const long double f = V[r][c];
        for (int k = 0; k < 2 * n; ++k) V[r][k] -= f * V[c][k];
      }
    }
    for (int pw = 0; pw < n; ++pw)
      for (int i = 0; i < n; ++i) h.n2m[pw][i] = (double)V[pw][n + i];
  }
  h.nGlobalElems = nG; h.nElems = g.nElems; h.offsetElem = g.offsetElem; h.N = g.N; h.nRanks = g.nRanks; h.myRank = g.myRank;
  CK(cudaMemcpyToSymbol(cst, &h, sizeof(h)));
  CK(cudaDeviceSynchronize());
  g.nPart = 0;
  g.hTailOff.assign(g.nRanks + 2, 0);
  g.ready = true;
  if (p->maxParticleNumber > 0 && reserve_particles(p->maxParticleNumber)) return 1;
  return 0;
}

// common part of an upload from the host and an emission on the device: `fill(c0, m)` stages records c0 .. c0+m-1 (PartState as
// [m][6], species, element, ParticleInside, IsNewPart, ids) in g.dStage / g.dStageI / g.dStageL; they are transposed into the particle
// arrays and the whole population is sorted by element.
extern "C++" {
namespace {
template <class Fill>
int upload_impl(const char* who, int64_t n, int32_t append, bool haveInside, bool haveIsNew, bool haveIds, bool haveRef, Fill fill) {
  if (g.binned) {   // back to the sorted arrays: the upload appends there and sorts; the next step re-plans the bins
    if (append) { if (bins_to_sorted()) return 1; }
    else { g.binned = false; g.sortedViewValid = false; }
  }
  const int64_t base = append ? g.nPart : 0;
  if (base + n >= (int64_t)0x7fffffff) return fail("%s: more than 2^31-1 particles on one GPU", who);
  if (reserve_particles(base + n)) return 1;
  begin_timing();
  const int64_t chunk = 1 << 22;
  if (reserve_stage(n < chunk ? (n ? n : 1) : chunk)) return 1;
  CK(cudaMemsetAsync(g.dCounters + 7, 0, sizeof(int), g.st));
  for (int64_t c0 = 0; c0 < n; c0 += chunk) {
    const int64_t m = (n - c0 < chunk) ? n - c0 : chunk;
    int32_t* dI = g.dStageI;
    if (fill(c0, m)) return 1;
    double* dRef = g.dStage + 6 * g.stageCap;
    k_aos_to_soa<<<(unsigned)((m + 255) / 256), 256, 0, g.st>>>(g.buf[g.cur], base + c0, m, g.dStage, dI, dI + g.stageCap,
                                                               haveInside ? dI + 2 * g.stageCap : nullptr,
                                                               haveIsNew ? dI + 3 * g.stageCap : nullptr,
                                                               (haveIds && g.carryIDs) ? g.dStageL : nullptr, base + c0,
                                                               haveRef ? dRef : nullptr, g.nGlobalElems, g.prm.nSpecies, g.dCounters + 7);
    ++g.lastLaunches;
    if (g.ref && !haveRef) {  // as at emission: PartPosRef from GetPositionInRefElem in the particle's element
      k_init_posref<<<(unsigned)((m + 127) / 128), 128, 0, g.st>>>(g.buf[g.cur], base + c0, m, g.dGeo);
      ++g.lastLaunches;
    }
    CK(cudaStreamSynchronize(g.st));
  }
  {
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, g.dCounters + 7, sizeof(int), cudaMemcpyDeviceToHost, g.st));
    CK(cudaStreamSynchronize(g.st));
    if (bad) {
      g.nPart = 0;   // the population is unusable: nothing of this upload is kept
      g.hTailOff.assign(g.nRanks + 2, 0);
      if (bad & 1) return fail("%s: GlobalElemID outside 1..nGlobalElems=%d", who, g.nGlobalElems);
      return fail("%s: PartSpecies outside 1..nSpecies=%d", who, g.prm.nSpecies);
    }
  }
  const int64_t nIn = base + n;
  if (nIn > 0) {
    k_keys_from_elem<<<(unsigned)((nIn + 255) / 256), 256, 0, g.st>>>(g.buf[g.cur].elem, g.dElemRank, g.dKeys, nIn, g.nElems,
                                                                      g.offsetElem, g.myRank, g.nRanks);
    ++g.lastLaunches;
  }
  if (sort_and_permute(nIn)) return 1;
  end_timing();
  if (g.nTotalSorted != g.nPart) return fail("%s: %lld particles lie in elements of other ranks", who, (long long)(g.nTotalSorted - g.nPart));
  return 0;
}
}  // namespace
}  // extern "C++"

int piclas_gpu_upload_particles(int64_t n, const double* PartState, const int32_t* PartSpecies, const int32_t* GlobalElemID,
                                const int32_t* ParticleInside, const int32_t* IsNewPart, const double* PartPosRef,
                                const int64_t* ids, int32_t append) {
  if (!g.ready) return fail("piclas_gpu_upload_particles: not initialised");
  if (g.exchangePending) return fail("piclas_gpu_upload_particles: the particle exchange of the last step is still open (piclas_gpu_exchange_finish)");
  CK(cudaSetDevice(g.device));
  if (n < 0) return fail("piclas_gpu_upload_particles: n < 0");
  if (n > 0 && (!PartState || !PartSpecies || !GlobalElemID)) return fail("piclas_gpu_upload_particles: null array");
  const bool haveRef = g.ref && PartPosRef;
  return upload_impl("piclas_gpu_upload_particles", n, append, ParticleInside != nullptr, IsNewPart != nullptr, ids != nullptr, haveRef,
                     [&](int64_t c0, int64_t m) {
    int32_t* dI = g.dStageI;
    CK(cudaMemcpyAsync(g.dStage, PartState + c0 * 6, m * 6 * 8, cudaMemcpyHostToDevice, g.st));
    CK(cudaMemcpyAsync(dI, PartSpecies + c0, m * 4, cudaMemcpyHostToDevice, g.st));
    CK(cudaMemcpyAsync(dI + g.stageCap, GlobalElemID + c0, m * 4, cudaMemcpyHostToDevice, g.st));
    if (ParticleInside) CK(cudaMemcpyAsync(dI + 2 * g.stageCap, ParticleInside + c0, m * 4, cudaMemcpyHostToDevice, g.st));
    if (IsNewPart) CK(cudaMemcpyAsync(dI + 3 * g.stageCap, IsNewPart + c0, m * 4, cudaMemcpyHostToDevice, g.st));
    if (ids && g.carryIDs) CK(cudaMemcpyAsync(g.dStageL, ids + c0, m * 8, cudaMemcpyHostToDevice, g.st));
    if (haveRef) CK(cudaMemcpyAsync(g.dStage + 6 * g.stageCap, PartPosRef + c0 * 3, m * 3 * 8, cudaMemcpyHostToDevice, g.st));
    return 0;
  });
}

// replaces the initial emission of one Part-Species[$]-Init[$] with SpaceIC = sin_deviation / cos_distribution and
// velocityDistribution = constant (SetParticlePosition + SetParticleVelocity, particle_emission.f90 -> particle_position_and_velocity.f90:
// 257-470, particle_emission_tools.f90:1235-1371): lattice positions, SinglePointToElement(doHALO = F) for each, particles outside the
// rank's elements dropped, IsNewPart set.
int piclas_gpu_emit_lattice(int32_t SpaceIC, int32_t iSpec, const int32_t* maxParticleNumber, double Amplitude, double WaveNumber,
                            const double* velocity, int32_t append, int64_t* nEmitted) {
  if (!g.ready) return fail("piclas_gpu_emit_lattice: not initialised");
  if (g.exchangePending) return fail("piclas_gpu_emit_lattice: the particle exchange of the last step is still open (piclas_gpu_exchange_finish)");
  CK(cudaSetDevice(g.device));
  if (SpaceIC != EMIT_SIN_DEVIATION && SpaceIC != EMIT_COS_DISTRIBUTION)
    return fail("piclas_gpu_emit_lattice: SpaceIC=%d; 1 (sin_deviation) and 2 (cos_distribution) are supported", SpaceIC);
  if (!maxParticleNumber || !velocity) return fail("piclas_gpu_emit_lattice: null array");
  if (iSpec < 1 || iSpec > g.prm.nSpecies) return fail("piclas_gpu_emit_lattice: species %d outside 1..nSpecies=%d", iSpec, g.prm.nSpecies);
  if (maxParticleNumber[0] < 1 || maxParticleNumber[1] < 1 || maxParticleNumber[2] < 1) return fail("piclas_gpu_emit_lattice: maxParticleNumberX/Y/Z < 1");
  if (SpaceIC == EMIT_COS_DISTRIBUTION && WaveNumber == 0.) return fail("piclas_gpu_emit_lattice: cos_distribution needs WaveNumber /= 0");
  if (!g.locT.FIBGM_nElems) return fail("piclas_gpu_emit_lattice: the mesh came without the FIBGM tables (GEO%%FIBGM*, FIBGM_Element)");
  EmitSpec s;
  s.kind = SpaceIC;
  s.nx = maxParticleNumber[0]; s.ny = maxParticleNumber[1]; s.nz = maxParticleNumber[2];
  s.amplitude = Amplitude; s.wavenumber = WaveNumber;
  for (int d = 0; d < 3; ++d) { s.velo[d] = velocity[d]; s.lo[d] = g.globLo[d]; s.len[d] = fabs(g.globHi[d] - g.globLo[d]); }
  s.species = iSpec;
  s.firstLocal = g.offsetElem + 1; s.lastLocal = g.offsetElem + g.nElems;
  const int64_t n = (int64_t)s.nx * s.ny * s.nz;
  unsigned long long* dAcc = nullptr;
  CK(cudaMalloc((void**)&dAcc, 8));
  CK(cudaMemsetAsync(dAcc, 0, 8, g.st));
  const int rc = upload_impl("piclas_gpu_emit_lattice", n, append, true, true, true, false, [&](int64_t c0, int64_t m) {
    int32_t* dI = g.dStageI;
    if (g.ref)
      k_emit_lattice<true><<<(unsigned)((m + 127) / 128), 128, 0, g.st>>>(g.locT, g.dTria, s, c0, m, g.dStage, dI, dI + g.stageCap, dI + 2 * g.stageCap,
                                                                         dI + 3 * g.stageCap, g.dStageL, dAcc);
    else
      k_emit_lattice<false><<<(unsigned)((m + 127) / 128), 128, 0, g.st>>>(g.locT, g.dTria, s, c0, m, g.dStage, dI, dI + g.stageCap, dI + 2 * g.stageCap,
                                                                          dI + 3 * g.stageCap, g.dStageL, dAcc);
    ++g.lastLaunches;
    CK(cudaGetLastError());
    return 0;
  });
  unsigned long long acc = 0;
  if (!rc) cudaMemcpy(&acc, dAcc, 8, cudaMemcpyDeviceToHost);
  cudaFree(dAcc);
  if (nEmitted) *nEmitted = (int64_t)acc;
  return rc;
}

int64_t piclas_gpu_num_particles(void) { return g.ready ? g.nPart : -1; }

int piclas_gpu_download_particles(int64_t nmax, double* PartState, int32_t* PartSpecies, int32_t* GlobalElemID, double* PartPosRef,
                                  int64_t* ids, int64_t* n_out) {
  if (!g.ready) return fail("piclas_gpu_download_particles: not initialised");
  if (g.exchangePending) return fail("piclas_gpu_download_particles: the particle exchange of the last step is still open (piclas_gpu_exchange_finish)");
  CK(cudaSetDevice(g.device));
  if (g.binned) {
    if (reserve_far(g.nPart, 0)) return 1;
    if (bins_sorted_view()) return 1;
  }
  const int64_t n = g.nPart;
  if (n_out) *n_out = n;
  if (n > nmax) return fail("piclas_gpu_download_particles: %lld particles do not fit into nmax=%lld", (long long)n, (long long)nmax);
  const int64_t chunk = 1 << 22;
  if (reserve_stage(n < chunk ? (n ? n : 1) : chunk)) return 1;
  for (int64_t c0 = 0; c0 < n; c0 += chunk) {
    const int64_t m = (n - c0 < chunk) ? n - c0 : chunk;
    double* dXi = g.dStage + 6 * g.stageCap;
    k_soa_to_aos<<<(unsigned)((m + 255) / 256), 256, 0, g.st>>>(g.buf[g.cur], c0, m, g.dStage, g.dStageI, g.dStageI + g.stageCap,
                                                               (PartPosRef && g.ref) ? dXi : nullptr, ids ? g.dStageL : nullptr);
    // TriaTracking keeps no PartPosRef: it is what GetPositionInRefElem yields for the current position and element, as in the
    // deposition (pic_depo_method.f90:479); the cached copy of the last deposition may be absent (closed-form path) or stale
    if (PartPosRef && !g.ref && m > 0)
      k_posref_download<<<(unsigned)((m + 127) / 128), 128, 0, g.st>>>(g.buf[g.cur], c0, m, g.dGeo, g.fast ? g.dAff : nullptr, dXi);
    if (PartState) CK(cudaMemcpyAsync(PartState + c0 * 6, g.dStage, m * 6 * 8, cudaMemcpyDeviceToHost, g.st));
    if (PartSpecies) CK(cudaMemcpyAsync(PartSpecies + c0, g.dStageI, m * 4, cudaMemcpyDeviceToHost, g.st));
    if (GlobalElemID) CK(cudaMemcpyAsync(GlobalElemID + c0, g.dStageI + g.stageCap, m * 4, cudaMemcpyDeviceToHost, g.st));
    if (PartPosRef) CK(cudaMemcpyAsync(PartPosRef + c0 * 3, dXi, m * 3 * 8, cudaMemcpyDeviceToHost, g.st));
    if (ids) CK(cudaMemcpyAsync(ids + c0, g.dStageL, m * 8, cudaMemcpyDeviceToHost, g.st));
    CK(cudaStreamSynchronize(g.st));
  }
  return 0;
}

int piclas_gpu_set_field(const double* E) {
  if (!g.ready) return fail("piclas_gpu_set_field: not initialised");
  if (!E) return fail("piclas_gpu_set_field: null field");
  CK(cudaSetDevice(g.device));
  CK(cudaMemcpyAsync(g.dE, E, (size_t)g.nElems * g.ND * 3 * 8, cudaMemcpyHostToDevice, g.st));
  if (g.dEmono && g.nElems > 0) {
    const int grid = g.nElems < g.nSMs * 16 ? g.nElems : g.nSMs * 16;
    switch (g.NP) {
      case 2: k_nodal_to_mono<2><<<grid, 128, 0, g.st>>>(g.dE, g.dEmono, g.nElems); break;
      case 3: k_nodal_to_mono<3><<<grid, 128, 0, g.st>>>(g.dE, g.dEmono, g.nElems); break;
      case 4: k_nodal_to_mono<4><<<grid, 128, 0, g.st>>>(g.dE, g.dEmono, g.nElems); break;
      case 5: k_nodal_to_mono<5><<<grid, 128, 0, g.st>>>(g.dE, g.dEmono, g.nElems); break;
      case 6: k_nodal_to_mono<6><<<grid, 128, 0, g.st>>>(g.dE, g.dEmono, g.nElems); break;
      case 7: k_nodal_to_mono<7><<<grid, 128, 0, g.st>>>(g.dE, g.dEmono, g.nElems); break;
      case 8: k_nodal_to_mono<8><<<grid, 128, 0, g.st>>>(g.dE, g.dEmono, g.nElems); break;
    }
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(g.st));
  g.haveField = true;
  return 0;
}

static int deposit_local() {
  const int grid = g.nElems < g.nSMs * 8 ? g.nElems : g.nSMs * 8;
  if (g.binEligible && ensure_binned()) return 1;
  const int gridB = g.nElems < g.nSMs * 16 ? g.nElems : g.nSMs * 16;   // k_bin_deposit_cvwm: one warp per CTA, 16 CTAs per SM
  cudaEventRecord(g.evp[0], g.st);
  if (grid > 0 && g.binned) {
    if (g.fast)
      k_bin_deposit_cvwm<true><<<gridB, DB_NT, 0, g.st>>>(g.bins, g.pool[g.binParity], bin_view(), g.binParity, g.offsetElem, g.dGeo, g.dTria, g.dAff,
                                                         g.dElemAcc);
    else
      k_bin_deposit_cvwm<false><<<gridB, DB_NT, 0, g.st>>>(g.bins, g.pool[g.binParity], bin_view(), g.binParity, g.offsetElem, g.dGeo, g.dTria, g.dAff,
                                                          g.dElemAcc);
    ++g.lastLaunches;
  } else if (grid > 0) {
    if (g.fast)
      k_deposit_cvwm<true><<<grid, STEP_NT, 0, g.st>>>(g.buf[g.cur], g.dElemOff, g.nElems, g.offsetElem, g.dGeo, g.dTria, g.dAff, g.dElemAcc,
                                                       g.ref ? 1 : 0);
    else
      k_deposit_cvwm<false><<<grid, STEP_NT, 0, g.st>>>(g.buf[g.cur], g.dElemOff, g.nElems, g.offsetElem, g.dGeo, g.dTria, g.dAff, g.dElemAcc, 1);
    ++g.lastLaunches;
  }
  cudaEventRecord(g.evp[1], g.st);
  k_node_sum<<<(g.nNodes * 4 + 255) / 256, 256, 0, g.st>>>(g.dAdjOff, g.dAdj, g.dElemAcc, g.dS, g.nNodes);
  ++g.lastLaunches;
  CK(cudaGetLastError());
  g.xiValid = !g.binned;
  return 0;
}

static int deposit_finish(double* PartSource, double* NodeSource) {
  k_node_final<<<(g.nNodes * 4 + 255) / 256, 256, 0, g.st>>>(g.dS, g.dPerN, g.dPerOff, g.dPerNodes, g.dNodeVolume, g.dNodeSource, g.nNodes,
                                                            1);
  ++g.lastLaunches;
  if (g.nElems > 0) {
    switch (g.NP) {
      case 2: launch_dofs<2>(); break;
      case 3: launch_dofs<3>(); break;
      case 4: launch_dofs<4>(); break;
      case 5: launch_dofs<5>(); break;
      case 6: launch_dofs<6>(); break;
      case 7: launch_dofs<7>(); break;
      case 8: launch_dofs<8>(); break;
    }
    ++g.lastLaunches;
  }
  CK(cudaGetLastError());
  cudaEventRecord(g.evp[2], g.st);
  end_timing();
  {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, g.evp[0], g.evp[1]);
    cudaEventElapsedTime(&b, g.evp[1], g.evp[2]);
    g.phaseMs[0] = a;
    g.phaseMs[1] = b;
  }
  if (PartSource) CK(cudaMemcpyAsync(PartSource, g.dPartSource, (size_t)g.nElems * g.ND * 4 * 8, cudaMemcpyDeviceToHost, g.st));
  if (NodeSource) CK(cudaMemcpyAsync(NodeSource, g.dNodeSource, (size_t)g.nNodes * 4 * 8, cudaMemcpyDeviceToHost, g.st));
  CK(cudaStreamSynchronize(g.st));
  return 0;
}

static int deposit_sf(double* PartSource) {
  if (g.nPart > g.sfFacCap) {
    for (int c = 0; c < 4; ++c) { cudaFree(g.dSfFac[c]); g.dSfFac[c] = nullptr; }
    g.sfFacCap = g.cap > g.nPart ? g.cap : g.nPart;
    for (int c = 0; c < 4; ++c) CK(cudaMalloc((void**)&g.dSfFac[c], (size_t)g.sfFacCap * 8));
  }
  CK(cudaMemsetAsync(g.dCounters, 0, 4 * sizeof(int), g.st));
  cudaEventRecord(g.evp[0], g.st);
  if (g.nPart > 0) {
    k_sf_prepare<<<(unsigned)((g.nPart + 127) / 128), 128, 0, g.st>>>(g.buf[g.cur], g.nPart, g.sfT, g.dSfFac[0], g.dSfFac[1], g.dSfFac[2],
                                                                    g.dSfFac[3], g.dCounters + 3);
    ++g.lastLaunches;
  }
  cudaEventRecord(g.evp[1], g.st);
  if (g.nSfTargets > 0) {
    int nt = ((g.ND + 31) / 32) * 32;
    if (nt < 64) nt = 64;
    const int grid = g.nSfTargets < g.nSMs * 16 ? g.nSfTargets : g.nSMs * 16;
    if (g.prm.arithmetic != 0 && g.prm.dim_sf == 3)
      k_sf_gather<true><<<grid, nt, 0, g.st>>>(g.buf[g.cur], g.dElemOff, g.nSfTargets, g.offsetElem, g.sfT, g.dSfFac[0], g.dSfFac[1], g.dSfFac[2],
                                               g.dSfFac[3], g.dPartSource);
    else
      k_sf_gather<false><<<grid, nt, 0, g.st>>>(g.buf[g.cur], g.dElemOff, g.nSfTargets, g.offsetElem, g.sfT, g.dSfFac[0], g.dSfFac[1], g.dSfFac[2],
                                                g.dSfFac[3], g.dPartSource);
    ++g.lastLaunches;
  }
  CK(cudaGetLastError());
  cudaEventRecord(g.evp[2], g.st);
  end_timing();
  {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, g.evp[0], g.evp[1]);
    cudaEventElapsedTime(&b, g.evp[1], g.evp[2]);
    g.phaseMs[0] = a;
    g.phaseMs[1] = b;
  }
  int hc[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(hc, g.dCounters, sizeof(hc), cudaMemcpyDeviceToHost, g.st));
  if (PartSource && g.nRanks == 1) CK(cudaMemcpyAsync(PartSource, g.dPartSource, (size_t)g.nElems * g.ND * 4 * 8, cudaMemcpyDeviceToHost, g.st));
  CK(cudaStreamSynchronize(g.st));
  if (hc[3]) return fail("piclas_gpu_deposit: charge-conserving shape function found no DOF within the radius of a particle");
  return 0;   // multi-rank: piclas_gpu_sf_halo_info -> exchange -> piclas_gpu_deposit_finish
}

int piclas_gpu_deposit(double* PartSource, double* NodeSource) {
  LbRange lb("LB_DEPO_SF: Deposition");
  if (!g.ready) return fail("piclas_gpu_deposit: not initialised");
  if (!g.prm.DoDeposition) return fail("piclas_gpu_deposit: PIC-DoDeposition=F");
  if (g.exchangePending) return fail("piclas_gpu_deposit: the particle exchange of the last step is still open (piclas_gpu_exchange_finish)");
  CK(cudaSetDevice(g.device));
  if (g.psCopyPending) CK(cudaStreamWaitEvent(g.st, g.evCopyDone, 0));   // the last PartSource is still on its way to the host
  begin_timing();
  if (g.sfActive) {
    if (NodeSource) return fail("piclas_gpu_deposit: NodeSource exists for cell_volweight_mean only");
    return deposit_sf(PartSource);
  }
  if (deposit_local()) return 1;
  if (g.nRanks > 1) {  // caller sums the node array over ranks, then piclas_gpu_deposit_finish
    end_timing();
    CK(cudaStreamSynchronize(g.st));
    return 0;
  }
  return deposit_finish(PartSource, NodeSource);
}

// ---- compact node halo (pic_depo_method.f90:565-673): only the nodes that several ranks contribute to travel ------------------------------
__global__ void k_halo_pack(const double* __restrict__ S, const int32_t* __restrict__ nodes, int n, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * 4) out[t] = S[(size_t)nodes[t >> 2] * 4 + (t & 3)];
}
// S[node] = contributions of rank 0 + rank 1 + ... in rank order (deterministic, as the receive loop of the reference, :659-673)
__global__ void k_halo_sum(const double* __restrict__ all, int nRanks, const int32_t* __restrict__ nodes, int n, double* __restrict__ S) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 4) return;
  double s = all[t];
  for (int r = 1; r < nRanks; ++r) s = s + all[(size_t)r * n * 4 + t];
  S[(size_t)nodes[t >> 2] * 4 + (t & 3)] = s;
}

__global__ void k_extract_component(const double* __restrict__ src4, double* __restrict__ dst, int64_t n, int comp) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src4[i * 4 + comp];
}

int piclas_gpu_get_charge(double* ChargeDensity) {
  if (!g.ready) return fail("piclas_gpu_get_charge: not initialised");
  if (!g.prm.DoDeposition) return fail("piclas_gpu_get_charge: PIC-DoDeposition=F");
  if (!ChargeDensity) return fail("piclas_gpu_get_charge: null array");
  CK(cudaSetDevice(g.device));
  const int64_t n = (int64_t)g.nElems * g.ND;
  if (n == 0) return 0;
  if (!g.dCharge) CK(cudaMalloc((void**)&g.dCharge, (size_t)n * 8));
  k_extract_component<<<(unsigned)((n + 255) / 256), 256, 0, g.st>>>(g.dPartSource, g.dCharge, n, 3);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ChargeDensity, g.dCharge, (size_t)n * 8, cudaMemcpyDeviceToHost, g.st));
  CK(cudaStreamSynchronize(g.st));
  return 0;
}

// PartSource of the last deposition to the host WITHOUT stopping the step: the copy runs on its own stream behind the deposition
// and beside what follows (piclas_gpu_set_field's host -> device copy uses the other direction of the link, the push kernels the
// SMs).  The Poisson solve needs the charge component only (piclas_gpu_get_charge); the current density is output / analysis data.
int piclas_gpu_get_partsource_async(double* PartSource) {
  if (!g.ready) return fail("piclas_gpu_get_partsource_async: not initialised");
  if (!g.prm.DoDeposition) return fail("piclas_gpu_get_partsource_async: PIC-DoDeposition=F");
  if (!PartSource) return fail("piclas_gpu_get_partsource_async: null array");
  CK(cudaSetDevice(g.device));
  if (!g.stCopy) {
    CK(cudaStreamCreateWithFlags(&g.stCopy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&g.evDepoDone, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.evCopyDone, cudaEventDisableTiming));
  }
  const size_t bytes = (size_t)g.nElems * g.ND * 4 * 8;
  CK(cudaEventRecord(g.evDepoDone, g.st));
  CK(cudaStreamWaitEvent(g.stCopy, g.evDepoDone, 0));
  if (bytes) CK(cudaMemcpyAsync(PartSource, g.dPartSource, bytes, cudaMemcpyDeviceToHost, g.stCopy));
  CK(cudaEventRecord(g.evCopyDone, g.stCopy));
  g.psCopyPending = true;
  return 0;
}

// blocks until the array of the last piclas_gpu_get_partsource_async is complete on the host
int piclas_gpu_partsource_wait(void) {
  if (!g.ready) return fail("piclas_gpu_partsource_wait: not initialised");
  if (!g.psCopyPending) return 0;
  CK(cudaSetDevice(g.device));
  CK(cudaEventSynchronize(g.evCopyDone));
  g.psCopyPending = false;
  return 0;
}

// ---- particle analysis reductions ---------------------------------------------------------------------------------------------
// CalcKineticEnergy (particle_analyze_tools.f90:709-842): per-species sums over the rank's particles.  One CTA sums a contiguous
// chunk, threads stride through it, fixed shuffle / shared-memory tree: the result does not depend on scheduling.
constexpr int AN_NT = 256;
__global__ void __launch_bounds__(AN_NT) k_kinetic_energy(PartBuf pb, int64_t n, int64_t chunk, int nSpecies, double* __restrict__ partE,
                                                          unsigned long long* __restrict__ partN) {
  __shared__ double sE[AN_NT / 32];
  __shared__ unsigned long long sN[AN_NT / 32];
  const int64_t p0 = (int64_t)blockIdx.x * chunk, p1 = (p0 + chunk < n) ? p0 + chunk : n;
  const double c2 = 1.0 / cst.c2_inv;
  const double RelativisticLimit = (1e6 / 299792458.0) * (1e6 / 299792458.0) * c2;   // globals_init.f90:111
  for (int sp = 0; sp < nSpecies; ++sp) {
    double e = 0.;
    unsigned long long cnt = 0;
    for (int64_t p = p0 + threadIdx.x; p < p1; p += AN_NT) {
      if ((pb.meta[p] & META_SPEC_MASK) != sp) continue;
      const double v0 = pb.v[0][p], v1 = pb.v[1][p], v2 = pb.v[2][p];
      const double partV2 = (v0 * v0 + v1 * v1) + v2 * v2;
      double Ekin_loc;
      if (partV2 < RelativisticLimit) Ekin_loc = 0.5 * cst.MassIC[sp] * partV2;
      else {
        double GammaFac = partV2 * cst.c2_inv;
        GammaFac = 1. / sqrt(1. - GammaFac);
        Ekin_loc = (GammaFac - 1.) * cst.MassIC[sp] * c2;
      }
      e = e + Ekin_loc * cst.MPF[sp];
      ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e = e + __shfl_down_sync(0xffffffffu, e, o);
      cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { sE[threadIdx.x >> 5] = e; sN[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double te = 0.;
      unsigned long long tn = 0;
      for (int w = 0; w < AN_NT / 32; ++w) { te = te + sE[w]; tn += sN[w]; }
      partE[(size_t)blockIdx.x * nSpecies + sp] = te;
      partN[(size_t)blockIdx.x * nSpecies + sp] = tn;
    }
    __syncthreads();
  }
}

int piclas_gpu_kinetic_energy(double* Ekin, int64_t* nPart) {
  if (!g.ready) return fail("piclas_gpu_kinetic_energy: not initialised");
  if (g.exchangePending) return fail("piclas_gpu_kinetic_energy: the particle exchange of the last step is still open (piclas_gpu_exchange_finish)");
  CK(cudaSetDevice(g.device));
  const int ns = g.prm.nSpecies;
  for (int s = 0; s < ns; ++s) { if (Ekin) Ekin[s] = 0.; if (nPart) nPart[s] = 0; }
  if (g.nPart == 0) return 0;
  if (g.binned) {
    if (reserve_far(g.nPart, 0)) return 1;
    if (bins_sorted_view()) return 1;
  }
  const int nb = g.nSMs * 8;
  const int64_t chunk = (g.nPart + nb - 1) / nb;
  double* dE = nullptr;
  unsigned long long* dN = nullptr;
  CK(cudaMalloc((void**)&dE, (size_t)nb * ns * 8));
  CK(cudaMalloc((void**)&dN, (size_t)nb * ns * 8));
  k_kinetic_energy<<<nb, AN_NT, 0, g.st>>>(g.buf[g.cur], g.nPart, chunk, ns, dE, dN);
  std::vector<double> hE((size_t)nb * ns);
  std::vector<unsigned long long> hN((size_t)nb * ns);
  cudaError_t e1 = cudaMemcpyAsync(hE.data(), dE, hE.size() * 8, cudaMemcpyDeviceToHost, g.st);
  cudaError_t e2 = cudaMemcpyAsync(hN.data(), dN, hN.size() * 8, cudaMemcpyDeviceToHost, g.st);
  cudaError_t e3 = cudaStreamSynchronize(g.st);
  cudaFree(dE); cudaFree(dN);
  CK(e1); CK(e2); CK(e3);
  for (int b = 0; b < nb; ++b)          // block order: fixed summation order
    for (int s = 0; s < ns; ++s) {
      if (Ekin) Ekin[s] = Ekin[s] + hE[(size_t)b * ns + s];
      if (nPart) nPart[s] += (int64_t)hN[(size_t)b * ns + s];
    }
  return 0;
}

int piclas_gpu_nodesource_device(void** devNodeSource) {
  if (!g.ready) return fail("piclas_gpu_nodesource_device: not initialised");
  *devNodeSource = g.dS;
  return 0;
}

int piclas_gpu_node_halo_info(int64_t* nDoubles, void** devSend, void** devRecvAll) {
  if (!g.ready) return fail("piclas_gpu_node_halo_info: not initialised");
  if (g.nRanks < 2 || !g.dHaloNodes) return fail("piclas_gpu_node_halo_info: needs several ranks and cell_volweight_mean");
  CK(cudaSetDevice(g.device));
  if (g.nHaloNodes > 0) {
    k_halo_pack<<<(g.nHaloNodes * 4 + 255) / 256, 256, 0, g.st>>>(g.dS, g.dHaloNodes, g.nHaloNodes, g.dHaloSend);
    ++g.lastLaunches;
    CK(cudaGetLastError());
  }
  g.haloPacked = true;
  *nDoubles = (int64_t)g.nHaloNodes * 4;
  *devSend = g.dHaloSend;
  *devRecvAll = g.dHaloRecv;
  return 0;
}

int piclas_gpu_set_stream(void* cudaStream) {
  if (!g.ready) return fail("piclas_gpu_set_stream: not initialised");
  CK(cudaSetDevice(g.device));
  CK(cudaStreamSynchronize(g.st));
  if (g.ownStream && g.st) cudaStreamDestroy(g.st);
  g.st = (cudaStream_t)cudaStream;
  g.ownStream = false;
  return 0;
}

int piclas_gpu_sf_halo_info(int64_t* nSendElemsPerRank, int64_t* nRecvElemsPerRank, int32_t* doublesPerElem, void** devSend,
                            void** devRecv) {
  if (!g.ready || !g.sfActive) return fail("piclas_gpu_sf_halo_info: shape-function deposition is not active");
  for (int r = 0; r < g.nRanks; ++r) { nSendElemsPerRank[r] = g.sfSendCount[r]; nRecvElemsPerRank[r] = g.sfRecvCount[r]; }
  *doublesPerElem = g.ND * 4;
  *devSend = g.dPartSource + (size_t)g.nElems * g.ND * 4;   // halo targets follow the local elements, grouped by owner rank
  *devRecv = g.dSfRecv;
  return 0;
}

int piclas_gpu_deposit_finish(double* PartSource, double* NodeSource) {
  LbRange lb("LB_DEPO_SF: halo sum + node -> DOF");
  if (!g.ready) return fail("piclas_gpu_deposit_finish: not initialised");
  CK(cudaSetDevice(g.device));
  if (g.sfActive) {
    if (g.nSfRecvElems > 0) {
      const int nd4 = g.ND * 4;
      k_sf_add_halo<<<g.nSfRecvElems, 128, 0, g.st>>>(g.dPartSource, g.dSfRecv, g.dSfRecvElem, g.dSfRecvOff, g.dSfRecvIdx, nd4);
      ++g.lastLaunches;
      CK(cudaGetLastError());
    }
    if (PartSource) CK(cudaMemcpyAsync(PartSource, g.dPartSource, (size_t)g.nElems * g.ND * 4 * 8, cudaMemcpyDeviceToHost, g.st));
    CK(cudaStreamSynchronize(g.st));
    return 0;
  }
  const int keep = g.lastLaunches;
  const double ms = g.lastMs;
  begin_timing();
  g.lastLaunches = keep;
  if (g.haloPacked) {   // the caller gathered every rank's compact contributions into dHaloRecv (piclas_gpu_node_halo_info)
    if (g.nHaloNodes > 0) {
      k_halo_sum<<<(g.nHaloNodes * 4 + 255) / 256, 256, 0, g.st>>>(g.dHaloRecv, g.nRanks, g.dHaloNodes, g.nHaloNodes, g.dS);
      ++g.lastLaunches;
    }
    g.haloPacked = false;
  }
  const int rc = deposit_finish(PartSource, NodeSource);
  g.lastMs += ms;
  return rc;
}

// the step on the bins: push + delivery, exact walk of the far list, pool for the next step
static double host_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int push_track_binned(double dt, int32_t* nLost) {
  const double t0 = host_ms();
  if (ensure_binned()) return 1;
  if (reserve_far(g.nPart, 0)) return 1;   // the far list (in the idle sorted buffers) can take every particle
  if (g.farIdxCap < g.cap) {               // (the sorted buffers may have been sized by an upload)
    cudaFree(g.dFarIdx);
    CK(cudaMalloc((void**)&g.dFarIdx, (size_t)(g.cap ? g.cap : 1) * 4));
    g.farIdxCap = g.cap;
  }
  begin_timing();
  CK(cudaMemsetAsync(g.dCounters, 0, 8 * sizeof(int), g.st));
  if (g.nRanks > 1) CK(cudaMemsetAsync(g.dEmigCnt, 0, (size_t)g.nRanks * sizeof(int), g.st));
  cudaEventRecord(g.evp[3], g.st);
  const int ne = g.nElems;
  if (ne > 0) {
    // far regions: element e may use as many slots of the far list as it holds particles
    k_bin_count<<<(ne + 255) / 256, 256, 0, g.st>>>(bin_view(), g.binParity, g.dBinTmp);
    ++g.lastLaunches;
    if (scan_i64(g.dBinTmp, ne, g.dFarBase)) return 1;
    switch (g.NP) {
      case 2: launch_bin_push<2>(dt); break;
      case 3: launch_bin_push<3>(dt); break;
      case 4: launch_bin_push<4>(dt); break;
      case 5: launch_bin_push<5>(dt); break;
      case 6: launch_bin_push<6>(dt); break;
      case 7: launch_bin_push<7>(dt); break;
      case 8: launch_bin_push<8>(dt); break;
    }
    k_far_total<<<(ne + 255) / 256, 256, 0, g.st>>>(g.dNFarE, ne, g.dBinTmp);
    ++g.lastLaunches;
    if (scan_i64(g.dBinTmp, ne, g.dFarDOff)) return 1;
    CK(cudaGetLastError());
  }
  int hc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t nFar = 0;
  {
    int* pc = reinterpret_cast<int*>(g.hPin);   // pinned: a true asynchronous copy, no staging in the driver
    g.hPin[8] = 0;
    CK(cudaMemcpyAsync(pc, g.dCounters, sizeof(hc), cudaMemcpyDeviceToHost, g.st));
    if (ne > 0) CK(cudaMemcpyAsync(&g.hPin[8], g.dFarDOff + ne, 8, cudaMemcpyDeviceToHost, g.st));
    CK(cudaStreamSynchronize(g.st));
    memcpy(hc, pc, sizeof(hc));
    nFar = g.hPin[8];
  }
  const double t1 = host_ms();
  if (nFar > 0) {
    const bool listed = hc[7] <= UNPUSHED_LIST;   // few: k_bin_push noted their far slots; many: scan of the indexed far list
    if (hc[7] > 0) {   // particles the call-free push kernel handed over unpushed
      if (!listed) {
        k_far_index<<<ne < g.nSMs * 16 ? ne : g.nSMs * 16, 128, 0, g.st>>>(g.dFarBase, g.dNFarE, g.dFarDOff, ne, g.dFarIdx);
        ++g.lastLaunches;
      }
      const int64_t nu = listed ? hc[7] : nFar;
      switch (g.NP) {
        case 2: launch_far_unpushed<2>(nu, dt, listed); break;
        case 3: launch_far_unpushed<3>(nu, dt, listed); break;
        case 4: launch_far_unpushed<4>(nu, dt, listed); break;
        case 5: launch_far_unpushed<5>(nu, dt, listed); break;
        case 6: launch_far_unpushed<6>(nu, dt, listed); break;
        case 7: launch_far_unpushed<7>(nu, dt, listed); break;
        case 8: launch_far_unpushed<8>(nu, dt, listed); break;
      }
    }
    if (launch_far_walk(nFar, hc[7] > 0 && !listed)) return 1;
  }
  CK(cudaGetLastError());
  cudaEventRecord(g.evp[4], g.st);
  g.nFar = nFar;
  g.farStats[0] = nFar; g.farStats[1] = hc[4]; g.farStats[2] = hc[5]; g.farStats[3] = hc[6];
  {
    // Particles that met a full region took the far list (correct, only slower: 0.1 ns per record, measured as 3.4 ms for 3.6e7 far
    // records).  Re-planning the capacities costs about 3.5 steps (130 ms at 5e8 particles), so it waits until the detours since
    // the last plan have cost as much (rent or buy: at most twice the cost of the best choice in hindsight).
    const int64_t diverted = (int64_t)hc[5] + hc[6];
    const double pushMs = host_ms() - t0;   // the push kernel has been waited for (nFar read back): about 3/4 of a step
    g.detourMs += 1.0e-7 * (double)diverted;
    bool replan = g.detourMs > 4.5 * pushMs;
    if (const char* v = getenv("PICLAS_GPU_REBIN_MIN")) replan = diverted > atoll(v);   // tests: re-plan on the first diverted particle
    if (hc[6] > hc[5] && g.inFracCapped && !getenv("PICLAS_GPU_REBIN_MIN")) replan = false;   // larger inboxes do not fit
    if (replan) {
      g.wantRebin = true;                                      // capacities from the new populations before the next step
      g.detourMs = 0.;
      if (hc[6] > hc[5] && g.inFrac < 0.5) g.inFrac *= 2.0;    // inboxes too small for this flow (drifting populations)
    }
  }
  g.sortedViewValid = false;
  if (g.nRanks > 1) {
    // several ranks: the emigrants are taken out of the far list (piclas_gpu_exchange_info), the immigrants join it, and
    // piclas_gpu_exchange_finish sorts it into the pool.  The walk counted the emigrants per destination rank.
    g.hEmigCnt.assign(g.nRanks, 0);
    if (g.nRanks <= 128) {
      int* pc = reinterpret_cast<int*>(g.hPin + 32);
      CK(cudaMemcpyAsync(pc, g.dEmigCnt, (size_t)g.nRanks * sizeof(int), cudaMemcpyDeviceToHost, g.st));
      CK(cudaStreamSynchronize(g.st));
      for (int r = 0; r < g.nRanks; ++r) g.hEmigCnt[r] = pc[r];
    } else {
      std::vector<int> tmp(g.nRanks);
      CK(cudaMemcpyAsync(tmp.data(), g.dEmigCnt, (size_t)g.nRanks * sizeof(int), cudaMemcpyDeviceToHost, g.st));
      CK(cudaStreamSynchronize(g.st));
      for (int r = 0; r < g.nRanks; ++r) g.hEmigCnt[r] = tmp[r];
    }
    g.exchangePending = true;
    g.nUnsorted = nFar;
  } else {
    if (far_finish(nFar)) return 1;
    g.nPart = g.nPart - nFar + (nFar > 0 ? g.hTailOff[0] : 0);
  }
  const double t2 = host_ms();
  {
    int* pc = reinterpret_cast<int*>(g.hPin);
    CK(cudaMemcpyAsync(pc, g.dCounters, sizeof(hc), cudaMemcpyDeviceToHost, g.st));
    CK(cudaStreamSynchronize(g.st));
    memcpy(hc, pc, sizeof(hc));
  }
  if (getenv("PICLAS_GPU_DEBUG"))
    fprintf(stderr, "[piclas_gpu] push_track (bins): %lld particles, %d delivered to face neighbours in-kernel, %lld through the far list, "
            "%d + %d diverted by full regions, %d walked with the determinant tests; host ms: push kernel %.2f, walk + far list %.2f, tail %.2f\n",
            (long long)g.nPart, (int)g.farStats[1], (long long)nFar, (int)g.farStats[2], (int)g.farStats[3], hc[2], t1 - t0, t2 - t1, host_ms() - t2);
  cudaEventRecord(g.evp[5], g.st);
  end_timing();
  {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, g.evp[3], g.evp[4]);
    cudaEventElapsedTime(&b, g.evp[4], g.evp[5]);
    g.phaseMs[2] = a;
    g.phaseMs[3] = b;
  }
  if (nLost) *nLost = hc[0];
  if (hc[1] == TRK_ERR_BC) return fail("piclas_gpu_push_track: particle hit a boundary condition that is not supported");
  if (hc[1] == TRK_ERR_ELEM) return fail("piclas_gpu_push_track: ERROR: Element not defined! Please increase the size of the halo region (HaloEpsVelo)!");
  if (hc[1] == TRK_ERR_LOOP) return fail("piclas_gpu_push_track: tracking loop did not terminate");
  return 0;
}

int piclas_gpu_push_track(double dt, int64_t iter, int32_t* nLost) {
  LbRange lb("LB_INTERPOLATION + LB_PUSH + LB_TRACK + LB_UNFP: push_track");
  (void)iter;
  if (!g.ready) return fail("piclas_gpu_push_track: not initialised");
  CK(cudaSetDevice(g.device));
  if (g.prm.DoInterpolation && !g.haveField) return fail("piclas_gpu_push_track: no field set (piclas_gpu_set_field)");
  if (g.exchangePending) return fail("piclas_gpu_push_track: the particle exchange of the last step is still open (piclas_gpu_exchange_finish)");
  if (g.binEligible) return push_track_binned(dt, nLost);
  begin_timing();
  CK(cudaMemsetAsync(g.dCounters, 0, 8 * sizeof(int), g.st));
  cudaEventRecord(g.evp[3], g.st);
  if (g.nPart > 0) {
    switch (g.NP) {
      case 2: launch_push_track<2>(dt); break;
      case 3: launch_push_track<3>(dt); break;
      case 4: launch_push_track<4>(dt); break;
      case 5: launch_push_track<5>(dt); break;
      case 6: launch_push_track<6>(dt); break;
      case 7: launch_push_track<7>(dt); break;
      case 8: launch_push_track<8>(dt); break;
    }
    CK(cudaGetLastError());
  }
  cudaEventRecord(g.evp[4], g.st);
  int hc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CK(cudaMemcpyAsync(hc, g.dCounters, sizeof(hc), cudaMemcpyDeviceToHost, g.st));
  const int64_t nBefore = g.nPart;
  if (g.nRanks > 1) {
    // several ranks: the emigrants are taken out of the unsorted arrays (piclas_gpu_exchange_info) and the one sort of the
    // step runs after the immigrants have arrived (piclas_gpu_exchange_finish)
    CK(cudaStreamSynchronize(g.st));
    g.exchangePending = true;
    g.nUnsorted = g.nPart;
    g.xiValid = false;
  } else if (sort_and_permute(g.nPart)) return 1;   // also synchronises
  if (getenv("PICLAS_GPU_DEBUG"))
    fprintf(stderr, "[piclas_gpu] push_track: %lld particles, first crossing in-kernel: %d by side planes, %d by determinants; %d to the walk kernel\n",
            (long long)nBefore, hc[4], hc[5], hc[2]);
  cudaEventRecord(g.evp[5], g.st);
  end_timing();
  {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, g.evp[3], g.evp[4]);
    cudaEventElapsedTime(&b, g.evp[4], g.evp[5]);
    g.phaseMs[2] = a;
    g.phaseMs[3] = b;
  }
  if (nLost) *nLost = hc[0];
  if (hc[1] == TRK_ERR_BC) return fail("piclas_gpu_push_track: particle hit a boundary condition that is not supported");
  if (hc[1] == TRK_ERR_ELEM) return fail("piclas_gpu_push_track: ERROR: Element not defined! Please increase the size of the halo region (HaloEpsVelo)!");
  if (hc[1] == TRK_ERR_LOOP) return fail("piclas_gpu_push_track: tracking loop did not terminate");
  return 0;
}

int piclas_gpu_exchange_info(int32_t* partCommSize, int64_t* nSendPerRank, void** devSendBuf) {
  LbRange lb("LB_PARTCOMM: emigrant extraction + pack");
  if (!g.ready) return fail("piclas_gpu_exchange_info: not initialised");
  CK(cudaSetDevice(g.device));
  if (partCommSize) *partCommSize = g.commSize;
  for (int r = 0; r < g.nRanks; ++r) nSendPerRank[r] = 0;
  if (devSendBuf) *devSendBuf = g.dCommSend;
  g.nEmig = 0;
  if (!g.exchangePending) return 0;   // no step since the last exchange: nothing to send
  cudaEventRecord(g.evp[6], g.st);
  const int64_t n = g.nUnsorted;
  const uint32_t lo = (uint32_t)g.nElems, hi = (uint32_t)(g.nElems + g.nRanks);
  const int64_t nTiles = (n + EM_TILE - 1) / EM_TILE;
  if (!g.dEmigOff) CK(cudaMalloc((void**)&g.dEmigOff, (size_t)(g.nRanks + 2) * sizeof(int64_t)));
  if (nTiles > g.tileCntCap) {
    cudaFree(g.dTileCnt);
    g.tileCntCap = nTiles + nTiles / 4 + 16;
    CK(cudaMalloc((void**)&g.dTileCnt, (size_t)g.tileCntCap * 4));
  }
  int64_t nEmig = 0;
  const bool known = g.binned && (int)g.hEmigCnt.size() == g.nRanks;   // bins: k_far_walk counted the emigrants of every rank
  if (n > 0) {
    k_emig_count<<<(unsigned)nTiles, EM_NT, 0, g.st>>>(g.dKeys, n, lo, hi, g.dTileCnt);
    k_emig_scan<<<1, 1024, 0, g.st>>>(g.dTileCnt, (uint32_t)nTiles, g.dEmigOff);
    g.lastLaunches += 2;
    if (known) {
      for (int r = 0; r < g.nRanks; ++r) nEmig += g.hEmigCnt[r];
    } else {
      CK(cudaMemcpyAsync(&nEmig, g.dEmigOff, sizeof(int64_t), cudaMemcpyDeviceToHost, g.st));
      CK(cudaStreamSynchronize(g.st));
    }
  }
  if (nEmig > g.emigCap) {
    cudaFree(g.dEmigIdx); cudaFree(g.dEmigKey);
    g.emigCap = nEmig + nEmig / 4 + 1024;
    CK(cudaMalloc((void**)&g.dEmigIdx, (size_t)g.emigCap * 4));
    CK(cudaMalloc((void**)&g.dEmigKey, (size_t)g.emigCap * 4));
  }
  if (nEmig > g.commSendCap) {
    cudaFree(g.dCommSend);
    g.commSendCap = nEmig + nEmig / 4 + 1024;
    CK(cudaMalloc((void**)&g.dCommSend, g.commSendCap * g.commSize * 8));
  }
  if (nEmig > 0) {
    k_emig_compact<<<(unsigned)nTiles, EM_NT, 0, g.st>>>(g.dKeys, n, lo, hi, g.dTileCnt, g.dEmigIdx, g.dEmigKey);
    ++g.lastLaunches;
    int rankBits = 1;
    while ((1 << rankBits) < g.nRanks) ++rankBits;
    uint32_t *sk = nullptr, *perm = nullptr;
    CK(radix_sort_by_key(g.sortws, g.dEmigKey, (size_t)nEmig, rankBits, g.st, &sk, &perm, &g.lastLaunches));   // stable: by rank, then particle order
    CK(segment_offsets(sk, (size_t)nEmig, (uint32_t)g.nRanks, g.dEmigOff, g.st));
    // (bins: the far list lies in buf[0] with the field layout of the sorted arrays)
    k_pack_emigrants_idx<<<(unsigned)((nEmig + 255) / 256), 256, 0, g.st>>>(g.buf[g.binned ? 0 : g.cur], g.dEmigIdx, perm, nEmig, g.commSize, g.ref ? 1 : 0,
                                                                             g.dCommSend, g.dKeys, (uint32_t)(g.nElems + g.nRanks),
                                                                             g.binned ? g.dFarIdx : nullptr);
    g.lastLaunches += 2;
    CK(cudaGetLastError());
    if (known) {
      for (int r = 0; r < g.nRanks; ++r) nSendPerRank[r] = g.hEmigCnt[r];
    } else {
      std::vector<int64_t> off(g.nRanks + 1, 0);
      CK(cudaMemcpyAsync(off.data(), g.dEmigOff, (size_t)(g.nRanks + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, g.st));
      CK(cudaStreamSynchronize(g.st));
      for (int r = 0; r < g.nRanks; ++r) nSendPerRank[r] = off[r + 1] - off[r];
    }
    if (nSendPerRank[g.myRank] != 0) return fail("piclas_gpu_exchange_info: internal error, emigrants addressed to the own rank");
  }
  // the counts on the device as well (piclas_gpu_exchange_device_info): a count exchange without a host->device copy
  for (int r = 0; r < g.nRanks && r < 200; ++r) g.hPin[48 + r] = nSendPerRank[r];
  if (g.nRanks <= 200) CK(cudaMemcpyAsync(g.dSendCounts, g.hPin + 48, (size_t)g.nRanks * sizeof(int64_t), cudaMemcpyHostToDevice, g.st));
  g.hEmigCnt.clear();
  g.nEmig = nEmig;
  cudaEventRecord(g.evp[7], g.st);
  if (devSendBuf) *devSendBuf = g.dCommSend;
  return 0;
}

int piclas_gpu_exchange_device_info(void** devSendCounts, int64_t* sendCapDoubles, int64_t* recvCapDoubles) {
  if (!g.ready) return fail("piclas_gpu_exchange_device_info: not initialised");
  if (devSendCounts) *devSendCounts = g.dSendCounts;
  if (sendCapDoubles) *sendCapDoubles = g.commSendCap * g.commSize;
  if (recvCapDoubles) *recvCapDoubles = g.commRecvCap * g.commSize;
  return 0;
}

int piclas_gpu_exchange_recv_buffer(int64_t nRecvTotal, void** devRecvBuf) {
  if (!g.ready) return fail("piclas_gpu_exchange_recv_buffer: not initialised");
  CK(cudaSetDevice(g.device));
  if (nRecvTotal > g.commRecvCap || !g.dCommRecv) {
    cudaFree(g.dCommRecv);
    g.commRecvCap = nRecvTotal + nRecvTotal / 4 + 1024;
    CK(cudaMalloc((void**)&g.dCommRecv, g.commRecvCap * g.commSize * 8));
  }
  *devRecvBuf = g.dCommRecv;
  return 0;
}

int piclas_gpu_exchange_finish(int64_t nRecvTotal) {
  LbRange lb("LB_PARTCOMM + LB_UNFP: unpack + far list sort");
  if (!g.ready) return fail("piclas_gpu_exchange_finish: not initialised");
  CK(cudaSetDevice(g.device));
  if (!g.exchangePending) {
    if (nRecvTotal != 0) return fail("piclas_gpu_exchange_finish: particles arrive but no step is open on this rank");
    return 0;
  }
  if (nRecvTotal > 0 && nRecvTotal > g.commRecvCap) return fail("piclas_gpu_exchange_finish: receive buffer too small");
  if (g.binned) {
    // the immigrants join the far list behind every element's region, then one sort by destination builds the pool
    const int64_t nFar = g.nFar, nAll = nFar + nRecvTotal, slot0 = g.nPart;
    if (slot0 + nRecvTotal >= (int64_t)0x7fffffff) return fail("piclas_gpu_exchange_finish: more than 2^31-1 particles on one GPU");
    if (reserve_far(slot0 + nRecvTotal, slot0 > nFar ? slot0 : nFar)) return 1;
    cudaEventRecord(g.evp[8], g.st);
    if (nRecvTotal > 0) {
      k_unpack_immigrants<<<(unsigned)((nRecvTotal + 255) / 256), 256, 0, g.st>>>(g.buf[0], slot0, nRecvTotal, g.commSize, 0, g.dCommRecv);
      k_keys_from_elem<<<(unsigned)((nRecvTotal + 255) / 256), 256, 0, g.st>>>(g.buf[0].elem + slot0, g.dElemRank, g.dKeys + nFar, nRecvTotal,
                                                                               g.nElems, g.offsetElem, g.myRank, g.nRanks);
      k_far_index_immigrants<<<(unsigned)((nRecvTotal + 255) / 256), 256, 0, g.st>>>(g.dFarIdx, nFar, nRecvTotal, (uint32_t)slot0);
      g.lastLaunches += 3;
      CK(cudaGetLastError());
    }
    g.exchangePending = false;
    if (far_finish(nAll)) return 1;
    const int64_t staying = nAll > 0 ? g.hTailOff[0] : 0;
    if (nAll > 0 && g.hTailOff[g.nRanks] != staying) return fail("piclas_gpu_exchange_finish: received particles that belong to another rank");
    g.nPart = g.nPart - nFar + staying;
    cudaEventRecord(g.evp[9], g.st);
    cudaEventSynchronize(g.evp[9]);
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, g.evp[6], g.evp[7]);
    cudaEventElapsedTime(&b, g.evp[8], g.evp[9]);
    g.phaseMs[3] = a + b;
    return 0;
  }
  const int64_t nIn0 = g.nUnsorted, nIn = nIn0 + nRecvTotal;
  if (nIn >= (int64_t)0x7fffffff) return fail("piclas_gpu_exchange_finish: more than 2^31-1 particles on one GPU");
  if (reserve_particles(nIn)) return 1;
  cudaEventRecord(g.evp[8], g.st);
  if (nRecvTotal > 0) {
    // immigrants behind the particles of the open step; emigrants and removed particles carry the "removed" key
    k_unpack_immigrants<<<(unsigned)((nRecvTotal + 255) / 256), 256, 0, g.st>>>(g.buf[g.cur], nIn0, nRecvTotal, g.commSize, g.ref ? 1 : 0, g.dCommRecv);
    k_keys_from_elem<<<(unsigned)((nRecvTotal + 255) / 256), 256, 0, g.st>>>(g.buf[g.cur].elem + nIn0, g.dElemRank, g.dKeys + nIn0, nRecvTotal,
                                                                             g.nElems, g.offsetElem, g.myRank, g.nRanks);
    g.lastLaunches += 2;
    CK(cudaGetLastError());
  }
  g.exchangePending = false;
  if (sort_and_permute(nIn)) return 1;   // the one sort of the step; also synchronises
  cudaEventRecord(g.evp[9], g.st);
  cudaEventSynchronize(g.evp[9]);
  {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, g.evp[6], g.evp[7]);   // emigrant extraction + pack
    cudaEventElapsedTime(&b, g.evp[8], g.evp[9]);   // unpack + sort + permute
    g.phaseMs[3] = a + b;
  }
  if (g.nTotalSorted != g.nPart) return fail("piclas_gpu_exchange_finish: received particles that belong to another rank");
  return 0;
}

int piclas_gpu_last_timing(double* ms_kernels, int32_t* nLaunches) {
  if (ms_kernels) *ms_kernels = g.lastMs;
  if (nLaunches) *nLaunches = g.lastLaunches;
  return 0;
}

int piclas_gpu_phase_timing(double* ms4) {
  for (int i = 0; i < 4; ++i) ms4[i] = g.phaseMs[i];
  return 0;
}

}  // extern "C"
