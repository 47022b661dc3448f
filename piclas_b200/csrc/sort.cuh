// sort.cuh — hand-written stable LSD radix sort (8- or 10-bit digits) of particles by element key (sm_100a).
//
// The particle step keeps the particle SoA sorted by (local) element so that one CTA can stage its element's
// field tile and geometry in shared memory and so that deposition is a deterministic segmented sum.  The sort
// is stable, hence bit-reproducible from run to run: equal keys keep their previous relative order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct SortWorkspace {
  uint32_t *keysA = nullptr, *keysB = nullptr;  // ping-pong keys
  uint32_t *permA = nullptr, *permB = nullptr;  // ping-pong source indices
  uint32_t *blockHist = nullptr;                // [nBlocks][RADIX] block-major
  uint32_t *groupSum = nullptr;                 // [nGroups][RADIX] column-scan partials
  size_t capacity = 0, histCapacity = 0;
};

// allocate for up to n keys
cudaError_t sort_workspace_reserve(SortWorkspace& ws, size_t n);
void sort_workspace_free(SortWorkspace& ws);

// Sorts keys (n entries, values < 2^bits except KEY_DEAD handled as all-ones within `bits`) stably.
// On return *sortedKeys / *perm point into the workspace: perm[i] = source index of the i-th smallest key.
// nLaunches is incremented by the number of kernels launched.
cudaError_t radix_sort_by_key(SortWorkspace& ws, const uint32_t* keys, size_t n, int bits, cudaStream_t st,
                              uint32_t** sortedKeys, uint32_t** perm, int* nLaunches);

// stable sort by (keysHi, keysLo) (keysLo < 2^bitsLo, keysHi < 2^bitsHi): *sortedHi = keysHi in sorted order, *perm as above
cudaError_t radix_sort_two_keys(SortWorkspace& ws, const uint32_t* keysLo, int bitsLo, const uint32_t* keysHi, int bitsHi, size_t n,
                                cudaStream_t st, uint32_t** sortedHi, uint32_t** perm, int* nLaunches);

// out[i] = in[perm[i]]
cudaError_t gather_f64(const double* in, double* out, const uint32_t* perm, size_t n, cudaStream_t st);
cudaError_t gather_i32(const int32_t* in, int32_t* out, const uint32_t* perm, size_t n, cudaStream_t st);
cudaError_t gather_u8(const uint8_t* in, uint8_t* out, const uint32_t* perm, size_t n, cudaStream_t st);
cudaError_t gather_i64(const int64_t* in, int64_t* out, const uint32_t* perm, size_t n, cudaStream_t st);

// out[i] = in[perm[i]] for every particle array in one kernel (id/idOut may be null)
cudaError_t gather_particles(const double* const x[3], const double* const v[3], const int32_t* elem, const uint8_t* meta,
                             const int64_t* id, double* const y[3], double* const w[3], int32_t* elemOut, uint8_t* metaOut,
                             int64_t* idOut, const uint32_t* perm, size_t n, cudaStream_t st);

// off[k] = first index i with sortedKeys[i] >= k, k = 0..nKeys (nKeys+1 entries)
cudaError_t segment_offsets(const uint32_t* sortedKeys, size_t n, uint32_t nKeys, int64_t* off, cudaStream_t st);
