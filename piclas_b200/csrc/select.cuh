// select.cuh — visiting a candidate list in the order of the reference's stable InsertionSort (utils.f90:52-101) without
// storing the list: each call scans all n entries and returns the one that follows the previously returned entry in
// ascending order of (distance, list position).  O(n^2) distance evaluations per list, no per-thread arrays; used by the
// RefMapping relocation when a FIBGM cell holds more elements than the stored-list path keeps (REF_MAX_BGM).
#pragma once

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

struct SortedVisit {
  bool first = true;
  double prevD = 0.;
  int prevI = -1;
  int count = 0;   // entries returned so far: never more than n, whatever the keys are (NaN keys cannot make the visit endless)
};

// dist(i) -> key of entry i, or `skip` for entries the caller filters out (they are never returned).
// Returns the list position of the next entry, -1 when all entries have been visited.
template <class DistFn>
__host__ __device__ inline int next_in_sorted_order(int n, const DistFn& dist, double skip, SortedVisit& s) {
  if (s.count >= n) return -1;
  int bestI = -1;
  double bestD = 0.;
  for (int i = 0; i < n; ++i) {
    const double d = dist(i);
    if (d == skip) continue;
    if (!s.first && (d < s.prevD || (d == s.prevD && i <= s.prevI))) continue;   // visited before
    if (bestI < 0 || d < bestD) { bestI = i; bestD = d; }                         // strict <: lowest position among equals
  }
  if (bestI >= 0) { s.first = false; s.prevD = bestD; s.prevI = bestI; ++s.count; }
  return bestI;
}
