// In-place ascending bitonic sorts of every room's keys (a power of two per room) in global memory, used by the per-room
// ordering passes (feature preparation: voxel keys, first-seen order, seed order; spatial index: Morton order).
// The compare-exchange network is the textbook one -- for k = 2, 4, .. P and j = k/2 .. 1 element i meets i | j, ascending where
// (i & k) == 0 -- but every pass with j < CH runs on a CH-element chunk staged in shared memory: a pass over global memory is
// one L2 round trip per element, and a room of 2^15 keys would pay 120 of them.  The first log2(CH) stages are one load /
// store per chunk, and every later stage is its j >= CH passes over global memory plus ONE staged sweep for the rest: 2^19
// keys take 36 trips through the array instead of 190, 2^15 keys 10 instead of 120.  The result is the sorted array either way
// (a sorting network on totally ordered keys; pairs are ordered lexicographically by (key, value)).
#pragma once
#include <algorithm>

#include "lrg_common.cuh"

// One launch per pass, the stream orders them, grid = (CTAs per room, rooms): every pass is independent work per
// compare-exchange (far passes) or per chunk (staged sweeps), so a large room spreads over the machine -- one CTA per room (round
// 1) is bound by what one SM can move through L2: 12 ms per sort of the 2^19 keys of a 300 k-point outdoor scene.  A room takes
// part in a stage only if the stage is within its own length.
namespace lrg {

struct RoomSort {
  unsigned long long* keys;     // the rooms' buffers, room r at keys + off[r]
  int* vals;                    // pairs only: room r at vals + off[r] * vals_mul
  int vals_mul;
  const long long* off;         // (R+1) buffer offsets; a buffer holds a power of two of elements
  const long long* cnt_off;     // optional (R+1): only the first pow2ceil(max(2, cnt_off[r+1] - cnt_off[r])) elements are sorted
};

__device__ __forceinline__ int room_sort_len(const RoomSort& s, int r) {
  const int P = (int)(s.off[r + 1] - s.off[r]);
  if (s.cnt_off == nullptr) return P;
  const int n = (int)(s.cnt_off[r + 1] - s.cnt_off[r]);
  int Pe = 2;
  while (Pe < n) Pe <<= 1;
  return min(Pe, P);
}

// stages k = 2 .. min(P, CH) of chunk blockIdx.x (k_sweep == 0), or the passes j = CH/2 .. 1 of stage k_sweep
template <bool PAIRS, int NT, int CH>
__global__ void __launch_bounds__(NT) room_sort_chunk_kernel(const RoomSort s, int k_sweep) {
  __shared__ unsigned long long sk[CH];
  __shared__ int sv[PAIRS ? CH : 1];
  const int r = blockIdx.y, tid = threadIdx.x;
  const int P = room_sort_len(s, r);
  const int base = blockIdx.x * CH;
  if (P < 2 || base >= P || (k_sweep != 0 && k_sweep > P)) return;
  unsigned long long* keys = s.keys + s.off[r];
  int* vals = PAIRS ? s.vals + s.off[r] * s.vals_mul : nullptr;
  const int n = min(CH, P - base);
  for (int t = tid; t < n; t += NT) { sk[t] = keys[base + t]; if (PAIRS) sv[t] = vals[base + t]; }
  __syncthreads();
  auto pass = [&](int k, int j) {
    for (int t = tid; t < (n >> 1); t += NT) {
      const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
      const unsigned long long a = sk[i], b = sk[i | j];
      bool gt = a > b;
      if (PAIRS) gt = gt || (a == b && sv[i] > sv[i | j]);
      if (gt == (((base + i) & k) == 0)) {
        sk[i] = b; sk[i | j] = a;
        if (PAIRS) { const int va = sv[i]; sv[i] = sv[i | j]; sv[i | j] = va; }
      }
    }
    __syncthreads();
  };
  if (k_sweep == 0) {
    for (int k = 2; k <= min(P, CH); k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) pass(k, j);
  } else {
    for (int j = CH >> 1; j > 0; j >>= 1) pass(k_sweep, j);
  }
  for (int t = tid; t < n; t += NT) { keys[base + t] = sk[t]; if (PAIRS) vals[base + t] = sv[t]; }
}

// pass (k, j) with j >= CH: partners in different chunks, over global memory
template <bool PAIRS>
__global__ void __launch_bounds__(256) room_sort_far_kernel(const RoomSort s, int k, int j) {
  const int r = blockIdx.y;
  const int P = room_sort_len(s, r);
  if (k > P) return;
  unsigned long long* keys = s.keys + s.off[r];
  int* vals = PAIRS ? s.vals + s.off[r] * s.vals_mul : nullptr;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (P >> 1); t += gridDim.x * blockDim.x) {
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const unsigned long long a = keys[i], b = keys[i | j];
    bool gt = a > b;
    int va = 0, vb = 0;
    if (PAIRS) { va = vals[i]; vb = vals[i | j]; gt = gt || (a == b && va > vb); }
    if (gt == ((i & k) == 0)) {
      keys[i] = b; keys[i | j] = a;
      if (PAIRS) { vals[i] = vb; vals[i | j] = va; }
    }
  }
}

// Sort every room's buffer (max_len = the longest sorted length over the rooms, a power of two).  Returns the launches made.
template <bool PAIRS>
inline int launch_room_sort(const RoomSort& s, int n_rooms, long long max_len, cudaStream_t stream) {
  constexpr int CH = PAIRS ? 2048 : 4096;
  if (n_rooms <= 0 || max_len < 2) return 0;
  const int chunks = (int)std::max<long long>(1, max_len / CH);
  int launches = 1;
  room_sort_chunk_kernel<PAIRS, 1024, CH><<<dim3(chunks, n_rooms), 1024, 0, stream>>>(s, 0);
  for (long long k = 2ll * CH; k <= max_len; k <<= 1) {
    const int far_ctas = (int)std::min<long long>(592, std::max<long long>(1, (max_len / 2) / (256 * 8)));
    for (long long j = k >> 1; j >= CH; j >>= 1) { room_sort_far_kernel<PAIRS><<<dim3(far_ctas, n_rooms), 256, 0, stream>>>(s, (int)k, (int)j); ++launches; }
    room_sort_chunk_kernel<PAIRS, 1024, CH><<<dim3(chunks, n_rooms), 1024, 0, stream>>>(s, (int)k);
    ++launches;
  }
  return launches;
}

}  // namespace lrg
