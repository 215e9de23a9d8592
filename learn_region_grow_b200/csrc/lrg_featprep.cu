// Feature preparation of /root/reference/test_region_grow.py:119-173 on the device: raw room points
// (x y z r g b ...) -> equalised points (first point of every voxel, in first-seen order, :125-136), room-normalised
// coordinates (:139), normal + curvature from the covariance of the raw points in the 27 surrounding voxels (:141-164)
// and the seed order argsort(curvatures) (:183).  One CTA per room for the ordering passes (bitonic sorts in global
// memory: a room is 10^4..10^5 keys and rooms run side by side), one thread per voxel / per point for the arithmetic.
//
// What is exact and what is not (tests/test_featprep_gpu.py): the equalisation maps, xyz, room coordinates and rgb are
// bit-identical to the reference; the covariance sums are accumulated like the reference (float32 products summed in
// float64, per voxel in insertion order, voxels in offset order); the 3x3 decomposition is a double-precision Jacobi
// eigen-solve instead of LAPACK's SVD, so normals / curvatures agree to rounding noise (except where two singular values
// coincide and the direction is undefined in the reference too).
#include <limits.h>

#include "lrg_featprep.cuh"
#include <algorithm>

#include "lrg_sort.cuh"
#include "lrg_step_body.cuh"

namespace lrg {

constexpr int kFpThreads = 1024;
constexpr unsigned long long kIdxMask = 0xFFFFFull;     // low 20 bits of a sort key: a room-local index

// (the per-room sorts: lrg_sort.cuh -- bitonic networks whose short-distance passes run on chunks staged in shared memory)
constexpr int kSortChunk = 4096, kSortChunkPairs = 2048;

// -------------------------------------------------------------------------------------------- phase 1: equalisation
__global__ void __launch_bounds__(kFpThreads) fp_keys_kernel(const __grid_constant__ FeatPrepArgs a) {
  __shared__ int s_mn[3], s_mx[3];
  const int room = blockIdx.x, tid = threadIdx.x;
  const long long base = a.raw_off[room];
  const int N = (int)(a.raw_off[room + 1] - base);
  const int P = (int)(a.sort_off[room + 1] - a.sort_off[room]);
  unsigned long long* keys = a.keys + a.sort_off[room];
  if (tid < 3) { s_mn[tid] = INT_MAX; s_mx[tid] = INT_MIN; }
  __syncthreads();
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = tid; i < N; i += kFpThreads)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v = voxel_of(a.raw[(base + i) * a.C + c], a.res);          // :126
      mn[c] = min(mn[c], v); mx[c] = max(mx[c], v);
    }
#pragma unroll
  for (int c = 0; c < 3; ++c) { atomicMin(&s_mn[c], mn[c]); atomicMax(&s_mx[c], mx[c]); }
  __syncthreads();
  const int o0 = s_mn[0], o1 = s_mn[1], o2 = s_mn[2];
  if (tid == 0) {
    a.raw_vmin[room] = make_int4(N ? o0 : 0, N ? o1 : 0, N ? o2 : 0, 0);
    if (N > 0 && (s_mx[0] - o0 > 1022 || s_mx[1] - o1 > 1022 || s_mx[2] - o2 > 1022)) atomicCAS(a.err, 0, room + 1);
  }
  for (int i = tid; i < P; i += kFpThreads) {
    unsigned long long key = ~0ull;
    if (i < N) {
      const float* p = a.raw + (base + i) * a.C;
      const unsigned x = (unsigned)(voxel_of(p[0], a.res) - o0) & 1023u, y = (unsigned)(voxel_of(p[1], a.res) - o1) & 1023u,
                     z = (unsigned)(voxel_of(p[2], a.res) - o2) & 1023u;
      key = ((unsigned long long)(x | (y << 10) | (z << 20)) << 20) | (unsigned long long)i;
    }
    keys[i] = key;
  }
  // (sorted by voxel, then by insertion index, by launch_room_sort behind this kernel)
}

__global__ void __launch_bounds__(kFpThreads) fp_unique_kernel(const __grid_constant__ FeatPrepArgs a) {
  __shared__ int s_scan[33];
  const int room = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = a.raw_off[room];
  const int N = (int)(a.raw_off[room + 1] - base);
  const int P = (int)(a.sort_off[room + 1] - a.sort_off[room]);
  const unsigned long long* keys = a.keys + a.sort_off[room];
  unsigned long long* keys2 = a.keys2 + a.sort_off[room];
  // rank of every sorted position = number of run starts at or before it - 1; runs = voxels
  int running = 0;
  for (int start = 0; start < N; start += kFpThreads * 4) {
    const int s0 = start + tid * 4;
    unsigned flags = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int s = s0 + q;
      if (s < N && (s == 0 || (keys[s] >> 20) != (keys[s - 1] >> 20))) flags |= 1u << q;
    }
    const int cnt = __popc(flags);
    const int incl = warp_incl_scan(cnt, lane);
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    scan_warp_totals<kFpThreads>(s_scan, warp, lane);
    __syncthreads();
    int u = running + s_scan[warp] + incl - cnt - 1;        // rank of the last run that started before s0
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int s = s0 + q;
      if (s < N) {
        if ((flags >> q) & 1u) {
          ++u;
          const unsigned long long key = keys[s];
          a.uniq_vox[base + u] = (unsigned)(key >> 20);
          a.uniq_start[base + u] = s;
          keys2[u] = ((key & kIdxMask) << 20) | (unsigned long long)u;   // first-seen index of the voxel, voxel rank
        }
        a.raw_rank[base + (int)(keys[s] & kIdxMask)] = u;
      }
    }
    running += s_scan[32];
    __syncthreads();
  }
  const int n_eq = running;
  if (tid == 0) a.n_eq[room] = n_eq;
  for (int i = n_eq + tid; i < P; i += kFpThreads) keys2[i] = ~0ull;
  // (keys2 is sorted -- voxels in first-seen order, :127-129 -- by launch_room_sort; fp_unique_post_kernel follows)
}

__global__ void fp_unique_post_kernel(const __grid_constant__ FeatPrepArgs a) {
  const int room = blockIdx.y;
  const long long base = a.raw_off[room];
  const int n_eq = a.n_eq[room];
  const unsigned long long* keys2 = a.keys2 + a.sort_off[room];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_eq; j += gridDim.x * blockDim.x) a.eq_of_uniq[base + (int)(keys2[j] & kIdxMask)] = j;
}

// per voxel: n, sum p, sum of the float32 outer products in float64 (:151-155)
__global__ void fp_voxel_sums_kernel(const __grid_constant__ FeatPrepArgs a) {
  const int room = blockIdx.y;
  const long long base = a.raw_off[room];
  const int N = (int)(a.raw_off[room + 1] - base);
  const int n_eq = a.n_eq[room];
  const unsigned long long* keys = a.keys + a.sort_off[room];
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_eq; u += gridDim.x * blockDim.x) {
    const int s0 = a.uniq_start[base + u], s1 = (u + 1 < n_eq) ? a.uniq_start[base + u + 1] : N;
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = s0; s < s1; ++s) {
      const float* p = a.raw + (base + (long long)(keys[s] & kIdxMask)) * a.C;
      const float x = p[0], y = p[1], z = p[2];
      acc[0] += 1.0;
      acc[1] += (double)x; acc[2] += (double)y; acc[3] += (double)z;
      acc[4] += (double)__fmul_rn(x, x); acc[5] += (double)__fmul_rn(x, y); acc[6] += (double)__fmul_rn(x, z);
      acc[7] += (double)__fmul_rn(y, y); acc[8] += (double)__fmul_rn(y, z); acc[9] += (double)__fmul_rn(z, z);
    }
    double* out = a.sums + (base + u) * 10;
#pragma unroll
    for (int c = 0; c < 10; ++c) out[c] = acc[c];
  }
}

// -------------------------------------------------------------------------------------------- phase 2: features
// cyclic Jacobi eigen-decomposition of a symmetric 3x3 matrix (double); V columns are the eigenvectors
__device__ void eig3(double A[3][3], double w[3], double V[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const double apq = A[p][q];
      if (apq == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const int r = 3 - p - q;
      const double app = A[p][p], aqq = A[q][q], arp = A[r][p], arq = A[r][q];
      A[p][p] = app - t * apq;
      A[q][q] = aqq + t * apq;
      A[p][q] = A[q][p] = 0.0;
      A[r][p] = A[p][r] = c * arp - s * arq;
      A[r][q] = A[q][r] = s * arp + c * arq;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double vip = V[i][p], viq = V[i][q];
        V[i][p] = c * vip - s * viq;
        V[i][q] = s * vip + c * viq;
      }
    }
  }
  w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

__device__ __forceinline__ unsigned long long dmax_bits(double v) { return (unsigned long long)__double_as_longlong(v); }

// Extent of every room's equalised points (:139) and reset of the per-room curvature maximum.
__global__ void __launch_bounds__(kFpThreads) fp_extent_kernel(const __grid_constant__ FeatPrepArgs a) {
  __shared__ float s_wlo[32][3], s_whi[32][3];
  const int room = blockIdx.x, tid = threadIdx.x;
  const long long rbase = a.raw_off[room], ebase = a.eq_off[room];
  const int n_eq = (int)(a.eq_off[room + 1] - ebase);
  const unsigned long long* keys2 = a.keys2 + a.sort_off[room];
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int j = tid; j < n_eq; j += kFpThreads) {
    const float* p = a.raw + (rbase + (long long)(keys2[j] >> 20)) * a.C;
#pragma unroll
    for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], p[c]); hi[c] = fmaxf(hi[c], p[c]); }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], d));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], d));
    }
  if ((tid & 31) == 0)
#pragma unroll
    for (int c = 0; c < 3; ++c) { s_wlo[tid >> 5][c] = lo[c]; s_whi[tid >> 5][c] = hi[c]; }
  __syncthreads();
  if (tid < 3) {
    float l = INFINITY, h = -INFINITY;
    for (int w = 0; w < kFpThreads / 32; ++w) { l = fminf(l, s_wlo[w][tid]); h = fmaxf(h, s_whi[w][tid]); }
    a.extent[room * 6 + tid] = l;
    a.extent[room * 6 + 3 + tid] = h;
  }
  if (tid == 0) { a.cmax[room] = 0ull; a.has_nan[room] = 0; }
}

// Per equalised point: covariance of the raw points in the 27 surrounding voxels, eigen-solve, feature row.  grid = (chunks,
// rooms): the points of a room are independent of each other up to the curvature maximum (one atomicMax per CTA).
__global__ void __launch_bounds__(kFpThreads) fp_features_kernel(const __grid_constant__ FeatPrepArgs a) {
  __shared__ unsigned long long s_cmax;
  __shared__ int s_nan;
  const int room = blockIdx.y, tid = threadIdx.x;
  const long long rbase = a.raw_off[room], ebase = a.eq_off[room];
  const int n_eq = (int)(a.eq_off[room + 1] - ebase);
  const unsigned long long* keys2 = a.keys2 + a.sort_off[room];
  const unsigned* uv = a.uniq_vox + rbase;
  if (tid == 0) { s_cmax = 0ull; s_nan = 0; }
  __syncthreads();
  const float lo0 = a.extent[room * 6 + 0], lo1 = a.extent[room * 6 + 1], lo2 = a.extent[room * 6 + 2];
  const float ex0 = __fsub_rn(a.extent[room * 6 + 3], lo0), ex1 = __fsub_rn(a.extent[room * 6 + 4], lo1), ex2 = __fsub_rn(a.extent[room * 6 + 5], lo2);

  for (int j = blockIdx.x * kFpThreads + tid; j < n_eq; j += gridDim.x * kFpThreads) {
    const int u = (int)(keys2[j] & kIdxMask);
    const float* p = a.raw + (rbase + (long long)(keys2[j] >> 20)) * a.C;
    // neighbourhood sums over the 27 surrounding voxels in itertools.product order (:146-155)
    const unsigned key = uv[u];
    const int vx = (int)(key & 1023u), vy = (int)((key >> 10) & 1023u), vz = (int)((key >> 20) & 1023u);
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          const int x = vx + dx, y = vy + dy, z = vz + dz;
          if ((unsigned)x > 1023u || (unsigned)y > 1023u || (unsigned)z > 1023u) continue;
          const unsigned q = (unsigned)x | ((unsigned)y << 10) | ((unsigned)z << 20);
          int lo = 0, hi = n_eq;                      // lower bound in the sorted voxel keys
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (uv[mid] < q) lo = mid + 1; else hi = mid; }
          if (lo < n_eq && uv[lo] == q) {
            const double* sv = a.sums + (rbase + lo) * 10;
#pragma unroll
            for (int c = 0; c < 10; ++c) acc[c] += sv[c];
          }
        }
    const double n = acc[0];
    double A[3][3], w[3], V[3][3];
    const double n2 = n * n;
    A[0][0] = acc[4] / n - (acc[1] * acc[1]) / n2; A[0][1] = acc[5] / n - (acc[1] * acc[2]) / n2; A[0][2] = acc[6] / n - (acc[1] * acc[3]) / n2;
    A[1][1] = acc[7] / n - (acc[2] * acc[2]) / n2; A[1][2] = acc[8] / n - (acc[2] * acc[3]) / n2; A[2][2] = acc[9] / n - (acc[3] * acc[3]) / n2;
    A[1][0] = A[0][1]; A[2][0] = A[0][2]; A[2][1] = A[1][2];
    eig3(A, w, V);
    // singular values = |eigenvalues|; V[2] of numpy.linalg.svd = direction of the smallest one (:157-161)
    int k = 0;
    if (fabs(w[1]) < fabs(w[k])) k = 1;
    if (fabs(w[2]) < fabs(w[k])) k = 2;
    const double ssum = fabs(w[0]) + fabs(w[1]) + fabs(w[2]);
    const double curv = fabs(fabs(w[k]) / ssum);          // 0/0 = NaN like the reference
    a.curv[ebase + j] = curv;
    if (curv != curv) s_nan = 1; else atomicMax(&s_cmax, dmax_bits(curv));
    float* f = a.feat + (ebase + j) * a.F;
    const float row[12] = {p[0], p[1], p[2],
                           __fdiv_rn(__fsub_rn(p[0], lo0), ex0), __fdiv_rn(__fsub_rn(p[1], lo1), ex1), __fdiv_rn(__fsub_rn(p[2], lo2), ex2),
                           p[3], p[4], p[5],
                           (float)fabs(V[0][k]), (float)fabs(V[1][k]), (float)fabs(V[2][k])};
#pragma unroll
    for (int c = 0; c < 12; ++c)
      if (c < a.F) f[c] = row[c];
    a.equalized_idx[ebase + j] = (int)(keys2[j] >> 20);
  }
  __syncthreads();
  if (tid == 0) {
    if (s_nan) a.has_nan[room] = 1;
    if (s_cmax) atomicMax(a.cmax + room, s_cmax);
  }
}

// Per room: curvature / max (:162-163), seed order = argsort(curvatures) (:183), unequalized_idx (:130).
__global__ void __launch_bounds__(kFpThreads) fp_order_kernel(const __grid_constant__ FeatPrepArgs a) {
  const int room = blockIdx.x, tid = threadIdx.x;
  const long long rbase = a.raw_off[room], ebase = a.eq_off[room];
  const int n_eq = (int)(a.eq_off[room + 1] - ebase);
  // a NaN anywhere makes the maximum -- and everything -- NaN in numpy
  const double cmax = a.has_nan[room] ? __longlong_as_double(0x7FF8000000000000ll) : __longlong_as_double((long long)a.cmax[room]);
  unsigned long long* okeys = a.keys + a.sort_off[room];      // the voxel sort keys are dead by now: reuse for the seed order
  const int P = (int)(a.sort_off[room + 1] - a.sort_off[room]);
  int Pe = 2;
  while (Pe < n_eq) Pe <<= 1;
  if (Pe > P) Pe = P;
  for (int j = tid; j < Pe; j += kFpThreads) {
    if (j < n_eq) {
      const double c = a.curv[ebase + j] / cmax;
      a.curv[ebase + j] = c;
      if (a.F > 12) a.feat[(ebase + j) * a.F + 12] = (float)c;
    }
  }
  __syncthreads();
  // seed order = argsort(curvatures) (:183): sort (curvature bits, index) pairs; non-negative doubles order like their bit
  // patterns, NaN sorts last like numpy; equal curvatures are ordered by index (numpy's unstable sort leaves that open).
  int* oidx = reinterpret_cast<int*>(a.keys2 + a.sort_off[room]);   // the first-seen sort keys are dead by now
  __syncthreads();
  for (int j = tid; j < Pe; j += kFpThreads) {
    okeys[j] = j < n_eq ? dmax_bits(a.curv[ebase + j]) : ~0ull;
    oidx[j] = j < n_eq ? j : INT_MAX;
  }
  // (the pairs are sorted by launch_room_sort; fp_order_post_kernel follows)
}

__global__ void fp_order_post_kernel(const __grid_constant__ FeatPrepArgs a) {
  const int room = blockIdx.y;
  const long long rbase = a.raw_off[room], ebase = a.eq_off[room];
  const int n_eq = (int)(a.eq_off[room + 1] - ebase);
  const int* oidx = reinterpret_cast<const int*>(a.keys2 + a.sort_off[room]);
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int j = t0; j < n_eq; j += nt) a.order[ebase + j] = oidx[j];
  for (int i = t0; i < (int)(a.raw_off[room + 1] - rbase); i += nt)
    a.unequalized_idx[rbase + i] = a.eq_of_uniq[rbase + a.raw_rank[rbase + i]];          // :130
}

// cluster_label[unequalized_idx] (:366): labels of the equalised points mapped back to every raw point
__global__ void fp_labels_raw_kernel(int n_rooms, const long long* __restrict__ raw_off, const long long* __restrict__ eq_off,
                                     const int* __restrict__ unequalized_idx, const int* __restrict__ label, int* __restrict__ out) {
  const int room = blockIdx.y;
  const long long rbase = raw_off[room], ebase = eq_off[room];
  const int N = (int)(raw_off[room + 1] - rbase);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    out[rbase + i] = label[ebase + unequalized_idx[rbase + i]];
}

int launch_labels_raw(int n_rooms, const long long* raw_off, const long long* eq_off, const int* unequalized_idx, const int* label,
                      int* out, cudaStream_t stream) {
  if (n_rooms <= 0) return LRG_OK;
  fp_labels_raw_kernel<<<dim3(32, n_rooms), 256, 0, stream>>>(n_rooms, raw_off, eq_off, unequalized_idx, label, out);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int launch_featprep_phase1(const FeatPrepArgs& a, cudaStream_t stream, int* n_launches) {
  if (a.n_rooms <= 0) return LRG_OK;
  int nl = 5;
  // (grid-stride kernels: enough CTAs per room to fill the machine when the upload holds few, large rooms)
  const int per_room = std::max(8, std::min(148, (4 * 148 + a.n_rooms - 1) / std::max(a.n_rooms, 1)));
  fp_keys_kernel<<<a.n_rooms, kFpThreads, 0, stream>>>(a);
  RoomSort s1{a.keys, nullptr, 0, a.sort_off, nullptr};
  nl += launch_room_sort<false>(s1, a.n_rooms, a.max_sort, stream);
  fp_unique_kernel<<<a.n_rooms, kFpThreads, 0, stream>>>(a);
  RoomSort s2{a.keys2, nullptr, 0, a.sort_off, nullptr};
  nl += launch_room_sort<false>(s2, a.n_rooms, a.max_sort, stream);
  fp_unique_post_kernel<<<dim3(per_room, a.n_rooms), 256, 0, stream>>>(a);
  fp_voxel_sums_kernel<<<dim3(2 * per_room, a.n_rooms), 256, 0, stream>>>(a);
  if (n_launches) *n_launches += nl - 1;               // (keys, unique, unique_post, voxel_sums + the sorts)
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int launch_featprep_phase2(const FeatPrepArgs& a, cudaStream_t stream, int* n_launches) {
  if (a.n_rooms <= 0) return LRG_OK;
  int nl = 4;                                          // extent, features, order, order_post
  fp_extent_kernel<<<a.n_rooms, kFpThreads, 0, stream>>>(a);
  const int per_room = std::max(8, std::min(148, (4 * 148 + a.n_rooms - 1) / std::max(a.n_rooms, 1)));
  fp_features_kernel<<<dim3(per_room, a.n_rooms), kFpThreads, 0, stream>>>(a);
  fp_order_kernel<<<a.n_rooms, kFpThreads, 0, stream>>>(a);
  RoomSort s3{a.keys, reinterpret_cast<int*>(a.keys2), 2, a.sort_off, a.eq_off};
  nl += launch_room_sort<true>(s3, a.n_rooms, a.max_order_sort, stream);
  fp_order_post_kernel<<<dim3(per_room, a.n_rooms), 256, 0, stream>>>(a);
  if (n_launches) *n_launches += nl;
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg
