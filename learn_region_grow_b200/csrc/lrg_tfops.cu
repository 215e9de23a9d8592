// sm_100a replacements for the PointNet++ custom ops of /root/reference/tf_ops (sampling, grouping, 3d_interpolation).
// Index-exact with the reference kernels: same distance expressions (so nvcc contracts them identically), same scan
// order, same tie rules.  All pointers are device pointers.
#include <cooperative_groups.h>
#include <float.h>

#include "lrg_common.cuh"

namespace lrg {

// ------------------------------------------------------------------------------------- farthest point sampling
// Reference: farthestpointsamplingKernel (tf_sampling_g.cu:105-170), <<<32,512>>>, min-distances in a global
// (32,n) workspace, 512-wide shared-memory tree argmax with 9 barriers per round.
// Here: one CTA per cloud, points and running min-distances in registers (PPT per thread), argmax by redux.sync
// reductions (fps_block_argmax) + one barrier per round.  The reference's tie rule -- smallest (k mod 512), then smallest k --
// is encoded in the low word of the key.
constexpr int kFpsThreads = 512;

__device__ __forceinline__ unsigned long long fps_key(float d, int k) {
  const unsigned tie = ((unsigned)(k & 511) << 22) | (unsigned)(k >> 9);
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0x7FFFFFFFu - tie);
}

// CTA-wide argmax of the 64-bit keys (distance bits ‖ tie word) as two 32-bit maxima per stage -- redux.sync on the
// distance bits (non-negative floats order like unsigned integers), then on the tie words of the lanes that hold the maximum
// -- instead of a 64-bit shuffle butterfly: 2 + 2 warp reductions and one barrier per round.  Returns the winning index.
template <int NT>
__device__ __forceinline__ int fps_block_argmax(float best, int besti, uint2 (*sred)[NT / 32], int j, int lane, int warp) {
  // a thread without points keeps (-1, 0) like the reference; it must lose against every real candidate: key 0
  const unsigned long long key = best < 0.f ? 0ull : fps_key(best, besti);
  const unsigned d = (unsigned)(key >> 32), t = (unsigned)key;
  const unsigned dmax = __reduce_max_sync(0xffffffffu, d);
  const unsigned tmax = __reduce_max_sync(0xffffffffu, d == dmax ? t : 0u);
  if (lane == 0) sred[j & 1][warp] = make_uint2(dmax, tmax);
  __syncthreads();
  const uint2 w = lane < NT / 32 ? sred[j & 1][lane] : make_uint2(0u, 0u);
  const unsigned d2 = __reduce_max_sync(0xffffffffu, w.x);
  const unsigned t2 = __reduce_max_sync(0xffffffffu, w.x == d2 ? w.y : 0u);
  if (d2 == 0u && t2 == 0u) return 0;
  const unsigned tie = 0x7FFFFFFFu - t2;
  return (int)(((tie & 0x3FFFFFu) << 9) | (tie >> 22));
}

// NT threads x PPT points per thread >= n.
template <int NT, int PPT>      // NT in {128, 256, 512}
__global__ void __launch_bounds__(NT) lrg_fps_kernel(int n, int m, const float* __restrict__ dataset, int* __restrict__ idxs) {
  __shared__ uint2 sred[2][NT / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pts = dataset + (size_t)b * n * 3;
  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int k = tid + q * NT;
    px[q] = py[q] = pz[q] = 0.f;
    if (k < n) { px[q] = pts[k * 3 + 0]; py[q] = pts[k * 3 + 1]; pz[q] = pts[k * 3 + 2]; }
    td[q] = 1e38f;
  }
  int old = 0;
  if (tid == 0) idxs[(size_t)b * m] = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
    float best = -1.f;
    int besti = 0;
    // a thread's own points are visited in the order of the tie rule -- (k mod 512, k) ascending -- so that the first of
    // equal distances is the one the reference keeps: with NT < 512 that is q = r, r + 512/NT, ... for r = 0 .. 512/NT - 1
    constexpr int R = 512 / NT;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = r; q < PPT; q += R) {
        const int k = tid + q * NT;
        if (k < n) {
          const float x2 = px[q], y2 = py[q], z2 = pz[q];
          const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
          const float d2 = min(d, td[q]);
          td[q] = d2;
          if (d2 > best) { best = d2; besti = k; }
        }
      }
    old = fps_block_argmax<NT>(best, besti, sred, j, lane, warp);
    if (tid == 0) idxs[(size_t)b * m + j] = old;
  }
}

// Clouds of 8,193 .. 65,536 points: a thread-block CLUSTER of CL CTAs per cloud keeps the points and the running minimum
// distances in registers (16 per thread) -- CTA r, thread t holds the points t + 512 (r + CL q), so that within a thread the
// tie rule's order is still ascending q -- where the one-CTA kernel above runs out of registers and the reference's layout
// (min-distances in a global workspace, lrg_fps_big_kernel below) re-reads 16 n bytes per round.  Per round: the CTA's argmax
// as above, then every CTA writes its (distance, tie) word into the round's slot of EVERY CTA's shared memory (distributed
// shared memory), one cluster barrier, and everybody takes the maximum of the CL words -- no global memory in the loop but the
// three coordinates of the last sample.  The slots are double-buffered by round parity: a CTA can only be one barrier ahead.
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kFpsThreads) lrg_fps_cluster_kernel(int n, int m, const float* __restrict__ dataset,
                                                                                               int* __restrict__ idxs) {
  namespace cg = cooperative_groups;
  constexpr int NT = kFpsThreads, PPT = 16;
  __shared__ uint2 sred[2][NT / 32];
  __shared__ uint2 ckey[2][8];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pts = dataset + (size_t)b * n * 3;
  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int k = tid + NT * (rank + CL * q);
    px[q] = py[q] = pz[q] = 0.f;
    if (k < n) { px[q] = pts[k * 3 + 0]; py[q] = pts[k * 3 + 1]; pz[q] = pts[k * 3 + 2]; }
    td[q] = 1e38f;
  }
  int old = 0;
  if (rank == 0 && tid == 0) idxs[(size_t)b * m] = 0;
  cluster.sync();                                     // (every CTA of the cluster is running before anybody writes into it)
  for (int j = 1; j < m; ++j) {
    const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
    float best = -1.f;
    int besti = 0;
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const int k = tid + NT * (rank + CL * q);
      if (k < n) {
        const float x2 = px[q], y2 = py[q], z2 = pz[q];
        const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
        const float d2 = min(d, td[q]);
        td[q] = d2;
        if (d2 > best) { best = d2; besti = k; }
      }
    }
    const unsigned long long key = best < 0.f ? 0ull : fps_key(best, besti);
    const unsigned d = (unsigned)(key >> 32), t = (unsigned)key;
    const unsigned dmax = __reduce_max_sync(0xffffffffu, d);
    const unsigned tmax = __reduce_max_sync(0xffffffffu, d == dmax ? t : 0u);
    if (lane == 0) sred[j & 1][warp] = make_uint2(dmax, tmax);
    __syncthreads();
    if (warp == 0) {
      const uint2 w = lane < NT / 32 ? sred[j & 1][lane] : make_uint2(0u, 0u);
      const unsigned d2 = __reduce_max_sync(0xffffffffu, w.x);
      const unsigned t2 = __reduce_max_sync(0xffffffffu, w.x == d2 ? w.y : 0u);
      if (lane < CL) *cluster.map_shared_rank(&ckey[j & 1][rank], lane) = make_uint2(d2, t2);
    }
    cluster.sync();
    const uint2 w = lane < CL ? ckey[j & 1][lane] : make_uint2(0u, 0u);
    const unsigned d3 = __reduce_max_sync(0xffffffffu, w.x);
    const unsigned t3 = __reduce_max_sync(0xffffffffu, w.x == d3 ? w.y : 0u);
    if (d3 == 0u && t3 == 0u) old = 0;
    else {
      const unsigned tie = 0x7FFFFFFFu - t3;
      old = (int)(((tie & 0x3FFFFFu) << 9) | (tie >> 22));
    }
    if (rank == 0 && tid == 0) idxs[(size_t)b * m + j] = old;
  }
}

// Large clouds: min-distances in the caller's workspace (same layout as the reference: one row per CTA).
__global__ void __launch_bounds__(kFpsThreads) lrg_fps_big_kernel(int b_total, int n, int m, const float* __restrict__ dataset,
                                                                float* __restrict__ temp, int* __restrict__ idxs) {
  __shared__ uint2 sred[2][kFpsThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = blockIdx.x; b < b_total; b += gridDim.x) {
    const float* pts = dataset + (size_t)b * n * 3;
    float* td = temp + (size_t)blockIdx.x * n;
    for (int k = tid; k < n; k += kFpsThreads) td[k] = 1e38f;
    int old = 0;
    if (tid == 0) idxs[(size_t)b * m] = 0;
    __syncthreads();
    for (int j = 1; j < m; ++j) {
      const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
        float best = -1.f;
      int besti = 0;
      for (int k = tid; k < n; k += kFpsThreads) {
        const float x2 = pts[k * 3 + 0], y2 = pts[k * 3 + 1], z2 = pts[k * 3 + 2];
        const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
        const float d2 = min(d, td[k]);
        td[k] = d2;
        if (d2 > best) { best = d2; besti = k; }
      }
      old = fps_block_argmax<kFpsThreads>(best, besti, sred, j, lane, warp);
      if (tid == 0) idxs[(size_t)b * m + j] = old;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------- prob_sample
// Reference: cumsumKernel + binarysearchKernel (tf_sampling_g.cu:7-104): an inclusive float32 cumulative sum of every row of
// probabilities, then for every query r in [0,1) the first category whose cumulative sum reaches r * total.  The indices
// depend on how the cumulative sums round, so the sums are formed in the reference's association:
//   * chunks of 8192 categories; a chunk = groups of four with the in-group prefixes v1, v1+v2, (v1+v2)+v3, (v3+v4)+(v1+v2)
//     (a trailing short group: plain running sums);
//   * the inclusive prefix P(x) of the first x group totals follows the reference's in-place sweep: with x = 2^b1 + 2^b2 + ...
//     (b1 > b2 > ...) it is ((S(0, 2^b1) + S(2^b1, 2^b1 + 2^b2)) + ...) where S(a, a + 2^b) is the balanced pairwise sum of
//     that aligned block -- so one up-sweep leaves every block sum a prefix needs in place (the blocks a prefix uses start
//     at even multiples of their size and are never overwritten by a higher level), and each thread assembles its prefix
//     from at most 11 of them instead of a second (down-)sweep with a barrier per level;
//   * element = (in-group prefix + P(groups before it)) + carry, the carry updated across chunks by the reference's
//     compensated sum (tf_sampling_g.cu:96-99).
// One CTA per row; the reference runs 32 CTAs x 512 threads over all rows.
constexpr int kProbThreads = 1024, kProbChunk = 8192;

__device__ __forceinline__ float prob_prefix(const float* T, int x) {      // inclusive prefix of the first x (>= 1) group totals
  float P = 0.f;
  int pos = 0;
  bool first = true;
  for (int b = 31 - __clz(x); b >= 0; --b)
    if ((x >> b) & 1) {
      const float S = T[pos + (1 << b) - 1];
      P = first ? S : __fadd_rn(S, P);
      first = false;
      pos += 1 << b;
    }
  return P;
}

__global__ void __launch_bounds__(kProbThreads) lrg_prob_cumsum_kernel(int n, const float* __restrict__ inp, float* __restrict__ out) {
  __shared__ float G[kProbChunk];            // in-group inclusive prefixes
  __shared__ float T[kProbChunk / 4];        // group totals -> pairwise block sums
  const int tid = threadIdx.x;
  const float* row = inp + (size_t)blockIdx.x * n;
  float* orow = out + (size_t)blockIdx.x * n;
  float carry = 0.f, comp = 0.f;
  for (int j = 0; j < n; j += kProbChunk) {
    const int len = min(n - j, kProbChunk), groups = (len + 3) >> 2;
    for (int g = tid; g < groups; g += kProbThreads) {
      const int k = g * 4;
      float a, b2, c, d;
      if (k + 3 < len) {
        const float v1 = row[j + k], v2 = row[j + k + 1], v3 = row[j + k + 2], v4 = row[j + k + 3];
        a = v1;
        b2 = __fadd_rn(v2, v1);
        c = __fadd_rn(v3, b2);
        d = __fadd_rn(__fadd_rn(v4, v3), b2);
      } else {
        float v = 0.f;
        float t[4];
        for (int q = 0; q < 4; ++q) {
          if (k + q < len) v = __fadd_rn(v, row[j + k + q]);
          t[q] = v;
        }
        a = t[0]; b2 = t[1]; c = t[2]; d = t[3];
      }
      G[k] = a; G[k + 1] = b2; G[k + 2] = c; G[k + 3] = d;
      T[g] = d;
    }
    for (int u = 0; (2 << u) <= groups; ++u) {       // up-sweep: T[(k+1) * 2^(u+1) - 1] = balanced sum of that aligned block
      __syncthreads();
      for (int k = tid; k < (groups >> (u + 1)); k += kProbThreads) {
        const int hi = (((k << 1) + 2) << u) - 1, lo = (((k << 1) + 1) << u) - 1;
        T[hi] = __fadd_rn(T[hi], T[lo]);
      }
    }
    __syncthreads();
    for (int g = tid; g < groups; g += kProbThreads) {
      const int k = g * 4;
      const float before = g > 0 ? prob_prefix(T, g) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k + q < len) {
          const float v = g > 0 ? __fadd_rn(G[k + q], before) : G[k + q];
          orow[j + k + q] = __fadd_rn(v, carry);
        }
    }
    // carry across chunks, compensated like the reference (every thread keeps its own copy)
    const float t = __fadd_rn(prob_prefix(T, groups), comp);
    const float r2 = __fadd_rn(carry, t);
    comp = __fsub_rn(t, __fsub_rn(r2, carry));
    carry = r2;
    __syncthreads();
  }
}

// result = n-1 walked down by descending powers of two while the cumulative sum k places below still reaches the query
// (tf_sampling_g.cu:91-103) -- the first category whose cumulative sum is >= r * total when the sums are monotone.
__global__ void lrg_prob_search_kernel(int b, int n, int m, const float* __restrict__ cum, const float* __restrict__ query, int* __restrict__ result) {
  int base = 1;
  while (base < n) base <<= 1;
  const long long total = (long long)b * m;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / m;
    const float* c = cum + i * n;
    const float q = __fmul_rn(query[t], c[n - 1]);
    int r = n - 1;
    for (int k = base; k >= 1; k >>= 1)
      if (r >= k && c[r - k] >= q) r -= k;
    result[t] = r;
  }
}

// ------------------------------------------------------------------------------------- gather / scatter-add
__global__ void lrg_gather_point_kernel(int b, int n, int m, const float* __restrict__ inp, const int* __restrict__ idx, float* __restrict__ out) {
  const long long total = (long long)b * m;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / m;
    const int a = idx[t];
    const float* src = inp + (i * n + a) * 3;
    out[t * 3 + 0] = src[0]; out[t * 3 + 1] = src[1]; out[t * 3 + 2] = src[2];
  }
}

__global__ void lrg_scatter_add_point_kernel(int b, int n, int m, const float* __restrict__ out_g, const int* __restrict__ idx, float* __restrict__ inp_g) {
  const long long total = (long long)b * m;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / m;
    const int a = idx[t];
    float* dst = inp_g + (i * n + a) * 3;
    atomicAdd(dst + 0, out_g[t * 3 + 0]); atomicAdd(dst + 1, out_g[t * 3 + 1]); atomicAdd(dst + 2, out_g[t * 3 + 2]);
  }
}

// ------------------------------------------------------------------------------------- ball query
// Reference: query_ball_point_gpu (tf_grouping_g.cu:3-36): one THREAD per query scanning all n points.
// Here: one WARP per query; 32 points per step, ballot-ordered append keeps "first nsample in index order".
__global__ void __launch_bounds__(256) lrg_query_ball_kernel(int b, int n, int m, float radius, int nsample, const float* __restrict__ xyz1,
                                                             const float* __restrict__ xyz2, int* __restrict__ idx, int* __restrict__ pts_cnt) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = wid; q < (long long)b * m; q += nwarps) {
    const long long bi = q / m;
    const float* p1 = xyz1 + bi * n * 3;
    const float x2 = xyz2[q * 3 + 0], y2 = xyz2[q * 3 + 1], z2 = xyz2[q * 3 + 2];
    int* out = idx + q * nsample;
    int cnt = 0, first = -1;
    for (int k0 = 0; k0 < n && cnt < nsample; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < n) {
        const float x1 = p1[k * 3 + 0], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
        const float d = max(sqrtf((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1)), 1e-20f);
        hit = d < radius;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        if (first < 0) first = k0 + __ffs(bal) - 1;
        const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < nsample) out[pos] = k;
        cnt = min(nsample, cnt + __popc(bal));
      }
    }
    if (first >= 0)
      for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;   // the reference pre-fills the row with the first hit
    if (lane == 0) pts_cnt[q] = cnt;
  }
}

// sample_and_group (train_pointnet.py:113-123) behind the sampling: gather_point of the sampled centres (:114), ball query,
// group_point of the coordinates with the translation normalisation (grouped_xyz -= new_xyz, :117), group_point of the
// features and the concat (:119-120) in ONE kernel -- a warp finds its query's nsample neighbours as above and writes their rows at once, instead of a (b,m,nsample)
// index tensor going through memory into two gather launches and two elementwise graph ops.
// TPQ threads per query: 32 (a warp does everything; narrow rows) or 256 (one CTA per query: its first warp runs the ball query,
// then all 256 threads copy the nsample x (3 + c) output rows -- wide feature rows, few queries).
template <int TPQ>
__global__ void __launch_bounds__(256) lrg_ball_group_kernel(int b, int n, int m, float radius, int nsample, int c, const float* __restrict__ xyz,
                                                             const float* __restrict__ points, const int* __restrict__ fps_idx,
                                                             float* __restrict__ new_xyz, int* __restrict__ idx, int* __restrict__ pts_cnt, float* __restrict__ grouped_xyz,
                                                             float* __restrict__ new_points) {
  const int lane = threadIdx.x & 31;
  const int tq = threadIdx.x % TPQ;                    // thread within the query's group
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / TPQ;
  const long long ngroups = ((long long)gridDim.x * blockDim.x) / TPQ;
  const int cw = 3 + (points != nullptr ? c : 0);
  for (long long q = gid; q < (long long)b * m; q += ngroups) {      // (uniform per CTA when TPQ == 256)
    const long long bi = q / m;
    const float* p1 = xyz + bi * n * 3;
    const int ci = fps_idx[q];                          // gather_point: the query is the sampled point itself
    const float x2 = p1[(size_t)ci * 3 + 0], y2 = p1[(size_t)ci * 3 + 1], z2 = p1[(size_t)ci * 3 + 2];
    int* out = idx + q * nsample;
    if (tq < 32) {
      if (lane < 3) new_xyz[q * 3 + lane] = lane == 0 ? x2 : lane == 1 ? y2 : z2;
      int cnt = 0, first = -1;
      for (int k0 = 0; k0 < n && cnt < nsample; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < n) {
          const float x1 = p1[k * 3 + 0], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
          const float d = max(sqrtf((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1)), 1e-20f);
          hit = d < radius;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
          if (first < 0) first = k0 + __ffs(bal) - 1;
          const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
          if (hit && pos < nsample) out[pos] = k;
          cnt = min(nsample, cnt + __popc(bal));
        }
      }
      if (first >= 0)
        for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;
      if (lane == 0) pts_cnt[q] = cnt;
    }
    // the row of indices was written by other threads of the group
    if constexpr (TPQ == 32) __syncwarp(); else { __threadfence_block(); __syncthreads(); }
    const float ctr[3] = {x2, y2, z2};
    for (int t = tq; t < nsample * cw; t += TPQ) {
      const int s = t / cw, l = t - s * cw;
      const int i = out[s];                            // (an empty ball leaves the row as the caller initialised it, like the reference)
      float v;
      if (l < 3) {
        v = __fsub_rn(p1[(size_t)i * 3 + l], ctr[l]);
        if (grouped_xyz != nullptr) grouped_xyz[(q * nsample + s) * 3 + l] = v;
      } else {
        v = points[(bi * n + i) * c + (l - 3)];
      }
      new_points[(q * nsample + s) * cw + l] = v;
    }
    if constexpr (TPQ != 32) __syncthreads();           // (the next query of this CTA rewrites nothing of this one; keeps the group together)
  }
}

// ------------------------------------------------------------------------------------- group / grad
__global__ void lrg_group_point_kernel(long long rows, int n, int c, int per_batch_rows, const float* __restrict__ points,
                                       const int* __restrict__ idx, float* __restrict__ out) {
  // rows = b*m*nsample gathered rows of c floats; vectorised when c % 4 == 0
  if ((c & 3) == 0) {
    const int c4 = c >> 2;
    const long long total = rows * c4;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
      const long long r = t / c4;
      const int l = (int)(t - r * c4);
      const long long bi = r / per_batch_rows;
      const float4 v = *reinterpret_cast<const float4*>(points + (bi * n + idx[r]) * c + l * 4);
      *reinterpret_cast<float4*>(out + r * c + l * 4) = v;
    }
  } else {
    const long long total = rows * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
      const long long r = t / c;
      const int l = (int)(t - r * c);
      const long long bi = r / per_batch_rows;
      out[t] = points[(bi * n + idx[r]) * c + l];
    }
  }
}

__global__ void lrg_group_point_grad_kernel(long long rows, int n, int c, int per_batch_rows, const float* __restrict__ grad_out,
                                            const int* __restrict__ idx, float* __restrict__ grad_points) {
  const long long total = rows * c;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / c;
    const int l = (int)(t - r * c);
    const long long bi = r / per_batch_rows;
    atomicAdd(grad_points + (bi * n + idx[r]) * c + l, grad_out[t]);
  }
}

// ------------------------------------------------------------------------------------- knn_point's distance matrix
// tf_grouping.py:66-68: dist[b][j][i] = sum_c (xyz1[b][i][c] - xyz2[b][j][c])^2 (tile + subtract + square + reduce_sum in the
// reference's graph), float32, channels summed left to right.  One thread per entry, rows of xyz1 contiguous across a warp.
__global__ void lrg_pairwise_sqdist_kernel(long long total, int n, int m, int c, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                           float* __restrict__ dist) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long bj = e / n;
    const int i = (int)(e - bj * n);
    const long long b = bj / m;
    const float* p1 = xyz1 + ((size_t)b * n + i) * c;
    const float* p2 = xyz2 + (size_t)bj * c;
    float s = 0.f;
    for (int k = 0; k < c; ++k) {
      const float d = __fsub_rn(p1[k], p2[k]);
      s = __fadd_rn(s, __fmul_rn(d, d));
    }
    dist[e] = s;
  }
}

// ------------------------------------------------------------------------------------- selection sort (top-k)
// Reference: selection_sort_gpu (tf_grouping_g.cu:83-123), one thread per row.  Here one warp per row: the row is
// copied, then k rounds of (argmin over the unsorted suffix with strict '<' => first minimum, swap).
__global__ void __launch_bounds__(256) lrg_selection_sort_kernel(long long rows, int n, int k, const float* __restrict__ dist,
                                                                 int* __restrict__ outi, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = wid; r < rows; r += nwarps) {
    const float* src = dist + r * n;
    float* o = out + r * n;
    int* oi = outi + r * n;
    for (int s = lane; s < n; s += 32) { o[s] = src[s]; oi[s] = s; }
    __syncwarp();
    for (int s = 0; s < k && s < n; ++s) {
      float best = o[s];
      int bi = s;
      for (int t = s + 1 + lane; t < n; t += 32) {
        const float v = o[t];
        if (v < best) { best = v; bi = t; }     // lane-local: ascending t, strict '<' keeps the first minimum
      }
#pragma unroll
      for (int dlt = 16; dlt > 0; dlt >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, dlt);
        const int oi2 = __shfl_xor_sync(0xffffffffu, bi, dlt);
        if (ob < best || (ob == best && oi2 < bi)) { best = ob; bi = oi2; }
      }
      // equal to o[s] never replaces s (strict '<' against p_dist[min] starting at min = s)
      if (!(best < o[s])) bi = s;
      if (lane == 0 && bi != s) {
        const float tv = o[bi]; o[bi] = o[s]; o[s] = tv;
        const int ti = oi[bi]; oi[bi] = oi[s]; oi[s] = ti;
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------- three_nn / interpolate
// Reference: threenn_cpu (tf_interpolate.cpp:60-103) -- a CPU op: float products summed left to right without
// contraction, compared as doubles, strict '<' cascade (earliest index wins ties), 1e40 -> inf / index 0 when m < 3.
__global__ void __launch_bounds__(256) lrg_three_nn_kernel(int b, int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                           float* __restrict__ dist, int* __restrict__ idx) {
  __shared__ float s2[256 * 3];
  const int bi = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const float* q = xyz1 + ((size_t)bi * n) * 3;
  const float* p2 = xyz2 + ((size_t)bi * m) * 3;
  float x1 = 0.f, y1 = 0.f, z1 = 0.f;
  if (j < n) { x1 = q[j * 3 + 0]; y1 = q[j * 3 + 1]; z1 = q[j * 3 + 2]; }
  float best1 = INFINITY, best2 = INFINITY, best3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int k0 = 0; k0 < m; k0 += 256) {
    const int cnt = min(256, m - k0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += 256) s2[t] = p2[(size_t)k0 * 3 + t];
    __syncthreads();
    if (j < n) {
      for (int k = 0; k < cnt; ++k) {
        const float dx = __fsub_rn(s2[k * 3 + 0], x1), dy = __fsub_rn(s2[k * 3 + 1], y1), dz = __fsub_rn(s2[k * 3 + 2], z1);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const int kk = k0 + k;
        if (d < best1) { best3 = best2; i3 = i2; best2 = best1; i2 = i1; best1 = d; i1 = kk; }
        else if (d < best2) { best3 = best2; i3 = i2; best2 = d; i2 = kk; }
        else if (d < best3) { best3 = d; i3 = kk; }
      }
    }
  }
  if (j < n) {
    const size_t o = ((size_t)bi * n + j) * 3;
    dist[o] = best1; dist[o + 1] = best2; dist[o + 2] = best3;
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
  }
}

__global__ void lrg_three_interpolate_kernel(int b, int m, int c, int n, const float* __restrict__ points, const int* __restrict__ idx,
                                             const float* __restrict__ weight, float* __restrict__ out) {
  const long long total = (long long)b * n * c;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / c;                 // b*n row
    const int l = (int)(t - r * c);
    const long long bi = r / n;
    const float* pb = points + bi * m * c;
    const float w1 = weight[r * 3], w2 = weight[r * 3 + 1], w3 = weight[r * 3 + 2];
    const int a1 = idx[r * 3], a2 = idx[r * 3 + 1], a3 = idx[r * 3 + 2];
    out[t] = __fadd_rn(__fadd_rn(__fmul_rn(pb[(size_t)a1 * c + l], w1), __fmul_rn(pb[(size_t)a2 * c + l], w2)), __fmul_rn(pb[(size_t)a3 * c + l], w3));
  }
}

__global__ void lrg_three_interpolate_grad_kernel(int b, int n, int c, int m, const float* __restrict__ grad_out, const int* __restrict__ idx,
                                                  const float* __restrict__ weight, float* __restrict__ grad_points) {
  const long long total = (long long)b * n * c;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / c;
    const int l = (int)(t - r * c);
    const long long bi = r / n;
    float* gp = grad_points + bi * m * c;
    const float g = grad_out[t];
    atomicAdd(gp + (size_t)idx[r * 3] * c + l, __fmul_rn(g, weight[r * 3]));
    atomicAdd(gp + (size_t)idx[r * 3 + 1] * c + l, __fmul_rn(g, weight[r * 3 + 1]));
    atomicAdd(gp + (size_t)idx[r * 3 + 2] * c + l, __fmul_rn(g, weight[r * 3 + 2]));
  }
}

static inline int grid_for(long long work, int threads, int max_blocks = 148 * 16) {
  long long blocks = (work + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > max_blocks) blocks = max_blocks;
  return (int)blocks;
}

// FPS by cloud size: registers of one CTA up to 8,192 points, of a cluster of 2 / 4 / 8 CTAs up to 65,536, else the
// reference's layout (min-distances in the caller's (32,n) workspace).  LRG_FPS_CLUSTER_MIN (tests) moves the first boundary.
static int g_fps_cluster_min = 512 * 16;
static int launch_fps(int b, int n, int m, const float* d_inp, float* d_temp, int* d_out, cudaStream_t st) {
  if (n <= g_fps_cluster_min && n <= 512 * 16) {
    if (n <= 512 * 1) lrg_fps_kernel<512, 1><<<b, 512, 0, st>>>(n, m, d_inp, d_out);
    else if (n <= 512 * 2) lrg_fps_kernel<512, 2><<<b, 512, 0, st>>>(n, m, d_inp, d_out);
    else if (n <= 512 * 4) lrg_fps_kernel<512, 4><<<b, 512, 0, st>>>(n, m, d_inp, d_out);
    else if (n <= 512 * 8) lrg_fps_kernel<512, 8><<<b, 512, 0, st>>>(n, m, d_inp, d_out);
    else lrg_fps_kernel<512, 16><<<b, 512, 0, st>>>(n, m, d_inp, d_out);
  } else if (n <= 8192 * 2) lrg_fps_cluster_kernel<2><<<b * 2, kFpsThreads, 0, st>>>(n, m, d_inp, d_out);
  else if (n <= 8192 * 4) lrg_fps_cluster_kernel<4><<<b * 4, kFpsThreads, 0, st>>>(n, m, d_inp, d_out);
  else if (n <= 8192 * 8) lrg_fps_cluster_kernel<8><<<b * 8, kFpsThreads, 0, st>>>(n, m, d_inp, d_out);
  else {
    LRG_REQUIRE(d_temp != nullptr, "FarthestPointSample with n=%d > %d needs the (32,n) temp workspace", n, 8192 * 8);
    lrg_fps_big_kernel<<<b < 32 ? b : 32, kFpsThreads, 0, st>>>(b, n, m, d_inp, d_temp, d_out);
  }
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg

using namespace lrg;

extern "C" {
#pragma GCC visibility push(default)

int lrg_fps_set_cluster_min(int n) { g_fps_cluster_min = n > 0 ? n : 512 * 16; return LRG_OK; }

int lrg_farthest_point_sampling(int b, int n, int m, const float* d_inp, float* d_temp, int* d_out, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && m >= 0, "FarthestPointSample expects b>=0, n>0, npoint>=0 (got b=%d n=%d m=%d)", b, n, m);
  if (b == 0 || m == 0) return LRG_OK;
  LRG_REQUIRE(d_inp && d_out, "NULL tensor pointer");
  cudaStream_t st = (cudaStream_t)s;
  // (one CTA: fewer, fatter threads were measured slower: 1024 points as 128 threads x 8 take 466 us for 1024 samples, 512 x 2 take 265 us)
  return launch_fps(b, n, m, d_inp, d_temp, d_out, st);
}

int lrg_sample_and_group(int b, int n, int npoint, float radius, int nsample, int c, const float* d_xyz, const float* d_points, float* d_temp,
                         int* d_fps_idx, float* d_new_xyz, float* d_new_points, int* d_idx, int* d_pts_cnt, float* d_grouped_xyz, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && npoint > 0 && nsample > 0 && c >= 0, "sample_and_group: bad shape (b=%d n=%d npoint=%d nsample=%d c=%d)", b, n, npoint, nsample, c);
  if (b == 0) return LRG_OK;
  LRG_REQUIRE(d_xyz && d_fps_idx && d_new_xyz && d_new_points && d_idx && d_pts_cnt, "NULL tensor pointer");
  LRG_REQUIRE(c == 0 || d_points != nullptr, "sample_and_group: c=%d feature channels but points is NULL", c);
  cudaStream_t st = (cudaStream_t)s;
  const int rc = launch_fps(b, n, npoint, d_xyz, d_temp, d_fps_idx, st);
  if (rc != LRG_OK) return rc;
  // wide rows and too few queries to fill the machine with one warp each: a CTA per query
  if ((3 + c) * nsample >= 2048 && (long long)b * npoint <= 148 * 64)
    lrg_ball_group_kernel<256><<<grid_for((long long)b * npoint * 256, 256), 256, 0, st>>>(b, n, npoint, radius, nsample, c, d_xyz, c > 0 ? d_points : nullptr,
                                                                                          d_fps_idx, d_new_xyz, d_idx, d_pts_cnt, d_grouped_xyz, d_new_points);
  else
    lrg_ball_group_kernel<32><<<grid_for((long long)b * npoint * 32, 256), 256, 0, st>>>(b, n, npoint, radius, nsample, c, d_xyz, c > 0 ? d_points : nullptr,
                                                                                        d_fps_idx, d_new_xyz, d_idx, d_pts_cnt, d_grouped_xyz, d_new_points);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_gather_point(int b, int n, int m, const float* d_inp, const int* d_idx, float* d_out, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && m >= 0, "GatherPoint: bad shape");
  if ((long long)b * m == 0) return LRG_OK;
  lrg_gather_point_kernel<<<grid_for((long long)b * m, 256), 256, 0, (cudaStream_t)s>>>(b, n, m, d_inp, d_idx, d_out);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_scatter_add_point(int b, int n, int m, const float* d_out_g, const int* d_idx, float* d_inp_g, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && m >= 0, "GatherPointGrad: bad shape");
  if ((long long)b * m == 0) return LRG_OK;
  lrg_scatter_add_point_kernel<<<grid_for((long long)b * m, 256), 256, 0, (cudaStream_t)s>>>(b, n, m, d_out_g, d_idx, d_inp_g);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_prob_sample(int b, int n, int m, const float* d_inp_p, const float* d_inp_r, float* d_temp, int* d_out, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && m >= 0, "ProbSample: bad shape (b %d, n %d, m %d)", b, n, m);
  if (b == 0) return LRG_OK;
  LRG_REQUIRE(d_inp_p != nullptr && d_temp != nullptr && (d_out != nullptr || m == 0) && (d_inp_r != nullptr || m == 0), "ProbSample: NULL argument (temp is a (b,n) workspace)");
  cudaStream_t st = (cudaStream_t)s;
  lrg_prob_cumsum_kernel<<<b, kProbThreads, 0, st>>>(n, d_inp_p, d_temp);
  if (m > 0) lrg_prob_search_kernel<<<grid_for((long long)b * m, 256), 256, 0, st>>>(b, n, m, d_temp, d_inp_r, d_out);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_query_ball_point(int b, int n, int m, float radius, int nsample, const float* d_xyz1, const float* d_xyz2, int* d_idx,
                         int* d_pts_cnt, lrg_stream_t s) {
  LRG_REQUIRE(radius > 0.f, "QueryBallPoint expects positive radius");            // tf_grouping.cpp:71
  LRG_REQUIRE(nsample > 0, "QueryBallPoint expects positive nsample");            // tf_grouping.cpp:74
  LRG_REQUIRE(b >= 0 && n > 0 && m >= 0, "QueryBallPoint: bad shape");
  if ((long long)b * m == 0) return LRG_OK;
  lrg_query_ball_kernel<<<grid_for((long long)b * m * 32, 256), 256, 0, (cudaStream_t)s>>>(b, n, m, radius, nsample, d_xyz1, d_xyz2, d_idx, d_pts_cnt);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_selection_sort(int b, int n, int m, int k, const float* d_dist, int* d_outi, float* d_out, lrg_stream_t s) {
  LRG_REQUIRE(k > 0, "SelectionSort expects positive k");                         // tf_grouping.cpp:113
  LRG_REQUIRE(b >= 0 && n > 0 && m >= 0, "SelectionSort: bad shape");
  if ((long long)b * m == 0) return LRG_OK;
  lrg_selection_sort_kernel<<<grid_for((long long)b * m * 32, 256), 256, 0, (cudaStream_t)s>>>((long long)b * m, n, k, d_dist, d_outi, d_out);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_pairwise_sqdist(int b, int n, int m, int c, const float* d_xyz1, const float* d_xyz2, float* d_dist, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && m >= 0 && c > 0, "pairwise_sqdist: bad shape");
  const long long total = (long long)b * m * n;
  if (total == 0) return LRG_OK;
  lrg_pairwise_sqdist_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>(total, n, m, c, d_xyz1, d_xyz2, d_dist);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_group_point(int b, int n, int c, int m, int nsample, const float* d_points, const int* d_idx, float* d_out, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && c > 0 && m >= 0 && nsample > 0, "GroupPoint: bad shape");
  const long long rows = (long long)b * m * nsample;
  if (rows == 0) return LRG_OK;
  lrg_group_point_kernel<<<grid_for(rows * c / ((c & 3) ? 1 : 4), 256), 256, 0, (cudaStream_t)s>>>(rows, n, c, m * nsample, d_points, d_idx, d_out);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_group_point_grad(int b, int n, int c, int m, int nsample, const float* d_grad_out, const int* d_idx, float* d_grad_points, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n > 0 && c > 0 && m >= 0 && nsample > 0, "GroupPointGrad: bad shape");
  const long long rows = (long long)b * m * nsample;
  if (rows == 0) return LRG_OK;
  lrg_group_point_grad_kernel<<<grid_for(rows * c, 256), 256, 0, (cudaStream_t)s>>>(rows, n, c, m * nsample, d_grad_out, d_idx, d_grad_points);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_three_nn(int b, int n, int m, const float* d_xyz1, const float* d_xyz2, float* d_dist, int* d_idx, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && n >= 0 && m >= 0, "ThreeNN: bad shape");
  if ((long long)b * n == 0) return LRG_OK;
  LRG_REQUIRE(b <= 65535, "ThreeNN: batch %d > 65535", b);
  lrg_three_nn_kernel<<<dim3((n + 255) / 256, b), 256, 0, (cudaStream_t)s>>>(b, n, m, d_xyz1, d_xyz2, d_dist, d_idx);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_three_interpolate(int b, int m, int c, int n, const float* d_points, const int* d_idx, const float* d_weight, float* d_out, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && m > 0 && c > 0 && n >= 0, "ThreeInterpolate: bad shape");
  const long long total = (long long)b * n * c;
  if (total == 0) return LRG_OK;
  lrg_three_interpolate_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>(b, m, c, n, d_points, d_idx, d_weight, d_out);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int lrg_three_interpolate_grad(int b, int n, int c, int m, const float* d_grad_out, const int* d_idx, const float* d_weight, float* d_grad_points, lrg_stream_t s) {
  LRG_REQUIRE(b >= 0 && m > 0 && c > 0 && n >= 0, "ThreeInterpolateGrad: bad shape");
  const long long total = (long long)b * n * c;
  if (total == 0) return LRG_OK;
  lrg_three_interpolate_grad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>(b, n, c, m, d_grad_out, d_idx, d_weight, d_grad_points);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

#pragma GCC visibility pop
}  // extern "C"
