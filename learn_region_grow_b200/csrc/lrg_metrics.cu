// Per-room segmentation statistics of /root/reference/test_region_grow.py:319-349 computed from ONE
// device pass over the points: the contingency table between the ground-truth object ids and the cluster labels.
//
//   reference                                                         here
//   numpy.unique(obj_id, return_counts)            (:325)             presence bitmap over the value range + block scan
//   sum(and(obj_id==i, cluster_label==j)) per i,j  (:331)             mt_contingency_kernel: one atomicAdd per point
//   sklearn normalized_mutual_info_score           (:346)             table -> host, closed form in double
//   sklearn adjusted_mutual_info_score             (:347)             expected mutual information on the device
//                                                                      (mt_emi_kernel: one warp per table cell, lgamma in double)
//   sklearn adjusted_rand_score                    (:348)             table -> host, pair counts in 128-bit integers
//   greedy IoU > 0.5 matching, PRC / RCL / mean IoU (:320-344)        table -> host
//
// The point arrays are read three times (range, presence, contingency) with coalesced 32-bit loads: 24 B per point.
// Everything that is O(classes x clusters) runs on the host in double precision, following scikit-learn's formulas
// (sklearn/metrics/cluster/_supervised.py, _expected_mutual_info_fast.pyx) term by term.
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "lrg_metrics.cuh"

namespace lrg {

constexpr int kMtThreads = 256;

struct MtRoomRange { int gmin, gmax, lmin, lmax; };

// ------------------------------------------------------------------------------------------------ pass 1: value ranges
__global__ void __launch_bounds__(kMtThreads) mt_range_kernel(const long long* __restrict__ off, const int* __restrict__ obj,
                                                               const int* __restrict__ lab, MtRoomRange* __restrict__ rng) {
  __shared__ int s[4][kMtThreads / 32];
  const int room = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = off[room], n = off[room + 1] - b;
  int gmin = INT_MAX, gmax = INT_MIN, lmin = INT_MAX, lmax = INT_MIN;
  for (long long i = tid; i < n; i += kMtThreads) {
    const int g = obj[b + i], l = lab[b + i];
    gmin = min(gmin, g); gmax = max(gmax, g); lmin = min(lmin, l); lmax = max(lmax, l);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    gmin = min(gmin, __shfl_xor_sync(0xffffffffu, gmin, d)); gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, d));
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, d)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
  }
  if (lane == 0) { s[0][warp] = gmin; s[1][warp] = gmax; s[2][warp] = lmin; s[3][warp] = lmax; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kMtThreads / 32; ++w) {
      gmin = min(gmin, s[0][w]); gmax = max(gmax, s[1][w]); lmin = min(lmin, s[2][w]); lmax = max(lmax, s[3][w]);
    }
    rng[room] = MtRoomRange{gmin, gmax, lmin, lmax};
  }
}

// ------------------------------------------------------------------------------------------------ pass 2: dense ids
// map_g / map_l: one int per value of the room's range, zeroed; presence marks, then an exclusive scan in value order
// turns the marks into dense indices (= the order of numpy.unique) and -1 for absent values.
__global__ void __launch_bounds__(kMtThreads) mt_presence_kernel(const long long* __restrict__ off, const int* __restrict__ obj,
                                                                  const int* __restrict__ lab, const MtRoomRange* __restrict__ rng,
                                                                  const long long* __restrict__ mg_off, const long long* __restrict__ ml_off,
                                                                  int* __restrict__ map_g, int* __restrict__ map_l) {
  const int room = blockIdx.y;
  const long long b = off[room], n = off[room + 1] - b;
  const MtRoomRange r = rng[room];
  for (long long i = (long long)blockIdx.x * kMtThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kMtThreads) {
    map_g[mg_off[room] + (obj[b + i] - r.gmin)] = 1;
    map_l[ml_off[room] + (lab[b + i] - r.lmin)] = 1;
  }
}

__global__ void __launch_bounds__(kMtThreads) mt_dense_kernel(const long long* __restrict__ m_off, int* __restrict__ map, int* __restrict__ n_dense) {
  __shared__ int s_warp[kMtThreads / 32 + 1];
  __shared__ int s_run;
  const int room = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long b = m_off[room], n = m_off[room + 1] - b;
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (long long start = 0; start < n; start += kMtThreads) {
    const long long i = start + tid;
    const int v = i < n ? map[b + i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    const int run = s_run;
    if (i < n) map[b + i] = v ? run + wbase + incl - 1 : -1;
    __syncthreads();
    if (tid == kMtThreads - 1) s_run = run + wbase + incl;
    __syncthreads();
  }
  if (tid == 0) n_dense[room] = s_run;
}

// ------------------------------------------------------------------------------------------------ pass 3: contingency
__global__ void __launch_bounds__(kMtThreads) mt_contingency_kernel(const long long* __restrict__ off, const int* __restrict__ obj,
                                                                     const int* __restrict__ lab, const MtRoomRange* __restrict__ rng,
                                                                     const long long* __restrict__ mg_off, const long long* __restrict__ ml_off,
                                                                     const int* __restrict__ map_g, const int* __restrict__ map_l,
                                                                     const int* __restrict__ n_clusters, const long long* __restrict__ tab_off,
                                                                     unsigned* __restrict__ tab) {
  const int room = blockIdx.y;
  const long long b = off[room], n = off[room + 1] - b;
  const MtRoomRange r = rng[room];
  const int K = n_clusters[room];
  unsigned* T = tab + tab_off[room];
  for (long long i = (long long)blockIdx.x * kMtThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kMtThreads) {
    const int gi = map_g[mg_off[room] + (obj[b + i] - r.gmin)];
    const int kj = map_l[ml_off[room] + (lab[b + i] - r.lmin)];
    atomicAdd(T + (size_t)gi * K + kj, 1u);
  }
}

// Row sums a_i (classes) and column sums b_j (clusters) of every room's table; sums[tab-independent offsets]:
// a at ab_off[room], b right after it.
__global__ void __launch_bounds__(kMtThreads) mt_marginals_kernel(const int* __restrict__ n_classes, const int* __restrict__ n_clusters,
                                                                   const long long* __restrict__ tab_off, const unsigned* __restrict__ tab,
                                                                   const long long* __restrict__ ab_off, long long* __restrict__ ab) {
  const int room = blockIdx.x, G = n_classes[room], K = n_clusters[room];
  const unsigned* T = tab + tab_off[room];
  long long* a = ab + ab_off[room];
  long long* b = a + G;
  for (int i = threadIdx.x; i < G; i += kMtThreads) {
    long long s = 0;
    for (int j = 0; j < K; ++j) s += T[(size_t)i * K + j];
    a[i] = s;
  }
  for (int j = threadIdx.x; j < K; j += kMtThreads) {
    long long s = 0;
    for (int i = 0; i < G; ++i) s += T[(size_t)i * K + j];
    b[j] = s;
  }
}

// Expected mutual information (sklearn _expected_mutual_info_fast.pyx): one warp per table cell (i, j); lanes stride over
// nij in [max(1, a+b-N), min(a,b)], fixed-shape warp reduction; cell results are summed per room in a fixed order by
// mt_emi_sum_kernel so the value does not depend on scheduling.
__global__ void __launch_bounds__(kMtThreads) mt_emi_kernel(int n_rooms, const long long* __restrict__ off, const int* __restrict__ n_classes,
                                                             const int* __restrict__ n_clusters, const long long* __restrict__ tab_off,
                                                             const long long* __restrict__ ab_off, const long long* __restrict__ ab,
                                                             double* __restrict__ cell) {
  const int room = blockIdx.y, G = n_classes[room], K = n_clusters[room];
  if (G <= 1 || K <= 1) return;
  const long long N = off[room + 1] - off[room];
  const long long* a = ab + ab_off[room];
  const long long* b = a + G;
  const int lane = threadIdx.x & 31;
  const long long cells = (long long)G * K;
  const double logN = log((double)N), glnN = lgamma((double)N + 1.0);
  for (long long c = (long long)blockIdx.x * (kMtThreads / 32) + (threadIdx.x >> 5); c < cells; c += (long long)gridDim.x * (kMtThreads / 32)) {
    const long long ai = a[c / K], bj = b[c % K];
    const double log_a = log((double)ai), log_b = log((double)bj);
    const double fixed = lgamma((double)ai + 1.0) + lgamma((double)bj + 1.0) + lgamma((double)(N - ai) + 1.0) + lgamma((double)(N - bj) + 1.0) - glnN;
    const long long start = max(1ll, ai - N + bj), end = min(ai, bj) + 1;
    double acc = 0.0;
    for (long long nij = start + lane; nij < end; nij += 32) {
      const double term1 = (double)nij / (double)N;
      const double term2 = logN + log((double)nij) - log_a - log_b;
      const double gln = fixed - lgamma((double)nij + 1.0) - lgamma((double)(ai - nij) + 1.0) - lgamma((double)(bj - nij) + 1.0) -
                         lgamma((double)(N - ai - bj + nij) + 1.0);
      acc += term1 * term2 * exp(gln);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) cell[tab_off[room] + c] = acc;
  }
}

__global__ void __launch_bounds__(kMtThreads) mt_emi_sum_kernel(const int* __restrict__ n_classes, const int* __restrict__ n_clusters,
                                                                 const long long* __restrict__ tab_off, const double* __restrict__ cell,
                                                                 double* __restrict__ emi) {
  __shared__ double s[kMtThreads];
  const int room = blockIdx.x, G = n_classes[room], K = n_clusters[room], tid = threadIdx.x;
  double acc = 0.0;
  if (G > 1 && K > 1)
    for (long long c = tid; c < (long long)G * K; c += kMtThreads) acc += cell[tab_off[room] + c];
  s[tid] = acc;
  __syncthreads();
  for (int d = kMtThreads / 2; d > 0; d >>= 1) {
    if (tid < d) s[tid] += s[tid + d];
    __syncthreads();
  }
  if (tid == 0) emi[room] = s[0];
}

// cluster_label2 (:323,335,339-341): matched clusters take the rank (k+1) of their object, the others j + obj_id.max().
__global__ void __launch_bounds__(kMtThreads) mt_relabel_kernel(const long long* __restrict__ off, const int* __restrict__ lab,
                                                                 const long long* __restrict__ relabel_off, const int* __restrict__ relabel,
                                                                 int* __restrict__ out) {
  const int room = blockIdx.y;
  const long long b = off[room], n = off[room + 1] - b;
  const int* R = relabel + relabel_off[room];
  const int lmax = (int)(relabel_off[room + 1] - relabel_off[room]) - 1;
  for (long long i = (long long)blockIdx.x * kMtThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kMtThreads) {
    const int l = lab[b + i];
    out[b + i] = (l >= 1 && l <= lmax) ? R[l] : 0;
  }
}

// ------------------------------------------------------------------------------------------------ host side
namespace {

// Scratch from the stream-ordered allocator (the device's default memory pool keeps the blocks between calls, so a call
// costs no cudaMalloc / cudaFree round trips after the first).
struct DevBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  ~DevBuf() { if (p) cudaFreeAsync(p, st); }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
  cudaError_t alloc(size_t bytes, cudaStream_t stream) { st = stream; return cudaMallocAsync(&p, std::max<size_t>(bytes, 16), stream); }
};

double entropy_of(const std::vector<long long>& counts) {
  // sklearn _entropy: 0 for a single cluster, else -sum(pi/sum * (log(pi) - log(sum)))
  if (counts.size() == 1) return 0.0;
  double total = 0.0;
  for (long long c : counts) total += (double)c;
  const double log_total = log(total);
  double h = 0.0;
  for (long long c : counts) h += ((double)c / total) * (log((double)c) - log_total);
  return -h;
}

}  // namespace

// Scores of one room from its contingency table T (G x K), emi from the device.
static void room_scores(long long N, int G, int K, const unsigned* T, const int* class_values, const int* cluster_values,
                        double emi, int obj_max, LrgRoomMetrics& m, int* relabel /* lmax+1 ints or NULL */) {
  const double kNaN = NAN;
  m.nmi = m.ami = m.ars = m.prc = m.rcl = m.iou = kNaN;
  m.n_points = (int)N; m.n_classes = G; m.n_clusters = 0; m.gt_match = 0;
  if (N <= 0) return;
  std::vector<long long> a(G, 0), b(K, 0);
  for (int i = 0; i < G; ++i)
    for (int j = 0; j < K; ++j) { a[i] += T[(size_t)i * K + j]; b[j] += T[(size_t)i * K + j]; }
  // ---- mutual information (sklearn mutual_info_score on the sparse contingency)
  double mi = 0.0;
  if (G > 1 && K > 1) {
    const double dN = (double)N, logN = log(dN);
    double pi_sum = 0.0, pj_sum = 0.0;
    for (long long v : a) pi_sum += (double)v;
    for (long long v : b) pj_sum += (double)v;
    const double log_sums = log(pi_sum) + log(pj_sum);
    for (int i = 0; i < G; ++i)
      for (int j = 0; j < K; ++j) {
        const unsigned nij = T[(size_t)i * K + j];
        if (nij == 0) continue;
        const double nm = (double)nij / dN;
        const double outer = (double)(a[i] * b[j]);
        double t = nm * (log((double)nij) - logN) + nm * (-log(outer) + log_sums);
        if (fabs(t) < 2.220446049250313e-16) t = 0.0;
        mi += t;
      }
    if (mi < 0.0) mi = 0.0;
  }
  const double h_true = entropy_of(a), h_pred = entropy_of(b);
  const double normalizer = 0.5 * (h_true + h_pred);
  // ---- NMI / AMI (average_method='arithmetic', the default since scikit-learn 0.22)
  if (G == 1 && K == 1) {
    m.nmi = 1.0; m.ami = 1.0;
  } else {
    m.nmi = mi == 0.0 ? 0.0 : mi / normalizer;
    if (G == 1 || K == 1) {
      m.ami = 0.0;
    } else {
      const double eps = 2.220446049250313e-16;
      double den = normalizer - emi, num = mi - emi;
      den = den < 0 ? std::min(den, -eps) : std::max(den, eps);
      num = num < 0 ? std::min(num, -eps) : std::max(num, eps);
      m.ami = num / den;
    }
  }
  // ---- adjusted Rand score from the pair confusion matrix (exact integers)
  {
    typedef __int128 i128;
    i128 sum_sq = 0, c01 = 0, c10 = 0;
    for (int i = 0; i < G; ++i)
      for (int j = 0; j < K; ++j) {
        const i128 v = T[(size_t)i * K + j];
        sum_sq += v * v; c01 += v * (i128)b[j]; c10 += v * (i128)a[i];
      }
    const i128 n = N;
    const i128 tp = sum_sq - n, fp = c01 - sum_sq, fn = c10 - sum_sq, tn = n * n - fp - fn - sum_sq;
    if (fn == 0 && fp == 0) m.ars = 1.0;
    else m.ars = 2.0 * ((double)tp * (double)tn - (double)fn * (double)fp) /
                 ((double)(tp + fn) * (double)(fn + tn) + (double)(tp + fp) * (double)(fp + tn));
  }
  // ---- greedy matching (:320-344): objects by descending point count; ties in the order of a stable ascending sort reversed
  int lmax = 0;
  for (int j = 0; j < K; ++j) lmax = std::max(lmax, cluster_values[j]);
  m.n_clusters = lmax;
  std::vector<int> col_of(lmax + 1, -1);
  for (int j = 0; j < K; ++j)
    if (cluster_values[j] >= 1) col_of[cluster_values[j]] = j;
  std::vector<int> by_count(G);
  for (int i = 0; i < G; ++i) by_count[i] = i;
  std::stable_sort(by_count.begin(), by_count.end(), [&](int x, int y) { return a[x] < a[y]; });
  std::reverse(by_count.begin(), by_count.end());
  std::vector<char> dt_match(lmax, 0);
  if (relabel != nullptr) for (int j = 0; j <= lmax; ++j) relabel[j] = 0;
  double iou_sum = 0.0;
  int gt_match = 0;
  for (int k = 0; k < G; ++k) {
    const int i = by_count[k];
    double best = 0.0;
    for (int j = 1; j <= lmax; ++j) {
      if (dt_match[j - 1]) continue;
      const int c = col_of[j];
      const long long inter = c >= 0 ? (long long)T[(size_t)i * K + c] : 0;
      const long long uni = a[i] + (c >= 0 ? b[c] : 0) - inter;
      const double iou = 1.0 * (double)inter / (double)uni;
      best = std::max(best, iou);
      if (iou > 0.5) {
        dt_match[j - 1] = 1; gt_match += 1;
        if (relabel != nullptr) relabel[j] = k + 1;
        break;
      }
    }
    iou_sum += best;
  }
  if (relabel != nullptr)
    for (int j = 1; j <= lmax; ++j)
      if (!dt_match[j - 1]) relabel[j] = j + obj_max;
  int matched = 0;
  for (char c : dt_match) matched += c;
  m.prc = lmax > 0 ? (double)matched / (double)lmax : kNaN;      // numpy.mean of an empty array is nan
  m.rcl = 1.0 * gt_match / G;
  m.iou = iou_sum / G;
  m.gt_match = gt_match;
  (void)class_values;
}

int segmentation_metrics(int n_rooms, const int64_t* room_offsets, const int32_t* d_obj_id, const int32_t* d_label,
                         LrgRoomMetrics* out, int32_t* d_label2, cudaStream_t st) {
  LRG_REQUIRE(n_rooms >= 0 && room_offsets != nullptr && (out != nullptr || n_rooms == 0), "metrics: bad arguments");
  if (n_rooms == 0) return LRG_OK;
  const long long total = room_offsets[n_rooms] - room_offsets[0];
  LRG_REQUIRE(room_offsets[0] == 0 && total >= 0, "metrics: room_offsets must start at 0 and be non-decreasing");
  LRG_REQUIRE(total == 0 || (d_obj_id != nullptr && d_label != nullptr), "metrics: NULL obj_id / cluster_label");
  const int R = n_rooms;
  long long maxN = 0;
  for (int r = 0; r < R; ++r) {
    LRG_REQUIRE(room_offsets[r + 1] >= room_offsets[r], "metrics: room_offsets must be non-decreasing");
    maxN = std::max<long long>(maxN, room_offsets[r + 1] - room_offsets[r]);
  }
  std::vector<long long> h_off(room_offsets, room_offsets + R + 1);
  {   // keep freed scratch in the device's default pool across calls (the default threshold releases it at every sync)
    int dev = 0;
    cudaMemPool_t mp = nullptr;
    unsigned long long keep = ~0ull;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess)
      cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  DevBuf d_off, d_rng;
  LRG_CUDA(d_off.alloc(sizeof(long long) * (R + 1), st));
  LRG_CUDA(d_rng.alloc(sizeof(MtRoomRange) * R, st));
  LRG_CUDA(cudaMemcpyAsync(d_off.p, h_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, st));
  mt_range_kernel<<<R, kMtThreads, 0, st>>>(d_off.as<long long>(), d_obj_id, d_label, d_rng.as<MtRoomRange>());
  std::vector<MtRoomRange> rng(R);
  LRG_CUDA(cudaMemcpyAsync(rng.data(), d_rng.p, sizeof(MtRoomRange) * R, cudaMemcpyDeviceToHost, st));
  LRG_CUDA(cudaStreamSynchronize(st));
  std::vector<long long> mg_off(R + 1, 0), ml_off(R + 1, 0);
  for (int r = 0; r < R; ++r) {
    const bool empty = h_off[r + 1] == h_off[r];
    const long long gspan = empty ? 0 : (long long)rng[r].gmax - rng[r].gmin + 1, lspan = empty ? 0 : (long long)rng[r].lmax - rng[r].lmin + 1;
    mg_off[r + 1] = mg_off[r] + gspan; ml_off[r + 1] = ml_off[r] + lspan;
  }
  const long long kMaxSpan = 1ll << 27;      // 512 MB of map per array at most
  LRG_REQUIRE(mg_off[R] <= kMaxSpan && ml_off[R] <= kMaxSpan,
              "metrics: the obj_id / cluster_label values of the rooms span %lld / %lld integers in total (limit %lld): remap them to small ids",
              mg_off[R], ml_off[R], kMaxSpan);
  DevBuf d_mg_off, d_ml_off, d_map_g, d_map_l, d_ng, d_nk;
  LRG_CUDA(d_mg_off.alloc(sizeof(long long) * (R + 1), st)); LRG_CUDA(d_ml_off.alloc(sizeof(long long) * (R + 1), st));
  LRG_CUDA(d_map_g.alloc(sizeof(int) * mg_off[R], st)); LRG_CUDA(d_map_l.alloc(sizeof(int) * ml_off[R], st));
  LRG_CUDA(d_ng.alloc(sizeof(int) * R, st)); LRG_CUDA(d_nk.alloc(sizeof(int) * R, st));
  LRG_CUDA(cudaMemcpyAsync(d_mg_off.p, mg_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, st));
  LRG_CUDA(cudaMemcpyAsync(d_ml_off.p, ml_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, st));
  LRG_CUDA(cudaMemsetAsync(d_map_g.p, 0, sizeof(int) * mg_off[R], st));
  LRG_CUDA(cudaMemsetAsync(d_map_l.p, 0, sizeof(int) * ml_off[R], st));
  const int chunks = (int)std::max<long long>(1, std::min<long long>(64, (maxN + kMtThreads * 4 - 1) / (kMtThreads * 4)));
  const dim3 grid_pts(chunks, R);
  mt_presence_kernel<<<grid_pts, kMtThreads, 0, st>>>(d_off.as<long long>(), d_obj_id, d_label, d_rng.as<MtRoomRange>(), d_mg_off.as<long long>(),
                                                      d_ml_off.as<long long>(), d_map_g.as<int>(), d_map_l.as<int>());
  mt_dense_kernel<<<R, kMtThreads, 0, st>>>(d_mg_off.as<long long>(), d_map_g.as<int>(), d_ng.as<int>());
  mt_dense_kernel<<<R, kMtThreads, 0, st>>>(d_ml_off.as<long long>(), d_map_l.as<int>(), d_nk.as<int>());
  std::vector<int> nG(R), nK(R), h_map_g(mg_off[R]), h_map_l(ml_off[R]);
  LRG_CUDA(cudaMemcpyAsync(nG.data(), d_ng.p, sizeof(int) * R, cudaMemcpyDeviceToHost, st));
  LRG_CUDA(cudaMemcpyAsync(nK.data(), d_nk.p, sizeof(int) * R, cudaMemcpyDeviceToHost, st));
  LRG_CUDA(cudaStreamSynchronize(st));
  std::vector<long long> tab_off(R + 1, 0), ab_off(R + 1, 0);
  for (int r = 0; r < R; ++r) { tab_off[r + 1] = tab_off[r] + (long long)nG[r] * nK[r]; ab_off[r + 1] = ab_off[r] + nG[r] + nK[r]; }
  LRG_REQUIRE(tab_off[R] <= (1ll << 28), "metrics: contingency tables need %lld cells (limit 2^28)", tab_off[R]);
  DevBuf d_tab_off, d_ab_off, d_tab, d_ab, d_cell, d_emi;
  LRG_CUDA(d_tab_off.alloc(sizeof(long long) * (R + 1), st)); LRG_CUDA(d_ab_off.alloc(sizeof(long long) * (R + 1), st));
  LRG_CUDA(d_tab.alloc(sizeof(unsigned) * tab_off[R], st)); LRG_CUDA(d_ab.alloc(sizeof(long long) * ab_off[R], st));
  LRG_CUDA(d_cell.alloc(sizeof(double) * tab_off[R], st)); LRG_CUDA(d_emi.alloc(sizeof(double) * R, st));
  LRG_CUDA(cudaMemcpyAsync(d_tab_off.p, tab_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, st));
  LRG_CUDA(cudaMemcpyAsync(d_ab_off.p, ab_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, st));
  LRG_CUDA(cudaMemsetAsync(d_tab.p, 0, sizeof(unsigned) * tab_off[R], st));
  mt_contingency_kernel<<<grid_pts, kMtThreads, 0, st>>>(d_off.as<long long>(), d_obj_id, d_label, d_rng.as<MtRoomRange>(), d_mg_off.as<long long>(),
                                                         d_ml_off.as<long long>(), d_map_g.as<int>(), d_map_l.as<int>(), d_nk.as<int>(),
                                                         d_tab_off.as<long long>(), d_tab.as<unsigned>());
  mt_marginals_kernel<<<R, kMtThreads, 0, st>>>(d_ng.as<int>(), d_nk.as<int>(), d_tab_off.as<long long>(), d_tab.as<unsigned>(),
                                                d_ab_off.as<long long>(), d_ab.as<long long>());
  long long max_cells = 1;
  for (int r = 0; r < R; ++r) max_cells = std::max(max_cells, (long long)nG[r] * nK[r]);
  const dim3 grid_emi((unsigned)std::min<long long>(1024, (max_cells + kMtThreads / 32 - 1) / (kMtThreads / 32)), R);
  mt_emi_kernel<<<grid_emi, kMtThreads, 0, st>>>(R, d_off.as<long long>(), d_ng.as<int>(), d_nk.as<int>(), d_tab_off.as<long long>(),
                                                 d_ab_off.as<long long>(), d_ab.as<long long>(), d_cell.as<double>());
  mt_emi_sum_kernel<<<R, kMtThreads, 0, st>>>(d_ng.as<int>(), d_nk.as<int>(), d_tab_off.as<long long>(), d_cell.as<double>(), d_emi.as<double>());
  std::vector<unsigned> tab(tab_off[R]);
  std::vector<double> emi(R);
  LRG_CUDA(cudaMemcpyAsync(tab.data(), d_tab.p, sizeof(unsigned) * tab_off[R], cudaMemcpyDeviceToHost, st));
  LRG_CUDA(cudaMemcpyAsync(emi.data(), d_emi.p, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
  LRG_CUDA(cudaMemcpyAsync(h_map_g.data(), d_map_g.p, sizeof(int) * mg_off[R], cudaMemcpyDeviceToHost, st));
  LRG_CUDA(cudaMemcpyAsync(h_map_l.data(), d_map_l.p, sizeof(int) * ml_off[R], cudaMemcpyDeviceToHost, st));
  LRG_CUDA(cudaStreamSynchronize(st));
  LRG_CUDA(cudaGetLastError());
  // ---- everything that is O(classes x clusters): host, double precision
  std::vector<long long> relabel_off(R + 1, 0);
  std::vector<int> relabel;
  for (int r = 0; r < R; ++r) {
    const bool empty = h_off[r + 1] == h_off[r];
    relabel_off[r + 1] = relabel_off[r] + (empty ? 1 : std::max(rng[r].lmax, 0) + 1);
  }
  if (d_label2 != nullptr) relabel.assign((size_t)relabel_off[R], 0);
  for (int r = 0; r < R; ++r) {
    const long long N = h_off[r + 1] - h_off[r];
    std::vector<int> class_values(nG[r]), cluster_values(nK[r]);
    if (N > 0) {
      for (long long v = 0; v < mg_off[r + 1] - mg_off[r]; ++v) { const int d = h_map_g[mg_off[r] + v]; if (d >= 0) class_values[d] = rng[r].gmin + (int)v; }
      for (long long v = 0; v < ml_off[r + 1] - ml_off[r]; ++v) { const int d = h_map_l[ml_off[r] + v]; if (d >= 0) cluster_values[d] = rng[r].lmin + (int)v; }
    }
    room_scores(N, nG[r], nK[r], tab.data() + tab_off[r], class_values.data(), cluster_values.data(), emi[r], N > 0 ? rng[r].gmax : 0,
                out[r], d_label2 != nullptr ? relabel.data() + relabel_off[r] : nullptr);
  }
  if (d_label2 != nullptr && total > 0) {
    DevBuf d_rel_off, d_rel;
    LRG_CUDA(d_rel_off.alloc(sizeof(long long) * (R + 1), st)); LRG_CUDA(d_rel.alloc(sizeof(int) * relabel.size(), st));
    LRG_CUDA(cudaMemcpyAsync(d_rel_off.p, relabel_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, st));
    LRG_CUDA(cudaMemcpyAsync(d_rel.p, relabel.data(), sizeof(int) * relabel.size(), cudaMemcpyHostToDevice, st));
    mt_relabel_kernel<<<grid_pts, kMtThreads, 0, st>>>(d_off.as<long long>(), d_label, d_rel_off.as<long long>(), d_rel.as<int>(), d_label2);
    LRG_CUDA(cudaStreamSynchronize(st));
    LRG_CUDA(cudaGetLastError());
  }
  return LRG_OK;
}

// obj_id[equalized_idx] (:136): ground-truth ids of the raw points -> ids of the equalised points
__global__ void __launch_bounds__(kMtThreads) mt_gather_eq_kernel(const long long* __restrict__ raw_off, const long long* __restrict__ eq_off,
                                                                   const int* __restrict__ equalized_idx, const int* __restrict__ obj_raw,
                                                                   int* __restrict__ obj_eq) {
  const int room = blockIdx.y;
  const long long rb = raw_off[room], eb = eq_off[room], n = eq_off[room + 1] - eb;
  for (long long i = (long long)blockIdx.x * kMtThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kMtThreads)
    obj_eq[eb + i] = obj_raw[rb + equalized_idx[eb + i]];
}

int launch_gather_equalized(int n_rooms, const long long* d_raw_off, const long long* d_eq_off, const int* d_equalized_idx, const int* d_obj_raw,
                            int* d_obj_eq, cudaStream_t st) {
  if (n_rooms <= 0) return LRG_OK;
  mt_gather_eq_kernel<<<dim3(32, n_rooms), kMtThreads, 0, st>>>(d_raw_off, d_eq_off, d_equalized_idx, d_obj_raw, d_obj_eq);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg

extern "C" {
#pragma GCC visibility push(default)

int lrg_segmentation_metrics(int n_rooms, const int64_t* room_offsets, const int32_t* d_obj_id, const int32_t* d_cluster_label,
                             LrgRoomMetrics* out, int32_t* d_cluster_label2, lrg_stream_t s) {
  return lrg::segmentation_metrics(n_rooms, room_offsets, d_obj_id, d_cluster_label, out, d_cluster_label2, (cudaStream_t)s);
}

#pragma GCC visibility pop
}  // extern "C"
