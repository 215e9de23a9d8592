// Host side of liblrg_b200.so: engine object (weights, workspaces, stream, CUDA graph of the lock-step grow loop) and
// the extern "C" entry points of include/lrg_b200.h that concern LrgNet and the region-grow driver.
#include <cuda_fp16.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "lrg_driver.cuh"
#include "lrg_featprep.cuh"
#include "lrg_metrics.cuh"
#include "lrg_persistent.cuh"
#include "lrg_tc.cuh"
#include "lrg_umma.cuh"

namespace lrg {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

template <class T>
static int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t err = cudaMalloc((void**)p, count * sizeof(T));
  if (err != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(err));
    return LRG_E_NOMEM;
  }
  return LRG_OK;
}

template <class T>
static int pool_alloc(LrgEngine* e, T** p, size_t count);
static void pool_free(LrgEngine* e, void* p);

#define LRG_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != LRG_OK) return rc__; \
  } while (0)

struct HostLayer { int K, N; size_t w_off, b_off; };

}  // namespace lrg

using namespace lrg;

struct PoolBlock { void* p; size_t bytes; bool used; };

struct LrgEngine {
  int device = 0;
  // device-memory pool for the per-upload arrays and workspaces: cudaMalloc / cudaFree cost milliseconds (and synchronise),
  // which would dominate an upload; blocks are recycled across calls and released when the engine is destroyed
  std::vector<PoolBlock> pool;
  int F = 13, Ni = 512, Nj = 512, lite = 0, max_batch = 1;
  std::vector<int> conv, conv2;
  size_t n_weights = 0;
  bool weights_loaded = false;
  cudaStream_t stream = nullptr;
  NetDesc net{};
  float* d_weights = nullptr;      // packed/padded device copy
  size_t packed_floats = 0;
  // tensor-core path (full model only): operand images + descriptor
  bool tc_available = false;
  int forward_mode = LRG_FORWARD_AUTO;
  unsigned char* d_tc_img = nullptr;   // operand images of the tensor tiles: 3xTF32 set, then 3xFP16 set
  int* d_range_flag = nullptr;         // 3xFP16 tiles: an activation left the fp16 range (TcNet::range_flag)
  bool force_tf32 = false;             // set while a call is repeated with 3xTF32 after a range overflow
  int last_kind = 0;                   // arithmetic of the last forward / segment call: 0 FMA, 1 3xTF32, 2 3xFP16
  int range_fallbacks = 0;             // calls repeated with 3xTF32 so far
  std::atomic<bool> busy{false};       // a segment call is running on this handle
  TcNet tcnet{};
  // forward workspaces for max_batch tile pairs (user-facing forward)
  int ws_batch = 0;
  float *d_x[2] = {nullptr, nullptr}, *d_h1[2] = {nullptr, nullptr}, *d_pooled = nullptr, *d_gproj = nullptr,
        *d_logits[2] = {nullptr, nullptr};
  // rooms
  int n_rooms = 0;
  long long total_pts = 0;
  int maxN = 0;
  float resolution = 0.1f;
  std::vector<long long> h_room_off;
  long long* d_room_off = nullptr;
  float* d_pts = nullptr;
  unsigned* d_pw = nullptr;              // packed per-point state words (rooms padded to 4 words)
  long long* d_pw_off = nullptr;
  // spatial index of the rooms (DriverArgs::sp_*, built at upload)
  long long* d_sp_off = nullptr;
  int* d_sp_perm = nullptr;
  unsigned* d_sp_vox = nullptr;
  uint2* d_sp_box = nullptr;
  int4* d_room_vmin = nullptr;
  long long total_words = 0;
  int *d_label = nullptr, *d_label_filled = nullptr, *d_order = nullptr;
  int *d_lab_list = nullptr, *d_unl_list = nullptr, *d_n_lab = nullptr, *d_n_unl = nullptr;
  LrgRoomStats* d_stats = nullptr;
  // rooms uploaded as raw points (device feature preparation): maps between raw and equalised points, dense features
  bool raw_mode = false;
  long long total_raw = 0;
  std::vector<long long> h_raw_off;
  long long* d_raw_off = nullptr;
  int *d_equalized_idx = nullptr, *d_unequalized_idx = nullptr;
  float* d_feat = nullptr;
  // slots
  int n_slots = 0, slots_maxN = 0;
  SlotState* d_slots = nullptr;
  int *d_listI = nullptr, *d_listJ = nullptr;
  unsigned *d_keyI = nullptr, *d_keyJ = nullptr;
  float* d_tile[2] = {nullptr, nullptr};
  int* d_tileidx[2] = {nullptr, nullptr};
  int* d_tilesrc[2] = {nullptr, nullptr};
  float *s_h1[2] = {nullptr, nullptr}, *s_pooled = nullptr, *s_gproj = nullptr, *s_logits[2] = {nullptr, nullptr};
  uint2* s_gproj_tagged = nullptr;      // (n_slots, 2, H0) {value, tag} words written by the projection servers
  int* d_counters = nullptr;       // [0] next_room, [1] finished_slots
  int* h_done = nullptr;           // mapped pinned
  int* d_done = nullptr;
  LrgStepTrace* d_trace = nullptr;
  int trace_capacity = 0, trace_rooms = 0;
  // random restarts: per-lane copies of the state words, group records, per-lane step counts of the last run
  unsigned* d_pw_lanes = nullptr;
  LaneGroup* d_groups = nullptr;
  int* d_parI = nullptr;        // beam search: index lists of the candidates in every group's queue
  SpecSync* d_spec_sync = nullptr;   // speculative lanes: per-group commit order; commit log
  int* d_room_order = nullptr;         // rooms in the order they are started (largest first)
  long long* d_pending_pts = nullptr;  // suffix sums of their point counts
  int* d_spec_est = nullptr;         // speculative lanes: per-group estimate of the grow steps its room still needs
  int* d_clog = nullptr;
  int* d_lane_steps = nullptr;
  int last_lanes = 1;
  // persistent grow kernel: work queue, per-slot stage counters, busy-time counters
  int sm_count = 0;
  unsigned long long* d_qring = nullptr;
  unsigned q_capacity = 0;
  unsigned* d_qctr = nullptr;            // [0] head, [1] tail of the high-priority ring, [2], [3] of the normal ring
  int* d_remaining = nullptr;
  SlotSync* d_sync = nullptr;
  int sync_slots = 0;
  unsigned long long* d_busy = nullptr;  // [24]
  unsigned long long h_busy[24] = {0};
  bool last_persistent = false;
  unsigned long long* d_tile_dbg = nullptr;   // [32] tile-stage cycle counters (diagnostics, LRG_TILE_TIMING=1)
  // profile of the last segment call
  float grow_ms = 0, fill_ms = 0, forward_ms = 0, prep_ms = 0;
  float kernel_ms[4] = {0, 0, 0, 0};
  long long iterations = 0, launches = 0;
  int prep_launches = 0;               // kernels of the last upload (feature preparation, pack, spatial index)
};

namespace lrg {

template <class T>
static int pool_alloc(LrgEngine* e, T** p, size_t count) {
  *p = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  int best = -1;
  for (size_t i = 0; i < e->pool.size(); ++i)
    if (!e->pool[i].used && e->pool[i].bytes >= bytes && e->pool[i].bytes <= 2 * bytes + 4096 &&
        (best < 0 || e->pool[i].bytes < e->pool[best].bytes))
      best = (int)i;
  if (best >= 0) {
    e->pool[best].used = true;
    *p = reinterpret_cast<T*>(e->pool[best].p);
    return LRG_OK;
  }
  void* q = nullptr;
  cudaError_t err = cudaMalloc(&q, bytes);
  if (err != cudaSuccess) {
    // release the idle blocks and retry once before giving up
    for (auto& b : e->pool)
      if (!b.used && b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    err = cudaMalloc(&q, bytes);
  }
  if (err != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(err));
    return LRG_E_NOMEM;
  }
  e->pool.push_back(PoolBlock{q, bytes, true});
  *p = reinterpret_cast<T*>(q);
  return LRG_OK;
}

static void pool_free(LrgEngine* e, void* p) {
  if (p == nullptr) return;
  for (auto& b : e->pool)
    if (b.p == p) { b.used = false; return; }
  cudaFree(p);     // not ours
}

static void channel_lists(int lite, std::vector<int>& conv, std::vector<int>& conv2) {
  if (lite == 1) { conv = {64, 64}; conv2 = {64}; }
  else if (lite == 2) { conv = {64, 64, 256}; conv2 = {64, 64}; }
  else { conv = {64, 64, 64, 128, 512}; conv2 = {256, 128}; }
}

static size_t count_weights(const LrgEngine* e) {
  size_t n = 0;
  for (int br = 0; br < 2; ++br)
    for (size_t i = 0; i < e->conv.size(); ++i) {
      int cin = i == 0 ? e->F : e->conv[i - 1];
      n += (size_t)cin * e->conv[i] + e->conv[i];
    }
  for (int h = 0; h < 2; ++h)
    for (size_t i = 0; i <= e->conv2.size(); ++i) {
      int cin = i == 0 ? e->conv.back() * 2 + e->conv[1] : e->conv2[i - 1];
      int cout = i == e->conv2.size() ? 2 : e->conv2[i];
      n += (size_t)cin * cout + cout;
    }
  return n;
}

static void free_forward_ws(LrgEngine* e) {
  for (int i = 0; i < 2; ++i) { cudaFree(e->d_x[i]); cudaFree(e->d_h1[i]); cudaFree(e->d_logits[i]); e->d_x[i] = e->d_h1[i] = e->d_logits[i] = nullptr; }
  cudaFree(e->d_pooled); cudaFree(e->d_gproj);
  e->d_pooled = e->d_gproj = nullptr;
  e->ws_batch = 0;
}

static int ensure_forward_ws(LrgEngine* e, int B) {
  if (B <= e->ws_batch) return LRG_OK;
  free_forward_ws(e);
  const int n[2] = {e->Ni, e->Nj};
  for (int i = 0; i < 2; ++i) {
    LRG_TRY(dev_alloc(&e->d_x[i], (size_t)B * n[i] * e->F));
    LRG_TRY(dev_alloc(&e->d_h1[i], (size_t)B * n[i] * e->net.C1));
    LRG_TRY(dev_alloc(&e->d_logits[i], (size_t)B * n[i] * 2));
  }
  LRG_TRY(dev_alloc(&e->d_pooled, (size_t)B * 2 * e->net.Clast));
  LRG_TRY(dev_alloc(&e->d_gproj, (size_t)B * 2 * e->net.H0));
  e->ws_batch = B;
  return LRG_OK;
}

static void free_rooms(LrgEngine* e) {
  pool_free(e, e->d_room_off); pool_free(e, e->d_pts); pool_free(e, e->d_pw); pool_free(e, e->d_pw_off); pool_free(e, e->d_room_vmin); pool_free(e, e->d_label);
  pool_free(e, e->d_sp_off); pool_free(e, e->d_sp_perm); pool_free(e, e->d_sp_vox); pool_free(e, e->d_sp_box);
  e->d_sp_off = nullptr; e->d_sp_perm = nullptr; e->d_sp_vox = nullptr; e->d_sp_box = nullptr;
  pool_free(e, e->d_label_filled); pool_free(e, e->d_order); pool_free(e, e->d_lab_list); pool_free(e, e->d_unl_list);
  pool_free(e, e->d_n_lab); pool_free(e, e->d_n_unl); pool_free(e, e->d_stats);
  pool_free(e, e->d_pw_lanes); pool_free(e, e->d_groups); pool_free(e, e->d_lane_steps); pool_free(e, e->d_parI);
  pool_free(e, e->d_spec_sync); pool_free(e, e->d_clog); pool_free(e, e->d_spec_est);
  pool_free(e, e->d_room_order); pool_free(e, e->d_pending_pts); e->d_room_order = nullptr; e->d_pending_pts = nullptr;
  e->d_pw_lanes = nullptr; e->d_groups = nullptr; e->d_lane_steps = nullptr; e->d_parI = nullptr; e->d_spec_sync = nullptr; e->d_clog = nullptr;
  e->d_spec_est = nullptr;
  pool_free(e, e->d_raw_off); pool_free(e, e->d_equalized_idx); pool_free(e, e->d_unequalized_idx); pool_free(e, e->d_feat);
  e->d_raw_off = nullptr; e->d_equalized_idx = e->d_unequalized_idx = nullptr; e->d_feat = nullptr; e->raw_mode = false; e->total_raw = 0;
  e->d_room_off = nullptr; e->d_pts = nullptr; e->d_pw = nullptr; e->d_pw_off = nullptr; e->d_room_vmin = nullptr; e->d_label = nullptr;
  e->d_label_filled = nullptr; e->d_order = nullptr; e->d_lab_list = e->d_unl_list = e->d_n_lab = e->d_n_unl = nullptr;
  e->d_stats = nullptr;
  e->n_rooms = 0; e->total_pts = 0; e->maxN = 0;
}

static void free_slots(LrgEngine* e) {
  cudaFree(e->d_slots); cudaFree(e->d_listI); cudaFree(e->d_listJ); cudaFree(e->d_keyI); cudaFree(e->d_keyJ);
  for (int i = 0; i < 2; ++i) {
    cudaFree(e->d_tile[i]); cudaFree(e->d_tileidx[i]); cudaFree(e->d_tilesrc[i]); cudaFree(e->s_h1[i]); cudaFree(e->s_logits[i]);
    e->d_tile[i] = nullptr; e->d_tileidx[i] = nullptr; e->d_tilesrc[i] = nullptr; e->s_h1[i] = nullptr; e->s_logits[i] = nullptr;
  }
  cudaFree(e->s_pooled); cudaFree(e->s_gproj); cudaFree(e->s_gproj_tagged); e->s_gproj_tagged = nullptr;
  e->d_slots = nullptr; e->d_listI = e->d_listJ = nullptr; e->d_keyI = e->d_keyJ = nullptr; e->s_pooled = e->s_gproj = nullptr;
  e->n_slots = 0; e->slots_maxN = 0;
}

static int ensure_slots(LrgEngine* e, int n_slots) {
  if (n_slots == e->n_slots && e->maxN <= e->slots_maxN) return LRG_OK;
  free_slots(e);
  const size_t S = n_slots, M = std::max(e->maxN, 1);
  LRG_TRY(dev_alloc(&e->d_slots, S));
  LRG_TRY(dev_alloc(&e->d_listI, S * M));
  LRG_TRY(dev_alloc(&e->d_listJ, S * M));
  LRG_TRY(dev_alloc(&e->d_keyI, S * M));
  LRG_TRY(dev_alloc(&e->d_keyJ, S * M));
  const int n[2] = {e->Ni, e->Nj};
  for (int i = 0; i < 2; ++i) {
    LRG_TRY(dev_alloc(&e->d_tile[i], S * n[i] * 16));     // slot tiles: rows padded to 16 floats
    LRG_TRY(dev_alloc(&e->d_tileidx[i], S * kMaxTilePts));
    LRG_TRY(dev_alloc(&e->d_tilesrc[i], S * kMaxTilePts));
    LRG_TRY(dev_alloc(&e->s_h1[i], S * n[i] * e->net.C1));
    LRG_TRY(dev_alloc(&e->s_logits[i], S * n[i] * 2));
  }
  LRG_TRY(dev_alloc(&e->s_pooled, S * 2 * e->net.Clast));
  LRG_TRY(dev_alloc(&e->s_gproj, S * 2 * e->net.H0));
  LRG_TRY(dev_alloc(&e->s_gproj_tagged, S * 2 * e->net.H0));
  e->n_slots = n_slots;
  e->slots_maxN = (int)M;
  return LRG_OK;
}

static bool use_tc(const LrgEngine* e) {
  return e->tc_available && e->forward_mode != LRG_FORWARD_FMA;
}
// Arithmetic of the tensor tiles: 3xFP16 unless 3xTF32 was asked for (or a call is being repeated after a range overflow).
static bool use_f16(const LrgEngine* e) {
  return use_tc(e) && e->forward_mode != LRG_FORWARD_TENSOR && !e->force_tf32;
}

// The LrgNet forward over fa.B tile pairs on `stream`: tensor-core kernels for the full model, fp32-FMA kernels otherwise.
static int run_forward(LrgEngine* e, const ForwardArgs& fa, cudaStream_t stream, cudaEvent_t* ev) {
  e->last_kind = use_tc(e) ? (use_f16(e) ? 2 : 1) : 0;
  if (use_tc(e)) return launch_forward_tc(e->tcnet, fa, use_f16(e), stream, ev);
  return launch_forward_timed(e->net, fa, stream, ev);
}
// After a synchronised 3xFP16 call: did an activation leave the fp16 range?  (Clears the flag.)
static int take_range_flag(LrgEngine* e, bool* overflow) {
  *overflow = false;
  if (e->d_range_flag == nullptr) return LRG_OK;
  int f = 0;
  LRG_CUDA(cudaMemcpy(&f, e->d_range_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (f != 0) {
    LRG_CUDA(cudaMemset(e->d_range_flag, 0, sizeof(int)));
    *overflow = true;
  }
  return LRG_OK;
}

// Operand image of W[k0:k0+Kc, n0:n0+Nc] (row-major [K][N] source) in the UMMA canonical no-swizzle K-major layout:
// element (n, k) at float offset (k/4)*(Nc*4) + n*4 + k%4; hi image followed by lo image.  Rows k >= K are zero.
static void pack_chunk(std::vector<float>& out, const float* W, int K, int N, int k0, int Kc, int n0, int Nc) {
  const size_t base = out.size();
  out.resize(base + (size_t)2 * Nc * Kc, 0.f);
  float* hi = out.data() + base;
  float* lo = hi + (size_t)Nc * Kc;
  for (int k = 0; k < Kc; ++k)
    for (int n = 0; n < Nc; ++n) {
      const float w = (k0 + k < K) ? W[(size_t)(k0 + k) * N + n0 + n] : 0.f;
      const size_t off = (size_t)(k / 4) * (Nc * 4) + (size_t)n * 4 + (k % 4);
      umma::split_tf32(w, hi[off], lo[off]);
    }
}

// The same chunk for the 3xFP16 tiles: W * scale split into fp16 hi + fp16 lo, element (n, k) at half offset
// (k/8)*(Nc*8) + n*8 + k%8 (core matrix = 8 rows x 8 halves = 128 bytes); hi image followed by lo image, appended as bytes.
static void pack_chunk_f16(std::vector<unsigned char>& out, const float* W, int K, int N, int k0, int Kc, int n0, int Nc, float scale) {
  const size_t base = out.size();
  out.resize(base + (size_t)2 * Nc * Kc * 2, 0);
  uint16_t* hi = reinterpret_cast<uint16_t*>(out.data() + base);
  uint16_t* lo = hi + (size_t)Nc * Kc;
  for (int k = 0; k < Kc; ++k)
    for (int n = 0; n < Nc; ++n) {
      const float w = (k0 + k < K) ? W[(size_t)(k0 + k) * N + n0 + n] * scale : 0.f;
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      const size_t off = (size_t)(k / 8) * (Nc * 8) + (size_t)n * 8 + (k % 8);
      memcpy(&hi[off], &h, 2);
      memcpy(&lo[off], &l, 2);
    }
}
// Power of two that brings the largest weight of a layer just below 2^15: the fp16 lo parts of all but negligible weights are
// then normal numbers, and hi never overflows.
static float f16_weight_scale(const float* W, size_t n) {
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) m = std::max(m, std::fabs(W[i]));
  if (!(m > 0.f) || !std::isfinite(m)) return 1.f;
  const int ex = (int)std::floor(std::log2(32768.0 / (double)m));
  return std::ldexp(1.f, std::max(-20, std::min(30, ex)));
}

}  // namespace lrg

extern "C" {
#pragma GCC visibility push(default)

const char* lrg_last_error(void) { return g_error.c_str(); }
int lrg_version(void) { return 100; }
int lrg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int lrg_engine_create(LrgEngine** out, int device, int feature_size, int num_inlier_points, int num_neighbor_points,
                      int lite, int max_batch) {
  LRG_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  LRG_REQUIRE(feature_size >= 3 && feature_size <= 16, "feature_size must be in [3,16], got %d", feature_size);
  LRG_REQUIRE(num_inlier_points > 0 && num_inlier_points <= kMaxTilePts && num_neighbor_points > 0 &&
                  num_neighbor_points <= kMaxTilePts,
              "num_inlier_points/num_neighbor_points must be in [1,%d]", kMaxTilePts);
  LRG_REQUIRE(lite >= 0 && lite <= 2, "lite must be 0, 1 or 2, got %d", lite);
  LRG_REQUIRE(max_batch >= 1, "max_batch must be >= 1");
  LRG_CUDA(cudaSetDevice(device));
  LrgEngine* e = new LrgEngine();
  e->device = device; e->F = feature_size; e->Ni = num_inlier_points; e->Nj = num_neighbor_points;
  e->lite = lite; e->max_batch = max_batch;
  channel_lists(lite, e->conv, e->conv2);
  e->n_weights = count_weights(e);
  cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->h_done, sizeof(int), cudaHostAllocMapped);
  if (err == cudaSuccess) err = cudaHostGetDevicePointer((void**)&e->d_done, e->h_done, 0);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->d_counters, 2 * sizeof(int));
  if (err == cudaSuccess) err = cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (err != cudaSuccess) {
    set_error("engine create: %s", cudaGetErrorString(err));
    delete e;
    return LRG_E_CUDA;
  }
  *out = e;
  return LRG_OK;
}

int lrg_engine_destroy(LrgEngine* e) {
  if (e == nullptr) return LRG_OK;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  free_forward_ws(e); free_rooms(e); free_slots(e);
  cudaFree(e->d_weights); cudaFree(e->d_tc_img); cudaFree(e->d_range_flag); cudaFree(e->d_counters); cudaFree(e->d_trace);
  cudaFree(e->d_qring); cudaFree(e->d_qctr); cudaFree(e->d_sync); cudaFree(e->d_busy); cudaFree(e->d_tile_dbg); cudaFree(e->d_remaining);
  for (auto& b : e->pool) cudaFree(b.p);
  e->pool.clear();
  cudaFreeHost(e->h_done);
  cudaStreamDestroy(e->stream);
  delete e;
  return LRG_OK;
}

size_t lrg_engine_weight_count(const LrgEngine* e) { return e ? e->n_weights : 0; }

int lrg_engine_set_forward_mode(LrgEngine* e, int mode) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  LRG_REQUIRE(mode == LRG_FORWARD_AUTO || mode == LRG_FORWARD_FMA || mode == LRG_FORWARD_TENSOR || mode == LRG_FORWARD_TENSOR_F16,
              "unknown forward mode %d", mode);
  if ((mode == LRG_FORWARD_TENSOR || mode == LRG_FORWARD_TENSOR_F16) && e->weights_loaded && !e->tc_available) {
    set_error("the tensor-core forward covers the full model only (lite=0, feature_size<=16)");
    return LRG_E_STATE;
  }
  e->forward_mode = mode;
  return LRG_OK;
}

int lrg_engine_forward_mode(const LrgEngine* e) {
  if (e == nullptr) return LRG_E_INVALID;
  return use_tc(e) ? (use_f16(e) ? LRG_FORWARD_TENSOR_F16 : LRG_FORWARD_TENSOR) : LRG_FORWARD_FMA;
}

int lrg_engine_range_overflow(LrgEngine* e, int* overflow, int* fallbacks) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  LRG_CUDA(cudaSetDevice(e->device));
  LRG_CUDA(cudaDeviceSynchronize());
  bool over = false;
  LRG_TRY(take_range_flag(e, &over));
  if (overflow) *overflow = over ? 1 : 0;
  if (fallbacks) *fallbacks = e->range_fallbacks;
  return LRG_OK;
}

int lrg_engine_load_weights(LrgEngine* e, const float* blob, size_t n_floats) {
  LRG_REQUIRE(e != nullptr && blob != nullptr, "engine/blob is NULL");
  LRG_REQUIRE(n_floats == e->n_weights, "weight blob has %zu floats, the network needs %zu", n_floats, e->n_weights);
  LRG_CUDA(cudaSetDevice(e->device));
  // repack: every matrix row-major [Kpad][N] (K padded to a multiple of 16 with zero rows); head layer 0 split into
  // its pooled part [2*Clast][H0] and its per-point part [C1][H0]
  std::vector<float> packed;
  auto add_matrix = [&](const float* src, int K, int N, int Kpad) {
    size_t off = packed.size();
    packed.resize(off + (size_t)Kpad * N, 0.f);
    memcpy(packed.data() + off, src, sizeof(float) * (size_t)K * N);
    while (packed.size() % 4) packed.push_back(0.f);
    return off;
  };
  auto add_vector = [&](const float* src, int N) {
    size_t off = packed.size();
    packed.insert(packed.end(), src, src + N);
    while (packed.size() % 4) packed.push_back(0.f);
    return off;
  };
  struct Off { size_t w, b; int K, Kpad, N; };
  Off conv[2][kMaxConv];
  Off head_local[2], hidden[2][kMaxHidden], outl[2];
  size_t w0g[2];
  const float* p = blob;
  const int nc = (int)e->conv.size(), nh = (int)e->conv2.size();
  for (int br = 0; br < 2; ++br)
    for (int i = 0; i < nc; ++i) {
      int K = i == 0 ? e->F : e->conv[i - 1], N = e->conv[i], Kpad = (K + 15) / 16 * 16;
      conv[br][i] = Off{add_matrix(p, K, N, Kpad), 0, K, Kpad, N};
      p += (size_t)K * N;
      conv[br][i].b = add_vector(p, N);
      p += N;
    }
  const int Clast = e->conv.back(), C1 = e->conv[1], H0 = e->conv2[0];
  // blob order: add head then remove head (util.py:138-162); device order: [0] = remove, [1] = add
  for (int hb = 0; hb < 2; ++hb) {
    const int h = hb == 0 ? 1 : 0;
    w0g[h] = add_matrix(p, 2 * Clast, H0, 2 * Clast);
    head_local[h] = Off{add_matrix(p + (size_t)2 * Clast * H0, C1, H0, C1), 0, C1, C1, H0};
    p += (size_t)(2 * Clast + C1) * H0;
    head_local[h].b = add_vector(p, H0);
    p += H0;
    for (int i = 1; i < nh; ++i) {
      int K = e->conv2[i - 1], N = e->conv2[i];
      hidden[h][i - 1] = Off{add_matrix(p, K, N, K), 0, K, K, N};
      p += (size_t)K * N;
      hidden[h][i - 1].b = add_vector(p, N);
      p += N;
    }
    int K = e->conv2.back();
    outl[h] = Off{add_matrix(p, K, 2, K), 0, K, K, 2};
    p += (size_t)K * 2;
    outl[h].b = add_vector(p, 2);
    p += 2;
  }
  if ((size_t)(p - blob) != n_floats) { set_error("internal: weight walk consumed %zu of %zu floats", (size_t)(p - blob), n_floats); return LRG_E_INVALID; }
  cudaFree(e->d_weights);
  e->d_weights = nullptr;
  LRG_TRY(dev_alloc(&e->d_weights, packed.size()));
  LRG_CUDA(cudaMemcpy(e->d_weights, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
  e->packed_floats = packed.size();
  NetDesc& net = e->net;
  memset(&net, 0, sizeof(net));
  net.F = e->F; net.n_conv = nc; net.n_hidden = nh; net.Clast = Clast; net.C1 = C1; net.H0 = H0;
  auto L = [&](const Off& o) { return LayerDesc{e->d_weights + o.w, e->d_weights + o.b, o.K, o.Kpad, o.N}; };
  for (int br = 0; br < 2; ++br)
    for (int i = 0; i < nc; ++i) net.conv[br][i] = L(conv[br][i]);
  for (int h = 0; h < 2; ++h) {
    net.W0g[h] = e->d_weights + w0g[h];
    net.head0_local[h] = L(head_local[h]);
    for (int i = 0; i + 1 < nh; ++i) net.hidden[h][i] = L(hidden[h][i]);
    net.out[h] = L(outl[h]);
  }
  LRG_TRY(forward_configure(net));
  // tensor-core path: pre-packed hi/lo operand images of the full model (lrg_forward_tc.cu)
  e->tc_available = false;
  if (e->lite == 0 && e->F <= 16) {
    std::vector<float> img;                                // 3xTF32 images
    std::vector<unsigned char> img16;                      // 3xFP16 images
    img.reserve((2 * kBranchImgBytesTf32 + 2 * kHeadImgBytesTf32) / 4);
    img16.reserve(2 * kBranchImgBytesF16 + 2 * kHeadImgBytesF16);
    size_t branch_off[2], head_off[2], branch_off16[2], head_off16[2];
    float branch_inv[2][5], head_inv[2][2];
    const float* q = blob;
    for (int br = 0; br < 2; ++br) {
      const float* Wl[kMaxConv];
      float sc[kMaxConv];
      for (int i = 0; i < nc; ++i) {
        const int K = i == 0 ? e->F : e->conv[i - 1];
        Wl[i] = q;
        sc[i] = f16_weight_scale(q, (size_t)K * e->conv[i]);
        branch_inv[br][i] = 1.f / sc[i];
        q += (size_t)K * e->conv[i] + e->conv[i];
      }
      branch_off[br] = img.size();
      pack_chunk(img, Wl[0], e->F, 64, 0, 16, 0, 64);
      pack_chunk(img, Wl[1], 64, 64, 0, 64, 0, 64);
      pack_chunk(img, Wl[2], 64, 64, 0, 64, 0, 64);
      pack_chunk(img, Wl[3], 64, 128, 0, 32, 0, 128);
      pack_chunk(img, Wl[3], 64, 128, 32, 32, 0, 128);
      for (int nb = 0; nb < 4; ++nb)
        for (int kc = 0; kc < 4; ++kc) pack_chunk(img, Wl[4], 128, 512, kc * 32, 32, nb * 128, 128);
      if ((img.size() - branch_off[br]) * 4 != kBranchImgBytesTf32) { set_error("internal: branch image size"); return LRG_E_INVALID; }
      branch_off16[br] = img16.size();
      pack_chunk_f16(img16, Wl[0], e->F, 64, 0, 16, 0, 64, sc[0]);
      pack_chunk_f16(img16, Wl[1], 64, 64, 0, 64, 0, 64, sc[1]);
      pack_chunk_f16(img16, Wl[2], 64, 64, 0, 64, 0, 64, sc[2]);
      pack_chunk_f16(img16, Wl[3], 64, 128, 0, 64, 0, 128, sc[3]);
      for (int nb = 0; nb < 4; ++nb)
        for (int kc = 0; kc < 2; ++kc) pack_chunk_f16(img16, Wl[4], 128, 512, kc * 64, 64, nb * 128, 128, sc[4]);
      if (img16.size() - branch_off16[br] != kBranchImgBytesF16) { set_error("internal: branch image size (fp16)"); return LRG_E_INVALID; }
    }
    for (int hb = 0; hb < 2; ++hb) {
      const int h = hb == 0 ? 1 : 0;                       // blob: add head first; device index 1 = add
      const float* K0local = q + (size_t)1024 * 256;       // rows 1024.. of kernel0: the per-point part [64][256]
      q += (size_t)1088 * 256 + 256;
      const float* K1 = q;                                 // [256][128]
      q += (size_t)256 * 128 + 128;
      q += 128 * 2 + 2;
      head_off[h] = img.size();
      auto W0 = [&](int nb) { pack_chunk(img, K0local, 64, 256, 0, 64, nb * 64, 64); };
      auto W1 = [&](int kc) {
        pack_chunk(img, K1, 256, 128, kc * 64, 32, 0, 128);
        pack_chunk(img, K1, 256, 128, kc * 64 + 32, 32, 0, 128);
      };
      W0(0); W0(1); W1(0); W0(2); W1(1); W0(3); W1(2); W1(3);   // the order lrg_tc_head_kernel consumes them in
      if ((img.size() - head_off[h]) * 4 != kHeadImgBytesTf32) { set_error("internal: head image size"); return LRG_E_INVALID; }
      const float s0 = f16_weight_scale(K0local, (size_t)64 * 256), s1 = f16_weight_scale(K1, (size_t)256 * 128);
      head_inv[h][0] = 1.f / s0; head_inv[h][1] = 1.f / s1;
      head_off16[h] = img16.size();
      auto W0h = [&](int nb) { pack_chunk_f16(img16, K0local, 64, 256, 0, 64, nb * 64, 64, s0); };
      auto W1h = [&](int kc) {
        pack_chunk_f16(img16, K1, 256, 128, kc * 64, 32, 0, 128, s1);
        pack_chunk_f16(img16, K1, 256, 128, kc * 64 + 32, 32, 0, 128, s1);
      };
      W0h(0); W0h(1); W1h(0); W0h(2); W1h(1); W0h(3); W1h(2); W1h(3);
      if (img16.size() - head_off16[h] != kHeadImgBytesF16) { set_error("internal: head image size (fp16)"); return LRG_E_INVALID; }
    }
    cudaFree(e->d_tc_img);
    e->d_tc_img = nullptr;
    const size_t tf32_bytes = img.size() * sizeof(float);
    LRG_TRY(dev_alloc(&e->d_tc_img, tf32_bytes + img16.size()));
    LRG_CUDA(cudaMemcpy(e->d_tc_img, img.data(), tf32_bytes, cudaMemcpyHostToDevice));
    LRG_CUDA(cudaMemcpy(e->d_tc_img + tf32_bytes, img16.data(), img16.size(), cudaMemcpyHostToDevice));
    if (e->d_range_flag == nullptr) LRG_TRY(dev_alloc(&e->d_range_flag, 1));
    LRG_CUDA(cudaMemset(e->d_range_flag, 0, sizeof(int)));
    TcNet& t = e->tcnet;
    memset(&t, 0, sizeof(t));
    t.F = e->F;
    t.range_flag = e->d_range_flag;
    for (int br = 0; br < 2; ++br) {
      t.branch_img[0][br] = e->d_tc_img + branch_off[br] * sizeof(float);
      t.branch_img[1][br] = e->d_tc_img + tf32_bytes + branch_off16[br];
      for (int i = 0; i < 5; ++i) { t.conv_bias[br][i] = net.conv[br][i].bias; t.branch_inv[br][i] = branch_inv[br][i]; }
    }
    for (int h = 0; h < 2; ++h) {
      t.W0g[h] = net.W0g[h];
      t.head_img[0][h] = e->d_tc_img + head_off[h] * sizeof(float);
      t.head_img[1][h] = e->d_tc_img + tf32_bytes + head_off16[h];
      t.head_inv[h][0] = head_inv[h][0]; t.head_inv[h][1] = head_inv[h][1];
      t.head_bias0[h] = net.head0_local[h].bias;
      t.head_bias1[h] = net.hidden[h][0].bias;
      t.head_W2[h] = net.out[h].W;
      t.head_bias2[h] = net.out[h].bias;
    }
    t.dbg = e->d_tile_dbg;
    LRG_TRY(tc_forward_configure());
    LRG_TRY(grow_configure());
    e->tc_available = true;
  }
  e->weights_loaded = true;
  return LRG_OK;
}

int lrg_forward_device(LrgEngine* e, int B, const float* d_inlier, const float* d_neighbor, float* d_add_out,
                       float* d_remove_out, lrg_stream_t stream) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  if (!e->weights_loaded) { set_error("lrg_forward: weights not loaded (call lrg_engine_load_weights first)"); return LRG_E_STATE; }
  LRG_REQUIRE(B >= 1, "batch must be >= 1");
  LRG_REQUIRE(d_inlier && d_neighbor && d_add_out && d_remove_out, "NULL tensor pointer");
  LRG_CUDA(cudaSetDevice(e->device));
  LRG_TRY(ensure_forward_ws(e, std::max(B, e->max_batch)));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  ForwardArgs fa{};
  fa.x[0] = d_inlier; fa.x[1] = d_neighbor; fa.x_stride = e->F;
  fa.n_pts[0] = e->Ni; fa.n_pts[1] = e->Nj;
  fa.h1[0] = e->d_h1[0]; fa.h1[1] = e->d_h1[1];
  fa.pooled = e->d_pooled; fa.gproj = e->d_gproj;
  fa.logits[0] = d_remove_out; fa.logits[1] = d_add_out;
  fa.active = nullptr; fa.active_stride = 0; fa.B = B;
  LRG_CUDA(cudaMemsetAsync(e->d_pooled, 0, sizeof(float) * (size_t)B * 2 * e->net.Clast, st));
  return run_forward(e, fa, st, nullptr);
}

int lrg_forward_host(LrgEngine* e, int B, const float* inlier, const float* neighbor, float* add_out, float* remove_out) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  if (!e->weights_loaded) { set_error("lrg_forward: weights not loaded (call lrg_engine_load_weights first)"); return LRG_E_STATE; }
  LRG_REQUIRE(B >= 1, "batch must be >= 1");
  LRG_REQUIRE(inlier && neighbor && add_out && remove_out, "NULL tensor pointer");
  LRG_CUDA(cudaSetDevice(e->device));
  LRG_TRY(ensure_forward_ws(e, std::max(B, e->max_batch)));
  LRG_CUDA(cudaMemcpyAsync(e->d_x[0], inlier, sizeof(float) * (size_t)B * e->Ni * e->F, cudaMemcpyHostToDevice, e->stream));
  LRG_CUDA(cudaMemcpyAsync(e->d_x[1], neighbor, sizeof(float) * (size_t)B * e->Nj * e->F, cudaMemcpyHostToDevice, e->stream));
  LRG_TRY(lrg_forward_device(e, B, e->d_x[0], e->d_x[1], e->d_logits[1], e->d_logits[0], e->stream));
  if (e->last_kind == 2) {
    // 3xFP16 tiles: an activation beyond the fp16 range voids the result -- repeat with 3xTF32 (or report it when 3xFP16 was asked for)
    LRG_CUDA(cudaStreamSynchronize(e->stream));
    bool over = false;
    LRG_TRY(take_range_flag(e, &over));
    if (over) {
      if (e->forward_mode == LRG_FORWARD_TENSOR_F16) { set_error("lrg_forward: an activation exceeds the fp16 range (use LRG_FORWARD_AUTO or LRG_FORWARD_TENSOR)"); return LRG_E_RANGE; }
      e->force_tf32 = true;
      e->range_fallbacks += 1;
      const int rc = lrg_forward_device(e, B, e->d_x[0], e->d_x[1], e->d_logits[1], e->d_logits[0], e->stream);
      e->force_tf32 = false;
      LRG_TRY(rc);
    }
  }
  LRG_CUDA(cudaMemcpyAsync(add_out, e->d_logits[1], sizeof(float) * (size_t)B * e->Nj * 2, cudaMemcpyDeviceToHost, e->stream));
  LRG_CUDA(cudaMemcpyAsync(remove_out, e->d_logits[0], sizeof(float) * (size_t)B * e->Ni * 2, cudaMemcpyDeviceToHost, e->stream));
  LRG_CUDA(cudaStreamSynchronize(e->stream));
  return LRG_OK;
}

// Allocates the per-room arrays for rooms of the given (equalised) sizes; frees whatever was uploaded before.
static int alloc_rooms(LrgEngine* e, int n_rooms, const int64_t* room_offsets, float resolution) {
  LRG_REQUIRE(room_offsets[0] == 0, "room_offsets[0] must be 0");
  long long total = room_offsets[n_rooms];
  int maxN = 0;
  for (int r = 0; r < n_rooms; ++r) {
    long long n = room_offsets[r + 1] - room_offsets[r];
    LRG_REQUIRE(n >= 0 && n < (1ll << 30), "room %d has an invalid point count", r);
    maxN = std::max(maxN, (int)n);
  }
  free_rooms(e);
  e->n_rooms = n_rooms; e->total_pts = total; e->maxN = maxN; e->resolution = resolution;
  e->h_room_off.assign(room_offsets, room_offsets + n_rooms + 1);
  const size_t T = (size_t)total;
  LRG_TRY(pool_alloc(e, &e->d_room_off, (size_t)n_rooms + 1));
  LRG_TRY(pool_alloc(e, &e->d_pts, T * 16));
  std::vector<long long> h_pw_off((size_t)n_rooms + 1, 0);
  for (int r = 0; r < n_rooms; ++r) h_pw_off[r + 1] = h_pw_off[r] + ((room_offsets[r + 1] - room_offsets[r] + 3) / 4) * 4;
  e->total_words = h_pw_off[n_rooms];
  LRG_TRY(pool_alloc(e, &e->d_pw, (size_t)e->total_words));
  LRG_TRY(pool_alloc(e, &e->d_pw_off, (size_t)n_rooms + 1));
  LRG_TRY(pool_alloc(e, &e->d_room_vmin, (size_t)std::max(n_rooms, 1)));
  LRG_CUDA(cudaMemcpy(e->d_pw_off, h_pw_off.data(), sizeof(long long) * (n_rooms + 1), cudaMemcpyHostToDevice));
  LRG_TRY(pool_alloc(e, &e->d_label, T));
  LRG_TRY(pool_alloc(e, &e->d_label_filled, T));
  LRG_TRY(pool_alloc(e, &e->d_order, T));
  LRG_TRY(pool_alloc(e, &e->d_lab_list, T));
  LRG_TRY(pool_alloc(e, &e->d_unl_list, T));
  LRG_TRY(pool_alloc(e, &e->d_n_lab, (size_t)n_rooms));
  LRG_TRY(pool_alloc(e, &e->d_n_unl, (size_t)n_rooms));
  LRG_TRY(pool_alloc(e, &e->d_stats, (size_t)n_rooms));
  LRG_CUDA(cudaMemcpy(e->d_room_off, e->h_room_off.data(), sizeof(long long) * (n_rooms + 1), cudaMemcpyHostToDevice));
  return LRG_OK;
}

// Dense (T, F) device feature rows -> padded rows + packed state words (e->d_order must already hold the seed order).
static int pack_rooms(LrgEngine* e, const float* d_dense, bool validate = false) {
  if (e->total_pts <= 0) return LRG_OK;
  LRG_CUDA(cudaMemsetAsync(e->d_counters, 0, 2 * sizeof(int), e->stream));
  LRG_TRY(launch_pack(d_dense, e->F, e->n_rooms, e->d_room_off, e->d_pw_off, e->resolution, e->d_pts, e->d_pw, e->d_room_vmin, e->d_counters, e->stream));
  e->prep_launches += 1;
  int bad_room = 0;
  LRG_CUDA(cudaMemcpyAsync(&bad_room, e->d_counters, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  LRG_CUDA(cudaStreamSynchronize(e->stream));
  LRG_REQUIRE(bad_room == 0, "room %d spans more than 1022 voxels along an axis at resolution %g (state words hold 10 bits per axis)", bad_room - 1, (double)e->resolution);
  {
    // spatial index: Morton order of every room's voxels + block boxes (the shell scans of the grow steps read only the blocks
    // that meet the shell)
    const int R = e->n_rooms;
    std::vector<long long> h_sp_off((size_t)R + 1, 0), h_key_off((size_t)R + 1, 0);
    long long max_keys = 0;
    for (int r = 0; r < R; ++r) {
      const long long n = e->h_room_off[r + 1] - e->h_room_off[r];
      long long P = 1;
      while (P < n) P <<= 1;
      h_sp_off[r + 1] = h_sp_off[r] + (n + kSpBlock - 1) / kSpBlock * kSpBlock;
      h_key_off[r + 1] = h_key_off[r] + (n > 0 ? std::max<long long>(P, 2) : 0);
      max_keys = std::max<long long>(max_keys, h_key_off[r + 1] - h_key_off[r]);
    }
    long long* d_key_off = nullptr;
    unsigned long long* d_keys = nullptr;
    LRG_TRY(pool_alloc(e, &e->d_sp_off, (size_t)R + 1));
    LRG_TRY(pool_alloc(e, &e->d_sp_perm, (size_t)h_sp_off[R]));
    LRG_TRY(pool_alloc(e, &e->d_sp_vox, (size_t)h_sp_off[R]));
    LRG_TRY(pool_alloc(e, &e->d_sp_box, (size_t)(h_sp_off[R] / kSpBlock)));
    LRG_TRY(pool_alloc(e, &d_key_off, (size_t)R + 1));
    int rc = pool_alloc(e, &d_keys, (size_t)h_key_off[R]);
    cudaError_t ce = cudaSuccess;
    if (rc == LRG_OK) {
      ce = cudaMemcpyAsync(e->d_sp_off, h_sp_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, e->stream);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_key_off, h_key_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, e->stream);
      if (ce == cudaSuccess)
        rc = launch_spatial_index(R, e->d_room_off, e->d_pw_off, e->d_pw, e->d_sp_off, d_key_off, d_keys, max_keys, e->d_sp_perm, e->d_sp_vox, e->d_sp_box, e->stream, &e->prep_launches);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);       // (the host vectors and the scratch go away)
    }
    pool_free(e, d_key_off);
    pool_free(e, d_keys);
    LRG_TRY(rc);
    LRG_CUDA(ce);
  }
  if (validate) {
    // caller-prepared features: the device applies the masks by point, the reference by voxel (test_region_grow.py:282-287)
    unsigned* d_scratch = nullptr;
    LRG_TRY(pool_alloc(e, &d_scratch, (size_t)3 * e->total_words));
    cudaError_t ce = cudaMemsetAsync(d_scratch, 0, sizeof(unsigned) * 3 * (size_t)e->total_words, e->stream);
    int rc = ce == cudaSuccess ? launch_validate_rooms(e->n_rooms, e->d_room_off, e->d_pw_off, e->d_pw, e->d_order, d_scratch, e->d_counters, e->stream) : LRG_E_CUDA;
    if (ce == cudaSuccess && rc == LRG_OK) ce = cudaMemcpyAsync(&bad_room, e->d_counters, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    pool_free(e, d_scratch);
    LRG_TRY(rc);
    LRG_CUDA(ce);
    if (bad_room != 0) {
      const int room = (bad_room & 0x0FFFFFFF) - 1;
      if ((bad_room >> 28) == 1) set_error("room %d: seed_order is not a permutation of 0..N-1", room);
      else set_error("room %d: two points share a voxel at resolution %g -- the driver needs one point per voxel (equalise like test_region_grow.py:125-136, or upload raw points with lrg_rooms_upload_raw)", room, (double)e->resolution);
      return LRG_E_INVALID;
    }
  }
  return LRG_OK;
}

int lrg_rooms_upload(LrgEngine* e, int n_rooms, const int64_t* room_offsets, const float* points, const int32_t* seed_order,
                     float resolution) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  LRG_REQUIRE(n_rooms >= 0 && room_offsets != nullptr, "bad room table");
  LRG_REQUIRE(resolution > 0.f, "resolution must be > 0");
  const long long total = room_offsets[n_rooms];
  LRG_REQUIRE(total == 0 || (points != nullptr && seed_order != nullptr), "NULL points/seed_order");
  LRG_CUDA(cudaSetDevice(e->device));
  LRG_TRY(alloc_rooms(e, n_rooms, room_offsets, resolution));
  e->prep_launches = 0;
  if (total > 0) {
    const size_t T = (size_t)total;
    float* d_raw = nullptr;     // staging for the dense (T, F) rows; freed after the pack kernel
    LRG_TRY(pool_alloc(e, &d_raw, T * e->F));
    cudaError_t ce = cudaMemcpyAsync(d_raw, points, sizeof(float) * T * e->F, cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->d_order, seed_order, sizeof(int) * T, cudaMemcpyHostToDevice, e->stream);
    int rc = ce == cudaSuccess ? pack_rooms(e, d_raw, true) : LRG_E_CUDA;
    if (ce != cudaSuccess) set_error("rooms upload -> %s", cudaGetErrorString(ce));
    cudaStreamSynchronize(e->stream);
    pool_free(e, d_raw);
    LRG_TRY(rc);
  }
  LRG_CUDA(cudaStreamSynchronize(e->stream));
  return LRG_OK;
}

static int upload_raw_impl(LrgEngine* e, int n_rooms, const int64_t* raw_offsets, const float* raw_points, int n_cols, float resolution,
                           bool device_src);

int lrg_rooms_upload_raw(LrgEngine* e, int n_rooms, const int64_t* raw_offsets, const float* raw_points, int n_cols, float resolution) {
  return upload_raw_impl(e, n_rooms, raw_offsets, raw_points, n_cols, resolution, false);
}

int lrg_rooms_upload_raw_device(LrgEngine* e, int n_rooms, const int64_t* raw_offsets, const float* d_raw_points, int n_cols, float resolution) {
  return upload_raw_impl(e, n_rooms, raw_offsets, d_raw_points, n_cols, resolution, true);
}

int lrg_last_prepare_launches(LrgEngine* e, int* n) {
  LRG_REQUIRE(e != nullptr && n != nullptr, "NULL argument");
  *n = e->prep_launches;
  return LRG_OK;
}

int lrg_last_prepare_ms(LrgEngine* e, float* ms) {
  LRG_REQUIRE(e != nullptr && ms != nullptr, "NULL argument");
  *ms = e->prep_ms;
  return LRG_OK;
}

static int upload_raw_impl(LrgEngine* e, int n_rooms, const int64_t* raw_offsets, const float* raw_points, int n_cols, float resolution,
                           bool device_src) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  LRG_REQUIRE(n_rooms >= 0 && raw_offsets != nullptr && raw_offsets[0] == 0, "bad room table");
  LRG_REQUIRE(resolution > 0.f, "resolution must be > 0");
  LRG_REQUIRE(n_cols >= 6, "raw points need at least x y z r g b (got %d columns)", n_cols);
  LRG_REQUIRE(e->F == 6 || e->F == 9 || e->F == 12 || e->F == 13, "feature_size %d has no definition from raw points (6, 9, 12 or 13)", e->F);
  const long long total_raw = raw_offsets[n_rooms];
  LRG_REQUIRE(total_raw == 0 || raw_points != nullptr, "NULL raw_points");
  LRG_CUDA(cudaSetDevice(e->device));
  std::vector<long long> sort_off((size_t)n_rooms + 1, 0);
  for (int r = 0; r < n_rooms; ++r) {
    const long long n = raw_offsets[r + 1] - raw_offsets[r];
    LRG_REQUIRE(n >= 0 && n < (1ll << 20), "room %d has %lld raw points (limit 1,048,575)", r, n);
    long long P = 2;
    while (P < n) P <<= 1;
    sort_off[r + 1] = sort_off[r] + P;
  }
  cudaStream_t st = e->stream;
  // workspace (freed before returning)
  float* d_raw = nullptr; long long *d_raw_off = nullptr, *d_sort_off = nullptr, *d_eq_off = nullptr;
  unsigned long long *d_keys = nullptr, *d_keys2 = nullptr; int4* d_vmin = nullptr; int *d_neq = nullptr, *d_err = nullptr;
  unsigned* d_uvox = nullptr; int *d_ustart = nullptr, *d_equ = nullptr, *d_rank = nullptr; double *d_sums = nullptr, *d_curv = nullptr;
  float* d_extent = nullptr; unsigned long long* d_cmax = nullptr; int* d_has_nan = nullptr;
  auto cleanup = [&]() {
    pool_free(e, d_raw); pool_free(e, d_sort_off); pool_free(e, d_eq_off); pool_free(e, d_keys); pool_free(e, d_keys2); pool_free(e, d_vmin); pool_free(e, d_neq);
    pool_free(e, d_extent); pool_free(e, d_cmax); pool_free(e, d_has_nan);
    pool_free(e, d_err); pool_free(e, d_uvox); pool_free(e, d_ustart); pool_free(e, d_equ); pool_free(e, d_rank); pool_free(e, d_sums); pool_free(e, d_curv);
  };
  const size_t TR = (size_t)total_raw, TS = (size_t)sort_off[n_rooms];
  int rc = LRG_OK;
  auto A = [&](int r) { if (rc == LRG_OK) rc = r; };
  if (!device_src) A(pool_alloc(e, &d_raw, TR * n_cols));
  A(pool_alloc(e, &d_raw_off, (size_t)n_rooms + 1)); A(pool_alloc(e, &d_sort_off, (size_t)n_rooms + 1));
  A(pool_alloc(e, &d_eq_off, (size_t)n_rooms + 1)); A(pool_alloc(e, &d_keys, TS)); A(pool_alloc(e, &d_keys2, TS)); A(pool_alloc(e, &d_vmin, (size_t)std::max(n_rooms, 1)));
  A(pool_alloc(e, &d_neq, (size_t)std::max(n_rooms, 1))); A(pool_alloc(e, &d_err, 1)); A(pool_alloc(e, &d_uvox, TR)); A(pool_alloc(e, &d_ustart, TR));
  A(pool_alloc(e, &d_equ, TR)); A(pool_alloc(e, &d_rank, TR)); A(pool_alloc(e, &d_sums, TR * 10));
  A(pool_alloc(e, &d_extent, (size_t)std::max(n_rooms, 1) * 6)); A(pool_alloc(e, &d_cmax, (size_t)std::max(n_rooms, 1)));
  A(pool_alloc(e, &d_has_nan, (size_t)std::max(n_rooms, 1)));
  if (rc != LRG_OK) { cleanup(); pool_free(e, d_raw_off); return rc; }
  std::vector<long long> h_raw_off(raw_offsets, raw_offsets + n_rooms + 1);
  cudaEvent_t pev0 = nullptr, pev1 = nullptr;      // device time of the preparation (includes the host round trip for the room sizes)
  cudaEventCreate(&pev0); cudaEventCreate(&pev1);
  cudaEventRecord(pev0, st);
  if (!device_src) cudaMemcpyAsync(d_raw, raw_points, sizeof(float) * TR * n_cols, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_raw_off, h_raw_off.data(), sizeof(long long) * (n_rooms + 1), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_sort_off, sort_off.data(), sizeof(long long) * (n_rooms + 1), cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(d_err, 0, sizeof(int), st);
  FeatPrepArgs fp{};
  fp.n_rooms = n_rooms; fp.C = n_cols; fp.F = e->F; fp.res = resolution;
  fp.raw_off = d_raw_off; fp.raw = device_src ? raw_points : d_raw; fp.sort_off = d_sort_off; fp.keys = d_keys; fp.keys2 = d_keys2; fp.raw_vmin = d_vmin;
  fp.n_eq = d_neq; fp.err = d_err; fp.uniq_vox = d_uvox; fp.uniq_start = d_ustart; fp.eq_of_uniq = d_equ; fp.sums = d_sums; fp.raw_rank = d_rank;
  for (int r = 0; r < n_rooms; ++r) fp.max_sort = std::max<long long>(fp.max_sort, sort_off[r + 1] - sort_off[r]);
  e->prep_launches = 0;
  rc = launch_featprep_phase1(fp, st, &e->prep_launches);
  std::vector<int> h_neq(std::max(n_rooms, 1), 0);
  int bad_room = 0;
  if (rc == LRG_OK) {
    cudaMemcpyAsync(h_neq.data(), d_neq, sizeof(int) * n_rooms, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&bad_room, d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { set_error("feature preparation -> %s", cudaGetErrorString(ce)); rc = LRG_E_CUDA; }
  }
  if (rc == LRG_OK && bad_room != 0) {
    set_error("room %d spans more than 1022 voxels along an axis at resolution %g (state words hold 10 bits per axis)", bad_room - 1, (double)resolution);
    rc = LRG_E_INVALID;
  }
  if (rc != LRG_OK) { cleanup(); pool_free(e, d_raw_off); cudaEventDestroy(pev0); cudaEventDestroy(pev1); return rc; }
  std::vector<int64_t> eq_off((size_t)n_rooms + 1, 0);
  for (int r = 0; r < n_rooms; ++r) eq_off[r + 1] = eq_off[r] + h_neq[r];
  rc = alloc_rooms(e, n_rooms, eq_off.data(), resolution);
  const size_t TE = (size_t)eq_off[n_rooms];
  if (rc == LRG_OK) rc = pool_alloc(e, &d_curv, TE);
  if (rc == LRG_OK) rc = pool_alloc(e, &e->d_feat, TE * e->F);
  if (rc == LRG_OK) rc = pool_alloc(e, &e->d_equalized_idx, TE);
  if (rc == LRG_OK) rc = pool_alloc(e, &e->d_unequalized_idx, TR);
  if (rc == LRG_OK) {
    std::vector<long long> h_eq(eq_off.begin(), eq_off.end());
    cudaMemcpyAsync(d_eq_off, h_eq.data(), sizeof(long long) * (n_rooms + 1), cudaMemcpyHostToDevice, st);
    fp.eq_off = d_eq_off; fp.feat = e->d_feat; fp.curv = d_curv; fp.order = e->d_order; fp.equalized_idx = e->d_equalized_idx;
    fp.unequalized_idx = e->d_unequalized_idx;
    fp.extent = d_extent; fp.cmax = d_cmax; fp.has_nan = d_has_nan;
    for (int r = 0; r < n_rooms; ++r) {
      long long pe = 2;
      while (pe < h_neq[r]) pe <<= 1;
      fp.max_order_sort = std::max<long long>(fp.max_order_sort, std::min<long long>(pe, sort_off[r + 1] - sort_off[r]));
    }
    rc = launch_featprep_phase2(fp, st, &e->prep_launches);
    if (rc == LRG_OK) rc = pack_rooms(e, e->d_feat);
    cudaEventRecord(pev1, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess) cudaEventElapsedTime(&e->prep_ms, pev0, pev1);
    if (rc == LRG_OK && ce != cudaSuccess) { set_error("feature preparation -> %s", cudaGetErrorString(ce)); rc = LRG_E_CUDA; }
  }
  cleanup();
  cudaEventDestroy(pev0); cudaEventDestroy(pev1);
  if (rc != LRG_OK) { pool_free(e, d_raw_off); return rc; }
  e->raw_mode = true; e->total_raw = total_raw; e->h_raw_off = h_raw_off; e->d_raw_off = d_raw_off;
  return LRG_OK;
}

int lrg_rooms_equalized_offsets(LrgEngine* e, int64_t* eq_offsets) {
  LRG_REQUIRE(e != nullptr && eq_offsets != nullptr, "NULL argument");
  for (int r = 0; r <= e->n_rooms; ++r) eq_offsets[r] = e->h_room_off.empty() ? 0 : e->h_room_off[r];
  return LRG_OK;
}

int lrg_rooms_features_download(LrgEngine* e, float* points, int32_t* seed_order, int32_t* equalized_idx, int32_t* unequalized_idx) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  if (!e->raw_mode) { set_error("rooms were not uploaded as raw points"); return LRG_E_STATE; }
  LRG_CUDA(cudaSetDevice(e->device));
  const size_t TE = (size_t)e->total_pts, TR = (size_t)e->total_raw;
  if (points && TE) LRG_CUDA(cudaMemcpy(points, e->d_feat, sizeof(float) * TE * e->F, cudaMemcpyDeviceToHost));
  if (seed_order && TE) LRG_CUDA(cudaMemcpy(seed_order, e->d_order, sizeof(int) * TE, cudaMemcpyDeviceToHost));
  if (equalized_idx && TE) LRG_CUDA(cudaMemcpy(equalized_idx, e->d_equalized_idx, sizeof(int) * TE, cudaMemcpyDeviceToHost));
  if (unequalized_idx && TR) LRG_CUDA(cudaMemcpy(unequalized_idx, e->d_unequalized_idx, sizeof(int) * TR, cudaMemcpyDeviceToHost));
  return LRG_OK;
}

int lrg_labels_download_raw(LrgEngine* e, int32_t* labels_raw, int filled) {
  LRG_REQUIRE(e != nullptr && (labels_raw != nullptr || e->total_raw == 0), "engine/labels is NULL");
  if (!e->raw_mode) { set_error("rooms were not uploaded as raw points"); return LRG_E_STATE; }
  LRG_CUDA(cudaSetDevice(e->device));
  if (e->total_raw <= 0) return LRG_OK;
  int* d_out = nullptr;
  LRG_TRY(pool_alloc(e, &d_out, (size_t)e->total_raw));
  int rc = launch_labels_raw(e->n_rooms, e->d_raw_off, e->d_room_off, e->d_unequalized_idx, filled ? e->d_label_filled : e->d_label, d_out, e->stream);
  cudaError_t ce = cudaSuccess;
  if (rc == LRG_OK) ce = cudaMemcpyAsync(labels_raw, d_out, sizeof(int) * (size_t)e->total_raw, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  pool_free(e, d_out);
  LRG_TRY(rc);
  LRG_CUDA(ce);
  return LRG_OK;
}

namespace {
// CUDA events that are destroyed on every return path.
struct EventSet {
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  int create() {
    for (auto& x : ev) LRG_CUDA(cudaEventCreate(&x));
    return LRG_OK;
  }
  ~EventSet() {
    for (auto x : ev)
      if (x) cudaEventDestroy(x);
  }
};
}  // namespace

static int segment_resident_impl(LrgEngine* e, const LrgGrowParams* params, LrgRoomStats* stats);

int lrg_segment_resident(LrgEngine* e, const LrgGrowParams* params, LrgRoomStats* stats) {
  LRG_REQUIRE(e != nullptr && params != nullptr, "engine/params is NULL");
  if (!e->weights_loaded) { set_error("lrg_segment: weights not loaded"); return LRG_E_STATE; }
  LRG_REQUIRE(params->resolution == 0.f || params->resolution == e->resolution,
              "params->resolution %g differs from the resolution the rooms were uploaded with (%g)", params->resolution, e->resolution);
  // one call at a time per engine handle (the slot arrays, queues and the stream are the engine's): fail fast instead of racing
  if (e->busy.exchange(true)) { set_error("lrg_segment_resident: the engine handle is in use by another call (one call at a time per handle)"); return LRG_E_STATE; }
  struct Release { LrgEngine* e; ~Release() { e->busy.store(false); } } release{e};
  int rc = segment_resident_impl(e, params, stats);
  if (rc == LRG_OK && e->last_kind == 2) {
    // 3xFP16 tiles: an activation beyond the fp16 range voids the run -- repeat it with 3xTF32 (or report it when 3xFP16 was asked for)
    bool over = false;
    LRG_TRY(take_range_flag(e, &over));
    if (over) {
      if (e->forward_mode == LRG_FORWARD_TENSOR_F16) { set_error("lrg_segment_resident: an activation exceeds the fp16 range (use LRG_FORWARD_AUTO or LRG_FORWARD_TENSOR)"); return LRG_E_RANGE; }
      e->force_tf32 = true;
      e->range_fallbacks += 1;
      rc = segment_resident_impl(e, params, stats);
      e->force_tf32 = false;
    }
  }
  return rc;
}

static int segment_resident_impl(LrgEngine* e, const LrgGrowParams* params, LrgRoomStats* stats) {
  LRG_CUDA(cudaSetDevice(e->device));
  const int n_rooms = e->n_rooms;
  const size_t T = (size_t)e->total_pts;
  // random restarts (test_random_restart.py): `lanes` consecutive slots grow the restarts of one seed side by side
  // beam search (test_beam_search.py): the beam_width x search_width expansions of a round are the lanes of a group
  const bool beam = params->beam_width > 0 || params->search_width > 0;
  if (beam) {
    LRG_REQUIRE(params->beam_width > 0 && params->search_width > 0, "beam_width %d and search_width %d must both be positive",
                params->beam_width, params->search_width);
    LRG_REQUIRE(params->num_restarts <= 1, "beam search and random restarts (num_restarts %d) exclude each other", params->num_restarts);
    LRG_REQUIRE((long long)params->beam_width * params->search_width <= kMaxLanes, "beam_width %d x search_width %d exceeds the limit of %d lanes",
                params->beam_width, params->search_width, kMaxLanes);
  }
  LRG_REQUIRE(beam || !(params->flags & LRG_FLAG_SCORE_ML), "LRG_FLAG_SCORE_ML ('ml' scoring) needs the beam-search driver (beam_width, search_width > 0): "
              "the reference's restart driver is broken with it (test_random_restart.py:196)");
  // speculative lanes: up to spec_lanes regions of one room side by side, committed in seed order (plain driver only)
  LRG_REQUIRE(params->spec_lanes >= 0 && params->spec_lanes <= kMaxLanes, "spec_lanes %d out of range [0,%d]", params->spec_lanes, kMaxLanes);
  // (0 = engine default: 4 lanes for the plain driver in the persistent kernel -- same labels, shorter chains; a traced run keeps
  // one lane so that lrg_trace_download sees one step sequence per room)
  const bool plain = !beam && params->num_restarts <= 1;
  const bool persistent_path = use_tc(e) && !(params->flags & (LRG_FLAG_KERNEL_TIMING | LRG_FLAG_NO_GRAPH | LRG_FLAG_LOCKSTEP));
  // (default: 4 lanes; 8 when the rooms are large -- a 177 k-point outdoor scene holds far more regions that do not touch one
  // another than a 12 k-point room: 19 scenes 1.34 -> 1.09 s with 8 lanes, while 8 lanes on rooms only add discarded steps)
  const int dflt_lanes = (e->n_rooms > 0 && e->total_pts / e->n_rooms >= 65536) ? 8 : 4;
  const int spec_lanes = params->spec_lanes != 0 ? params->spec_lanes : (plain && persistent_path && params->trace_capacity == 0) ? dflt_lanes : 1;
  const bool spec = spec_lanes > 1 && plain;
  LRG_REQUIRE(params->spec_lanes <= 1 || spec, "spec_lanes %d needs the plain driver (no restarts, no beam search)", params->spec_lanes);
  const int lanes = beam ? params->beam_width * params->search_width : params->num_restarts > 1 ? params->num_restarts : spec ? spec_lanes : 1;
  const bool grouped = lanes > 1 || beam;     // slots form groups with a LaneGroup record (a 1 x 1 beam is a group of one lane)
  LRG_REQUIRE(lanes <= kMaxLanes, "num_restarts %d exceeds the limit of %d", lanes, kMaxLanes);
  int n_slots = params->max_slots > 0 ? std::max(params->max_slots, lanes) : spec ? 148 * lanes : (lanes > 1 ? 296 : 148);
  int n_groups = std::max(1, std::min(n_slots / lanes, std::max(n_rooms, 1)));
  n_slots = n_groups * lanes;
  LRG_TRY(ensure_slots(e, n_slots));
  e->last_lanes = lanes;
  if (grouped) {
    pool_free(e, e->d_pw_lanes); pool_free(e, e->d_groups); pool_free(e, e->d_lane_steps); pool_free(e, e->d_parI);
    pool_free(e, e->d_spec_sync); pool_free(e, e->d_clog);
    e->d_pw_lanes = nullptr; e->d_groups = nullptr; e->d_lane_steps = nullptr; e->d_parI = nullptr; e->d_spec_sync = nullptr; e->d_clog = nullptr;
    if (spec) {
      pool_free(e, e->d_spec_est); e->d_spec_est = nullptr;
      LRG_TRY(pool_alloc(e, &e->d_spec_est, (size_t)2 * n_groups));       // [0,n): work estimates, [n,2n): the room is critical
      LRG_TRY(pool_alloc(e, &e->d_spec_sync, (size_t)n_groups));
      LRG_TRY(pool_alloc(e, &e->d_clog, (size_t)n_groups * std::max(e->slots_maxN, 1)));
    }
    LRG_TRY(pool_alloc(e, &e->d_pw_lanes, (size_t)std::max<long long>(e->total_words, 4) * lanes));
    LRG_TRY(pool_alloc(e, &e->d_groups, (size_t)n_groups));
    LRG_TRY(pool_alloc(e, &e->d_lane_steps, (size_t)std::max(n_rooms, 1) * lanes));
    if (beam) LRG_TRY(pool_alloc(e, &e->d_parI, (size_t)n_groups * params->beam_width * std::max(e->slots_maxN, 1)));
  }
  cudaStream_t st = e->stream;
  if (params->trace_capacity > 0 && (e->d_trace == nullptr || e->trace_capacity != params->trace_capacity || e->trace_rooms < n_rooms * lanes)) {
    cudaFree(e->d_trace);
    e->d_trace = nullptr;
    LRG_TRY(dev_alloc(&e->d_trace, (size_t)std::max(n_rooms, 1) * lanes * params->trace_capacity));
    e->trace_rooms = n_rooms * lanes;
  }
  *e->h_done = 0;
  EventSet evs;
  LRG_TRY(evs.create());
  const cudaEvent_t ev0 = evs.ev[0], ev1 = evs.ev[1], ev2 = evs.ev[2];
  LRG_CUDA(cudaEventRecord(ev0, st));
  // reset per-run state (inside the timed region: it is part of one pass over the rooms)
  unsigned* d_words = e->d_pw;
  if (grouped) {
    d_words = e->d_pw_lanes;
    if (e->total_words > 0) LRG_CUDA(cudaMemcpyAsync(d_words, e->d_pw, sizeof(unsigned) * (size_t)e->total_words, cudaMemcpyDeviceToDevice, st));
    std::vector<LaneGroup> ginit(n_groups);
    memset(ginit.data(), 0, sizeof(LaneGroup) * n_groups);
    for (auto& g : ginit) { g.room = -1; g.cluster_id = 1; }
    LRG_CUDA(cudaMemcpyAsync(e->d_groups, ginit.data(), sizeof(LaneGroup) * n_groups, cudaMemcpyHostToDevice, st));
    LRG_CUDA(cudaMemsetAsync(e->d_lane_steps, 0, sizeof(int) * (size_t)std::max(n_rooms, 1) * lanes, st));
    if (spec) {
      LRG_CUDA(cudaMemsetAsync(e->d_spec_sync, 0, sizeof(SpecSync) * (size_t)n_groups, st));
      LRG_CUDA(cudaMemsetAsync(e->d_spec_est, 0, sizeof(int) * (size_t)2 * n_groups, st));
    }
    LRG_CUDA(cudaStreamSynchronize(st));      // (ginit is a host temporary)
  }
  LRG_TRY(launch_reset_words(n_rooms, e->d_room_off, e->d_pw_off, d_words, lanes, e->total_words, st));
  LRG_CUDA(cudaMemsetAsync(e->d_label, 0, T * sizeof(int), st));
  LRG_CUDA(cudaMemsetAsync(e->d_stats, 0, sizeof(LrgRoomStats) * std::max(n_rooms, 1), st));
  LRG_CUDA(cudaMemsetAsync(e->d_counters, 0, 2 * sizeof(int), st));
  std::vector<SlotState> init(n_slots);
  memset(init.data(), 0, sizeof(SlotState) * n_slots);
  for (auto& s : init) s.room = -1;
  if (grouped)
    for (int s = 0; s < n_slots; ++s) init[s].parked = (s % lanes) != 0;      // lane 0 of every group takes the first room
  LRG_CUDA(cudaMemcpyAsync(e->d_slots, init.data(), sizeof(SlotState) * n_slots, cudaMemcpyHostToDevice, st));
  if (params->trace_capacity > 0) {
    LRG_CUDA(cudaMemsetAsync(e->d_trace, 0, sizeof(LrgStepTrace) * (size_t)std::max(n_rooms, 1) * lanes * params->trace_capacity, st));
    e->trace_capacity = params->trace_capacity;
  } else {
    e->trace_capacity = 0;
  }

  DriverArgs da{};
  da.n_rooms = n_rooms; da.room_off = e->d_room_off; da.pts = e->d_pts; da.pw = d_words; da.pw_off = e->d_pw_off; da.room_vmin = e->d_room_vmin;
  da.tune_step = (params->flags & LRG_FLAG_NO_STEP_OVERLAP) ? 1 : 0;
  const bool no_sp = (params->flags & LRG_FLAG_NO_SPATIAL_INDEX) != 0;
  da.sp_off = no_sp ? nullptr : e->d_sp_off; da.sp_perm = no_sp ? nullptr : e->d_sp_perm; da.sp_vox = no_sp ? nullptr : e->d_sp_vox; da.sp_box = no_sp ? nullptr : e->d_sp_box;
  da.label = e->d_label; da.order = e->d_order; da.slots = e->d_slots; da.n_slots = n_slots; da.maxN = e->slots_maxN;
  da.listI = e->d_listI; da.listJ = e->d_listJ; da.keyI = e->d_keyI; da.keyJ = e->d_keyJ;
  da.tile[0] = e->d_tile[0]; da.tile[1] = e->d_tile[1]; da.tileidx[0] = e->d_tileidx[0]; da.tileidx[1] = e->d_tileidx[1];
  da.tilesrc[0] = e->d_tilesrc[0]; da.tilesrc[1] = e->d_tilesrc[1];
  da.logits[0] = e->s_logits[0]; da.logits[1] = e->s_logits[1];
  da.pooled = e->s_pooled; da.pooled_per_slot = 2 * e->net.Clast;
  da.Ni = e->Ni; da.Nj = e->Nj; da.F = e->F;
  da.resolution = e->resolution; da.cluster_threshold = params->cluster_threshold; da.seed = params->seed;
  da.max_steps = params->max_steps_per_region; da.room_id_base = params->room_id_base;
  da.next_room = e->d_counters; da.finished_slots = e->d_counters + 1; da.done_flag = e->d_done;
  if (n_rooms > 1 && !(params->flags & LRG_FLAG_ROOMS_IN_ORDER)) {
    // start the largest rooms first (more points = more grow steps, roughly): the tail of the run is then made of short rooms
    std::vector<int> ord((size_t)n_rooms);
    for (int r = 0; r < n_rooms; ++r) ord[r] = r;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
      return e->h_room_off[a + 1] - e->h_room_off[a] > e->h_room_off[b + 1] - e->h_room_off[b];
    });
    std::vector<long long> pend((size_t)n_rooms + 1, 0);
    for (int k = n_rooms - 1; k >= 0; --k) pend[k] = pend[k + 1] + (e->h_room_off[ord[k] + 1] - e->h_room_off[ord[k]]);
    pool_free(e, e->d_room_order); pool_free(e, e->d_pending_pts);
    e->d_room_order = nullptr; e->d_pending_pts = nullptr;
    LRG_TRY(pool_alloc(e, &e->d_room_order, (size_t)n_rooms));
    LRG_TRY(pool_alloc(e, &e->d_pending_pts, (size_t)n_rooms + 1));
    LRG_CUDA(cudaMemcpyAsync(e->d_room_order, ord.data(), sizeof(int) * n_rooms, cudaMemcpyHostToDevice, st));
    LRG_CUDA(cudaMemcpyAsync(e->d_pending_pts, pend.data(), sizeof(long long) * (n_rooms + 1), cudaMemcpyHostToDevice, st));
    LRG_CUDA(cudaStreamSynchronize(st));      // (host temporaries)
    da.room_order = e->d_room_order; da.pending_pts = e->d_pending_pts;
  }
  da.dbg = e->d_tile_dbg ? e->d_tile_dbg + 32 : nullptr;
  da.stats = e->d_stats; da.trace = e->trace_capacity > 0 ? e->d_trace : nullptr; da.trace_capacity = e->trace_capacity;
  da.lanes = lanes; da.groups = e->d_groups; da.pw_lane_stride = e->total_words; da.lane_steps = grouped ? e->d_lane_steps : nullptr;
  da.score_ml = (beam && (params->flags & LRG_FLAG_SCORE_ML)) ? 1 : 0;
  da.beam_width = beam ? params->beam_width : 0; da.search_width = beam ? params->search_width : 0; da.parI = beam ? e->d_parI : nullptr;
  da.spec = spec ? 1 : 0; da.spec_sync = e->d_spec_sync; da.clog = e->d_clog; da.q_ctr = nullptr;
  // which rooms speculate: the spec_top rooms with the most estimated work left, and anybody while CTAs idle (the tail)
  da.spec_top = params->spec_top == 0 ? 8 : params->spec_top < 0 ? (1 << 30) : params->spec_top;
  da.spec_min_idle = params->spec_min_idle == 0 ? 96 : params->spec_min_idle < 0 ? (1 << 30) : params->spec_min_idle;
  da.spec_crit = params->spec_crit == 0 ? 40 : params->spec_crit < 0 ? 0 : params->spec_crit;
  da.total_pts = e->total_pts;
  da.spec_est = e->d_spec_est;

  ForwardArgs fa{};
  fa.x[0] = e->d_tile[0]; fa.x[1] = e->d_tile[1]; fa.x_stride = 16;
  fa.n_pts[0] = e->Ni; fa.n_pts[1] = e->Nj;
  fa.h1[0] = e->s_h1[0]; fa.h1[1] = e->s_h1[1];
  fa.pooled = e->s_pooled; fa.gproj = e->s_gproj;
  fa.logits[0] = e->s_logits[0]; fa.logits[1] = e->s_logits[1];
  fa.active = &e->d_slots[0].active; fa.active_stride = (int)(sizeof(SlotState) / sizeof(int)); fa.B = n_slots;
  fa.n_valid = &e->d_slots[0].n_in; fa.n_valid_stride = fa.active_stride;   // n_in, n_nb are adjacent in SlotState

  e->iterations = 0; e->launches = 0; e->forward_ms = 0;
  for (int i = 0; i < 4; ++i) e->kernel_ms[i] = 0;
  const bool kernel_timing = (params->flags & LRG_FLAG_KERNEL_TIMING) != 0;
  const bool use_graph = !kernel_timing && !(params->flags & LRG_FLAG_NO_GRAPH);
  const bool persistent = use_tc(e) && use_graph && !(params->flags & LRG_FLAG_LOCKSTEP) && n_slots <= kMaxGrowSlots;
  e->last_persistent = persistent && n_rooms > 0;
  int rc = LRG_OK;
  if (n_rooms > 0) {
    if (persistent) {
      // one launch for the whole run: every slot starts with a STEP item, the device queue does the rest
      unsigned cap = 1024;
      while (cap < (unsigned)n_slots * 16u + 1024u) cap <<= 1;
      if (cap > e->q_capacity) {
        cudaFree(e->d_qring);
        e->d_qring = nullptr;
        e->q_capacity = 0;
        LRG_TRY(dev_alloc(&e->d_qring, (size_t)3 * cap));     // three rings: reserved CTAs, everybody, projection requests
        e->q_capacity = cap;
      }
      if (e->d_qctr == nullptr) LRG_TRY(dev_alloc(&e->d_qctr, 8));
      if (e->d_busy == nullptr) LRG_TRY(dev_alloc(&e->d_busy, 24));
      if (n_slots > e->sync_slots) {
        cudaFree(e->d_sync); cudaFree(e->d_remaining);
        e->d_sync = nullptr; e->d_remaining = nullptr;
        e->sync_slots = 0;
        LRG_TRY(dev_alloc(&e->d_sync, (size_t)n_slots));
        LRG_TRY(dev_alloc(&e->d_remaining, (size_t)n_slots + 1));
        e->sync_slots = n_slots;
      }
      // (random restarts: only lane 0 of every group starts; it wakes the other lanes once it holds a seed)
      std::vector<unsigned long long> first(n_groups);
      for (int g = 0; g < n_groups; ++g) first[g] = (1ull << 32) | make_item(ITEM_STEP, g * lanes, 0, 0);
      const unsigned ctr[8] = {0u, 0u, 0u, (unsigned)n_groups, 0u, 0u, 0u, 0u};
      LRG_CUDA(cudaMemsetAsync(e->d_qring, 0, sizeof(unsigned long long) * 3 * e->q_capacity, st));
      LRG_CUDA(cudaMemcpyAsync(e->d_qring + e->q_capacity, first.data(), sizeof(unsigned long long) * n_groups, cudaMemcpyHostToDevice, st));
      LRG_CUDA(cudaMemcpyAsync(e->d_qctr, ctr, sizeof(ctr), cudaMemcpyHostToDevice, st));
      LRG_CUDA(cudaMemsetAsync(e->d_busy, 0, sizeof(unsigned long long) * 24, st));
      LRG_CUDA(cudaMemsetAsync(e->d_sync, 0, sizeof(SlotSync) * n_slots, st));
      LRG_CUDA(cudaMemsetAsync(e->d_remaining, 0, sizeof(int) * ((size_t)n_slots + 1), st));
      GrowArgs ga{};
      ga.da = da;
      ga.da.done_flag = nullptr;
      ga.da.q_ctr = e->d_qctr + 2;                  // head / tail of the work queue everybody pops (load hint of the speculative window)
      ga.fa = fa;
      ga.fa.active = nullptr;
      ga.net = e->tcnet;
      for (int k = 0; k < 2; ++k) {
        ga.q[k].ring = e->d_qring + (size_t)k * e->q_capacity; ga.q[k].cap_mask = e->q_capacity - 1;
        ga.q[k].head = e->d_qctr + 2 * k; ga.q[k].tail = e->d_qctr + 2 * k + 1;
      }
      ga.sync = e->d_sync;
      ga.progress = e->d_qctr + 5; ga.abort = reinterpret_cast<int*>(e->d_qctr + 6);
      ga.busy_ns = e->d_busy;
      ga.remaining = e->d_remaining;
      // reserved CTAs for the slots with the most work left (LRG_FLAG_PRIORITY; LRG_HI="slots,ctas" overrides for experiments)
      ga.hi_slots = 0; ga.hi_ctas = 0;
      if (spec && (params->flags & LRG_FLAG_PRIORITY)) {
        // speculative lanes: 16 reserved CTAs serve the CRITICAL rooms (DriverArgs::spec_crit).  Off by default: it pays only when a
        // GPU holds few enough rooms to be chain-bound but enough to queue (34 rooms: 204 -> 190 ms) and costs throughput elsewhere
        // (68 rooms 210 -> 217 ms, 272 rooms 632 -> 711 ms; profiles/r2an_*)
        ga.hi_slots = 1; ga.hi_ctas = 16; ga.hi_crit = 1;
      } else if (lanes == 1 && !beam) {
        const int hs = (params->flags & LRG_FLAG_PRIORITY) ? 2 : 0, hc = (params->flags & LRG_FLAG_PRIORITY) ? 24 : 0;
        const int n_ctas = e->sm_count > 0 ? e->sm_count : 148;
        if (hs > 0 && hc > 0 && hc < n_ctas) { ga.hi_slots = std::min(hs, 8); ga.hi_ctas = hc; }
      }
      // pooled-projection servers: 16 of the CTAs keep W0[:1024] of both heads in shared memory (LRG_GSERVERS=0 turns them off)
      const int n_ctas_total = e->sm_count > 0 ? e->sm_count : 148;
      ga.n_servers = !(params->flags & LRG_FLAG_NO_PROJ_SERVERS) && n_ctas_total >= 4 * kProjServers ? kProjServers : 0;
      ga.greq_ring = e->d_qring + (size_t)2 * e->q_capacity; ga.greq_mask = e->q_capacity - 1; ga.greq_tail = e->d_qctr + 4;
      ga.gproj_tagged = e->s_gproj_tagged;
      if (ga.n_servers > 0) LRG_CUDA(cudaMemsetAsync(e->s_gproj_tagged, 0, sizeof(uint2) * (size_t)n_slots * 2 * e->net.H0, st));   // (tags restart with the run)
      if (ga.hi_ctas + ga.n_servers >= n_ctas_total) { ga.hi_ctas = 0; ga.hi_slots = 0; }
      // both measured positive (profiles/README.md): bit 0 = split branch tiles over idle CTAs, bit 1 = head tiles go out with the projection
      ga.tune = ((params->flags & LRG_FLAG_NO_TILE_SPLIT) ? 0 : 1) | ((params->flags & LRG_FLAG_HEADS_AFTER_PROJ) ? 0 : 2);
      e->last_kind = use_f16(e) ? 2 : 1;
      rc = launch_grow(ga, e->sm_count > 0 ? e->sm_count : 148, use_f16(e), st);
      if (rc == LRG_OK) {
        cudaError_t se = cudaStreamSynchronize(st);
        if (se != cudaSuccess) { set_error("persistent grow kernel -> %s", cudaGetErrorString(se)); rc = LRG_E_CUDA; }
      }
      if (rc == LRG_OK) LRG_CUDA(cudaMemcpy(e->h_busy, e->d_busy, sizeof(e->h_busy), cudaMemcpyDeviceToHost));
      if (rc == LRG_OK) {
        unsigned aborted = 0;
        LRG_CUDA(cudaMemcpy(&aborted, e->d_qctr + 6, sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (aborted != 0) { set_error("persistent grow kernel: no work item retired anywhere for 20 s -- the run was abandoned"); rc = LRG_E_STATE; }
      }
      e->iterations = 0;
      e->launches = 1;
    } else if (use_graph) {
      constexpr int kIterPerGraph = 8;
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      LRG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      for (int it = 0; it < kIterPerGraph && rc == LRG_OK; ++it) {
        rc = launch_step(da, st);
        if (rc == LRG_OK) rc = run_forward(e, fa, st, nullptr);
      }
      cudaError_t cerr = cudaStreamEndCapture(st, &graph);
      if (rc != LRG_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      LRG_CUDA(cerr);
      LRG_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      cudaEvent_t evq[2];
      LRG_CUDA(cudaEventCreateWithFlags(&evq[0], cudaEventDisableTiming));
      LRG_CUDA(cudaEventCreateWithFlags(&evq[1], cudaEventDisableTiming));
      long long launched = 0;
      while (true) {
        cudaError_t le = cudaGraphLaunch(exec, st);
        if (le != cudaSuccess) { set_error("cudaGraphLaunch -> %s", cudaGetErrorString(le)); rc = LRG_E_CUDA; break; }
        cudaEventRecord(evq[launched & 1], st);
        ++launched;
        if (launched >= 2) {
          cudaError_t se = cudaEventSynchronize(evq[launched & 1]);   // the launch before the one just queued
          if (se != cudaSuccess) { set_error("grow loop -> %s", cudaGetErrorString(se)); rc = LRG_E_CUDA; break; }
          if (*(volatile int*)e->h_done) break;
        }
      }
      e->iterations = launched * kIterPerGraph;
      e->launches = e->iterations * 4;
      cudaStreamSynchronize(st);
      cudaEventDestroy(evq[0]); cudaEventDestroy(evq[1]);
      cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
    } else {
      cudaEvent_t kev[5];
      for (auto& k : kev) LRG_CUDA(cudaEventCreate(&k));
      while (rc == LRG_OK) {
        if (kernel_timing) cudaEventRecord(kev[0], st);
        rc = launch_step(da, st);
        if (rc != LRG_OK) break;
        rc = run_forward(e, fa, st, kernel_timing ? kev + 1 : nullptr);
        e->iterations += 1; e->launches += 4;
        if (kernel_timing || (e->iterations % 16) == 0) {
          cudaError_t se = cudaStreamSynchronize(st);
          if (se != cudaSuccess) { set_error("grow loop -> %s", cudaGetErrorString(se)); rc = LRG_E_CUDA; break; }
          if (kernel_timing) {
            for (int i = 0; i < 4; ++i) { float ms = 0; cudaEventElapsedTime(&ms, kev[i], kev[i + 1]); e->kernel_ms[i] += ms; }
            e->forward_ms = e->kernel_ms[1] + e->kernel_ms[2] + e->kernel_ms[3];
          }
          if (*(volatile int*)e->h_done) break;
        }
      }
      for (auto& k : kev) cudaEventDestroy(k);
    }
  }
  if (rc != LRG_OK) return rc;
  LRG_CUDA(cudaEventRecord(ev1, st));
  FillArgs fl{};
  fl.n_rooms = n_rooms; fl.room_off = e->d_room_off; fl.pts = e->d_pts; fl.label = e->d_label; fl.label_filled = e->d_label_filled;
  fl.lab_list = e->d_lab_list; fl.unl_list = e->d_unl_list; fl.n_lab = e->d_n_lab; fl.n_unl = e->d_n_unl; fl.F = e->F;
  if (!(params->flags & LRG_FLAG_NO_SPATIAL_INDEX)) {     // nearest labelled point through the rooms' spatial index (same labels)
    fl.sp_off = e->d_sp_off; fl.sp_perm = e->d_sp_perm; fl.sp_box = e->d_sp_box; fl.room_vmin = e->d_room_vmin; fl.resolution = e->resolution;
  }
  LRG_TRY(launch_fill(fl, st));
  e->launches += 2;
  LRG_CUDA(cudaEventRecord(ev2, st));
  LRG_CUDA(cudaStreamSynchronize(st));
  LRG_CUDA(cudaEventElapsedTime(&e->grow_ms, ev0, ev1));
  LRG_CUDA(cudaEventElapsedTime(&e->fill_ms, ev1, ev2));
  if (stats != nullptr && n_rooms > 0)
    LRG_CUDA(cudaMemcpy(stats, e->d_stats, sizeof(LrgRoomStats) * n_rooms, cudaMemcpyDeviceToHost));
  return LRG_OK;
}

int lrg_labels_download(LrgEngine* e, int32_t* labels, int filled) {
  LRG_REQUIRE(e != nullptr && (labels != nullptr || e->total_pts == 0), "engine/labels is NULL");
  LRG_CUDA(cudaSetDevice(e->device));
  if (e->total_pts > 0)
    LRG_CUDA(cudaMemcpy(labels, filled ? e->d_label_filled : e->d_label, sizeof(int) * (size_t)e->total_pts, cudaMemcpyDeviceToHost));
  return LRG_OK;
}

int lrg_trace_download_lane(LrgEngine* e, int room, int lane, LrgStepTrace* out, int capacity, int* n_steps) {
  LRG_REQUIRE(e != nullptr && out != nullptr && n_steps != nullptr, "NULL argument");
  LRG_REQUIRE(room >= 0 && room < e->n_rooms, "room %d out of range", room);
  LRG_REQUIRE(lane >= 0 && lane < e->last_lanes, "lane %d out of range (the last run had %d)", lane, e->last_lanes);
  if (e->last_lanes <= 1) return lrg_trace_download(e, room, out, capacity, n_steps);
  if (e->trace_capacity <= 0 || e->d_trace == nullptr) { set_error("no trace recorded (trace_capacity was 0)"); return LRG_E_STATE; }
  LRG_CUDA(cudaSetDevice(e->device));
  int steps = 0;
  LRG_CUDA(cudaMemcpy(&steps, e->d_lane_steps + (size_t)room * e->last_lanes + lane, sizeof(int), cudaMemcpyDeviceToHost));
  const int n = std::min(std::min(steps, e->trace_capacity), capacity);
  *n_steps = steps;
  if (n > 0)
    LRG_CUDA(cudaMemcpy(out, e->d_trace + ((size_t)room * e->last_lanes + lane) * e->trace_capacity, sizeof(LrgStepTrace) * n, cudaMemcpyDeviceToHost));
  return LRG_OK;
}

int lrg_trace_download(LrgEngine* e, int room, LrgStepTrace* out, int capacity, int* n_steps) {
  LRG_REQUIRE(e != nullptr && out != nullptr && n_steps != nullptr, "NULL argument");
  LRG_REQUIRE(room >= 0 && room < e->n_rooms, "room %d out of range", room);
  LRG_REQUIRE(e->last_lanes <= 1, "the last run used %d restart lanes: use lrg_trace_download_lane", e->last_lanes);
  if (e->trace_capacity <= 0 || e->d_trace == nullptr) { set_error("no trace recorded (trace_capacity was 0)"); return LRG_E_STATE; }
  LRG_CUDA(cudaSetDevice(e->device));
  LrgRoomStats st;
  LRG_CUDA(cudaMemcpy(&st, e->d_stats + room, sizeof(st), cudaMemcpyDeviceToHost));
  int n = std::min(std::min(st.grow_steps, e->trace_capacity), capacity);
  *n_steps = st.grow_steps;
  if (n > 0) LRG_CUDA(cudaMemcpy(out, e->d_trace + (size_t)room * e->trace_capacity, sizeof(LrgStepTrace) * n, cudaMemcpyDeviceToHost));
  return LRG_OK;
}

int lrg_segment_rooms_host(LrgEngine* e, int n_rooms, const int64_t* room_offsets, const float* points, const int32_t* seed_order,
                           const LrgGrowParams* params, int32_t* labels_filled, LrgRoomStats* stats) {
  LRG_REQUIRE(params != nullptr, "params is NULL");
  LRG_TRY(lrg_rooms_upload(e, n_rooms, room_offsets, points, seed_order, params->resolution));
  LRG_TRY(lrg_segment_resident(e, params, stats));
  return lrg_labels_download(e, labels_filled, 1);
}

int lrg_last_segment_profile(LrgEngine* e, float* grow_ms, float* fill_ms, int64_t* iterations, int64_t* kernel_launches,
                             float* forward_ms) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  if (grow_ms) *grow_ms = e->grow_ms;
  if (fill_ms) *fill_ms = e->fill_ms;
  if (iterations) *iterations = e->iterations;
  if (kernel_launches) *kernel_launches = e->launches;
  if (forward_ms) *forward_ms = e->forward_ms;
  return LRG_OK;
}

int lrg_last_grow_queue_delay(LrgEngine* e, double delay_ms[4]) {
  LRG_REQUIRE(e != nullptr && delay_ms != nullptr, "NULL argument");
  const int types[4] = {ITEM_STEP, ITEM_BRANCH, ITEM_GPROJ, ITEM_HEAD};
  for (int i = 0; i < 4; ++i) delay_ms[i] = e->last_persistent ? (double)e->h_busy[16 + types[i]] * 1e-6 : 0.0;
  return LRG_OK;
}

int lrg_last_grow_profile(LrgEngine* e, int* persistent, double busy_ms[4], int64_t items[4]) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  if (persistent) *persistent = e->last_persistent ? 1 : 0;
  const int types[4] = {ITEM_STEP, ITEM_BRANCH, ITEM_GPROJ, ITEM_HEAD};
  for (int i = 0; i < 4; ++i) {
    if (busy_ms) busy_ms[i] = e->last_persistent ? (double)e->h_busy[types[i]] * 1e-6 : 0.0;
    if (items) items[i] = e->last_persistent ? (int64_t)e->h_busy[8 + types[i]] : 0;
  }
  return LRG_OK;
}

int lrg_engine_set_tile_timing(LrgEngine* e, int on) {
  LRG_REQUIRE(e != nullptr, "engine is NULL");
  LRG_CUDA(cudaSetDevice(e->device));
  if (on && e->d_tile_dbg == nullptr) {
    LRG_TRY(dev_alloc(&e->d_tile_dbg, 64));
    LRG_CUDA(cudaMemset(e->d_tile_dbg, 0, 64 * sizeof(unsigned long long)));
  } else if (!on && e->d_tile_dbg != nullptr) {
    LRG_CUDA(cudaDeviceSynchronize());
    cudaFree(e->d_tile_dbg);
    e->d_tile_dbg = nullptr;
  }
  e->tcnet.dbg = e->d_tile_dbg;
  return LRG_OK;
}

int lrg_tile_timing(LrgEngine* e, uint64_t out[64], int reset) {
  LRG_REQUIRE(e != nullptr && out != nullptr, "NULL argument");
  if (e->d_tile_dbg == nullptr) { set_error("tile timing is off (call lrg_engine_set_tile_timing(e, 1) first)"); return LRG_E_STATE; }
  LRG_CUDA(cudaSetDevice(e->device));
  LRG_CUDA(cudaMemcpy(out, e->d_tile_dbg, 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (reset) LRG_CUDA(cudaMemset(e->d_tile_dbg, 0, 64 * sizeof(uint64_t)));
  return LRG_OK;
}

int lrg_last_kernel_times(LrgEngine* e, float out_ms[4]) {
  LRG_REQUIRE(e != nullptr && out_ms != nullptr, "NULL argument");
  for (int i = 0; i < 4; ++i) out_ms[i] = e->kernel_ms[i];
  return LRG_OK;
}

int lrg_labels_device_ptr(LrgEngine* e, int filled, void** d_ptr) {
  LRG_REQUIRE(e != nullptr && d_ptr != nullptr, "NULL argument");
  *d_ptr = filled ? (void*)e->d_label_filled : (void*)e->d_label;
  return LRG_OK;
}

int lrg_room_metrics(LrgEngine* e, const int32_t* obj_id, int raw, int filled, LrgRoomMetrics* out, int32_t* cluster_label2) {
  LRG_REQUIRE(e != nullptr && (out != nullptr || e->n_rooms == 0), "engine/out is NULL");
  if (e->d_label == nullptr && e->total_pts > 0) { set_error("no rooms uploaded"); return LRG_E_STATE; }
  if (raw && !e->raw_mode) { set_error("rooms were not uploaded as raw points"); return LRG_E_STATE; }
  LRG_CUDA(cudaSetDevice(e->device));
  if (e->n_rooms == 0) return LRG_OK;
  const size_t TE = (size_t)e->total_pts, TIN = raw ? (size_t)e->total_raw : TE;
  LRG_REQUIRE(obj_id != nullptr || TIN == 0, "obj_id is NULL");
  int *d_in = nullptr, *d_eq = nullptr, *d_l2 = nullptr;
  int rc = pool_alloc(e, &d_in, TIN);
  if (rc == LRG_OK && raw) rc = pool_alloc(e, &d_eq, TE);
  if (rc == LRG_OK && cluster_label2 != nullptr) rc = pool_alloc(e, &d_l2, TE);
  cudaError_t ce = cudaSuccess;
  if (rc == LRG_OK && TIN > 0) ce = cudaMemcpyAsync(d_in, obj_id, sizeof(int) * TIN, cudaMemcpyHostToDevice, e->stream);
  if (rc == LRG_OK && ce == cudaSuccess && raw)
    rc = launch_gather_equalized(e->n_rooms, e->d_raw_off, e->d_room_off, e->d_equalized_idx, d_in, d_eq, e->stream);
  if (rc == LRG_OK && ce == cudaSuccess)
    rc = segmentation_metrics(e->n_rooms, reinterpret_cast<const int64_t*>(e->h_room_off.data()), raw ? d_eq : d_in, filled ? e->d_label_filled : e->d_label, out, d_l2, e->stream);
  if (rc == LRG_OK && ce == cudaSuccess && cluster_label2 != nullptr && TE > 0) {
    ce = cudaMemcpyAsync(cluster_label2, d_l2, sizeof(int) * TE, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  }
  pool_free(e, d_in); pool_free(e, d_eq); pool_free(e, d_l2);
  LRG_TRY(rc);
  LRG_CUDA(ce);
  return LRG_OK;
}

int lrg_malloc(void** d_ptr, size_t bytes) {
  LRG_REQUIRE(d_ptr != nullptr, "d_ptr is NULL");
  LRG_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
  return LRG_OK;
}
int lrg_free(void* d_ptr) { LRG_CUDA(cudaFree(d_ptr)); return LRG_OK; }
int lrg_memcpy_h2d(void* d_dst, const void* src, size_t bytes) { LRG_CUDA(cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice)); return LRG_OK; }
int lrg_memcpy_d2h(void* dst, const void* d_src, size_t bytes) { LRG_CUDA(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost)); return LRG_OK; }
int lrg_memset(void* d_ptr, int value, size_t bytes) { LRG_CUDA(cudaMemset(d_ptr, value, bytes)); return LRG_OK; }
int lrg_device_synchronize(void) { LRG_CUDA(cudaDeviceSynchronize()); return LRG_OK; }
int lrg_set_device(int device) { LRG_CUDA(cudaSetDevice(device)); return LRG_OK; }
int lrg_host_alloc(void** ptr, size_t bytes) { LRG_REQUIRE(ptr != nullptr, "ptr is NULL"); LRG_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault)); return LRG_OK; }
int lrg_host_free(void* ptr) { LRG_CUDA(cudaFreeHost(ptr)); return LRG_OK; }

#pragma GCC visibility pop
}  // extern "C"
