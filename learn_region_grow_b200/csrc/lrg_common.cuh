// Shared declarations of the sm_100a LRGNet grow engine (internal; the public surface is include/lrg_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/lrg_b200.h"

namespace lrg {

void set_error(const char* fmt, ...);

#define LRG_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t err__ = (call);                                                                 \
    if (err__ != cudaSuccess) {                                                                 \
      lrg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__));   \
      return LRG_E_CUDA;                                                                        \
    }                                                                                           \
  } while (0)

#define LRG_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      lrg::set_error(__VA_ARGS__);      \
      return LRG_E_INVALID;             \
    }                                   \
  } while (0)

constexpr int kMaxConv = 5;     // per-branch 1x1 conv layers (learn_region_grow_util.py:78)
constexpr int kMaxHidden = 2;   // hidden head layers (util.py:79)
constexpr int kTileRows = 128;  // points per CTA tile in the MLP kernels

// One dense layer: W is row-major [Kpad][N] (rows >= K are zero), bias [N].
struct LayerDesc {
  const float* W;
  const float* bias;
  int K, Kpad, N;
};

// Device-side description of the network (passed by value as a kernel parameter).
struct NetDesc {
  int F;                 // feature_size
  int n_conv;            // conv layers per branch
  int n_hidden;          // hidden head layers
  int Clast;             // CONV_CHANNELS[-1]
  int C1;                // CONV_CHANNELS[1]: the per-point feature concatenated after pooling (util.py:130,134)
  int H0;                // CONV2_CHANNELS[0]
  LayerDesc conv[2][kMaxConv];        // [0]=inlier branch lrg_*, [1]=neighbor branch lrg_neighbor_*
  const float* W0g[2];                // head layer 0, pooled part: [2*Clast][H0]; [0]=remove head (inlier rows), [1]=add head (neighbor rows)
  LayerDesc head0_local[2];           // head layer 0, per-point part: [C1][H0] + bias0
  LayerDesc hidden[2][kMaxHidden];    // hidden[h][i] = head layer i+1 (i < n_hidden-1)
  LayerDesc out[2];                   // final [Hlast][2]
};

// Launch description of one forward over `B` tile pairs.
struct ForwardArgs {
  const float* x[2];     // [0] inlier (B, N0, x_stride), [1] neighbor (B, N1, x_stride): rows of F features
  int x_stride;          // floats between consecutive rows: F for user tensors, 16 (padded, 16-byte aligned) for slot tiles
  int n_pts[2];
  float* h1[2];          // (B, n_pts, C1) scratch
  float* pooled;         // (B, 2*Clast) as int-ordered floats, must be zero on entry
  float* gproj;          // (B, 2 heads, H0)
  float* logits[2];      // [0] remove_output (B, N0, 2), [1] add_output (B, N1, 2)
  const int* active;     // optional (B) flags; NULL = all active
  int active_stride;     // stride in ints between consecutive flags
  // optional: n_valid[b * n_valid_stride + s] = rows of set s (0 inlier, 1 neighbor) of tile pair b that carry distinct
  // points; rows beyond are padding duplicates (test_region_grow.py:239-240,251-252) and are not evaluated.  NULL = all rows.
  const int* n_valid;
  int n_valid_stride;
  int B;
};

__device__ __forceinline__ int forward_valid_rows(const ForwardArgs& fa, int b, int s) {
  const int n = fa.n_pts[s];
  if (fa.n_valid == nullptr) return n;
  const int v = __ldcg(fa.n_valid + (size_t)b * fa.n_valid_stride + s);
  return v < n ? v : n;
}

int launch_forward(const NetDesc& net, const ForwardArgs& fa, cudaStream_t stream);
// same launches with an event recorded before the branch kernel and after each of the three kernels (ev[0..3])
int launch_forward_timed(const NetDesc& net, const ForwardArgs& fa, cudaStream_t stream, cudaEvent_t* ev);
size_t forward_smem_branch(const NetDesc& net);
size_t forward_smem_head(const NetDesc& net);
int forward_configure(const NetDesc& net);

// ---------------------------------------------------------------------------------------------- Philox4x32-10
// Same generator as oracle/philox.py; draw = word 0 of philox(key=seed, counter=(element, stream, step, room)) with
// stream = (seed point of the region << 8) | (8 * restart lane + kStream*), step = step within the region.
enum {
  kStreamInlierKey = 0, kStreamNeighborKey = 1, kStreamAddUniform = 2, kStreamRemoveUniform = 3,
  kStreamInlierPad = 4, kStreamNeighborPad = 5
};

__host__ __device__ inline uint32_t philox_draw(uint64_t seed, uint32_t room, uint32_t step, uint32_t stream, uint32_t element) {
  uint32_t c0 = element, c1 = stream, c2 = step, c3 = room;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c0;
}

}  // namespace lrg
