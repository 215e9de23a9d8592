// Tensor-core (tcgen05) forward of the full LrgNet model: shared declarations (internal).
#pragma once
#include "lrg_common.cuh"

namespace lrg {

constexpr int kBranchChunks = 21;   // L0 (8 KB) + L1 + L2 + 2 (L3, K halves) + 16 (L4: 4 column blocks x 4 K quarters), 32 KB each
constexpr int kHeadChunks = 12;     // 4 x W0 column block [64x64] interleaved with 4 x 2 W1 K-halves [128x32]
constexpr size_t kBranchImgFloats = 2048 + 20 * 8192;
constexpr size_t kHeadImgFloats = 12 * 8192;

// Device pointers of the pre-packed operand images and the fp32 vectors the epilogues read.
struct TcNet {
  int F;
  const float* branch_img[2];       // [0] inlier branch, [1] neighbor branch
  const float* conv_bias[2][5];
  const float* W0g[2];              // head layer 0, pooled part [1024][256] fp32 ([0] remove head, [1] add head)
  const float* head_img[2];
  const float* head_bias0[2];       // [256]
  const float* head_bias1[2];       // [128]
  const float* head_W2[2];          // [128][2]
  const float* head_bias2[2];       // [2]
  // diagnostics (NULL = off): summed clock64 cycles per tile stage, [0..15] branch tile, [16..31] head tile;
  // the last entry of each half counts tiles
  unsigned long long* dbg;
};

int tc_forward_configure();
int launch_forward_tc(const TcNet& net, const ForwardArgs& fa, cudaStream_t stream, cudaEvent_t* ev);

}  // namespace lrg
