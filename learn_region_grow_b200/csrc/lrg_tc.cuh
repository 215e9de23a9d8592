// Tensor-core (tcgen05) forward of the full LrgNet model: shared declarations (internal).
#pragma once
#include "lrg_common.cuh"

namespace lrg {

// Operand images (B operands of the tensor tiles, hi image followed by lo image per chunk, consumption order), one set per
// arithmetic kind:
//   3xTF32 (kind::tf32, K = 8 per MMA):  branch = L0 (8 KB) + L1 + L2 + 2 (L3, K halves) + 16 (L4: 4 column blocks x 4 K
//            quarters), 32 KB each; head = 4 x W0 column block [64x64] interleaved with 4 x 2 W1 K-halves [128x32], 32 KB each
//   3xFP16 (kind::f16, K = 16 per MMA, half the bytes and half the MMAs): branch = L0 (4 KB) + L1 + L2 (16 KB each) + L3
//            (32 KB) + 8 (L4: 4 column blocks x 2 K halves, 32 KB each); head = the same 12 chunks at 16 KB each.  The
//            weights of every layer are scaled by a power of two before the split (so that their lo parts are normal
//            fp16 numbers); the epilogue multiplies the accumulator by the inverse.
constexpr int kBranchChunksTf32 = 21, kBranchChunksF16 = 12;
constexpr int kHeadChunks = 12;
constexpr size_t kBranchImgBytesTf32 = 8192 + 20 * 32768, kBranchImgBytesF16 = 4096 + 2 * 16384 + 9 * 32768;
constexpr size_t kHeadImgBytesTf32 = 12 * 32768, kHeadImgBytesF16 = 12 * 16384;

// Device pointers of the pre-packed operand images and the fp32 vectors the epilogues read.
struct TcNet {
  int F;
  const unsigned char* branch_img[2][2];   // [kind: 0 = tf32, 1 = f16][0 inlier branch, 1 neighbor branch]
  const float* conv_bias[2][5];
  const float* W0g[2];              // head layer 0, pooled part [1024][256] fp32 ([0] remove head, [1] add head)
  const unsigned char* head_img[2][2];     // [kind][head]
  const float* head_bias0[2];       // [256]
  const float* head_bias1[2];       // [128]
  const float* head_W2[2];          // [128][2]
  const float* head_bias2[2];       // [2]
  float branch_inv[2][5];           // 3xFP16: 1 / (power of two the layer's weights were scaled by); [branch][layer]
  float head_inv[2][2];             // [head][layer 0 (per-point part), layer 1]
  // 3xFP16: set to 1 by any tile that saw an activation (or input) beyond the fp16 range -- the caller repeats the call with
  // 3xTF32 (never observed on real inputs: the activations of the shipped model stay below ~500)
  int* range_flag;
  // diagnostics (NULL = off): summed clock64 cycles per tile stage, [0..15] branch tile, [16..31] head tile;
  // the last entry of each half counts tiles
  unsigned long long* dbg;
};

int tc_forward_configure();
int launch_forward_tc(const TcNet& net, const ForwardArgs& fa, bool f16, cudaStream_t stream, cudaEvent_t* ev);

}  // namespace lrg
