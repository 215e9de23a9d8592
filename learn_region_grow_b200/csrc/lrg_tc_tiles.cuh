// Device-side bodies of the tensor-core (tcgen05) LrgNet forward: one 128-point tile of one branch / one head, and one
// 64-column block of the pooled projection.  Shared by the stand-alone kernels of the lock-step loop
// (lrg_forward_tc.cu) and by the persistent grow kernel (lrg_persistent.cu).  See lrg_forward_tc.cu for the design.
//
// Calling convention: every thread of the CTA calls the function (blockDim.x >= 192, a multiple of 32); warps 0-3 are the
// epilogue warps (warp w owns TMEM lanes 32w..32w+31 = tile rows), warp 4 issues the MMAs, warp 5 streams the weights,
// further warps only take part in the CTA barriers.  `tmem` is the base of 256 allocated TMEM columns.  Data produced by
// other CTAs during the same launch (tiles, h1, pooled, gproj) is read with ld.global.cg, never through the
// non-coherent path.
#pragma once
#include "lrg_common.cuh"
#include "lrg_tc.cuh"
#include "lrg_umma.cuh"

namespace lrg {

using namespace umma;

constexpr uint32_t kSlotBytes = 32768;
constexpr uint32_t kActBytes = 131072;                 // two 64 KB activation regions (or one 128-channel hi/lo pair)
constexpr uint32_t kTcSmem = kActBytes + 3 * kSlotBytes;
constexpr uint32_t kKdir = 2048;                       // bytes between K-adjacent core matrices of a 128-row operand
constexpr uint32_t kMNdir = 128;                       // bytes between 8-row groups
constexpr uint32_t kTmemCols = 256;

struct TcBarriers {
  uint64_t full[3], empty[3];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t act_ready;          // branch: activations of the next layer written; head: h1 tile + sG written
  uint64_t c_ready, c_free;    // head only
  uint64_t acc1_full;          // head only
};
constexpr int kTcBarrierCount = sizeof(TcBarriers) / 8;

// Small per-CTA scratch next to the barriers: biases (branch: layers 0-3 = 320 floats; head: gproj row 256 + bias1 128 +
// W2 256 + bias2 2 = 642 floats).
struct TcStatic {
  TcBarriers bars;
  float vec[648];
};

__device__ __forceinline__ void mbar_inval(uint32_t bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

// One weight chunk = hi image + lo image of an [Nc x Kc] K-major operand; D[tmem] (+)= A(hi,lo)[128 x Kc] . chunk^T.
__device__ __forceinline__ void mma_chunk(uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, int Nc, int Kc, uint32_t d_tmem,
                                          uint32_t idesc, bool first) {
  const uint32_t b_lo = b_hi + (uint32_t)(Nc * Kc * 4);
  const uint32_t kdirB = (uint32_t)Nc * 16;
  uint32_t acc = first ? 0u : 1u;
#pragma unroll
  for (int term = 0; term < 3; ++term) {
    const uint32_t a0 = (term == 1) ? a_lo : a_hi;
    const uint32_t b0 = (term == 2) ? b_lo : b_hi;
    for (int ks = 0; ks < Kc / 8; ++ks) {
      umma_tf32(d_tmem, make_desc(a0 + ks * 2 * kKdir, kKdir, kMNdir), make_desc(b0 + ks * 2 * kdirB, kdirB, kMNdir), idesc, acc);
      acc = 1u;
    }
  }
}

// bias + ReLU + hi/lo split of 32 accumulator columns of this thread's row, written as 8 canonical 16-byte chunks.
__device__ __forceinline__ void store_act32(const uint32_t (&v)[32], const float* s_bias, float* s_hi, float* s_lo, int chunk0,
                                            int r, float* g_row) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(s_bias + q * 4);
    float4 x, hi, lo;
    x.x = fmaxf(__uint_as_float(v[q * 4 + 0]) + b.x, 0.f);
    x.y = fmaxf(__uint_as_float(v[q * 4 + 1]) + b.y, 0.f);
    x.z = fmaxf(__uint_as_float(v[q * 4 + 2]) + b.z, 0.f);
    x.w = fmaxf(__uint_as_float(v[q * 4 + 3]) + b.w, 0.f);
    split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(s_hi + (chunk0 + q) * 512 + r * 4) = hi;
    *reinterpret_cast<float4*>(s_lo + (chunk0 + q) * 512 + r * 4) = lo;
    if (g_row != nullptr) *reinterpret_cast<float4*>(g_row + q * 4) = x;
  }
}

__device__ __forceinline__ void tc_init_barriers(TcBarriers& bars, bool head) {
  for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(&bars.full[i]), 1); mbar_init(smem_u32(&bars.empty[i]), 1); }
  for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars.acc_full[i]), 1); mbar_init(smem_u32(&bars.acc_empty[i]), 128); }
  mbar_init(smem_u32(&bars.act_ready), head ? 128 + 32 : 128);   // head: 128 rows of h1 + the 32 loader lanes that fill vec[]
  mbar_init(smem_u32(&bars.c_ready), 128);
  mbar_init(smem_u32(&bars.c_free), 1);
  mbar_init(smem_u32(&bars.acc1_full), 1);
  fence_barrier_init();
}
__device__ __forceinline__ void tc_inval_barriers(TcBarriers& bars) {
  uint64_t* p = reinterpret_cast<uint64_t*>(&bars);
  for (int i = 0; i < kTcBarrierCount; ++i) mbar_inval(smem_u32(p + i));
}

// ------------------------------------------------------------------------------------------------------ branch tile
// x (rows x F) -> 64 -> 64 -> 64 -> 128 -> 512 -> column max merged into pooled (learn_region_grow_util.py:106-123).
__device__ __forceinline__ void tc_branch_tile(const TcNet& net, const ForwardArgs& fa, int b, int br, int tile,
                                               unsigned char* smem, TcStatic& st, uint32_t tmem) {
  const int n = fa.n_pts[br];
  const int row0 = tile * 128;
  const int rows = min(128, n - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TcBarriers& bars = st.bars;
  float* const act = reinterpret_cast<float*>(smem);
  const uint32_t act_u32 = smem_u32(smem);
  const uint32_t ring_u32 = act_u32 + kActBytes;

  if (tid == 0) tc_init_barriers(bars, false);
  if (tid >= 192 && tid < 192 + 80) {                  // biases of layers 0-3 (64,64,64,128) -> st.vec[0..320)
    const int i = (tid - 192) * 4;
    const int l = i < 64 ? 0 : i < 128 ? 1 : i < 192 ? 2 : 3;
    const int o = i - (l == 0 ? 0 : l == 1 ? 64 : l == 2 ? 128 : 192);
    *reinterpret_cast<float4*>(&st.vec[i]) = __ldg(reinterpret_cast<const float4*>(net.conv_bias[br][l] + o));
  } else if (blockDim.x < 192 + 80 && tid < 80) {      // (CTA without spare warps: the epilogue warps do it)
    const int i = tid * 4;
    const int l = i < 64 ? 0 : i < 128 ? 1 : i < 192 ? 2 : 3;
    const int o = i - (l == 0 ? 0 : l == 1 ? 64 : l == 2 ? 128 : 192);
    *reinterpret_cast<float4*>(&st.vec[i]) = __ldg(reinterpret_cast<const float4*>(net.conv_bias[br][l] + o));
  }
  __syncthreads();

  // activation regions (float offsets): 64-channel tensors use hi = region, lo = region + 8192 floats (32 KB);
  // x (16 channels) uses hi = 0, lo = 2048 floats; h3 (128 channels) uses hi = 0, lo = 16384 floats (64 KB).
  constexpr int kR0 = 0, kR1 = 16384, kLo64 = 8192, kLoX = 2048, kLo128 = 16384;

  if (warp < 4) {
    // ===================================================================== epilogue warps: thread = tile row
    const int r = tid;
    const bool valid = r < rows;
    {
      const float* xrow = fa.x[br] + ((size_t)b * n + row0 + r) * net.F;
      float xv[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) xv[c] = (valid && c < net.F) ? __ldcg(xrow + c) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 hi, lo;
        split_tf32(xv[q * 4 + 0], hi.x, lo.x); split_tf32(xv[q * 4 + 1], hi.y, lo.y);
        split_tf32(xv[q * 4 + 2], hi.z, lo.z); split_tf32(xv[q * 4 + 3], hi.w, lo.w);
        *reinterpret_cast<float4*>(act + kR0 + q * 512 + r * 4) = hi;
        *reinterpret_cast<float4*>(act + kR0 + kLoX + q * 512 + r * 4) = lo;
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars.act_ready));
    }
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    float* g_h1 = valid ? fa.h1[br] + ((size_t)b * n + row0 + r) * 64 : nullptr;
#pragma unroll 1
    for (int l = 0; l < 4; ++l) {
      const int buf = l & 1;
      mbar_wait(smem_u32(&bars.acc_full[buf]), (uint32_t)(l >> 1) & 1u);
      tcgen05_fence_after();
      const int N = (l == 3) ? 128 : 64;
      float* s_hi = act + ((l == 0 || l == 2) ? kR1 : kR0);
      float* s_lo = s_hi + ((l == 3) ? kLo128 : kLo64);
      const float* s_bias = st.vec + l * 64;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + buf * 128 + c0, v);
        tmem_ld_wait();
        store_act32(v, s_bias + c0, s_hi, s_lo, c0 / 4, r, (l == 1 && g_h1 != nullptr) ? g_h1 + c0 : nullptr);
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.acc_empty[buf]));
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars.act_ready));
    }
    // last layer: column max over the tile's rows, bias and ReLU after the max (both monotone)
    int* gmax = reinterpret_cast<int*>(fa.pooled) + (size_t)b * 1024 + br * 512;
    const float* bias4 = net.conv_bias[br][4];
#pragma unroll 1
    for (int nb = 0; nb < 4; ++nb) {
      const int j = 4 + nb, buf = nb & 1;
      float b4[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) b4[c] = __ldg(bias4 + nb * 128 + c * 32 + lane);
      mbar_wait(smem_u32(&bars.acc_full[buf]), (uint32_t)(j >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tlane + buf * 128 + c * 32, v);
        tmem_ld_wait();
        float m[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) m[i] = valid ? __uint_as_float(v[i]) : -INFINITY;
        // butterfly: after the step with distance d each lane keeps the half of its columns selected by bit d of the lane
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
          const bool upper = (lane & d) != 0;
#pragma unroll
          for (int i = 0; i < d; ++i) {
            const float send = upper ? m[i] : m[i + d];
            const float keep = upper ? m[i + d] : m[i];
            m[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, d));
          }
        }
        const int col = nb * 128 + c * 32 + lane;    // lane L ends up with column c*32 + L
        atomicMax(gmax + col, __float_as_int(fmaxf(m[0] + b4[c], 0.f)));
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.acc_empty[buf]));
    }
  } else if (warp == 4) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc64 = make_idesc_tf32(128, 64), idesc128 = make_idesc_tf32(128, 128);
      int chunk = 0;
      auto next_chunk = [&](uint32_t a_hi, uint32_t a_lo, int Nc, int Kc, uint32_t d, uint32_t idesc, bool first) {
        const int slot = chunk % 3;
        mbar_wait(smem_u32(&bars.full[slot]), (uint32_t)(chunk / 3) & 1u);
        tcgen05_fence_after();
        mma_chunk(a_hi, a_lo, ring_u32 + slot * kSlotBytes, Nc, Kc, d, idesc, first);
        umma_commit(smem_u32(&bars.empty[slot]));
        ++chunk;
      };
      const uint32_t R0 = act_u32, R1 = act_u32 + kR1 * 4;
      for (int l = 0; l < 4; ++l) {
        const int buf = l & 1;
        mbar_wait(smem_u32(&bars.act_ready), (uint32_t)l & 1u);
        mbar_wait(smem_u32(&bars.acc_empty[buf]), ((uint32_t)(l >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t d = tmem + buf * 128;
        if (l == 0) next_chunk(R0, R0 + kLoX * 4, 64, 16, d, idesc64, true);
        else if (l == 1) next_chunk(R1, R1 + kLo64 * 4, 64, 64, d, idesc64, true);
        else if (l == 2) next_chunk(R0, R0 + kLo64 * 4, 64, 64, d, idesc64, true);
        else {
          next_chunk(R1, R1 + kLo64 * 4, 128, 32, d, idesc128, true);
          next_chunk(R1 + 8 * kKdir, R1 + kLo64 * 4 + 8 * kKdir, 128, 32, d, idesc128, false);
        }
        umma_commit(smem_u32(&bars.acc_full[buf]));
      }
      mbar_wait(smem_u32(&bars.act_ready), 0u);        // h3 (fifth completion of act_ready)
      tcgen05_fence_after();
      for (int nb = 0; nb < 4; ++nb) {
        const int j = 4 + nb, buf = nb & 1;
        mbar_wait(smem_u32(&bars.acc_empty[buf]), ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        for (int kc = 0; kc < 4; ++kc)
          next_chunk(act_u32 + kc * 8 * kKdir, act_u32 + kLo128 * 4 + kc * 8 * kKdir, 128, 32, tmem + buf * 128, idesc128, kc == 0);
        umma_commit(smem_u32(&bars.acc_full[buf]));
      }
    }
  } else if (warp == 5) {
    // ===================================================================== weight loader
    if (lane == 0) {
      const float* img = net.branch_img[br];
      size_t off = 0;
      for (int i = 0; i < kBranchChunks; ++i) {
        const int slot = i % 3;
        const uint32_t bytes = (i == 0) ? 8192u : kSlotBytes;
        mbar_wait(smem_u32(&bars.empty[slot]), ((uint32_t)(i / 3) & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(&bars.full[slot]), bytes);
        bulk_g2s(ring_u32 + slot * kSlotBytes, img + off, bytes, smem_u32(&bars.full[slot]));
        off += bytes / 4;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) tc_inval_barriers(bars);
}

// ------------------------------------------------------------------------------------------------------ pooled projection
// gproj[b][h][cb*64 + c] = bias0_h[c] + sum_k pooled[b][k] * W0g_h[k][c], k summed in 8 groups of 128 combined in fixed
// order (deterministic).  Needs blockDim.x == 512 and 4 KB + 2 KB of scratch (float sP[1024], float sR[8][64]).
__device__ __forceinline__ void tc_gproj_block(const TcNet& net, const ForwardArgs& fa, int b, int h, int cb, float* sP, float* sR) {
  const int tid = threadIdx.x, col = tid & 63, kg = tid >> 6;
  for (int i = tid; i < 1024; i += 512) sP[i] = __ldcg(fa.pooled + (size_t)b * 1024 + i);
  __syncthreads();
  const float* W = net.W0g[h] + (size_t)(kg * 128) * 256 + cb * 64 + col;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 4
  for (int k = 0; k < 128; k += 4) {
    acc0 = fmaf(sP[kg * 128 + k + 0], __ldg(W + (size_t)(k + 0) * 256), acc0);
    acc1 = fmaf(sP[kg * 128 + k + 1], __ldg(W + (size_t)(k + 1) * 256), acc1);
    acc2 = fmaf(sP[kg * 128 + k + 2], __ldg(W + (size_t)(k + 2) * 256), acc2);
    acc3 = fmaf(sP[kg * 128 + k + 3], __ldg(W + (size_t)(k + 3) * 256), acc3);
  }
  sR[kg * 64 + col] = (acc0 + acc1) + (acc2 + acc3);
  __syncthreads();
  if (tid < 64) {
    float s = __ldg(net.head_bias0[h] + cb * 64 + tid);
#pragma unroll
    for (int g = 0; g < 8; ++g) s += sR[g * 64 + tid];
    fa.gproj[((size_t)b * 2 + h) * 256 + cb * 64 + tid] = s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------ head tile
// [gproj row as bias] + h1 . W0[1024:] -> ReLU -> 256 -> 128 -> ReLU -> 2 (learn_region_grow_util.py:138-162).
__device__ __forceinline__ void tc_head_tile(const TcNet& net, const ForwardArgs& fa, int b, int h, int tile,
                                             unsigned char* smem, TcStatic& st, uint32_t tmem) {
  const int n = fa.n_pts[h];
  const int row0 = tile * 128;
  const int rows = min(128, n - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TcBarriers& bars = st.bars;
  float* const act = reinterpret_cast<float*>(smem);
  const uint32_t act_u32 = smem_u32(smem);
  const uint32_t ring_u32 = act_u32 + kActBytes;
  constexpr int kA0 = 0, kC = 16384, kLo64 = 8192;
  float* const sG = st.vec;            // [256] bias0 + pooled . W0[:1024]
  float* const sB1 = st.vec + 256;     // [128]
  float* const sW2 = st.vec + 384;     // [128][2]
  float* const sB2 = st.vec + 640;     // [2]

  if (tid == 0) tc_init_barriers(bars, true);
  __syncthreads();

  if (warp < 4) {
    const int r = tid;
    const bool valid = r < rows;
    {
      const float4* hrow = reinterpret_cast<const float4*>(fa.h1[h] + ((size_t)b * n + row0 + r) * 64);
#pragma unroll 4
      for (int q = 0; q < 16; ++q) {
        float4 x = valid ? __ldcg(hrow + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 hi, lo;
        split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(act + kA0 + q * 512 + r * 4) = hi;
        *reinterpret_cast<float4*>(act + kA0 + kLo64 + q * 512 + r * 4) = lo;
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars.act_ready));
    }
    mbar_wait(smem_u32(&bars.act_ready), 0u);          // st.vec is complete as well
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int nb = 0; nb < 4; ++nb) {
      const int buf = nb & 1;
      mbar_wait(smem_u32(&bars.acc_full[buf]), (uint32_t)(nb >> 1) & 1u);
      if (nb >= 1) mbar_wait(smem_u32(&bars.c_free), (uint32_t)(nb - 1) & 1u);   // H1(nb-1) has consumed the C buffer
      tcgen05_fence_after();
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + buf * 64 + c0, v);
        tmem_ld_wait();
        store_act32(v, sG + nb * 64 + c0, act + kC, act + kC + kLo64, c0 / 4, r, nullptr);
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.acc_empty[buf]));
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars.c_ready));
    }
    // hidden layer 2 (+bias, ReLU) and the 128 -> 2 output layer in registers (util.py:145-149 / :158-162)
    mbar_wait(smem_u32(&bars.acc1_full), 0u);
    tcgen05_fence_after();
    float o0 = sB2[0], o1 = sB2[1];
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tlane + 128 + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = fmaxf(__uint_as_float(v[i]) + sB1[c0 + i], 0.f);
        const float2 w = *reinterpret_cast<const float2*>(sW2 + 2 * (c0 + i));
        o0 = fmaf(x, w.x, o0);
        o1 = fmaf(x, w.y, o1);
      }
    }
    if (valid) *reinterpret_cast<float2*>(fa.logits[h] + ((size_t)b * n + row0 + r) * 2) = make_float2(o0, o1);
  } else if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc64 = make_idesc_tf32(128, 64), idesc128 = make_idesc_tf32(128, 128);
      int chunk = 0;
      auto next_chunk = [&](uint32_t a_hi, uint32_t a_lo, int Nc, int Kc, uint32_t d, uint32_t idesc, bool first) {
        const int slot = chunk % 3;
        mbar_wait(smem_u32(&bars.full[slot]), (uint32_t)(chunk / 3) & 1u);
        tcgen05_fence_after();
        mma_chunk(a_hi, a_lo, ring_u32 + slot * kSlotBytes, Nc, Kc, d, idesc, first);
        umma_commit(smem_u32(&bars.empty[slot]));
        ++chunk;
      };
      const uint32_t A0 = act_u32 + kA0 * 4, Cb = act_u32 + kC * 4;
      auto H0 = [&](int nb) {
        const int buf = nb & 1;
        mbar_wait(smem_u32(&bars.acc_empty[buf]), ((uint32_t)(nb >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        next_chunk(A0, A0 + kLo64 * 4, 64, 64, tmem + buf * 64, idesc64, true);
        umma_commit(smem_u32(&bars.acc_full[buf]));
      };
      auto H1 = [&](int kc) {
        mbar_wait(smem_u32(&bars.c_ready), (uint32_t)kc & 1u);
        tcgen05_fence_after();
        next_chunk(Cb, Cb + kLo64 * 4, 128, 32, tmem + 128, idesc128, kc == 0);
        next_chunk(Cb + 8 * kKdir, Cb + kLo64 * 4 + 8 * kKdir, 128, 32, tmem + 128, idesc128, false);
        umma_commit(smem_u32(&bars.c_free));
        if (kc == 3) umma_commit(smem_u32(&bars.acc1_full));
      };
      mbar_wait(smem_u32(&bars.act_ready), 0u);
      tcgen05_fence_after();
      H0(0); H0(1); H1(0); H0(2); H1(1); H0(3); H1(2); H1(3);
    }
  } else if (warp == 5) {
    // loader warp: first fill the ring, then stage the small vectors, then keep the ring fed
    const float* img = net.head_img[h];
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) {
        mbar_expect_tx(smem_u32(&bars.full[i]), kSlotBytes);
        bulk_g2s(ring_u32 + i * kSlotBytes, img + (size_t)i * (kSlotBytes / 4), kSlotBytes, smem_u32(&bars.full[i]));
      }
    }
    const float* g = fa.gproj + ((size_t)b * 2 + h) * 256;
    for (int c = lane; c < 256; c += 32) sG[c] = __ldcg(g + c);
    for (int c = lane; c < 128; c += 32) sB1[c] = __ldg(net.head_bias1[h] + c);
    for (int c = lane; c < 256; c += 32) sW2[c] = __ldg(net.head_W2[h] + c);
    if (lane < 2) sB2[lane] = __ldg(net.head_bias2[h] + lane);
    mbar_arrive(smem_u32(&bars.act_ready));
    if (lane == 0) {
      for (int i = 3; i < kHeadChunks; ++i) {
        const int slot = i % 3;
        mbar_wait(smem_u32(&bars.empty[slot]), ((uint32_t)(i / 3) & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(&bars.full[slot]), kSlotBytes);
        bulk_g2s(ring_u32 + slot * kSlotBytes, img + (size_t)i * (kSlotBytes / 4), kSlotBytes, smem_u32(&bars.full[slot]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) tc_inval_barriers(bars);
}

}  // namespace lrg
