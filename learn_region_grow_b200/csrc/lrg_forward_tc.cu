// LrgNet forward on the 5th-generation tensor cores (tcgen05 / TMEM) for the full model
// (/root/reference/learn_region_grow_util.py:77-79: conv 64,64,64,128,512; heads 256,128,2).
//
// Precision: the parity bar is the fp32 TF graph, so every contraction runs as 3xTF32 -- each fp32 operand is split
// into hi (TF32-exact) + lo (the fp32 remainder) and D += hi.hi + lo.hi + hi.lo with fp32 accumulation in TMEM, which
// carries ~21 mantissa bits per product (measured by tools/umma_probe.cu: max error 5e-6 where an fp32 FMA chain has 2e-6).
//
//   lrg_tc_branch_kernel  one CTA per (128-point tile, branch, tile pair): x -> 64 -> 64 -> 64 -> 128 -> 512 -> column max.
//                         Activations never leave the SM: the epilogue warps read the accumulator from TMEM, add bias,
//                         ReLU, split hi/lo and write the next layer's A operand straight into shared memory in the
//                         UMMA canonical K-major layout.  Weights stream from L2 as pre-packed 32 KB operand images
//                         through a 3-slot ring fed by 1-D bulk async copies (TMA without a tensor map).  The 128->512
//                         layer is never materialised: its accumulator (2 x 128 TMEM columns, double buffered) is
//                         reduced to the column max by a warp butterfly and merged with atomicMax.
//   lrg_gproj_partial_kernel  pooled(1024) . W0[:1024] for both heads, split over K into 16 deterministic partials.
//   lrg_tc_head_kernel    one CTA per (128-point tile, head, tile pair): [pooled part via the partials] + h1.W0[1024:]
//                         -> 256 -> 128 -> 2, the 256-wide hidden layer produced 64 channels at a time and consumed as
//                         K-chunks of the next layer so that it fits beside the operand ring.
//
// Warp roles (192 threads): warps 0-3 = epilogue (warp w owns TMEM lanes 32w..32w+31 = tile rows), warp 4 = MMA issuer
// (one elected thread), warp 5 = weight loader (one elected thread).  All hand-offs are mbarriers.
#include "lrg_common.cuh"
#include "lrg_tc.cuh"
#include "lrg_umma.cuh"

namespace lrg {

using namespace umma;

constexpr int kTcThreads = 192;
constexpr uint32_t kSlotBytes = 32768;
constexpr uint32_t kActBytes = 131072;                 // two 64 KB activation regions (or one 128-channel hi/lo pair)
constexpr uint32_t kTcSmem = kActBytes + 3 * kSlotBytes;
constexpr uint32_t kKdir = 2048;                       // bytes between K-adjacent core matrices of a 128-row operand
constexpr uint32_t kMNdir = 128;                       // bytes between 8-row groups
constexpr uint32_t kTmemCols = 256;

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
__device__ __forceinline__ void store_act32(const uint32_t (&v)[32], const float* __restrict__ bias, float* s_hi, float* s_lo,
                                            int chunk0, int r, float* g_row) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 b = __ldg(reinterpret_cast<const float4*>(bias) + q);
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

struct TcBarriers {
  uint64_t full[3], empty[3];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t act_ready;          // branch: activations of the next layer written; head: h1 tile written
  uint64_t c_ready, c_free;    // head only
  uint64_t acc1_full;          // head only
  uint32_t tmem_base;
};

// ------------------------------------------------------------------------------------------------------ branch
__global__ void __launch_bounds__(kTcThreads, 1) lrg_tc_branch_kernel(const __grid_constant__ TcNet net, const __grid_constant__ ForwardArgs fa) {
  const int b = blockIdx.z, br = blockIdx.y, tile = blockIdx.x;
  if (fa.active != nullptr && fa.active[(size_t)b * fa.active_stride] == 0) return;
  const int n = fa.n_pts[br];
  const int row0 = tile * 128;
  if (row0 >= n) return;
  const int rows = min(128, n - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) TcBarriers bars;
  float* const act = reinterpret_cast<float*>(smem);
  const uint32_t act_u32 = smem_u32(smem);
  const uint32_t ring_u32 = act_u32 + kActBytes;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(&bars.full[i]), 1); mbar_init(smem_u32(&bars.empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars.acc_full[i]), 1); mbar_init(smem_u32(&bars.acc_empty[i]), 128); }
    mbar_init(smem_u32(&bars.act_ready), 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&bars.tmem_base), kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = bars.tmem_base;

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
      for (int c = 0; c < 16; ++c) xv[c] = (valid && c < net.F) ? __ldg(xrow + c) : 0.f;
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
      const float* bias = net.conv_bias[br][l];
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + buf * 128 + c0, v);
        tmem_ld_wait();
        store_act32(v, bias + c0, s_hi, s_lo, c0 / 4, r, (l == 1 && g_h1 != nullptr) ? g_h1 + c0 : nullptr);
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.acc_empty[buf]));
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars.act_ready));
    }
    // last layer: column max over the tile's rows, bias and ReLU after the max (both monotone)
    int* gmax = reinterpret_cast<int*>(fa.pooled) + (size_t)b * 1024 + br * 512;
#pragma unroll 1
    for (int nb = 0; nb < 4; ++nb) {
      const int j = 4 + nb, buf = nb & 1;
      mbar_wait(smem_u32(&bars.acc_full[buf]), (uint32_t)(j >> 1) & 1u);
      tcgen05_fence_after();
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + buf * 128 + c0, v);
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
        const int col = nb * 128 + c0 + lane;        // lane L ends up with column c0 + L
        const float p = fmaxf(m[0] + __ldg(net.conv_bias[br][4] + col), 0.f);
        atomicMax(gmax + col, __float_as_int(p));
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
  } else {
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
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------ pooled projection
// part[ks][b][h][c] = sum_{k in split ks} pooled[b][k] * W0g_h[k][c]      (16 splits of 64 rows; summed in order by the heads)
__global__ void __launch_bounds__(256) lrg_gproj_partial_kernel(const __grid_constant__ TcNet net, const __grid_constant__ ForwardArgs fa,
                                                                float* __restrict__ part) {
  __shared__ float sW[64][64];
  __shared__ float sP[16][64];
  const int cb = blockIdx.x, h = blockIdx.y, ks = blockIdx.z;
  const int tid = threadIdx.x, col = tid & 63, g = tid >> 6;
  const float* W = net.W0g[h] + (size_t)(ks * 64) * 256 + cb * 64;
  for (int i = tid; i < 64 * 16; i += 256) {
    const int k = i >> 4, c4 = i & 15;
    *reinterpret_cast<float4*>(&sW[k][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * 256) + c4);
  }
  for (int b0 = 0; b0 < fa.B; b0 += 16) {
    __syncthreads();
    for (int i = tid; i < 16 * 64; i += 256) {
      const int bb = i >> 6, k = i & 63, b = b0 + bb;
      const bool act = b < fa.B && (fa.active == nullptr || fa.active[(size_t)b * fa.active_stride] != 0);
      sP[bb][k] = act ? fa.pooled[(size_t)b * 1024 + ks * 64 + k] : 0.f;
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float w = sW[k][col];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(sP[g * 4 + q][k], w, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int b = b0 + g * 4 + q;
      if (b < fa.B) part[(((size_t)ks * fa.B + b) * 2 + h) * 256 + cb * 64 + col] = acc[q];
    }
  }
}

// ------------------------------------------------------------------------------------------------------ heads
__global__ void __launch_bounds__(kTcThreads, 1) lrg_tc_head_kernel(const __grid_constant__ TcNet net, const __grid_constant__ ForwardArgs fa,
                                                                    const float* __restrict__ part) {
  const int b = blockIdx.z, h = blockIdx.y, tile = blockIdx.x;   // h: 0 = remove head on inlier rows, 1 = add head on neighbor rows
  if (fa.active != nullptr && fa.active[(size_t)b * fa.active_stride] == 0) return;
  const int n = fa.n_pts[h];
  const int row0 = tile * 128;
  if (row0 >= n) return;
  const int rows = min(128, n - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) TcBarriers bars;
  __shared__ __align__(16) float sG[256];            // bias0 + pooled . W0[:1024] of this (tile pair, head)
  float* const act = reinterpret_cast<float*>(smem);
  const uint32_t act_u32 = smem_u32(smem);
  const uint32_t ring_u32 = act_u32 + kActBytes;
  constexpr int kA0 = 0, kC = 16384, kLo64 = 8192;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(&bars.full[i]), 1); mbar_init(smem_u32(&bars.empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars.acc_full[i]), 1); mbar_init(smem_u32(&bars.acc_empty[i]), 128); }
    mbar_init(smem_u32(&bars.act_ready), 128 + 32);   // 128 rows of h1 + the 32 lanes that build sG
    mbar_init(smem_u32(&bars.c_ready), 128);
    mbar_init(smem_u32(&bars.c_free), 1);
    mbar_init(smem_u32(&bars.acc1_full), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&bars.tmem_base), kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp < 4) {
    const int r = tid;
    const bool valid = r < rows;
    {
      const float4* hrow = reinterpret_cast<const float4*>(fa.h1[h] + ((size_t)b * n + row0 + r) * 64);
#pragma unroll 4
      for (int q = 0; q < 16; ++q) {
        float4 x = valid ? hrow[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 hi, lo;
        split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(act + kA0 + q * 512 + r * 4) = hi;
        *reinterpret_cast<float4*>(act + kA0 + kLo64 + q * 512 + r * 4) = lo;
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars.act_ready));
    }
    mbar_wait(smem_u32(&bars.act_ready), 0u);          // sG is complete as well
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
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 g4 = *reinterpret_cast<const float4*>(&sG[nb * 64 + c0 + q * 4]);
          float4 x, hi, lo;
          x.x = fmaxf(__uint_as_float(v[q * 4 + 0]) + g4.x, 0.f);
          x.y = fmaxf(__uint_as_float(v[q * 4 + 1]) + g4.y, 0.f);
          x.z = fmaxf(__uint_as_float(v[q * 4 + 2]) + g4.z, 0.f);
          x.w = fmaxf(__uint_as_float(v[q * 4 + 3]) + g4.w, 0.f);
          split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
          *reinterpret_cast<float4*>(act + kC + (c0 / 4 + q) * 512 + r * 4) = hi;
          *reinterpret_cast<float4*>(act + kC + kLo64 + (c0 / 4 + q) * 512 + r * 4) = lo;
        }
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.acc_empty[buf]));
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars.c_ready));
    }
    // hidden layer 2 (+bias, ReLU) and the 128 -> 2 output layer in registers (util.py:145-149 / :158-162)
    mbar_wait(smem_u32(&bars.acc1_full), 0u);
    tcgen05_fence_after();
    const float* b1 = net.head_bias1[h];
    const float* W2 = net.head_W2[h];
    float o0 = __ldg(net.head_bias2[h]), o1 = __ldg(net.head_bias2[h] + 1);
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tlane + 128 + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = fmaxf(__uint_as_float(v[i]) + __ldg(b1 + c0 + i), 0.f);
        const float2 w = __ldg(reinterpret_cast<const float2*>(W2) + c0 + i);
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
  } else {
    // loader warp: first fill the ring, then fold the pooled projection partials, then keep the ring fed
    const float* img = net.head_img[h];
    int i = 0;
    if (lane == 0) {
      for (; i < 3; ++i) {
        mbar_expect_tx(smem_u32(&bars.full[i]), kSlotBytes);
        bulk_g2s(ring_u32 + i * kSlotBytes, img + (size_t)i * (kSlotBytes / 4), kSlotBytes, smem_u32(&bars.full[i]));
      }
    }
    for (int c = lane; c < 256; c += 32) {
      float s = __ldg(net.head_bias0[h] + c);
      for (int ks = 0; ks < kGprojSplits; ++ks) s += part[(((size_t)ks * fa.B + b) * 2 + h) * 256 + c];
      sG[c] = s;
    }
    mbar_arrive(smem_u32(&bars.act_ready));
    if (lane == 0) {
      for (i = 3; i < kHeadChunks; ++i) {
        const int slot = i % 3;
        mbar_wait(smem_u32(&bars.empty[slot]), ((uint32_t)(i / 3) & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(&bars.full[slot]), kSlotBytes);
        bulk_g2s(ring_u32 + slot * kSlotBytes, img + (size_t)i * (kSlotBytes / 4), kSlotBytes, smem_u32(&bars.full[slot]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------ host
int tc_forward_configure() {
  LRG_CUDA(cudaFuncSetAttribute(lrg_tc_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LRG_CUDA(cudaFuncSetAttribute(lrg_tc_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  return LRG_OK;
}

// pooled must be zero for the active tile pairs on entry; gproj_part holds kGprojSplits * B * 2 * 256 floats.
int launch_forward_tc(const TcNet& net, const ForwardArgs& fa, float* gproj_part, cudaStream_t stream, cudaEvent_t* ev) {
  if (fa.B <= 0) return LRG_OK;
  const int nmax = fa.n_pts[0] > fa.n_pts[1] ? fa.n_pts[0] : fa.n_pts[1];
  dim3 grid((nmax + 127) / 128, 2, fa.B);
  if (ev) cudaEventRecord(ev[0], stream);
  lrg_tc_branch_kernel<<<grid, kTcThreads, kTcSmem, stream>>>(net, fa);
  if (ev) cudaEventRecord(ev[1], stream);
  lrg_gproj_partial_kernel<<<dim3(4, 2, kGprojSplits), 256, 0, stream>>>(net, fa, gproj_part);
  if (ev) cudaEventRecord(ev[2], stream);
  lrg_tc_head_kernel<<<grid, kTcThreads, kTcSmem, stream>>>(net, fa, gproj_part);
  if (ev) cudaEventRecord(ev[3], stream);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg
