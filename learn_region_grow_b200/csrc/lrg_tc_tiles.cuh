// Device-side bodies of the tensor-core (tcgen05) LrgNet forward: one 128-point tile of one branch / one head, and one
// 64-column block of the pooled projection.  Shared by the stand-alone kernels of the lock-step loop
// (lrg_forward_tc.cu) and by the persistent grow kernel (lrg_persistent.cu).  See lrg_forward_tc.cu for the design.
//
// Calling convention: every thread of the CTA calls the function (blockDim.x >= 192, a multiple of 32); warps 0-3 are the
// epilogue warps (warp w owns TMEM lanes 32w..32w+31 = tile rows), warp 4 issues the MMAs, warp 5 streams the weights,
// further warps only take part in the CTA barriers.  `tmem` is the base of 512 allocated TMEM columns.  Data produced by
// other CTAs during the same launch (tiles, h1, pooled, gproj) is read with ld.global.cg, never through the
// non-coherent path.
#pragma once
#include "lrg_common.cuh"
#include "lrg_tc.cuh"
#include "lrg_umma.cuh"

namespace lrg {

using namespace umma;

constexpr uint32_t kSlotBytes = 32768;
constexpr int kRingSlots = 6;                          // weight-operand ring: the only dynamic shared memory the tiles use
constexpr uint32_t kTcSmem = kRingSlots * kSlotBytes;
constexpr uint32_t kMNdir = 128;                       // bytes between 8-row groups of a K-major operand image
constexpr uint32_t kTmemCols = 512;
// Tensor-memory map (columns).  Activations are the A operand of the next layer and live in TMEM (tcgen05.mma with A in
// tensor memory: A[m][k] = lane m, column base + k; settled by tools/umma_probe.cu), written in place by the epilogue
// with tcgen05.st: only the weights (B operand) go through shared memory, which halves the shared-memory traffic per
// MMA and leaves the whole 192 KB ring to the weight stream.
//   branch: [0,128) / [128,256) accumulator ping-pong; [256,384) activation hi; [384,512) activation lo
//   head:   [0,64) / [64,128) layer-0 accumulators; [128,256) layer-1 accumulator; [256,320) / [320,384) h1 tile hi / lo;
//           [384,448) / [448,512) 64-channel slice of the hidden layer hi / lo
constexpr uint32_t kTmAhi = 256, kTmAlo = 384;
constexpr uint32_t kTmH1hi = 256, kTmH1lo = 320, kTmChi = 384, kTmClo = 448;
constexpr float kF16Limit = 65000.f;                   // 3xFP16: activations must stay below this (fp16 max = 65504)

struct TcBarriers {
  uint64_t full[kRingSlots], empty[kRingSlots];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t act_ready;               // branch: activations of the next layer written; head: h1 tile written
  uint64_t c_ready[2], c_free[2];   // head only: the two 32-channel halves of the hidden-layer slice
  uint64_t acc1_full;               // head only
  uint64_t vec_ready;               // head only: vec[] (pooled projection row, bias1, W2, bias2) staged
};
constexpr int kTcBarrierCount = sizeof(TcBarriers) / 8;

// Small per-CTA scratch next to the barriers: biases (branch: layers 0-3 = 320 floats; head: gproj row 256 + bias1 128 +
// W2 256 + bias2 2 = 642 floats).
struct TcStatic {
  TcBarriers bars;
  alignas(16) float vec[648];
};

// diagnostics: thread 0 adds the cycles since *t to dbg[stage] and restarts the interval
__device__ __forceinline__ void tc_stamp(unsigned long long* dbg, int stage, long long& t) {
  if (dbg != nullptr && threadIdx.x == 0) {
    const long long now = clock64();
    atomicAdd(dbg + stage, (unsigned long long)(now - t));
    t = now;
  }
}

__device__ __forceinline__ void mbar_inval(uint32_t bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

// One weight chunk = hi image + lo image of an [Nc x Kc] K-major operand in shared memory; the activations' hi / lo parts
// sit in TMEM columns a_hi.. / a_lo..:  D[tmem] (+)= A(hi,lo)[128 x Kc] . chunk^T as hi.hi + lo.hi + hi.lo.
// F16 = false: kind::tf32, one 32-bit TMEM column and 4 bytes of the image per element, K = 8 per instruction;
// F16 = true:  kind::f16, two elements per TMEM column (packed fp16 pairs) and 2 bytes per element, K = 16 per instruction
// -- either way one instruction consumes 8 TMEM columns of A and two K-adjacent core matrices of B, so the descriptor
// arithmetic is the same; the fp16 form needs half the instructions (and half the weight bytes) for the same K.
// Fully unrolled with compile-time shapes: the single issuing thread must sustain one tcgen05.mma per <= 32..64 cycles,
// so per instruction there is only a descriptor add (the start-address field advances by a constant) left to do.
template <bool F16, int Nc, int Kc>
__device__ __forceinline__ void mma_chunk(uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t d_tmem, bool first) {
  constexpr uint32_t kdirB = (uint32_t)Nc * 16;                 // bytes between K-adjacent core matrices of the image
  constexpr uint32_t idesc = (F16 ? 0u : ((2u << 7) | (2u << 10))) | (1u << 4) | ((uint32_t)(Nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int KS = F16 ? Kc / 16 : Kc / 8;
  const uint64_t bd_hi = make_desc(b_hi, kdirB, kMNdir);
  const uint64_t bd_lo = make_desc(b_hi + (uint32_t)(Nc * Kc * (F16 ? 2 : 4)), kdirB, kMNdir);
#pragma unroll
  for (int term = 0; term < 3; ++term) {
    const uint32_t a0 = (term == 1) ? a_lo : a_hi;
    const uint64_t bd = (term == 2) ? bd_lo : bd_hi;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      if (F16) umma_f16_ts(d_tmem, a0 + ks * 8, bd + (uint64_t)((ks * 2 * kdirB) >> 4), idesc, (term == 0 && ks == 0 && first) ? 0u : 1u);
      else umma_tf32_ts(d_tmem, a0 + ks * 8, bd + (uint64_t)((ks * 2 * kdirB) >> 4), idesc, (term == 0 && ks == 0 && first) ? 0u : 1u);
    }
  }
}

// hi = x rounded to TF32; lo = x - hi (exact; the tensor core truncates it to TF32, an error of 2^-23 relative to x)
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(__fsub_rn(x, __uint_as_float(hi)));
}

// scale + bias + ReLU + hi/lo split of 32 accumulator columns (channels c0..c0+31) of this thread's row, stored as the A
// operand of the next layer: TMEM columns t_hi + c0.. / t_lo + c0.. (3xTF32, one column per channel) or t_hi + c0/2.. /
// t_lo + c0/2.. (3xFP16, packed pairs); optionally mirrored to global memory as fp32.  inv = 1 for 3xTF32 (the fmaf is then
// an exactly rounded add).  over: set when a value leaves the fp16 range.
template <bool F16>
__device__ __forceinline__ void store_act32(uint32_t (&v)[32], const float* s_bias, float inv, uint32_t t_hi, uint32_t t_lo, int c0,
                                            float* g_row, bool& over) {
  float x[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(s_bias + q * 4);
    x[q * 4 + 0] = fmaxf(fmaf(__uint_as_float(v[q * 4 + 0]), inv, b.x), 0.f);
    x[q * 4 + 1] = fmaxf(fmaf(__uint_as_float(v[q * 4 + 1]), inv, b.y), 0.f);
    x[q * 4 + 2] = fmaxf(fmaf(__uint_as_float(v[q * 4 + 2]), inv, b.z), 0.f);
    x[q * 4 + 3] = fmaxf(fmaf(__uint_as_float(v[q * 4 + 3]), inv, b.w), 0.f);
    if (g_row != nullptr) *reinterpret_cast<float4*>(g_row + q * 4) = make_float4(x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
  }
  if (F16) {
    uint32_t hi[16], lo[16];
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      mx = fmaxf(mx, fmaxf(x[2 * i], x[2 * i + 1]));
      split_f16x2(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
    }
    over |= !(mx < kF16Limit);                    // (also catches NaN)
    tmem_st16(t_hi + (uint32_t)(c0 >> 1), hi);
    tmem_st16(t_lo + (uint32_t)(c0 >> 1), lo);
  } else {
    uint32_t lo[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) split_fast(x[i], v[i], lo[i]);
    tmem_st32(t_hi + (uint32_t)c0, v);
    tmem_st32(t_lo + (uint32_t)c0, lo);
  }
}

__device__ __forceinline__ void tc_init_barriers(TcBarriers& bars, bool head) {
  for (int i = 0; i < kRingSlots; ++i) { mbar_init(smem_u32(&bars.full[i]), 1); mbar_init(smem_u32(&bars.empty[i]), 1); }
  for (int i = 0; i < 2; ++i) {
    mbar_init(smem_u32(&bars.acc_full[i]), 1);
    mbar_init(smem_u32(&bars.acc_empty[i]), 128);
    mbar_init(smem_u32(&bars.c_ready[i]), 128);
    mbar_init(smem_u32(&bars.c_free[i]), 1);
  }
  mbar_init(smem_u32(&bars.act_ready), 128);
  mbar_init(smem_u32(&bars.acc1_full), 1);
  mbar_init(smem_u32(&bars.vec_ready), 32);
  (void)head;
  fence_barrier_init();
}
__device__ __forceinline__ void tc_inval_barriers(TcBarriers& bars) {
  uint64_t* p = reinterpret_cast<uint64_t*>(&bars);
  for (int i = 0; i < kTcBarrierCount; ++i) mbar_inval(smem_u32(p + i));
}

// Chunk c of a branch / head operand image: size and offset in bytes (lrg_tc.cuh has the chunk lists).
template <bool F16> __device__ __forceinline__ uint32_t branch_chunk_bytes(int c) {
  if (F16) return c == 0 ? 4096u : c <= 2 ? 16384u : 32768u;
  return c == 0 ? 8192u : 32768u;
}
template <bool F16> __device__ __forceinline__ uint32_t branch_chunk_off(int c) {
  if (F16) return c == 0 ? 0u : c == 1 ? 4096u : c == 2 ? 20480u : 36864u + (uint32_t)(c - 3) * 32768u;
  return c == 0 ? 0u : 8192u + (uint32_t)(c - 1) * 32768u;
}
template <bool F16> __device__ __forceinline__ uint32_t head_chunk_bytes() { return F16 ? 16384u : 32768u; }

// Weight loader (one thread): streams the operand-image chunks [first, n_head) and [gap_to, gap_to + n_tail) through the
// ring in that order (the ring position is the running count).
template <bool F16, bool HEAD>
__device__ __forceinline__ void tc_stream_weights(TcBarriers& bars, uint32_t ring_u32, const unsigned char* img, int first, int n_head,
                                                  int gap_to, int n_tail) {
  int i = first;                                   // position in the ring sequence
  const int total = n_head + n_tail;
  for (; i < total; ++i) {
    const int c = i < n_head ? i : gap_to + (i - n_head);
    const int slot = i % kRingSlots;
    const uint32_t bytes = HEAD ? head_chunk_bytes<F16>() : branch_chunk_bytes<F16>(c);
    const uint32_t off = HEAD ? (uint32_t)c * head_chunk_bytes<F16>() : branch_chunk_off<F16>(c);
    mbar_wait(smem_u32(&bars.empty[slot]), ((uint32_t)(i / kRingSlots) & 1u) ^ 1u);
    mbar_expect_tx(smem_u32(&bars.full[slot]), bytes);
    bulk_g2s(ring_u32 + slot * kSlotBytes, img + off, bytes, smem_u32(&bars.full[slot]));
  }
}

// ------------------------------------------------------------------------------------------------------ branch tile
// x (rows x F) -> 64 -> 64 -> 64 -> 128 -> 512 -> column max merged into pooled (learn_region_grow_util.py:106-123).
// nb0..nb1: which 128-column blocks of the last layer this call reduces (a tile may be split over several CTAs, each
// recomputing the cheap first four layers, to shorten the critical path when SMs are idle); h1 is written when nb0 == 0.
template <bool F16>
__device__ __forceinline__ void tc_branch_tile(const TcNet& net, const ForwardArgs& fa, int b, int br, int tile, int nvalid,
                                               int nb0, int nb1, unsigned char* smem, TcStatic& st, uint32_t tmem) {
  const int n = fa.n_pts[br];
  const int row0 = tile * 128;
  const int rows = min(128, nvalid - row0);       // rows >= nvalid are padding duplicates: never read, never pooled
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TcBarriers& bars = st.bars;
  const uint32_t ring_u32 = smem_u32(smem);

  long long tstamp = clock64();
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
  tc_stamp(net.dbg, 0, tstamp);

  if (warp < 4) {
    // ===================================================================== epilogue warps: thread = tile row = TMEM lane
    const int r = tid;
    const bool valid = r < rows;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    bool over = false;
    {
      const float* xrow = fa.x[br] + ((size_t)b * n + row0 + r) * fa.x_stride;
      float xv[16];
      if (fa.x_stride == 16) {                       // slot tiles: padded rows, four 128-bit loads
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = valid ? __ldcg(reinterpret_cast<const float4*>(xrow) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          xv[q * 4 + 0] = t.x; xv[q * 4 + 1] = t.y; xv[q * 4 + 2] = t.z; xv[q * 4 + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) xv[c] = (valid && c < net.F) ? __ldcg(xrow + c) : 0.f;
      }
      if (F16) {
        uint32_t hi[16], lo[16];
        float mx = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < 8) { split_f16x2(xv[2 * c], xv[2 * c + 1], hi[c], lo[c]); mx = fmaxf(mx, fmaxf(fabsf(xv[2 * c]), fabsf(xv[2 * c + 1]))); }
          else { hi[c] = 0u; lo[c] = 0u; }
        }
        over |= !(mx < kF16Limit);
        tmem_st16(tlane + kTmAhi, hi);
        tmem_st16(tlane + kTmAlo, lo);
      } else {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c < 16) split_fast(xv[c], hi[c], lo[c]); else { hi[c] = 0u; lo[c] = 0u; }
        }
        tmem_st32(tlane + kTmAhi, hi);
        tmem_st32(tlane + kTmAlo, lo);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.act_ready));
      tc_stamp(net.dbg, 1, tstamp);
    }
    float* g_h1 = (valid && nb0 == 0) ? fa.h1[br] + ((size_t)b * n + row0 + r) * 64 : nullptr;
#pragma unroll 1
    for (int l = 0; l < 4; ++l) {
      const int buf = l & 1;
      mbar_wait(smem_u32(&bars.acc_full[buf]), (uint32_t)(l >> 1) & 1u);
      tcgen05_fence_after();
      const int N = (l == 3) ? 128 : 64;
      const float* s_bias = st.vec + l * 64;
      const float inv = F16 ? net.branch_inv[br][l] : 1.f;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + buf * 128 + c0, v);
        tmem_ld_wait();
        store_act32<F16>(v, s_bias + c0, inv, tlane + kTmAhi, tlane + kTmAlo, c0, (l == 1 && g_h1 != nullptr) ? g_h1 + c0 : nullptr, over);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.acc_empty[buf]));
      mbar_arrive(smem_u32(&bars.act_ready));
      tc_stamp(net.dbg, 2 + l, tstamp);
    }
    // last layer: column max over the tile's rows, bias and ReLU after the max (both monotone)
    int* gmax = reinterpret_cast<int*>(fa.pooled) + (size_t)b * 1024 + br * 512;
    const float* bias4 = net.conv_bias[br][4];
    const float inv4 = F16 ? net.branch_inv[br][4] : 1.f;
#pragma unroll 1
    for (int nb = nb0; nb < nb1; ++nb) {
      const int j = 4 + (nb - nb0), buf = j & 1;
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
        atomicMax(gmax + col, __float_as_int(fmaxf(fmaf(m[0], inv4, b4[c]), 0.f)));
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.acc_empty[buf]));
      tc_stamp(net.dbg, 6 + (nb - nb0), tstamp);
    }
    if (F16 && over && net.range_flag != nullptr) *reinterpret_cast<volatile int*>(net.range_flag) = 1;
  } else if (warp == 4) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      int chunk = 0;
      long long w_full = 0, w_acc = 0, w_act = 0;      // diagnostics: cycles this thread waited for weights / TMEM / activations
      const bool timing = net.dbg != nullptr;
      auto timed_wait = [&](uint32_t bar, uint32_t parity, long long& acc) {
        if (timing) { const long long t0 = clock64(); mbar_wait(bar, parity); acc += clock64() - t0; }
        else mbar_wait(bar, parity);
      };
      auto wait_chunk = [&]() {
        const int slot = chunk % kRingSlots;
        timed_wait(smem_u32(&bars.full[slot]), (uint32_t)(chunk / kRingSlots) & 1u, w_full);
        tcgen05_fence_after();
        return ring_u32 + slot * kSlotBytes;
      };
      auto done_chunk = [&]() {
        umma_commit(smem_u32(&bars.empty[chunk % kRingSlots]));
        ++chunk;
      };
      const uint32_t a_hi = tmem + kTmAhi, a_lo = tmem + kTmAlo;
      for (int l = 0; l < 4; ++l) {
        const int buf = l & 1;
        timed_wait(smem_u32(&bars.act_ready), (uint32_t)l & 1u, w_act);
        timed_wait(smem_u32(&bars.acc_empty[buf]), ((uint32_t)(l >> 1) & 1u) ^ 1u, w_acc);
        tcgen05_fence_after();
        const uint32_t d = tmem + buf * 128;
        if (l == 0) { mma_chunk<F16, 64, 16>(a_hi, a_lo, wait_chunk(), d, true); done_chunk(); }
        else if (l < 3) { mma_chunk<F16, 64, 64>(a_hi, a_lo, wait_chunk(), d, true); done_chunk(); }
        else if (F16) { mma_chunk<true, 128, 64>(a_hi, a_lo, wait_chunk(), d, true); done_chunk(); }
        else {
          mma_chunk<false, 128, 32>(a_hi, a_lo, wait_chunk(), d, true); done_chunk();
          mma_chunk<false, 128, 32>(a_hi + 32, a_lo + 32, wait_chunk(), d, false); done_chunk();
        }
        umma_commit(smem_u32(&bars.acc_full[buf]));
      }
      timed_wait(smem_u32(&bars.act_ready), 0u, w_act);   // h3 (fifth completion of act_ready)
      tcgen05_fence_after();
      for (int nb = nb0; nb < nb1; ++nb) {
        const int j = 4 + (nb - nb0), buf = j & 1;
        timed_wait(smem_u32(&bars.acc_empty[buf]), ((uint32_t)(j >> 1) & 1u) ^ 1u, w_acc);
        tcgen05_fence_after();
        // (a chunk covers 32 TMEM columns of A in both kinds: 32 channels as tf32, 64 channels as packed fp16 pairs)
#pragma unroll 1
        for (int kc = 0; kc < (F16 ? 2 : 4); ++kc) {
          mma_chunk<F16, 128, F16 ? 64 : 32>(a_hi + kc * 32, a_lo + kc * 32, wait_chunk(), tmem + buf * 128, kc == 0);
          done_chunk();
        }
        umma_commit(smem_u32(&bars.acc_full[buf]));
      }
      if (timing) {
        atomicAdd(net.dbg + 11, (unsigned long long)w_full);
        atomicAdd(net.dbg + 12, (unsigned long long)w_acc);
        atomicAdd(net.dbg + 13, (unsigned long long)w_act);
      }
    }
  } else if (warp == 5) {
    // ===================================================================== weight loader
    if (lane == 0) {
      if (F16) tc_stream_weights<true, false>(bars, ring_u32, net.branch_img[1][br], 0, 4, 4 + 2 * nb0, 2 * (nb1 - nb0));
      else tc_stream_weights<false, false>(bars, ring_u32, net.branch_img[0][br], 0, 5, 5 + 4 * nb0, 4 * (nb1 - nb0));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) tc_inval_barriers(bars);
  tc_stamp(net.dbg, 10, tstamp);
  if (net.dbg != nullptr && tid == 0) atomicAdd(net.dbg + 15, 1ull);
}

// ------------------------------------------------------------------------------------------------------ pooled projection
// gproj[b][h][cb*64 + c] = bias0_h[c] + sum_k pooled[b][k] * W0g_h[k][c]: 512 threads = 16 column quads x 32 K-groups of
// 32 rows; every thread keeps 16-32 128-bit weight loads in flight (the item is ~two L2 round trips + 256 KB of traffic), partial sums are combined in fixed order (deterministic).  Scratch: float sP[1024], float sR[32][64].
__device__ __forceinline__ void tc_gproj_block(const TcNet& net, const ForwardArgs& fa, int b, int h, int cb, float* sP, float* sR) {
  const int tid = threadIdx.x, c4 = tid & 15, kg = tid >> 4;
  const float4* W = reinterpret_cast<const float4*>(net.W0g[h] + (size_t)(kg * 32) * 256 + cb * 64) + c4;
  float4 w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = __ldg(W + (size_t)i * 64);
  for (int i = tid; i < 1024; i += 512) sP[i] = __ldcg(fa.pooled + (size_t)b * 1024 + i);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float4 wn[16];
    if (half == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) wn[i] = __ldg(W + (size_t)(16 + i) * 64);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float p = sP[kg * 32 + half * 16 + i];
      acc.x = fmaf(p, w[i].x, acc.x); acc.y = fmaf(p, w[i].y, acc.y); acc.z = fmaf(p, w[i].z, acc.z); acc.w = fmaf(p, w[i].w, acc.w);
    }
    if (half == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = wn[i];
    }
  }
  *reinterpret_cast<float4*>(sR + kg * 64 + c4 * 4) = acc;
  __syncthreads();
  if (tid < 64) {
    float s = __ldg(net.head_bias0[h] + cb * 64 + tid);
#pragma unroll
    for (int g = 0; g < 32; ++g) s += sR[g * 64 + tid];
    fa.gproj[((size_t)b * 2 + h) * 256 + cb * 64 + tid] = s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------ head tile
// [gproj row as bias] + h1 . W0[1024:] -> ReLU -> 256 -> 128 -> ReLU -> 2 (learn_region_grow_util.py:138-162).
// The 256-wide hidden layer is produced as four 64-channel slices; each slice is handed to the 256->128 layer as two
// 32-channel K-chunks the moment its epilogue has written them (c_ready / c_free per half).
// gproj_pending: optional counter that reaches 0 when this tile pair's pooled projection has been written (the
// persistent kernel publishes head tiles together with the projection blocks so that the tile's prologue and first MMAs
// overlap them); NULL = the projection is already there.  gproj_tagged (persistent kernel with projection servers): the
// 256 {value, tag} words of this (slot, head) instead of fa.gproj; a word is valid once its tag equals `tag`.
template <bool F16>
__device__ __forceinline__ void tc_head_tile(const TcNet& net, const ForwardArgs& fa, int b, int h, int tile, int nvalid,
                                             const int* gproj_pending, const uint2* gproj_tagged, unsigned tag, unsigned char* smem,
                                             TcStatic& st, uint32_t tmem) {
  const int n = fa.n_pts[h];
  const int row0 = tile * 128;
  const int rows = min(128, nvalid - row0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TcBarriers& bars = st.bars;
  const uint32_t ring_u32 = smem_u32(smem);
  float* const sG = st.vec;            // [256] bias0 + pooled . W0[:1024]
  float* const sB1 = st.vec + 256;     // [128]
  float* const sW2 = st.vec + 384;     // [128][2]
  float* const sB2 = st.vec + 640;     // [2]

  long long tstamp = clock64();
  if (tid == 0) tc_init_barriers(bars, true);
  __syncthreads();
  tc_stamp(net.dbg, 16, tstamp);

  if (warp < 4) {
    const int r = tid;
    const bool valid = r < rows;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    {
      const float4* hrow = reinterpret_cast<const float4*>(fa.h1[h] + ((size_t)b * n + row0 + r) * 64);
      float4 x[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) x[q] = valid ? __ldcg(hrow + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (F16) {
        // (h1 was produced by a branch tile of this kind, which already checked its range)
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          split_f16x2(x[q].x, x[q].y, hi[q * 2 + 0], lo[q * 2 + 0]);
          split_f16x2(x[q].z, x[q].w, hi[q * 2 + 1], lo[q * 2 + 1]);
        }
        tmem_st32(tlane + kTmH1hi, hi);
        tmem_st32(tlane + kTmH1lo, lo);
      } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = x[half * 8 + q];
            split_fast(v.x, hi[q * 4 + 0], lo[q * 4 + 0]); split_fast(v.y, hi[q * 4 + 1], lo[q * 4 + 1]);
            split_fast(v.z, hi[q * 4 + 2], lo[q * 4 + 2]); split_fast(v.w, hi[q * 4 + 3], lo[q * 4 + 3]);
          }
          tmem_st32(tlane + kTmH1hi + half * 32, hi);
          tmem_st32(tlane + kTmH1lo + half * 32, lo);
        }
      }
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bars.act_ready));
    }
    mbar_wait(smem_u32(&bars.vec_ready), 0u);          // st.vec is complete
    tc_stamp(net.dbg, 17, tstamp);
    bool over = false;
    const float inv0 = F16 ? net.head_inv[h][0] : 1.f, inv1 = F16 ? net.head_inv[h][1] : 1.f;
#pragma unroll 1
    for (int nb = 0; nb < 4; ++nb) {
      const int buf = nb & 1;
      mbar_wait(smem_u32(&bars.acc_full[buf]), (uint32_t)(nb >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(tlane + buf * 64 + half * 32, v);
        tmem_ld_wait();
        if (nb >= 1) {                                 // the previous slice's K-chunk of this half has been consumed
          mbar_wait(smem_u32(&bars.c_free[half]), (uint32_t)(nb - 1) & 1u);
          tcgen05_fence_after();
        }
        store_act32<F16>(v, sG + nb * 64 + half * 32, inv0, tlane + kTmChi, tlane + kTmClo, half * 32, nullptr, over);
        tmem_st_wait();
        tcgen05_fence_before();
        if (half == 1) mbar_arrive(smem_u32(&bars.acc_empty[buf]));
        mbar_arrive(smem_u32(&bars.c_ready[half]));
      }
      tc_stamp(net.dbg, 18 + nb, tstamp);
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
        const float x = fmaxf(fmaf(__uint_as_float(v[i]), inv1, sB1[c0 + i]), 0.f);
        const float2 w = *reinterpret_cast<const float2*>(sW2 + 2 * (c0 + i));
        o0 = fmaf(x, w.x, o0);
        o1 = fmaf(x, w.y, o1);
      }
    }
    if (valid) *reinterpret_cast<float2*>(fa.logits[h] + ((size_t)b * n + row0 + r) * 2) = make_float2(o0, o1);
    if (F16 && over && net.range_flag != nullptr) *reinterpret_cast<volatile int*>(net.range_flag) = 1;
    tc_stamp(net.dbg, 22, tstamp);
  } else if (warp == 4) {
    if (lane == 0) {
      int chunk = 0;
      auto wait_chunk = [&]() {
        const int slot = chunk % kRingSlots;
        mbar_wait(smem_u32(&bars.full[slot]), (uint32_t)(chunk / kRingSlots) & 1u);
        tcgen05_fence_after();
        return ring_u32 + slot * kSlotBytes;
      };
      auto done_chunk = [&]() {
        umma_commit(smem_u32(&bars.empty[chunk % kRingSlots]));
        ++chunk;
      };
      auto H0 = [&](int nb) {
        const int buf = nb & 1;
        mbar_wait(smem_u32(&bars.acc_empty[buf]), ((uint32_t)(nb >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        mma_chunk<F16, 64, 64>(tmem + kTmH1hi, tmem + kTmH1lo, wait_chunk(), tmem + buf * 64, true);
        done_chunk();
        umma_commit(smem_u32(&bars.acc_full[buf]));
      };
      auto H1 = [&](int kc) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          mbar_wait(smem_u32(&bars.c_ready[half]), (uint32_t)kc & 1u);
          tcgen05_fence_after();
          constexpr uint32_t hc = F16 ? 16 : 32;             // TMEM columns of a 32-channel half of the hidden slice
          mma_chunk<F16, 128, 32>(tmem + kTmChi + half * hc, tmem + kTmClo + half * hc, wait_chunk(), tmem + 128, kc == 0 && half == 0);
          done_chunk();
          umma_commit(smem_u32(&bars.c_free[half]));
        }
        if (kc == 3) umma_commit(smem_u32(&bars.acc1_full));
      };
      mbar_wait(smem_u32(&bars.act_ready), 0u);
      tcgen05_fence_after();
      H0(0); H0(1); H1(0); H0(2); H1(1); H0(3); H1(2); H1(3);
    }
  } else if (warp == 5) {
    // loader warp: first fill the ring, then stage the small vectors, then keep the ring fed
    const unsigned char* img = net.head_img[F16 ? 1 : 0][h];
    if (lane == 0) {
      for (int i = 0; i < kRingSlots; ++i) {
        mbar_expect_tx(smem_u32(&bars.full[i]), head_chunk_bytes<F16>());
        bulk_g2s(ring_u32 + i * kSlotBytes, img + (size_t)i * head_chunk_bytes<F16>(), head_chunk_bytes<F16>(), smem_u32(&bars.full[i]));
      }
    }
    const float* g = fa.gproj + ((size_t)b * 2 + h) * 256;
    float tg[8], tb[4], tw[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) tb[i] = __ldg(net.head_bias1[h] + lane + 32 * i);
#pragma unroll
    for (int i = 0; i < 8; ++i) tw[i] = __ldg(net.head_W2[h] + lane + 32 * i);
    if (gproj_tagged != nullptr) {
      // answered by the projection servers: every value arrives as one 8-byte {value, tag} word; a lane is done when its eight
      // words carry the tag of this forward
      const long long t0 = clock64();
      unsigned pending = 0xFFu;
      while (pending != 0u) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if ((pending >> i) & 1u) {
            const uint2 x = __ldcg(gproj_tagged + lane + 32 * i);
            if (x.y == tag) { tg[i] = __uint_as_float(x.x); pending &= ~(1u << i); }
          }
        if (pending != 0u) {
          __nanosleep(40);
          if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
        }
      }
    } else {
      if (gproj_pending != nullptr) {
        if (lane == 0) {
          const long long t0 = clock64();
          while (*reinterpret_cast<const volatile int*>(gproj_pending) != 0) {
            __nanosleep(40);
            if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
          }
          __threadfence();
        }
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) tg[i] = __ldcg(g + lane + 32 * i);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sG[lane + 32 * i] = tg[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) sB1[lane + 32 * i] = tb[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) sW2[lane + 32 * i] = tw[i];
    if (lane < 2) sB2[lane] = __ldg(net.head_bias2[h] + lane);
    mbar_arrive(smem_u32(&bars.vec_ready));
    if (lane == 0) tc_stream_weights<F16, true>(bars, ring_u32, img, kRingSlots, kHeadChunks, 0, 0);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) tc_inval_barriers(bars);
  tc_stamp(net.dbg, 23, tstamp);
  if (net.dbg != nullptr && tid == 0) atomicAdd(net.dbg + 31, 1ull);
}

}  // namespace lrg
