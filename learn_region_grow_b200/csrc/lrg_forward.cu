// LrgNet forward for sm_100a: the 16 Conv1D+BiasAdd+ReLU, 2 Max, Tile and ConcatV2 nodes of
// /root/reference/learn_region_grow_util.py:106-162 as three kernels.
//
//   lrg_branch_kernel  one CTA per (128-point tile, branch, tile pair): the whole per-point conv stack with the
//                      activations resident in shared memory, weights streamed from L2 through a 3-stage cp.async
//                      ring, conv[1] written out for the heads and the last layer reduced to its column max on the
//                      fly (never materialised; util.py:122-123).
//   lrg_gproj_kernel   the pooled half of head layer 0, once per tile pair instead of once per point: the reference
//                      tiles the 1024-wide pooled row in front of every point (util.py:128-135) and multiplies it
//                      by kernel0 512 times; algebraically Z.K0 = g.K0[:1024] + h1.K0[1024:].
//   lrg_head_kernel    one CTA per (128-point tile, head, tile pair): 64->256->128->2 with the projected pooled row
//                      as the bias of the first layer (util.py:138-162).
//
// fp32 FMA throughout (the parity bar is the fp32 TF graph); 8x8 register tiles, 128-bit shared-memory loads.
#include "lrg_common.cuh"

namespace lrg {

constexpr int kThreads = 256;
constexpr int kSlice = 16;            // k rows per weight stage
constexpr int kStages = 3;
constexpr int kStageFloats = kSlice * 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[8][CN] += sIn[128 x Kpad] * W[Kpad x N][:, n0 : n0 + 16*CN]; thread (ty, tx) owns rows ty*8.., cols n0 + tx*CN..
template <int CN>
__device__ __forceinline__ void gemm_chunk(const float* __restrict__ sIn, int ldIn, int Kpad,
                                           const float* __restrict__ Wg, int N, int n0, float* sW,
                                           float (&acc)[8][CN], int tid) {
  constexpr int NC = 16 * CN;
  const int ty = tid >> 4, tx = tid & 15;
  const int ns = Kpad / kSlice;
  auto prefetch = [&](int s) {
    float* dst = sW + (s % kStages) * kStageFloats;
    const float* src = Wg + (size_t)(s * kSlice) * N + n0;
#pragma unroll
    for (int c = tid; c < kSlice * (NC / 4); c += kThreads) {
      int r = c / (NC / 4), q = c % (NC / 4);
      cp_async16(dst + r * NC + q * 4, src + (size_t)r * N + q * 4);
    }
    cp_async_commit();
  };
  __syncthreads();   // previous chunk's readers of sW are done and the previous epilogue's smem stores are visible
  prefetch(0);
  if (ns > 1) prefetch(1);
  for (int s = 0; s < ns; ++s) {
    if (s + 1 < ns) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    if (s + 2 < ns) prefetch(s + 2);
    const float* wbuf = sW + (s % kStages) * kStageFloats + tx * CN;
    const float* arow = sIn + (ty * 8) * ldIn + s * kSlice;
#pragma unroll
    for (int kk = 0; kk < kSlice; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(arow + i * ldIn + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[CN];
#pragma unroll
        for (int j = 0; j < CN; j += 4) {
          float4 t = *reinterpret_cast<const float4*>(wbuf + (kk + q) * NC + j);
          w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
#pragma unroll
          for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(av, w[j], acc[i][j]);
        }
      }
    }
  }
}

// One dense layer over the CTA's 128-row tile: sOut = relu(sIn . W + bias0), optionally mirrored to global memory
// (conv[1]) and/or reduced to the column max.  bias0 may point at a per-tile-pair vector (head layer 0).
template <int CN>
__device__ __forceinline__ void dense_layer(const float* sIn, int ldIn, const LayerDesc& L, const float* bias0,
                                            float* sOut, int ldOut, float* sW, int tid, int rows,
                                            float* gOut, int ldG, int* sMax, int* gMax) {
  constexpr int NC = 16 * CN;
  const int ty = tid >> 4, tx = tid & 15;
  for (int n0 = 0; n0 < L.N; n0 += NC) {
    float acc[8][CN];
#pragma unroll
    for (int j = 0; j < CN; ++j) {
      float bv = bias0[n0 + tx * CN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = bv;
    }
    if (gMax != nullptr && tid < 128) sMax[tid] = 0;
    gemm_chunk<CN>(sIn, ldIn, L.Kpad, L.W, L.N, n0, sW, acc, tid);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < CN; ++j) acc[i][j] = fmaxf(acc[i][j], 0.f);
    if (sOut != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CN; j += 4)
          *reinterpret_cast<float4*>(sOut + (ty * 8 + i) * ldOut + n0 + tx * CN + j) =
              make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
    }
    if (gOut != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (ty * 8 + i < rows) {
#pragma unroll
          for (int j = 0; j < CN; j += 4)
            *reinterpret_cast<float4*>(gOut + (size_t)(ty * 8 + i) * ldG + n0 + tx * CN + j) =
                make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
        }
    }
    if (gMax != nullptr) {
      // column max over the tile's valid rows; activations are >= 0 so the int ordering of the bits is the float ordering
#pragma unroll
      for (int j = 0; j < CN; ++j) {
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (ty * 8 + i < rows) m = fmaxf(m, acc[i][j]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        if ((ty & 1) == 0) atomicMax(&sMax[tx * CN + j], __float_as_int(m));
      }
      __syncthreads();
      if (tid < NC) atomicMax(&gMax[n0 + tid], sMax[tid]);
    }
  }
}

__device__ __forceinline__ void run_layer(const float* sIn, int ldIn, const LayerDesc& L, const float* bias0,
                                          float* sOut, int ldOut, float* sW, int tid, int rows, float* gOut,
                                          int ldG, int* sMax, int* gMax) {
  if ((L.N & 127) == 0)
    dense_layer<8>(sIn, ldIn, L, bias0, sOut, ldOut, sW, tid, rows, gOut, ldG, sMax, gMax);
  else
    dense_layer<4>(sIn, ldIn, L, bias0, sOut, ldOut, sW, tid, rows, gOut, ldG, sMax, gMax);
}

struct SmemPlan { int buf0_floats, buf1_floats; };

__host__ __device__ inline SmemPlan plan_branch(const NetDesc& net) {
  int m0 = 16, m1 = 0;   // buf0: input tile + odd layers' outputs; buf1: even layers' outputs
  for (int l = 0; l + 1 < net.n_conv; ++l) {
    int n = net.conv[0][l].N;
    if (l & 1) m0 = n > m0 ? n : m0; else m1 = n > m1 ? n : m1;
  }
  return SmemPlan{kTileRows * (m0 + 4), kTileRows * (m1 + 4)};
}

__host__ __device__ inline SmemPlan plan_head(const NetDesc& net) {
  int m0 = net.C1, m1 = net.H0;   // buf0: h1 tile + odd layers' outputs; buf1: layer 0 and even layers' outputs
  for (int i = 0; i + 1 < net.n_hidden; ++i) {
    int n = net.hidden[0][i].N;
    if (i & 1) m1 = n > m1 ? n : m1; else m0 = n > m0 ? n : m0;
  }
  return SmemPlan{kTileRows * (m0 + 4), kTileRows * (m1 + 4)};
}

__global__ void __launch_bounds__(kThreads, 1)
lrg_branch_kernel(const __grid_constant__ NetDesc net, const __grid_constant__ ForwardArgs fa) {
  const int b = blockIdx.z, br = blockIdx.y, tile = blockIdx.x;
  if (fa.active != nullptr && fa.active[(size_t)b * fa.active_stride] == 0) return;
  const int n = fa.n_pts[br];
  const int nv = forward_valid_rows(fa, b, br);
  const int row0 = tile * kTileRows;
  if (row0 >= nv) return;
  const int rows = min(kTileRows, nv - row0);
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) float smem[];
  const SmemPlan plan = plan_branch(net);
  float* buf0 = smem;
  float* buf1 = buf0 + plan.buf0_floats;
  float* sW = buf1 + plan.buf1_floats;
  int* sMax = reinterpret_cast<int*>(sW + kStages * kStageFloats);

  // input tile, zero padded to 16 features and 128 rows (row-major, ld 20)
  const float* x = fa.x[br] + ((size_t)b * n + row0) * fa.x_stride;
  for (int idx = tid; idx < kTileRows * 16; idx += kThreads) {
    int r = idx >> 4, c = idx & 15;
    buf0[r * 20 + c] = (r < rows && c < net.F) ? x[(size_t)r * fa.x_stride + c] : 0.f;
  }
  const float* cur = buf0;
  int ld = 20;
  for (int l = 0; l < net.n_conv; ++l) {
    const LayerDesc& L = net.conv[br][l];
    const bool last = (l == net.n_conv - 1);
    float* out = last ? nullptr : ((l & 1) ? buf0 : buf1);
    float* gOut = (l == 1) ? fa.h1[br] + ((size_t)b * n + row0) * net.C1 : nullptr;
    int* gMax = last ? reinterpret_cast<int*>(fa.pooled) + (size_t)b * 2 * net.Clast + br * net.Clast : nullptr;
    run_layer(cur, ld, L, L.bias, out, L.N + 4, sW, tid, rows, gOut, net.C1, sMax, gMax);
    cur = out;
    ld = L.N + 4;
  }
}

// gproj[b][h][c] = bias0_h[c] + sum_k pooled[b][k] * W0g_h[k][c]
__global__ void __launch_bounds__(256)
lrg_gproj_kernel(const __grid_constant__ NetDesc net, const __grid_constant__ ForwardArgs fa) {
  constexpr int BB = 8;
  const int h = blockIdx.y, c0 = blockIdx.x * 64, b0 = blockIdx.z * BB;
  const int tid = threadIdx.x, col = tid & 63, kq = tid >> 6;
  const int Kg = 2 * net.Clast;
  extern __shared__ __align__(16) float smem[];
  float* sg = smem;                 // [BB][Kg]
  float* sred = sg + BB * Kg;       // [4][BB][64]
  bool any = false;
  for (int bb = 0; bb < BB; ++bb) {
    int b = b0 + bb;
    bool act = b < fa.B && (fa.active == nullptr || fa.active[(size_t)b * fa.active_stride] != 0);
    any |= act;
    for (int k = tid; k < Kg; k += 256) sg[bb * Kg + k] = act ? fa.pooled[(size_t)b * Kg + k] : 0.f;
  }
  if (!any) return;
  __syncthreads();
  float acc[BB];
#pragma unroll
  for (int bb = 0; bb < BB; ++bb) acc[bb] = 0.f;
  const float* W = net.W0g[h] + c0 + col;
  const int kper = Kg / 4;
  for (int k = kq * kper; k < (kq + 1) * kper; ++k) {
    float w = W[(size_t)k * net.H0];
#pragma unroll
    for (int bb = 0; bb < BB; ++bb) acc[bb] = fmaf(sg[bb * Kg + k], w, acc[bb]);
  }
#pragma unroll
  for (int bb = 0; bb < BB; ++bb) sred[(kq * BB + bb) * 64 + col] = acc[bb];
  __syncthreads();
  for (int idx = tid; idx < BB * 64; idx += 256) {
    int bb = idx >> 6, c = idx & 63, b = b0 + bb;
    if (b < fa.B) {
      float v = net.head0_local[h].bias[c0 + c];
      v += sred[(0 * BB + bb) * 64 + c];
      v += sred[(1 * BB + bb) * 64 + c];
      v += sred[(2 * BB + bb) * 64 + c];
      v += sred[(3 * BB + bb) * 64 + c];
      fa.gproj[((size_t)b * 2 + h) * net.H0 + c0 + c] = v;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
lrg_head_kernel(const __grid_constant__ NetDesc net, const __grid_constant__ ForwardArgs fa) {
  const int b = blockIdx.z, h = blockIdx.y, tile = blockIdx.x;   // h: 0 = remove head on inlier rows, 1 = add head on neighbor rows
  if (fa.active != nullptr && fa.active[(size_t)b * fa.active_stride] == 0) return;
  const int n = fa.n_pts[h];
  const int nv = forward_valid_rows(fa, b, h);
  const int row0 = tile * kTileRows;
  if (row0 >= nv) return;
  const int rows = min(kTileRows, nv - row0);
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) float smem[];
  const SmemPlan plan = plan_head(net);
  float* buf0 = smem;
  float* buf1 = buf0 + plan.buf0_floats;
  float* sW = buf1 + plan.buf1_floats;

  const int C1 = net.C1;
  const float* h1 = fa.h1[h] + ((size_t)b * n + row0) * C1;
  for (int idx = tid; idx < kTileRows * (C1 / 4); idx += kThreads) {
    int r = idx / (C1 / 4), c4 = idx % (C1 / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) v = *reinterpret_cast<const float4*>(h1 + (size_t)r * C1 + c4 * 4);
    *reinterpret_cast<float4*>(buf0 + r * (C1 + 4) + c4 * 4) = v;
  }
  // layer 0: per-point part, the pooled part + bias arrive through gproj
  run_layer(buf0, C1 + 4, net.head0_local[h], fa.gproj + ((size_t)b * 2 + h) * net.H0, buf1, net.H0 + 4, sW, tid,
            rows, nullptr, 0, nullptr, nullptr);
  const float* cur = buf1;
  int ld = net.H0 + 4;
  for (int i = 0; i + 1 < net.n_hidden; ++i) {
    const LayerDesc& L = net.hidden[h][i];
    float* out = (i & 1) ? buf1 : buf0;
    run_layer(cur, ld, L, L.bias, out, L.N + 4, sW, tid, rows, nullptr, 0, nullptr, nullptr);
    cur = out;
    ld = L.N + 4;
  }
  __syncthreads();
  // final layer Hlast -> 2 logits, no activation (util.py:145-149 / :158-162)
  const LayerDesc& Lo = net.out[h];
  const int r = tid >> 1, o = tid & 1;
  float acc = Lo.bias[o];
  const float* in = cur + r * ld;
  for (int k = 0; k < Lo.K; ++k) acc = fmaf(in[k], Lo.W[k * 2 + o], acc);
  if (r < rows) fa.logits[h][((size_t)b * n + row0 + r) * 2 + o] = acc;
}

size_t forward_smem_branch(const NetDesc& net) {
  SmemPlan p = plan_branch(net);
  return sizeof(float) * (size_t)(p.buf0_floats + p.buf1_floats + kStages * kStageFloats) + 128 * sizeof(int);
}

size_t forward_smem_head(const NetDesc& net) {
  SmemPlan p = plan_head(net);
  return sizeof(float) * (size_t)(p.buf0_floats + p.buf1_floats + kStages * kStageFloats);
}

static size_t gproj_smem(const NetDesc& net) { return sizeof(float) * (size_t)(8 * 2 * net.Clast + 4 * 8 * 64); }

int forward_configure(const NetDesc& net) {
  LRG_REQUIRE(net.F <= 16, "feature_size %d > 16 is not supported", net.F);
  size_t sb = forward_smem_branch(net), sh = forward_smem_head(net), sg = gproj_smem(net);
  LRG_REQUIRE(sb <= 232448 && sh <= 232448 && sg <= 232448, "network does not fit the 227 KB shared memory tile plan");
  LRG_CUDA(cudaFuncSetAttribute(lrg_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
  LRG_CUDA(cudaFuncSetAttribute(lrg_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  LRG_CUDA(cudaFuncSetAttribute(lrg_gproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg));
  return LRG_OK;
}

// pooled must be zero for the active tile pairs on entry.
int launch_forward_timed(const NetDesc& net, const ForwardArgs& fa, cudaStream_t stream, cudaEvent_t* ev) {
  if (fa.B <= 0) return LRG_OK;
  const int nmax = fa.n_pts[0] > fa.n_pts[1] ? fa.n_pts[0] : fa.n_pts[1];
  const int tiles = (nmax + kTileRows - 1) / kTileRows;
  dim3 grid(tiles, 2, fa.B);
  if (ev) cudaEventRecord(ev[0], stream);
  lrg_branch_kernel<<<grid, kThreads, forward_smem_branch(net), stream>>>(net, fa);
  if (ev) cudaEventRecord(ev[1], stream);
  dim3 ggrid(net.H0 / 64, 2, (fa.B + 7) / 8);
  lrg_gproj_kernel<<<ggrid, 256, gproj_smem(net), stream>>>(net, fa);
  if (ev) cudaEventRecord(ev[2], stream);
  lrg_head_kernel<<<grid, kThreads, forward_smem_head(net), stream>>>(net, fa);
  if (ev) cudaEventRecord(ev[3], stream);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int launch_forward(const NetDesc& net, const ForwardArgs& fa, cudaStream_t stream) {
  return launch_forward_timed(net, fa, stream, nullptr);
}

}  // namespace lrg
