// LrgNet forward on the 5th-generation tensor cores (tcgen05 / TMEM) for the full model
// (/root/reference/learn_region_grow_util.py:77-79: conv 64,64,64,128,512; heads 256,128,2): the stand-alone kernels of
// the lock-step loop and of lrg_forward_device.  The tile bodies live in lrg_tc_tiles.cuh (shared with the persistent
// grow kernel).
//
// Precision: the parity bar is the fp32 TF graph, so every contraction runs as a three-term split product -- each fp32
// operand is split into hi + lo and D += hi.hi + lo.hi + hi.lo with fp32 accumulation in TMEM, which carries ~21-22
// mantissa bits per product.  Two kinds: 3xFP16 (kind::f16, hi / lo are fp16, weights pre-scaled by a power of two per
// layer; half the MMAs and half the weight bytes; tools/umma_f16_probe.cu: 2e-7..9e-7 relative per GEMM) is the default;
// 3xTF32 (kind::tf32; tools/umma_probe.cu: 5e-6..4e-5 absolute per GEMM where an fp32 FMA chain has 2e-6..1e-5) is what a
// call falls back to when an activation leaves the fp16 range (TcNet::range_flag).
//
//   lrg_tc_branch_kernel  one CTA per (128-point tile, branch, tile pair): x -> 64 -> 64 -> 64 -> 128 -> 512 -> column max.
//                         Activations never leave the SM: the epilogue warps read the accumulator from TMEM, add bias,
//                         ReLU, split hi/lo and write the next layer's A operand straight into shared memory in the
//                         UMMA canonical K-major layout.  Weights stream from L2 as pre-packed 32 KB operand images
//                         through a 3-slot ring fed by 1-D bulk async copies (TMA without a tensor map).  The 128->512
//                         layer is never materialised: its accumulator (2 x 128 TMEM columns, double buffered) is
//                         reduced to the column max by a warp butterfly and merged with atomicMax.
//   lrg_tc_gproj_kernel   pooled(1024) . W0[:1024] + bias0 for both heads (the reference multiplies the tiled pooled
//                         row 512 times, util.py:128-135; algebraically Z.K0 = g.K0[:1024] + h1.K0[1024:]).
//   lrg_tc_head_kernel    one CTA per (128-point tile, head, tile pair): gproj row as bias + h1.W0[1024:] -> 256 -> 128
//                         -> 2, the 256-wide hidden layer produced 64 channels at a time and consumed as K-chunks of the
//                         next layer so that it fits beside the operand ring.
//
// Warp roles (192 threads): warps 0-3 = epilogue (warp w owns TMEM lanes 32w..32w+31 = tile rows), warp 4 = MMA issuer
// (one elected thread), warp 5 = weight loader (one elected thread).  All hand-offs are mbarriers.
#include "lrg_tc_tiles.cuh"

namespace lrg {

constexpr int kTcThreads = 192;

template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1) lrg_tc_branch_kernel(const __grid_constant__ TcNet net, const __grid_constant__ ForwardArgs fa) {
  const int b = blockIdx.z, br = blockIdx.y, tile = blockIdx.x;
  if (fa.active != nullptr && fa.active[(size_t)b * fa.active_stride] == 0) return;
  const int nvalid = forward_valid_rows(fa, b, br);
  if (tile * 128 >= nvalid) return;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(16) TcStatic st;
  __shared__ uint32_t tmem_base;
  if ((threadIdx.x >> 5) == 4) tmem_alloc(smem_u32(&tmem_base), kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base;
  tc_branch_tile<F16>(net, fa, b, br, tile, nvalid, 0, 4, smem, st, tmem);
  if ((threadIdx.x >> 5) == 4) tmem_dealloc(tmem, kTmemCols);
}

__global__ void __launch_bounds__(512) lrg_tc_gproj_kernel(const __grid_constant__ TcNet net, const __grid_constant__ ForwardArgs fa) {
  __shared__ float sP[1024];
  __shared__ float sR[32 * 64];
  const int cb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (fa.active != nullptr && fa.active[(size_t)b * fa.active_stride] == 0) return;
  tc_gproj_block(net, fa, b, h, cb, sP, sR);
}

template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1) lrg_tc_head_kernel(const __grid_constant__ TcNet net, const __grid_constant__ ForwardArgs fa) {
  const int b = blockIdx.z, h = blockIdx.y, tile = blockIdx.x;   // h: 0 = remove head on inlier rows, 1 = add head on neighbor rows
  if (fa.active != nullptr && fa.active[(size_t)b * fa.active_stride] == 0) return;
  const int nvalid = forward_valid_rows(fa, b, h);
  if (tile * 128 >= nvalid) return;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(16) TcStatic st;
  __shared__ uint32_t tmem_base;
  if ((threadIdx.x >> 5) == 4) tmem_alloc(smem_u32(&tmem_base), kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base;
  tc_head_tile<F16>(net, fa, b, h, tile, nvalid, nullptr, nullptr, 0u, smem, st, tmem);
  if ((threadIdx.x >> 5) == 4) tmem_dealloc(tmem, kTmemCols);
}

int tc_forward_configure() {
  LRG_CUDA(cudaFuncSetAttribute(lrg_tc_branch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LRG_CUDA(cudaFuncSetAttribute(lrg_tc_head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LRG_CUDA(cudaFuncSetAttribute(lrg_tc_branch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LRG_CUDA(cudaFuncSetAttribute(lrg_tc_head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  return LRG_OK;
}

// pooled must be zero for the active tile pairs on entry.  f16: 3xFP16 tiles (net.range_flag reports a range overflow).
int launch_forward_tc(const TcNet& net, const ForwardArgs& fa, bool f16, cudaStream_t stream, cudaEvent_t* ev) {
  if (fa.B <= 0) return LRG_OK;
  const int nmax = fa.n_pts[0] > fa.n_pts[1] ? fa.n_pts[0] : fa.n_pts[1];
  dim3 grid((nmax + 127) / 128, 2, fa.B);
  if (ev) cudaEventRecord(ev[0], stream);
  if (f16) lrg_tc_branch_kernel<true><<<grid, kTcThreads, kTcSmem, stream>>>(net, fa);
  else lrg_tc_branch_kernel<false><<<grid, kTcThreads, kTcSmem, stream>>>(net, fa);
  if (ev) cudaEventRecord(ev[1], stream);
  lrg_tc_gproj_kernel<<<dim3(4, 2, fa.B), 512, 0, stream>>>(net, fa);
  if (ev) cudaEventRecord(ev[2], stream);
  if (f16) lrg_tc_head_kernel<true><<<grid, kTcThreads, kTcSmem, stream>>>(net, fa);
  else lrg_tc_head_kernel<false><<<grid, kTcThreads, kTcSmem, stream>>>(net, fa);
  if (ev) cudaEventRecord(ev[3], stream);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg
