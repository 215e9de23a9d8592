// Stand-alone probe of the tcgen05 building blocks the LrgNet tensor kernels rely on (sm_100a):
//   * TMEM alloc / dealloc, tcgen05.mma kind::tf32 with both operands in shared memory (no-swizzle K-major canonical
//     layout), tcgen05.commit -> mbarrier, tcgen05.ld 32x32b, cp.async.bulk -> mbarrier, fence.proxy.async;
//   * the 3xTF32 split (hi.hi + lo.hi + hi.lo) against a float64 host evaluation.
// It answers, on the GPU, the questions that cannot be settled by reading: which descriptor field is the K-direction
// stride (variant 0: LBO = K direction, SBO = M/N direction; variant 1: swapped) and how accurate the split is.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../learn_region_grow_b200/csrc/lrg_umma.cuh"

using namespace lrg::umma;

struct ProbeArgs {
  const float* A;      // [128][K] row-major fp32
  const float* Bimg;   // hi image then lo image, canonical layout, N*K floats each
  float* D;            // [128][N]
  int K, N, variant, terms;   // variant 0: A from shared memory; variant 2: A from tensor memory (written with tcgen05.st)
};

__global__ void __launch_bounds__(192, 1) probe_kernel(ProbeArgs pa) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int K = pa.K, N = pa.N;
  float* sA_hi = reinterpret_cast<float*>(smem);
  float* sA_lo = sA_hi + 128 * K;
  float* sB = sA_lo + 128 * K;                       // hi image, lo image
  __shared__ __align__(8) uint64_t bars[3];          // 0: B landed, 1: A written, 2: accumulator complete
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_b = smem_u32(&bars[0]), bar_a = smem_u32(&bars[1]), bar_d = smem_u32(&bars[2]);
  if (tid == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_a, 128);
    mbar_init(bar_d, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_smem), 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;

  if (warp < 4) {
    // A tile: thread = row, 16-byte chunks of 4 K-elements, canonical offset (k/4)*2048 + row*16
    const int r = tid;
    for (int k4 = 0; k4 < K / 4; ++k4) {
      float4 v = *reinterpret_cast<const float4*>(pa.A + (size_t)r * K + k4 * 4);
      float4 hi, lo;
      split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
      *reinterpret_cast<float4*>(sA_hi + k4 * 512 + r * 4) = hi;
      *reinterpret_cast<float4*>(sA_lo + k4 * 512 + r * 4) = lo;
    }
    if (pa.variant == 2) {
      // A operand in TMEM: hi at columns 128.., lo at columns 192.. (K <= 64), thread = row = lane
      for (int k0 = 0; k0 < K; k0 += 32) {
        uint32_t vh[32], vl[32];
        for (int j = 0; j < 32; ++j) {
          float hi = 0.f, lo = 0.f;
          if (k0 + j < K) split_tf32(pa.A[(size_t)r * K + k0 + j], hi, lo);
          vh[j] = __float_as_uint(hi); vl[j] = __float_as_uint(lo);
        }
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 128 + k0, vh);
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 192 + k0, vl);
      }
      tmem_st_wait();
      tcgen05_fence_before();
    }
    fence_proxy_async();
    mbar_arrive(bar_a);
    // epilogue
    mbar_wait(bar_d, 0);
    tcgen05_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) pa.D[(size_t)r * N + c0 + j] = __uint_as_float(v[j]);
    }
    tcgen05_fence_before();
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(2 * N * K * sizeof(float));
      mbar_expect_tx(bar_b, bytes);
      bulk_g2s(smem_u32(sB), pa.Bimg, bytes, bar_b);
    }
  } else {   // warp 4: MMA issuer
    if (lane == 0) {
      mbar_wait(bar_b, 0);
      mbar_wait(bar_a, 0);
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_tf32(128, N);
      const uint32_t kdirA = 128 * 16, kdirB = (uint32_t)N * 16, mndir = 128;   // byte strides between core matrices
      const uint32_t a_hi = smem_u32(sA_hi), a_lo = smem_u32(sA_lo), b_hi = smem_u32(sB), b_lo = smem_u32(sB + N * K);
      uint32_t acc = 0;
      for (int t = 0; t < pa.terms; ++t) {
        const uint32_t a0 = (t == 1) ? a_lo : a_hi, b0 = (t == 2) ? b_lo : b_hi;
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint32_t aaddr = a0 + ks * 2 * kdirA, baddr = b0 + ks * 2 * kdirB;
          const uint64_t da = make_desc(aaddr, kdirA, mndir), db = make_desc(baddr, kdirB, mndir);
          if (pa.variant == 2) umma_tf32_ts(tmem, tmem + ((t == 1) ? 192 : 128) + ks * 8, db, idesc, acc);
          else umma_tf32(tmem, da, db, idesc, acc);
          acc = 1;
        }
      }
      umma_commit(bar_d);
    }
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------- MMA rate probe
// One thread issues `reps` x 8 tf32 MMAs (M=128, N, K=8) with A in TMEM and B in shared memory, either all into one
// accumulator or alternating between two, then commits and waits; cycles per MMA tell whether back-to-back MMAs into
// the same accumulator run at the nominal N/2 cycles.
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int alternate, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;    // 64 KB of zeros as the B operand
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint32_t kdirB = (uint32_t)N * 16;
    const uint64_t bd = make_desc(smem_u32(smem), kdirB, 128);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t d = tmem + ((alternate && (ks & 1)) ? 256u : 0u);
        umma_tf32_ts(d, tmem + 448 + ks * 8, bd + (uint64_t)(((ks & 3) * 2 * kdirB) >> 4), idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[0] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void pack_image(const std::vector<float>& W /* [N][K] */, int N, int K, float* hi, float* lo) {
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float h, l;
      split_tf32(W[(size_t)n * K + k], h, l);
      size_t off = (size_t)(k / 4) * (N * 4) + (size_t)n * 4 + (k % 4);
      hi[off] = h;
      lo[off] = l;
    }
}

int main() {
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { printf("no CUDA device\n"); return 2; }
  int fails = 0;
  const int cfgs[][2] = {{16, 64}, {64, 64}, {64, 128}, {32, 128}};
  for (auto& cfg : cfgs) {
    const int K = cfg[0], N = cfg[1];
    std::vector<float> A(128 * K), W((size_t)N * K), img(2 * (size_t)N * K), D(128 * N);
    srand(1234 + K + N);
    for (auto& x : A) x = (float)rand() / RAND_MAX * 4.f - 1.f;
    for (auto& x : W) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    pack_image(W, N, K, img.data(), img.data() + (size_t)N * K);
    std::vector<double> ref(128 * N);
    std::vector<float> ref32(128 * N);
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        double s = 0; float s32 = 0.f;
        for (int k = 0; k < K; ++k) { s += (double)A[r * K + k] * (double)W[(size_t)n * K + k]; s32 = fmaf(A[r * K + k], W[(size_t)n * K + k], s32); }
        ref[r * N + n] = s; ref32[r * N + n] = s32;
      }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, img.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * 4;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int variant = 0; variant <= 2; variant += 2)   // (variant 1, descriptor fields swapped, faults with an illegal address: settled)
      for (int terms = 1; terms <= 3; terms += 2) {
        cudaMemset(dD, 0xff, D.size() * 4);
        ProbeArgs pa{dA, dB, dD, K, N, variant, terms};
        probe_kernel<<<1, 192, smem>>>(pa);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("K=%d N=%d variant=%d terms=%d: CUDA error %s\n", K, N, variant, terms, cudaGetErrorString(err)); return 3; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double e = 0, e32 = 0;
        for (size_t i = 0; i < D.size(); ++i) {
          double d = fabs((double)D[i] - ref[i]);
          if (!(d <= e)) e = d;                       // NaN-propagating max
          e32 = fmax(e32, fabs((double)ref32[i] - ref[i]));
        }
        const bool ok = terms == 3 ? e < 1e-4 : e < 5e-2;
        printf("K=%3d N=%3d variant=%d terms=%d  max|D-ref64| = %.3e   (fp32 fmaf chain: %.3e)  %s\n", K, N, variant, terms, e, e32, ok ? "OK" : "MISMATCH");
        if (!ok) ++fails;
      }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int cases[][2] = {{64, 0}, {64, 1}, {128, 0}, {128, 1}, {256, 0}, {256, 1}};
    for (auto& c : cases) {
      long long cyc = 0;
      for (int it = 0; it < 2; ++it) {
        rate_kernel<<<1, 128, 65536>>>(c[0], c[1], 64, d_out);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("rate probe N=%d: CUDA error %s\n", c[0], cudaGetErrorString(err)); return 3; }
        cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      }
      printf("rate: M=128 N=%3d K=8 tf32, %s accumulator(s): %.1f cycles per MMA (nominal %d)\n", c[0], c[1] ? "two alternating" : "one", cyc / 512.0, c[0] / 2);
    }
    cudaFree(d_out);
  }
  printf(fails ? "probe: FAILED in %d case(s)\n" : "probe: all cases pass (LBO = K-direction stride, SBO = M/N-direction stride; A from TMEM: lane = row, column = k)\n", fails);
  return fails ? 1 : 0;
}
