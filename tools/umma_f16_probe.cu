// Stand-alone probe (sm_100a) for the 3xFP16 variant of the LrgNet tensor tiles: tcgen05.mma kind::f16 with the A operand in
// TENSOR MEMORY (written by tcgen05.st as packed pairs of fp16) and B in shared memory (no-swizzle K-major canonical layout,
// core matrix = 8 rows x 16 bytes = 8 fp16 along K).  It settles on the GPU what cannot be settled by reading:
//   * how a 16-bit A operand is laid out in TMEM (variant 0: column c of lane m holds A[m][2c] in its low half and A[m][2c+1]
//     in its high half, one MMA of K = 16 consuming 8 columns; variant 1: halves swapped),
//   * the accuracy of x = hi + lo with both parts fp16 (weights pre-scaled by a power of two so that their lo parts stay
//     normal), D = hi.hi + lo.hi + hi.lo with fp32 accumulation, against a float64 host evaluation,
//   * the issue rate of back-to-back MMAs (M = 128, K = 16) for N = 64 / 128 / 256.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_f16_probe tools/umma_f16_probe.cu && tools/umma_f16_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../learn_region_grow_b200/csrc/lrg_umma.cuh"

using namespace lrg::umma;

struct ProbeArgs {
  const float* A;        // [128][K] row-major fp32
  const uint16_t* Bimg;  // hi image then lo image (fp16 bits), canonical layout, N*K halves each
  float* D;              // [128][N]
  int K, N, variant, terms;
  float inv_scale;       // 1 / (power of two the weights were scaled by)
};

__global__ void __launch_bounds__(192, 1) probe_kernel(ProbeArgs pa) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int K = pa.K, N = pa.N;
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
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t col_hi = 256, col_lo = 256 + 64;      // K <= 128: K/2 <= 64 columns each

  if (warp < 4) {
    const int r = tid;
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t vh[32], vl[32];
      for (int j = 0; j < 32; ++j) {
        uint32_t ph = 0, pl = 0;
        if (c0 + j < K / 2) {
          const float x0 = pa.A[(size_t)r * K + 2 * (c0 + j)], x1 = pa.A[(size_t)r * K + 2 * (c0 + j) + 1];
          split_f16x2(x0, x1, ph, pl);
          if (pa.variant == 1) { ph = (ph >> 16) | (ph << 16); pl = (pl >> 16) | (pl << 16); }
        }
        vh[j] = ph; vl[j] = pl;
      }
      tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + col_hi + c0, vh);
      tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + col_lo + c0, vl);
    }
    tmem_st_wait();
    tcgen05_fence_before();
    mbar_arrive(bar_a);
    mbar_wait(bar_d, 0);
    tcgen05_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) pa.D[(size_t)r * N + c0 + j] = __uint_as_float(v[j]) * pa.inv_scale;
    }
    tcgen05_fence_before();
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(2 * N * K * 2);
      mbar_expect_tx(bar_b, bytes);
      bulk_g2s(smem_u32(smem), pa.Bimg, bytes, bar_b);
    }
  } else {
    if (lane == 0) {
      mbar_wait(bar_b, 0);
      mbar_wait(bar_a, 0);
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_f16(128, N);
      const uint32_t kdirB = (uint32_t)N * 16, mndir = 128;
      const uint32_t b_hi = smem_u32(smem), b_lo = b_hi + (uint32_t)(N * K * 2);
      uint32_t acc = 0;
      for (int t = 0; t < pa.terms; ++t) {
        const uint32_t a0 = tmem + ((t == 1) ? col_lo : col_hi), b0 = (t == 2) ? b_lo : b_hi;
        for (int ks = 0; ks < K / 16; ++ks) {
          umma_f16_ts(tmem, a0 + ks * 8, make_desc(b0 + ks * 2 * kdirB, kdirB, mndir), idesc, acc);
          acc = 1;
        }
      }
      umma_commit(bar_d);
    }
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int alternate, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t kdirB = (uint32_t)N * 16;
    const uint64_t bd = make_desc(smem_u32(smem), kdirB, 128);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t d = tmem + ((alternate && (ks & 1)) ? 256u : 0u);
        umma_f16_ts(d, tmem + 448 + ks * 8, bd + (uint64_t)(((ks & 3) * 2 * kdirB) >> 4), idesc, 1u);
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

static uint16_t h_bits(float x) { __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h_val(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

static void pack_image(const std::vector<float>& W /* [N][K] */, int N, int K, float scale, uint16_t* hi, uint16_t* lo) {
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float w = W[(size_t)n * K + k] * scale;
      const uint16_t h = h_bits(w);
      const uint16_t l = h_bits(w - h_val(h));
      const size_t off = (size_t)(k / 8) * (N * 8) + (size_t)n * 8 + (k % 8);
      hi[off] = h;
      lo[off] = l;
    }
}

int main() {
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { printf("no CUDA device\n"); return 2; }
  int fails = 0, layout = -1;
  const int cfgs[][2] = {{16, 64}, {64, 64}, {64, 128}, {128, 128}};
  for (auto& cfg : cfgs) {
    const int K = cfg[0], N = cfg[1];
    std::vector<float> A(128 * K), W((size_t)N * K), D(128 * N);
    std::vector<uint16_t> img(2 * (size_t)N * K);
    srand(1234 + K + N);
    for (auto& x : A) x = ((float)rand() / RAND_MAX * 4.f - 1.f) * ((rand() & 7) == 0 ? 100.f : 1.f);   // activations up to a few hundred
    for (auto& x : W) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * ((rand() & 3) == 0 ? 0.01f : 1.f) * 5.f;
    float wmax = 0.f;
    for (auto x : W) wmax = fmaxf(wmax, fabsf(x));
    const float scale = exp2f(floorf(log2f(32768.f / wmax)));
    pack_image(W, N, K, scale, img.data(), img.data() + (size_t)N * K);
    std::vector<double> ref(128 * N);
    double refmax = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * (double)W[(size_t)n * K + k];
        ref[r * N + n] = s;
        refmax = fmax(refmax, fabs(s));
      }
    float *dA, *dD;
    uint16_t* dB;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, img.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * N * K) * 2;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int variant = 0; variant < 2; ++variant)
      for (int terms = 1; terms <= 3; terms += 2) {
        cudaMemset(dD, 0xff, D.size() * 4);
        ProbeArgs pa{dA, dB, dD, K, N, variant, terms, 1.f / scale};
        probe_kernel<<<1, 192, smem>>>(pa);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("K=%d N=%d variant=%d terms=%d: CUDA error %s\n", K, N, variant, terms, cudaGetErrorString(err)); return 3; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double e = 0;
        for (size_t i = 0; i < D.size(); ++i) {
          double d = fabs((double)D[i] - ref[i]);
          if (!(d <= e)) e = d;
        }
        const double rel = e / refmax;
        const bool ok = terms == 3 ? rel < 2e-6 : rel < 3e-3;
        printf("K=%3d N=%3d variant=%d terms=%d  max|D-ref64| = %.3e  (max|ref| %.1f, relative %.2e, weight scale 2^%d)  %s\n", K, N, variant,
               terms, e, refmax, rel, (int)log2f(scale), ok ? "OK" : "mismatch");
        if (ok && terms == 3) { if (layout < 0) layout = variant; else if (layout != variant) layout = 99; }
        if (!ok && variant == (layout < 0 ? 0 : layout) && layout != 99) ++fails;
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
      printf("rate: M=128 N=%3d K=16 f16, A in TMEM, %s accumulator(s): %.1f cycles per MMA (nominal %d)\n", c[0], c[1] ? "two alternating" : "one",
             cyc / 512.0, c[0] / 2);
    }
    cudaFree(d_out);
  }
  printf("probe: TMEM layout of a 16-bit A operand: %s\n", layout == 0 ? "variant 0 (low half = even k)" : layout == 1 ? "variant 1 (low half = odd k)" : "UNDETERMINED");
  return (fails || layout < 0 || layout == 99) ? 1 : 0;
}
