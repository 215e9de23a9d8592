// Stand-alone kernels of the region-grow driver (/root/reference/test_region_grow.py:175-316): feature-row packing,
// the lock-step step kernel (one CTA per room in flight advances its room by one grow step per launch; the body is
// lrg_step_body.cuh, shared with the persistent grow kernel) and the final nearest-neighbour fill.
#include "lrg_step_body.cuh"

namespace lrg {

// ----------------------------------------------------------------------------------------------------- pack
__global__ void lrg_pack_kernel(const float* __restrict__ points, int F, long long total, float res,
                                float* __restrict__ pts16, int4* __restrict__ vox) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float row[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) row[c] = c < F ? points[i * F + c] : 0.f;
#pragma unroll
  for (int c = 0; c < 16; c += 4)
    *reinterpret_cast<float4*>(pts16 + i * 16 + c) = make_float4(row[c], row[c + 1], row[c + 2], row[c + 3]);
  vox[i] = make_int4(voxel_of(row[0], res), voxel_of(row[1], res), voxel_of(row[2], res), 0);
}

int launch_pack(const float* d_points, int F, long long total, float resolution, float* d_pts16, int4* d_vox,
                cudaStream_t stream) {
  if (total <= 0) return LRG_OK;
  lrg_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_points, F, total, resolution, d_pts16, d_vox);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

// ----------------------------------------------------------------------------------------------------- step
__global__ void __launch_bounds__(kStepThreads, 1) lrg_step_kernel(const __grid_constant__ DriverArgs da) {
  __shared__ StepShared sh;
  step_body<kStepThreads>(da, blockIdx.x, sh);
}

int launch_step(const DriverArgs& da, cudaStream_t stream) {
  lrg_step_kernel<<<da.n_slots, kStepThreads, 0, stream>>>(da);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

// ----------------------------------------------------------------------------------------------------- fill
// test_region_grow.py:308-316: every unlabeled point takes the label of the nearest labelled point of its room,
// squared L2 over all F feature columns in float32, first minimum wins.
__global__ void __launch_bounds__(kStepThreads) lrg_fill_lists_kernel(const __grid_constant__ FillArgs fa) {
  __shared__ int scan[33];
  const int room = blockIdx.x;
  const long long base = fa.room_off[room];
  const int N = (int)(fa.room_off[room + 1] - base);
  const int* label = fa.label + base;
  for (int i = threadIdx.x; i < N; i += kStepThreads) fa.label_filled[base + i] = label[i];
  const int nl = block_compact<kStepThreads>(N, [&](int i) { return label[i] != 0; }, fa.lab_list + base, scan);
  const int nu = block_compact<kStepThreads>(N, [&](int i) { return label[i] == 0; }, fa.unl_list + base, scan);
  if (threadIdx.x == 0) { fa.n_lab[room] = nl; fa.n_unl[room] = nu; }
}

// numpy.sum((a - b)**2) over a contiguous float32 row: pairwise-sum order of numpy's add.reduce inner loop
// (8 accumulators for 8 <= n <= 128, sequential remainder; plain left-to-right below 8).
__device__ __forceinline__ float numpy_sqdist(const float* __restrict__ a, const float* __restrict__ b, int F) {
  float e[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float d = __fsub_rn(a[c], b[c]);
    e[c] = __fmul_rn(d, d);
  }
  if (F < 8) {
    float s = 0.f;
    for (int c = 0; c < F; ++c) s = __fadd_rn(s, e[c]);
    return s;
  }
  float s = __fadd_rn(__fadd_rn(__fadd_rn(e[0], e[1]), __fadd_rn(e[2], e[3])),
                      __fadd_rn(__fadd_rn(e[4], e[5]), __fadd_rn(e[6], e[7])));
  for (int c = 8; c < F; ++c) s = __fadd_rn(s, e[c]);
  return s;
}

__global__ void __launch_bounds__(256) lrg_fill_kernel(const __grid_constant__ FillArgs fa) {
  const int room = blockIdx.y;
  const long long base = fa.room_off[room];
  const int nl = fa.n_lab[room], nu = fa.n_unl[room];
  if (nl == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* pts = fa.pts + base * 16;
  const int* lab_list = fa.lab_list + base;
  for (int u = blockIdx.x * 8 + warp; u < nu; u += gridDim.x * 8) {
    const int i = fa.unl_list[base + u];
    float q[16];
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      float4 t = *reinterpret_cast<const float4*>(pts + (size_t)i * 16 + c);
      q[c] = t.x; q[c + 1] = t.y; q[c + 2] = t.z; q[c + 3] = t.w;
    }
    float best = INFINITY;
    int bestpos = INT_MAX;
    for (int pos = lane; pos < nl; pos += 32) {
      const float* row = pts + (size_t)lab_list[pos] * 16;
      float rv[16];
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        float4 t = *reinterpret_cast<const float4*>(row + c);
        rv[c] = t.x; rv[c + 1] = t.y; rv[c + 2] = t.z; rv[c + 3] = t.w;
      }
      const float d = numpy_sqdist(rv, q, fa.F);
      if (d < best) { best = d; bestpos = pos; }
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, dlt);
      const int op = __shfl_xor_sync(0xffffffffu, bestpos, dlt);
      if (ob < best || (ob == best && op < bestpos)) { best = ob; bestpos = op; }
    }
    if (lane == 0 && bestpos != INT_MAX) fa.label_filled[base + i] = fa.label[base + lab_list[bestpos]];
  }
}

int launch_fill(const FillArgs& fa, cudaStream_t stream) {
  if (fa.n_rooms <= 0) return LRG_OK;
  lrg_fill_lists_kernel<<<fa.n_rooms, kStepThreads, 0, stream>>>(fa);
  lrg_fill_kernel<<<dim3(64, fa.n_rooms), 256, 0, stream>>>(fa);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg
