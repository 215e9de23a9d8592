// Stand-alone kernels of the region-grow driver (/root/reference/test_region_grow.py:175-316): feature-row packing,
// the lock-step step kernel (one CTA per room in flight advances its room by one grow step per launch; the body is
// lrg_step_body.cuh, shared with the persistent grow kernel) and the final nearest-neighbour fill.
#include "lrg_step_body.cuh"
#include "lrg_sort.cuh"

namespace lrg {

// ----------------------------------------------------------------------------------------------------- pack
// One CTA per room: voxelise (numpy.round(xyz / resolution), :175), find the room's voxel origin, then write the padded
// 16-float feature rows and the packed state words (coordinates relative to the origin, flags clear, padding VISITED).
__global__ void __launch_bounds__(1024) lrg_pack_kernel(const float* __restrict__ points, int F, const long long* __restrict__ room_off,
                                                        const long long* __restrict__ pw_off, float res, float* __restrict__ pts16,
                                                        unsigned* __restrict__ pw, int4* __restrict__ room_vmin, int* err) {
  __shared__ int s_mn[3], s_mx[3];
  const int room = blockIdx.x, tid = threadIdx.x;
  const long long base = room_off[room];
  const int N = (int)(room_off[room + 1] - base);
  if (tid < 3) { s_mn[tid] = INT_MAX; s_mx[tid] = INT_MIN; }
  __syncthreads();
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = tid; i < N; i += 1024)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int v = voxel_of(points[(base + i) * F + a], res);
      mn[a] = min(mn[a], v); mx[a] = max(mx[a], v);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) { atomicMin(&s_mn[a], mn[a]); atomicMax(&s_mx[a], mx[a]); }
  __syncthreads();
  const int o0 = s_mn[0], o1 = s_mn[1], o2 = s_mn[2];
  if (tid == 0) {
    room_vmin[room] = make_int4(N ? o0 : 0, N ? o1 : 0, N ? o2 : 0, 0);
    if (N > 0 && (s_mx[0] - o0 > 1022 || s_mx[1] - o1 > 1022 || s_mx[2] - o2 > 1022)) atomicCAS(err, 0, room + 1);
  }
  unsigned* w = pw + pw_off[room];
  const int n4 = (N + 3) & ~3;
  for (int i = tid; i < n4; i += 1024) {
    if (i >= N) { w[i] = PW_VIS; continue; }
    float row[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) row[c] = c < F ? points[(base + i) * F + c] : 0.f;
#pragma unroll
    for (int c = 0; c < 16; c += 4)
      *reinterpret_cast<float4*>(pts16 + (base + i) * 16 + c) = make_float4(row[c], row[c + 1], row[c + 2], row[c + 3]);
    const unsigned x = (unsigned)(voxel_of(row[0], res) - o0) & 1023u, y = (unsigned)(voxel_of(row[1], res) - o1) & 1023u,
                   z = (unsigned)(voxel_of(row[2], res) - o2) & 1023u;
    w[i] = x | (y << 10) | (z << 20);
  }
}

__global__ void __launch_bounds__(1024) lrg_reset_words_kernel(const long long* __restrict__ room_off, const long long* __restrict__ pw_off,
                                                               unsigned* __restrict__ pw, int lanes, long long lane_stride) {
  const int room = blockIdx.x;
  const int N = (int)(room_off[room + 1] - room_off[room]);
  unsigned* w = pw + pw_off[room];
  const int n4 = (N + 3) & ~3;
  for (int i = threadIdx.x; i < n4; i += 1024) {
    const unsigned v = i < N ? (w[i] & PW_XYZ) : PW_VIS;
    for (int l = 0; l < lanes; ++l) w[(long long)l * lane_stride + i] = v;
  }
}

// Spatial index of a room (DriverArgs::sp_*): the neighbour shell of a grow step (test_region_grow.py:221-229) is a small box,
// but the reference's point order (first-seen voxels of the raw file, :125-136) is not spatial -- a whole-room scan reads all
// N state words per step.  One CTA per room sorts the points by the Morton code of their voxel (bitonic sort of
// code << 32 | index in global scratch), stores the permutation and the coordinates in that order, and the bounding box
// of every run of kSpBlock points; a step then reads the block boxes, and only the blocks that meet its shell.
__device__ __forceinline__ unsigned morton_spread10(unsigned v) {       // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void lrg_spatial_keys_kernel(const long long* __restrict__ room_off, const long long* __restrict__ pw_off, const unsigned* __restrict__ pw,
                                        const long long* __restrict__ key_off, unsigned long long* __restrict__ keys_all) {
  const int room = blockIdx.y;
  const int N = (int)(room_off[room + 1] - room_off[room]);
  if (N <= 0) return;
  const unsigned* w = pw + pw_off[room];
  unsigned long long* keys = keys_all + key_off[room];
  const int P = (int)(key_off[room + 1] - key_off[room]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < N) {
      const unsigned v = w[i];
      const unsigned code = morton_spread10(v & 1023u) | (morton_spread10((v >> 10) & 1023u) << 1) | (morton_spread10((v >> 20) & 1023u) << 2);
      k = ((unsigned long long)code << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
}

// behind the sort: permutation and coordinates in Morton order, one warp per block of kSpBlock points for its bounding box
__global__ void __launch_bounds__(256) lrg_spatial_blocks_kernel(const long long* __restrict__ room_off, const long long* __restrict__ pw_off,
                                                                const unsigned* __restrict__ pw, const long long* __restrict__ sp_off,
                                                                const long long* __restrict__ key_off, const unsigned long long* __restrict__ keys_all,
                                                                int* __restrict__ sp_perm, unsigned* __restrict__ sp_vox, uint2* __restrict__ sp_box) {
  const int room = blockIdx.y;
  const int N = (int)(room_off[room + 1] - room_off[room]);
  if (N <= 0) return;
  const unsigned* w = pw + pw_off[room];
  const unsigned long long* keys = keys_all + key_off[room];
  const long long so = sp_off[room];
  const int nblk = (int)(sp_off[room + 1] - so) / kSpBlock;
  const int lane = threadIdx.x & 31;
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nblk; b += (gridDim.x * blockDim.x) >> 5) {
    int mn[3] = {1023, 1023, 1023}, mx[3] = {0, 0, 0};
    for (int q = lane; q < kSpBlock; q += 32) {
      const int m = b * kSpBlock + q;
      const int i = m < N ? (int)(keys[m] & 0xFFFFFFFFull) : 0;
      const unsigned v = m < N ? (w[i] & 0x3FFFFFFFu) : 0x3FFFFFFFu;
      sp_perm[so + m] = i;
      sp_vox[so + m] = v;
      if (m < N) {
        const int c[3] = {(int)(v & 1023u), (int)((v >> 10) & 1023u), (int)((v >> 20) & 1023u)};
#pragma unroll
        for (int a = 0; a < 3; ++a) { mn[a] = min(mn[a], c[a]); mx[a] = max(mx[a], c[a]); }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
        mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
      }
    if (lane == 0) sp_box[so / kSpBlock + b] = make_uint2((unsigned)mn[0] | ((unsigned)mn[1] << 10) | ((unsigned)mn[2] << 20),
                                                          (unsigned)mx[0] | ((unsigned)mx[1] << 10) | ((unsigned)mx[2] << 20));
  }
}

int launch_spatial_index(int n_rooms, const long long* d_room_off, const long long* d_pw_off, const unsigned* d_pw, const long long* d_sp_off,
                         const long long* d_key_off, unsigned long long* d_keys, long long max_keys, int* d_sp_perm, unsigned* d_sp_vox,
                         uint2* d_sp_box, cudaStream_t stream, int* n_launches) {
  if (n_rooms <= 0) return LRG_OK;
  const int per_room = std::max(4, std::min(148, (4 * 148 + n_rooms - 1) / n_rooms));
  lrg_spatial_keys_kernel<<<dim3(per_room, n_rooms), 256, 0, stream>>>(d_room_off, d_pw_off, d_pw, d_key_off, d_keys);
  RoomSort rs{d_keys, nullptr, 0, d_key_off, nullptr};
  const int ns = launch_room_sort<false>(rs, n_rooms, max_keys, stream);
  if (n_launches) *n_launches += 2 + ns;
  lrg_spatial_blocks_kernel<<<dim3(per_room, n_rooms), 256, 0, stream>>>(d_room_off, d_pw_off, d_pw, d_sp_off, d_key_off, d_keys, d_sp_perm, d_sp_vox, d_sp_box);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

// Preconditions of the index-based driver (caller-prepared features, lrg_rooms_upload): the reference applies its masks by
// VOXEL (every point whose voxel is in the add / remove set, test_region_grow.py:282-287), the device by point -- the two agree
// only when every voxel holds exactly one point, which the reference's equalisation guarantees (:125-136) but a caller's
// array may not.  One CTA per room: (a) seed_order must be a permutation of 0..N-1 (find_seed indexes the words with it),
// (b) no two points of the room may share a voxel at the upload resolution (open-addressing hash set of the packed voxel
// words in scratch, 2 entries per point).  err[0] = (room + 1) | kind << 28, kind 1 = seed order, 2 = duplicate voxel.
__global__ void __launch_bounds__(1024) lrg_validate_rooms_kernel(const long long* __restrict__ room_off, const long long* __restrict__ pw_off,
                                                                  const unsigned* __restrict__ pw, const int* __restrict__ order,
                                                                  unsigned* __restrict__ scratch /* 3 * total_words, zeroed */, int* err) {
  const int room = blockIdx.x, tid = threadIdx.x;
  const long long base = room_off[room];
  const int N = (int)(room_off[room + 1] - base);
  if (N <= 0) return;
  const unsigned* w = pw + pw_off[room];
  unsigned* seen = scratch + 3 * pw_off[room];          // [N] marks of the seed order
  unsigned* table = seen + ((N + 3) & ~3);               // [2 * n4] voxel word + 1, 0 = empty
  const unsigned cap = 2u * (unsigned)((N + 3) & ~3);
  int bad = 0;
  for (int i = tid; i < N; i += 1024) {
    const int o = order[base + i];
    if (o < 0 || o >= N || atomicExch(&seen[o], 1u) != 0u) bad |= 1;
    const unsigned key = (w[i] & PW_XYZ) + 1u;
    unsigned h = (key * 2654435761u) % cap;
    while (true) {
      const unsigned prev = atomicCAS(&table[h], 0u, key);
      if (prev == 0u) break;
      if (prev == key) { bad |= 2; break; }
      h = h + 1 == cap ? 0 : h + 1;
    }
  }
  if (bad) atomicCAS(err, 0, (room + 1) | ((bad & 1 ? 1 : 2) << 28));
}

int launch_validate_rooms(int n_rooms, const long long* d_room_off, const long long* d_pw_off, const unsigned* d_pw, const int* d_order,
                          unsigned* d_scratch, int* d_err, cudaStream_t stream) {
  if (n_rooms <= 0) return LRG_OK;
  lrg_validate_rooms_kernel<<<n_rooms, 1024, 0, stream>>>(d_room_off, d_pw_off, d_pw, d_order, d_scratch, d_err);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int launch_pack(const float* d_points, int F, int n_rooms, const long long* d_room_off, const long long* d_pw_off, float resolution,
                float* d_pts16, unsigned* d_pw, int4* d_room_vmin, int* d_err, cudaStream_t stream) {
  if (n_rooms <= 0) return LRG_OK;
  lrg_pack_kernel<<<n_rooms, 1024, 0, stream>>>(d_points, F, d_room_off, d_pw_off, resolution, d_pts16, d_pw, d_room_vmin, d_err);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

int launch_reset_words(int n_rooms, const long long* d_room_off, const long long* d_pw_off, unsigned* d_pw, int lanes, long long lane_stride,
                       cudaStream_t stream) {
  if (n_rooms <= 0) return LRG_OK;
  lrg_reset_words_kernel<<<n_rooms, 1024, 0, stream>>>(d_room_off, d_pw_off, d_pw, lanes < 1 ? 1 : lanes, lane_stride);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

// ----------------------------------------------------------------------------------------------------- step
__global__ void __launch_bounds__(kStepThreads, 1) lrg_step_kernel(const __grid_constant__ DriverArgs da) {
  extern __shared__ __align__(16) unsigned char step_smem[];
  step_body<kStepThreads>(da, blockIdx.x, *reinterpret_cast<StepShared*>(step_smem));
}

size_t step_smem_bytes() { return sizeof(StepShared); }

int launch_step(const DriverArgs& da, cudaStream_t stream) {
  // (per call: the attribute belongs to the current device / context, and an engine may live on any device)
  LRG_CUDA(cudaFuncSetAttribute(lrg_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepShared)));
  lrg_step_kernel<<<da.n_slots, kStepThreads, sizeof(StepShared), stream>>>(da);
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

// The same search through the room's spatial index (lrg_spatial_index_kernel): the 13-D squared distance is at least its x, y, z
// part, and that is at least the distance from the query to the bounding box of a block of Morton-ordered points -- a block
// whose bound exceeds the best distance found so far cannot hold the nearest labelled point.  One warp per unlabeled point:
// (1) the bound of every block, the block with the smallest bound is searched first (it usually holds the query's own
// surroundings); (2) every other block whose bound does not exceed the best distance so far is searched, best first updated
// after each.  Distances are numpy_sqdist as above and ties go to the smallest point index, so the result is the brute-force
// kernel's, bit for bit -- with ~1 % of its distance evaluations on a 12 k-point room and ~0.1 % on a 177 k-point scene.
// Bounds are computed from voxel boxes (coordinate = rint(x / resolution), so x lies within 0.51 voxels of its voxel's centre
// -- 0.5 plus the rounding of the division) and shaved by 1e-4 so that no float rounding can put a bound above a distance.
__global__ void __launch_bounds__(256) lrg_fill_spatial_kernel(const __grid_constant__ FillArgs fa) {
  const int room = blockIdx.y;
  const long long base = fa.room_off[room];
  const int N = (int)(fa.room_off[room + 1] - base);
  const int nl = fa.n_lab[room], nu = fa.n_unl[room];
  if (nl == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* pts = fa.pts + base * 16;
  const int* label = fa.label + base;
  const long long so = fa.sp_off[room];
  const int nblk = (N + kSpBlock - 1) / kSpBlock;
  const uint2* box = fa.sp_box + so / kSpBlock;
  const int4* perm4 = reinterpret_cast<const int4*>(fa.sp_perm + so);
  const int4 vm = fa.room_vmin[room];
  const float res = fa.resolution;
  for (int u = blockIdx.x * 8 + warp; u < nu; u += gridDim.x * 8) {
    const int i = fa.unl_list[base + u];
    float q[16];
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      float4 t = *reinterpret_cast<const float4*>(pts + (size_t)i * 16 + c);
      q[c] = t.x; q[c + 1] = t.y; q[c + 2] = t.z; q[c + 3] = t.w;
    }
    auto bound = [&](int b) -> float {
      const uint2 bb = __ldg(box + b);
      float s = 0.f;
      const int lo[3] = {(int)(bb.x & 1023u) + vm.x, (int)((bb.x >> 10) & 1023u) + vm.y, (int)((bb.x >> 20) & 1023u) + vm.z};
      const int hi[3] = {(int)(bb.y & 1023u) + vm.x, (int)((bb.y >> 10) & 1023u) + vm.y, (int)((bb.y >> 20) & 1023u) + vm.z};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float l = ((float)lo[a] - 0.51f) * res, h = ((float)hi[a] + 0.51f) * res;
        const float d = fmaxf(0.f, fmaxf(l - q[a], q[a] - h));
        s += d * d;
      }
      return s * 0.9999f;
    };
    float best = INFINITY;
    int bestidx = INT_MAX;
    auto search = [&](int b) {                             // (warp-uniform b) the labelled points of block b against the query
      const int m0 = b * kSpBlock + lane * 4;
      const int4 p = __ldg(perm4 + (m0 >> 2));
      const int idx[4] = {p.x, p.y, p.z, p.w};
      int lab[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) lab[k] = (m0 + k < N) ? label[idx[k]] : 0;
      float lb = INFINITY;
      int li = INT_MAX;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lab[k] != 0) {
          const float* row = pts + (size_t)idx[k] * 16;
          float rv[16];
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            float4 t = *reinterpret_cast<const float4*>(row + c);
            rv[c] = t.x; rv[c + 1] = t.y; rv[c + 2] = t.z; rv[c + 3] = t.w;
          }
          const float d = numpy_sqdist(rv, q, fa.F);
          if (d < lb || (d == lb && idx[k] < li)) { lb = d; li = idx[k]; }
        }
#pragma unroll
      for (int dlt = 16; dlt > 0; dlt >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, lb, dlt);
        const int oi = __shfl_xor_sync(0xffffffffu, li, dlt);
        if (ob < lb || (ob == lb && oi < li)) { lb = ob; li = oi; }
      }
      if (lb < best || (lb == best && li < bestidx)) { best = lb; bestidx = li; }
    };
    // (1) the most promising block first
    float mb = INFINITY;
    int mblk = INT_MAX;
    for (int b = lane; b < nblk; b += 32) {
      const float s = bound(b);
      if (s < mb) { mb = s; mblk = b; }
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, mb, dlt);
      const int oi = __shfl_xor_sync(0xffffffffu, mblk, dlt);
      if (ob < mb || (ob == mb && oi < mblk)) { mb = ob; mblk = oi; }
    }
    search(mblk);
    // (2) every other block that can still hold a nearer (or equally near, lower-indexed) labelled point
    for (int b0 = 0; b0 < nblk; b0 += 32) {
      const int b = b0 + lane;
      const float s = (b < nblk && b != mblk) ? bound(b) : INFINITY;
      unsigned todo = __ballot_sync(0xffffffffu, s <= best);
      while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const float sl = __shfl_sync(0xffffffffu, s, l);
        if (sl <= best) search(b0 + l);                    // (best may have dropped since the ballot)
      }
    }
    if (lane == 0 && bestidx != INT_MAX) fa.label_filled[base + i] = label[bestidx];
  }
}

int launch_fill(const FillArgs& fa, cudaStream_t stream) {
  if (fa.n_rooms <= 0) return LRG_OK;
  lrg_fill_lists_kernel<<<fa.n_rooms, kStepThreads, 0, stream>>>(fa);
  if (fa.sp_perm != nullptr) lrg_fill_spatial_kernel<<<dim3(64, fa.n_rooms), 256, 0, stream>>>(fa);
  else lrg_fill_kernel<<<dim3(64, fa.n_rooms), 256, 0, stream>>>(fa);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg
