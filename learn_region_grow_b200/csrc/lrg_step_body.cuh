// The region-grow driver of /root/reference/test_region_grow.py:175-316 as a device function: one CTA advances one
// room slot by one grow step -- apply the previous step's add/remove logits, recompute the bounding box, run the
// stuck/stop logic, advance to the next seed / next room when a region ends, scan the neighbour shell, take the
// 9-channel median, sample 512+512 points and gather the centred tiles for the next forward.
// Templated on the CTA size NT (a multiple of 32 that divides 1024) so that both the stand-alone lock-step kernel
// (1024 threads) and the persistent grow kernel (512 threads) run the same code.
#pragma once
#include <limits.h>

#include "lrg_driver.cuh"

namespace lrg {

// ----------------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ int voxel_of(float x, float res) {
  // numpy.round(points[:, :3] / resolution).astype(int)  (:175): float32 division, round-half-even
  return __float2int_rn(__fdiv_rn(x, res));
}

__device__ __forceinline__ unsigned sortable(float f) {
  unsigned u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unsortable(unsigned u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Exclusive scan of one value per warp (s_scan[0..NT/32)) by warp 0; total in s_scan[32].
template <int NT>
__device__ __forceinline__ void scan_warp_totals(int* s_scan, int warp, int lane) {
  if (warp == 0) {
    int v = lane < NT / 32 ? s_scan[lane] : 0;
    int inc2 = warp_incl_scan(v, lane);
    s_scan[lane] = inc2 - v;
    if (lane == 31) s_scan[32] = inc2;
  }
}

// Ordered block-wide compaction of {i in [0,n) : pred(i)} into out (ascending); returns the count to every thread.
// s_scan: 33 ints of shared memory.
template <int NT, class Pred>
__device__ int block_compact(int n, Pred pred, int* __restrict__ out, int* s_scan) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int running = 0;
  for (int start = 0; start < n; start += NT * 4) {
    const int i0 = start + tid * 4;
    unsigned flags = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i0 + q < n && pred(i0 + q)) flags |= 1u << q;
    const int cnt = __popc(flags);
    const int incl = warp_incl_scan(cnt, lane);
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    scan_warp_totals<NT>(s_scan, warp, lane);
    __syncthreads();
    int off = running + s_scan[warp] + incl - cnt;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if ((flags >> q) & 1u) out[off++] = i0 + q;
    running += s_scan[32];
    __syncthreads();
  }
  return running;
}

// Block-wide radix select.  NS key streams, RPK ranks per stream (virtual channel v = s*RPK + r, NS*RPK <= 32).
// keyfn(j, k[NS], valid[NS]) yields the sortable keys of element j.  On return s_prefix[v] is the key of rank
// s_rank_in[v] and s_rank[v] the rank *within* the run of keys equal to it.
template <int NT, int NS, int RPK, class KeyFn>
__device__ void block_radix_select(int nmax, KeyFn keyfn, unsigned* s_prefix, int* s_rank, int* s_hist) {
  constexpr int NV = NS * RPK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < NV) s_prefix[tid] = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < NV * 256; i += NT) s_hist[i] = 0;
    __syncthreads();
    const unsigned himask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int j = tid; j < nmax; j += NT) {
      unsigned k[NS];
      bool valid[NS];
      keyfn(j, k, valid);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (!valid[s]) continue;
#pragma unroll
        for (int r = 0; r < RPK; ++r) {
          const int v = s * RPK + r;
          if (((k[s] ^ s_prefix[v]) & himask) == 0) atomicAdd(&s_hist[v * 256 + ((k[s] >> shift) & 255)], 1);
        }
      }
    }
    __syncthreads();
    for (int v = warp; v < NV; v += NT / 32) {
      int c[8], sum = 0;
#pragma unroll
      for (int t = 0; t < 8; ++t) { c[t] = s_hist[v * 256 + lane * 8 + t]; sum += c[t]; }
      const int incl = warp_incl_scan(sum, lane);
      int acc = incl - sum;
      const int rank = s_rank[v];
      if (acc <= rank && rank < incl) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (rank < acc + c[t]) {
            s_prefix[v] |= (unsigned)(lane * 8 + t) << shift;
            s_rank[v] = rank - acc;
            break;
          }
          acc += c[t];
        }
      }
    }
    __syncthreads();
  }
}

// choice(n, K, replace=False) over the Philox keys: the K smallest (key, index) pairs in ascending index order
// (oracle/lrg_driver.py PhiloxRng.sample).  T = key of rank K-1, E = how many keys equal to T are taken.
template <int NT>
__device__ void block_select_smallest(int n, const unsigned* __restrict__ keys, unsigned T, int E, int* s_out,
                                      int* s_scan /* 66 ints */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int run_sel = 0, run_eq = 0;
  for (int start = 0; start < n; start += NT * 4) {
    const int i0 = start + tid * 4;
    unsigned fl = 0, fe = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i0 + q < n) {
        unsigned k = keys[i0 + q];
        if (k < T) fl |= 1u << q;
        if (k == T) fe |= 1u << q;
      }
    const int ce = __popc(fe);
    const int incl_e = warp_incl_scan(ce, lane);
    if (lane == 31) s_scan[warp] = incl_e;
    __syncthreads();
    scan_warp_totals<NT>(s_scan, warp, lane);
    __syncthreads();
    int eq_before = run_eq + s_scan[warp] + incl_e - ce;
    const int eq_total = s_scan[32];
    unsigned sel = fl;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if ((fe >> q) & 1u) {
        if (eq_before < E) sel |= 1u << q;
        ++eq_before;
      }
    __syncthreads();
    const int cs = __popc(sel);
    const int incl_s = warp_incl_scan(cs, lane);
    if (lane == 31) s_scan[warp] = incl_s;
    __syncthreads();
    scan_warp_totals<NT>(s_scan, warp, lane);
    __syncthreads();
    int off = run_sel + s_scan[warp] + incl_s - cs;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if ((sel >> q) & 1u) {
        if (off < kMaxTilePts) s_out[off] = i0 + q;
        ++off;
      }
    run_sel += s_scan[32];
    run_eq += eq_total;
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------------- step
struct StepShared {
  SlotState S;
  int scan[68];
  unsigned prefix[32];
  int rank[32];
  int hist[18 * 256];
  int sel[2][kMaxTilePts];      // sampled list positions: [0] inlier, [1] neighbor
  int4 odd[2 * kMaxTilePts];    // re-rounded voxels that do not match their source point (x, y, z, kind)
  int n_odd;
  int red[32 * 6];
  int flag;
  int all_done;                 // set when this call retired the last slot of the run
};

enum { MODE_NEW_REGION = 0, MODE_SCAN = 1 };

__device__ __forceinline__ float confidence(float l0, float l1) {
  // scipy.special.softmax over the two logits, column 1 (test_region_grow.py:262-263)
  if (l1 >= l0) return 1.f / (expf(l0 - l1) + 1.f);
  float e = expf(l1 - l0);
  return e / (1.f + e);
}

// On return sh.S holds the slot's state (also written back): S.active != 0 means tiles are ready for a forward,
// S.finished != 0 means the slot has retired (no rooms left).
template <int NT>
__device__ void step_body(const DriverArgs& da, const int slot, StepShared& sh) {
  constexpr int VT = 2 * kMaxTilePts / NT;       // tile rows (512 inlier + 512 neighbor) handled per thread
  static_assert(VT * NT == 2 * kMaxTilePts, "NT must divide 1024");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  SlotState* gS = da.slots + slot;
  for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
    reinterpret_cast<int*>(&sh.S)[i] = __ldcg(reinterpret_cast<const int*>(gS) + i);
  if (tid == 0) { sh.n_odd = 0; sh.flag = 0; sh.all_done = 0; }
  __syncthreads();
  SlotState& S = sh.S;
  if (S.finished) return;

  int* listI = da.listI + (size_t)slot * da.maxN;
  int* listJ = da.listJ + (size_t)slot * da.maxN;
  unsigned* keyI = da.keyI + (size_t)slot * da.maxN;
  unsigned* keyJ = da.keyJ + (size_t)slot * da.maxN;
  const float res = da.resolution;

  // room-dependent pointers (re-derived whenever the slot moves to another room)
  long long base = 0;
  int N = 0;
  const float* pts = nullptr;
  const int4* vox = nullptr;
  unsigned char* state = nullptr;
  auto bind_room = [&]() {
    base = da.room_off[S.room];
    N = (int)(da.room_off[S.room + 1] - base);
    pts = da.pts + base * 16;
    vox = da.vox + base;
    state = da.state + base;
  };
  if (S.room >= 0) bind_room();

  // stop_growing (:210-217): visited |= current; label when the region is larger than the threshold
  auto stop_region = [&](int reason) {
    int cnt = 0;
    for (int i = tid; i < N; i += NT) cnt += (state[i] & ST_CUR) ? 1 : 0;
    int w = cnt;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
    if (lane == 0) sh.red[warp] = w;
    __syncthreads();
    int total = 0;
    for (int i = 0; i < NT / 32; ++i) total += sh.red[i];
    const bool labelled = total > da.cluster_threshold;
    int* label = da.label + base;
    for (int i = tid; i < N; i += NT) {
      unsigned char st = state[i];
      if (st & ST_CUR) {
        state[i] = ST_VISITED;
        if (labelled) label[i] = S.cluster_id;
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (labelled) S.cluster_id += 1;
      S.regions += 1;
      int slot_r = reason == STOP_NONEIGHBOR ? 0 : reason == STOP_NOEXPAND ? 1 : reason == STOP_STUCK ? 2 : 3;
      S.stops[slot_r] += 1;
      S.active = 0;
    }
    __syncthreads();
    return total;
  };

  int mode = MODE_NEW_REGION;

  // ------------------------------------------------------------------ apply the pending step (:262-306)
  if (S.active) {
    const int room_rng = da.room_id_base + S.room;
    const unsigned step_rng = (unsigned)S.total_steps;
    LrgStepTrace* tr = nullptr;
    if (da.trace != nullptr && S.total_steps < da.trace_capacity)
      tr = da.trace + (size_t)S.room * da.trace_capacity + S.total_steps;
    bool m_[VT], normal_[VT];
    int p_[VT];
    int upd = 0;
#pragma unroll
    for (int k = 0; k < VT; ++k) {
      const int vt = tid + k * NT;                        // virtual thread: 0..511 inlier rows (remove), 512.. neighbor rows (add)
      const bool is_add = vt >= kMaxTilePts;
      const int r = is_add ? vt - kMaxTilePts : vt;
      const int nrows = is_add ? da.Nj : da.Ni;
      bool m = false;
      int p = -1;
      if (r < nrows) {
        // padding rows duplicate a distinct row (:239-240,251-252): same input, same logits, own uniform draw
        const int src = __ldcg(da.tilesrc[is_add ? 1 : 0] + (size_t)slot * kMaxTilePts + r);
        const float2 lg = __ldcg(reinterpret_cast<const float2*>(da.logits[is_add ? 1 : 0] + ((size_t)slot * nrows + src) * 2));
        const float conf = confidence(lg.x, lg.y);
        const unsigned draw = philox_draw(da.seed, room_rng, step_rng, is_add ? kStreamAddUniform : kStreamRemoveUniform, r);
        const float u = (float)(draw >> 8) * (1.0f / 16777216.0f);
        m = u < conf;                                          // :266-267
        p = __ldcg(da.tileidx[is_add ? 1 : 0] + (size_t)slot * kMaxTilePts + r);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, m);
      if (tr != nullptr && lane == 0) {
        const int vwarp = vt >> 5;
        if (is_add) tr->add_mask[vwarp - kMaxTilePts / 32] = bal; else tr->remove_mask[vwarp] = bal;
      }
      // un-centre x,y, re-voxelise (:270-277); a row whose voxel no longer equals its source point's voxel goes
      // through the exact set-membership path below
      bool normal = false;
      if (m) {
        const float cx = S.center[0], cy = S.center[1];
        const float x = __fadd_rn(__fsub_rn(pts[(size_t)p * 16 + 0], cx), cx);
        const float y = __fadd_rn(__fsub_rn(pts[(size_t)p * 16 + 1], cy), cy);
        const int4 v = vox[p];
        const int vx = voxel_of(x, res), vy = voxel_of(y, res);
        normal = (vx == v.x && vy == v.y);
        if (!normal) {
          int o = atomicAdd(&sh.n_odd, 1);
          sh.odd[o] = make_int4(vx, vy, v.z, is_add ? 1 : 0);
        }
      }
      // adds first, removes second (:283-286)
      if (m && normal && is_add) { state[p] = (unsigned char)(state[p] | ST_CUR); upd = 1; }
      m_[k] = m; normal_[k] = normal; p_[k] = p;
    }
    __syncthreads();
    const int n_odd = sh.n_odd;
    if (n_odd > 0) {
      for (int i = tid; i < N; i += NT) {
        const int4 v = vox[i];
        bool hit = false;
        for (int o = 0; o < n_odd; ++o) hit |= (sh.odd[o].w == 1 && sh.odd[o].x == v.x && sh.odd[o].y == v.y && sh.odd[o].z == v.z);
        if (hit && !(state[i] & ST_CUR)) { state[i] = (unsigned char)(state[i] | ST_CUR); upd = 1; }
      }
    }
    const int updated = __syncthreads_or(upd);
#pragma unroll
    for (int k = 0; k < VT; ++k) {
      const bool is_add = (tid + k * NT) >= kMaxTilePts;
      if (m_[k] && normal_[k] && !is_add) state[p_[k]] = (unsigned char)(state[p_[k]] & ~ST_CUR);
    }
    if (n_odd > 0) {
      __syncthreads();
      for (int i = tid; i < N; i += NT) {
        const int4 v = vox[i];
        bool hit = false;
        for (int o = 0; o < n_odd; ++o) hit |= (sh.odd[o].w == 0 && sh.odd[o].x == v.x && sh.odd[o].y == v.y && sh.odd[o].z == v.z);
        if (hit) state[i] = (unsigned char)(state[i] & ~ST_CUR);
      }
    }
    __syncthreads();
    if (tid == 0) { S.steps += 1; S.total_steps += 1; }     // :288
    __syncthreads();

    int reason = STOP_NONE;
    int size_after = -1;
    if (!updated) {
      reason = STOP_NOEXPAND;                                // :304-306
    } else {
      // inlier list + bounding box of the updated region (:292-293)
      int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
      const int n_in = block_compact<NT>(N, [&](int i) {
        if (!(state[i] & ST_CUR)) return false;
        const int4 v = vox[i];
        mn[0] = min(mn[0], v.x); mn[1] = min(mn[1], v.y); mn[2] = min(mn[2], v.z);
        mx[0] = max(mx[0], v.x); mx[1] = max(mx[1], v.y); mx[2] = max(mx[2], v.z);
        return true;
      }, listI, sh.scan);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
          mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
        }
      if (lane == 0)
        for (int a = 0; a < 3; ++a) { sh.red[warp * 6 + a] = mn[a]; sh.red[warp * 6 + 3 + a] = mx[a]; }
      __syncthreads();
      size_after = n_in;
      if (tid == 0) {
        S.n_in = n_in;
        if (n_in > 0) {
          for (int a = 0; a < 3; ++a) {
            int lo = INT_MAX, hi = INT_MIN;
            for (int w = 0; w < NT / 32; ++w) { lo = min(lo, sh.red[w * 6 + a]); hi = max(hi, sh.red[w * 6 + 3 + a]); }
            S.minD[a] = lo; S.maxD[a] = hi;
          }
          bool expanded = false;
          for (int a = 0; a < 3; ++a) expanded |= (S.minD[a] < S.seqMin[a]) || (S.maxD[a] > S.seqMax[a]);
          int rsn = STOP_NONE;
          if (!expanded) {                                   // :294-299
            if (S.stuck >= 1) rsn = STOP_STUCK; else S.stuck += 1;
          } else {
            S.stuck = 0;                                     // :300-301
          }
          for (int a = 0; a < 3; ++a) { S.seqMin[a] = min(S.seqMin[a], S.minD[a]); S.seqMax[a] = max(S.seqMax[a], S.maxD[a]); }
          if (rsn == STOP_NONE && da.max_steps > 0 && S.steps >= da.max_steps) rsn = STOP_MAXSTEPS;
          sh.flag = rsn;
        } else {
          sh.flag = STOP_EMPTY;
        }
      }
      __syncthreads();
      reason = sh.flag;
      __syncthreads();
    }
    if (tr != nullptr && tid == 0) { tr->stop_reason = reason; tr->size_after = size_after; }
    if (reason != STOP_NONE) {
      int total = stop_region(reason);
      if (tr != nullptr && tid == 0 && size_after < 0) tr->size_after = total;
      mode = MODE_NEW_REGION;
    } else {
      mode = MODE_SCAN;
    }
  }

  // ------------------------------------------------------------------ find the next region that needs a forward
  while (true) {
    if (mode == MODE_NEW_REGION) {
      // next unvisited seed in curvature order (:183-188)
      int found = -1;
      if (S.room >= 0) {
        const int* order = da.order + base;
        for (int start = S.cursor; start < N; start += NT) {
          const int pos = start + tid;
          const bool ok = pos < N && !(state[order[pos]] & ST_VISITED);
          const unsigned bal = __ballot_sync(0xffffffffu, ok);
          if (lane == 0) sh.red[warp] = bal ? (start + warp * 32 + __ffs(bal) - 1) : INT_MAX;
          __syncthreads();
          int best = INT_MAX;
          for (int w = 0; w < NT / 32; ++w) best = min(best, sh.red[w]);
          __syncthreads();
          if (best != INT_MAX) { found = best; break; }
        }
      }
      if (found < 0) {
        // room exhausted (or first launch): publish its stats and fetch the next room from the queue
        if (tid == 0) {
          if (S.room >= 0) {
            LrgRoomStats& st = da.stats[S.room];
            st.n_points = N; st.grow_steps = S.total_steps; st.regions = S.regions; st.clusters = S.cluster_id - 1;
            st.stop_noneighbor = S.stops[0]; st.stop_noexpand = S.stops[1]; st.stop_stuck = S.stops[2]; st.stop_other = S.stops[3];
          }
          const int nr = atomicAdd(da.next_room, 1);
          S.room = nr < da.n_rooms ? nr : -1;
          S.cursor = 0; S.cluster_id = 1; S.total_steps = 0; S.regions = 0;
          S.stops[0] = S.stops[1] = S.stops[2] = S.stops[3] = 0;
          S.active = 0;
        }
        __syncthreads();
        if (S.room < 0) {
          if (tid == 0) {
            S.finished = 1;
            const int fin = atomicAdd(da.finished_slots, 1) + 1;
            if (fin == da.n_slots) {
              sh.all_done = 1;
              if (da.done_flag != nullptr) { *da.done_flag = 1; __threadfence_system(); }
            }
          }
          __syncthreads();
          break;
        }
        bind_room();
        continue;
      }
      // begin a region at the seed (:189-205)
      const int seed = da.order[base + found];
      if (tid == 0) {
        const int4 v = vox[seed];
        S.cursor = found + 1;
        S.seed = seed;
        S.minD[0] = S.maxD[0] = S.seqMin[0] = S.seqMax[0] = v.x;
        S.minD[1] = S.maxD[1] = S.seqMin[1] = S.seqMax[1] = v.y;
        S.minD[2] = S.maxD[2] = S.seqMin[2] = S.seqMax[2] = v.z;
        S.stuck = 0; S.steps = 0; S.n_in = 1;
        state[seed] = (unsigned char)(state[seed] | ST_CUR);
        listI[0] = seed;
      }
      __syncthreads();
      mode = MODE_SCAN;
    }
    // neighbour shell: bbox +- 1 voxel, not current, not visited (:222-229)
    const int lo0 = S.minD[0] - 1, lo1 = S.minD[1] - 1, lo2 = S.minD[2] - 1;
    const int hi0 = S.maxD[0] + 1, hi1 = S.maxD[1] + 1, hi2 = S.maxD[2] + 1;
    const int n_nb = block_compact<NT>(N, [&](int i) {
      if (state[i] & (ST_CUR | ST_VISITED)) return false;
      const int4 v = vox[i];
      return v.x >= lo0 && v.x <= hi0 && v.y >= lo1 && v.y <= hi1 && v.z >= lo2 && v.z <= hi2;
    }, listJ, sh.scan);
    if (n_nb == 0) {                                          // :233-235
      stop_region(STOP_NONEIGHBOR);
      mode = MODE_NEW_REGION;
      continue;
    }
    if (tid == 0) S.n_nb = n_nb;
    __syncthreads();
    break;
  }

  if (!S.finished) {
    // ---------------------------------------------------------------- tiles for the next forward (:237-254)
    const int n_in = S.n_in, n_nb = S.n_nb;
    const int room_rng = da.room_id_base + S.room;
    const unsigned step_rng = (unsigned)S.total_steps;
    // median of every centred channel over ALL current points (:241): channels 0,1 and 6..F-1
    const int nch = 2 + (da.F > 6 ? da.F - 6 : 0);
    if (tid < 18) sh.rank[tid] = (tid & 1) ? (n_in / 2) : ((n_in - 1) / 2);
    __syncthreads();
    block_radix_select<NT, 9, 2>(n_in, [&](int j, unsigned (&k)[9], bool (&valid)[9]) {
      const float* row = pts + (size_t)listI[j] * 16;
      const float4 a = *reinterpret_cast<const float4*>(row);
      const float4 b = *reinterpret_cast<const float4*>(row + 4);
      const float4 c = *reinterpret_cast<const float4*>(row + 8);
      const float4 d = *reinterpret_cast<const float4*>(row + 12);
      const float vals[9] = {a.x, a.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x};
#pragma unroll
      for (int s = 0; s < 9; ++s) { k[s] = sortable(vals[s]); valid[s] = s < nch; }
    }, sh.prefix, sh.rank, sh.hist);
    if (tid < 16) {
      float cval = 0.f;
      const int ch = tid < 2 ? tid : tid - 4;                 // feature column -> median channel
      if ((tid < 2 || tid >= 6) && tid < da.F) {
        const float lo = unsortable(sh.prefix[ch * 2]), hi = unsortable(sh.prefix[ch * 2 + 1]);
        cval = (n_in & 1) ? lo : __fmul_rn(__fadd_rn(lo, hi), 0.5f);   // numpy.median: mean of the two middle values
      }
      S.center[tid] = cval;
    }
    __syncthreads();

    // sampling (:237-240, :249-252) with the Philox stream of oracle/lrg_driver.py PhiloxRng
    const bool fullI = n_in >= da.Ni, fullJ = n_nb >= da.Nj;
    if (fullI) for (int j = tid; j < n_in; j += NT) keyI[j] = philox_draw(da.seed, room_rng, step_rng, kStreamInlierKey, j);
    if (fullJ) for (int j = tid; j < n_nb; j += NT) keyJ[j] = philox_draw(da.seed, room_rng, step_rng, kStreamNeighborKey, j);
    if (tid == 0) { sh.rank[0] = da.Ni - 1; sh.rank[1] = da.Nj - 1; }
    __syncthreads();
    if (fullI || fullJ) {
      const int nmax = max(fullI ? n_in : 0, fullJ ? n_nb : 0);
      block_radix_select<NT, 2, 1>(nmax, [&](int j, unsigned (&k)[2], bool (&valid)[2]) {
        valid[0] = fullI && j < n_in; valid[1] = fullJ && j < n_nb;
        k[0] = valid[0] ? keyI[j] : 0u; k[1] = valid[1] ? keyJ[j] : 0u;
      }, sh.prefix, sh.rank, sh.hist);
    }
    const unsigned TI = sh.prefix[0], TJ = sh.prefix[1];
    const int EI = sh.rank[0] + 1, EJ = sh.rank[1] + 1;
    __syncthreads();
    if (fullI) block_select_smallest<NT>(n_in, keyI, TI, EI, sh.sel[0], sh.scan);
    else
      for (int r = tid; r < da.Ni; r += NT)
        sh.sel[0][r] = r < n_in ? r : (int)__umulhi(philox_draw(da.seed, room_rng, step_rng, kStreamInlierPad, r - n_in), (unsigned)n_in);
    if (fullJ) block_select_smallest<NT>(n_nb, keyJ, TJ, EJ, sh.sel[1], sh.scan);
    else
      for (int r = tid; r < da.Nj; r += NT)
        sh.sel[1][r] = r < n_nb ? r : (int)__umulhi(philox_draw(da.seed, room_rng, step_rng, kStreamNeighborPad, r - n_nb), (unsigned)n_nb);
    __syncthreads();

    // gather + centre (:242-247, :253): columns 0:2 and 6: are centred, z and the room coordinates are not
    {
      const bool tracing = da.trace != nullptr && S.total_steps < da.trace_capacity;
#pragma unroll
      for (int k = 0; k < VT; ++k) {
        const int vt = tid + k * NT;
        const bool is_nb = vt >= kMaxTilePts;
        const int r = is_nb ? vt - kMaxTilePts : vt;
        const int nrows = is_nb ? da.Nj : da.Ni;
        unsigned crc_term = 0;
        if (r < nrows) {
          const int nset = is_nb ? n_nb : n_in;
          const int pos = sh.sel[is_nb ? 1 : 0][r];
          const int p = (is_nb ? listJ : listI)[pos];
          if (r < nset) {                                    // a distinct point: materialise its tile row
            const float* row = pts + (size_t)p * 16;
            float* out = da.tile[is_nb ? 1 : 0] + ((size_t)slot * nrows + r) * da.F;
            for (int c = 0; c < da.F; ++c) {
              float v = row[c];
              if (c < 2 || c >= 6) v = __fsub_rn(v, S.center[c]);
              out[c] = v;
            }
          }
          da.tileidx[is_nb ? 1 : 0][(size_t)slot * kMaxTilePts + r] = p;
          da.tilesrc[is_nb ? 1 : 0][(size_t)slot * kMaxTilePts + r] = r < nset ? r : pos;   // padding: pos < nset is the row it copies
          crc_term = (unsigned)(r + 1) * (unsigned)p;
        }
        if (tracing) {
          unsigned w = crc_term;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
          if (lane == 0) sh.red[vt >> 5] = (int)w;
        }
      }
      if (tracing) {
        __syncthreads();
        if (tid == 0) {
          LrgStepTrace* tr = da.trace + (size_t)S.room * da.trace_capacity + S.total_steps;
          unsigned ci = 0, cj = 0;
          for (int w2 = 0; w2 < 16; ++w2) { ci += (unsigned)sh.red[w2]; cj += (unsigned)sh.red[16 + w2]; }
          tr->seed_point = S.seed; tr->step_in_region = S.steps; tr->n_inlier = n_in; tr->n_neighbor = n_nb;
          tr->inlier_idx_crc = ci; tr->neighbor_idx_crc = cj;
          for (int c = 0; c < 16; ++c) tr->center[c] = S.center[c];
        }
      }
    }
    for (int i = tid; i < da.pooled_per_slot; i += NT) da.pooled[(size_t)slot * da.pooled_per_slot + i] = 0.f;
    if (tid == 0) S.active = 1;
  }
  __syncthreads();
  for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
    reinterpret_cast<int*>(gS)[i] = reinterpret_cast<const int*>(&sh.S)[i];
}

}  // namespace lrg
