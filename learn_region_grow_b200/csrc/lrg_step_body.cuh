// The region-grow driver of /root/reference/test_region_grow.py:175-316 as a device function: one CTA advances one
// room slot by one grow step -- apply the previous step's add/remove logits, recompute the bounding box, run the
// stuck/stop logic, advance to the next seed / next room when a region ends, scan the neighbour shell, take the
// 9-channel median, sample 512+512 points and gather the centred tiles for the next forward.
// Templated on the CTA size NT (a multiple of 32 that divides 1024) so that both the stand-alone lock-step kernel
// (1024 threads) and the persistent grow kernel (512 threads) run the same code.
//
// A step is a chain of dependent phases, so it is built for latency: the per-point state is ONE packed 32-bit word
// (10+10+10 bits of room-relative voxel coordinates, CURRENT and VISITED flags) that a whole-room scan reads with every
// 128-bit load in flight at once (one L2 round trip per scan), the lists are produced by a single ordered block
// compaction per scan, and the median works on keys staged once in shared memory.
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

// ----------------------------------------------------------------------------------------------------- packed state words
constexpr unsigned PW_CUR = 1u << 30, PW_VIS = 1u << 31, PW_XYZ = 0x3FFFFFFFu;
__device__ __forceinline__ int pw_x(unsigned w) { return (int)(w & 1023u); }
__device__ __forceinline__ int pw_y(unsigned w) { return (int)((w >> 10) & 1023u); }
__device__ __forceinline__ int pw_z(unsigned w) { return (int)((w >> 20) & 1023u); }

// Ordered block-wide compaction over a room's state words: out[] receives, ascending, every i with pred(word_i);
// visit(word) is called for each selected word (bounding boxes).  Every thread issues all of its 128-bit loads of a
// 32*NT-point chunk before using any (one memory round trip per chunk); warp w owns a contiguous run of the chunk, read
// lane-interleaved so the loads coalesce (a thread-contiguous assignment needs one warp scan instead of eight but its
// 32-lines-per-instruction loads were measured 4k cycles slower); per-warp scans plus one scan of the warp totals give
// the global order.  The
// array is padded to a multiple of four words with VISITED words that no predicate selects.  The first cap_s entries are
// mirrored in shared memory (out_s) so that the phases that follow do not pay a global round trip for the list.
// (Keeping the words in registers for the second scan of a step was tried: the 32 extra live registers spill and cost more
// than the L2 round trip they save.)
// s_scan: 33 ints.  Returns the count to every thread.
// (Plain loads on purpose: the words are rewritten by this CTA between scans and by other CTAs between steps; the
// CTA barrier orders the former, the acquire fence at the start of a work item drops stale L1 lines for the latter.)
constexpr int kScanU = 8;
template <int NT, class Pred, class Visit>
__device__ int scan_words(const unsigned* pw, int N, Pred pred, Visit visit, int* out, int* out_s, int cap_s, int* s_scan, bool l2 = false) {
  constexpr int U = kScanU;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n4 = (N + 3) >> 2;
  int running = 0;
  for (int c4 = 0; c4 < n4; c4 += NT * U) {
    // warp w owns the 32*U consecutive 128-bit words from q0 - lane; lane-interleaved so that every load is coalesced
    const int q0 = c4 + warp * 32 * U + lane;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = q0 + u * 32;
      // (l2: speculative lanes update the words with atomics, which act in L2 -- read them there)
      v[u] = q >= n4 ? make_uint4(PW_VIS, PW_VIS, PW_VIS, PW_VIS) : l2 ? __ldcg(reinterpret_cast<const uint4*>(pw) + q) : reinterpret_cast<const uint4*>(pw)[q];
    }
    unsigned flags = 0;
    int offs[U], wtotal = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned f = 0;
      if (pred(v[u].x)) f |= 1u;
      if (pred(v[u].y)) f |= 2u;
      if (pred(v[u].z)) f |= 4u;
      if (pred(v[u].w)) f |= 8u;
      flags |= f << (4 * u);
      const int cnt = __popc(f);
      const int incl = warp_incl_scan(cnt, lane);
      offs[u] = wtotal + incl - cnt;
      wtotal += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_scan[warp] = wtotal;
    __syncthreads();
    scan_warp_totals<NT>(s_scan, warp, lane);
    __syncthreads();
    const int wbase = running + s_scan[warp];
    if (flags) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned f = (flags >> (4 * u)) & 15u;
        if (f) {
          int o = wbase + offs[u];
          const int i0 = (q0 + u * 32) * 4;
          if (f & 1u) { out[o] = i0; if (o < cap_s) out_s[o] = i0; ++o; visit(v[u].x); }
          if (f & 2u) { out[o] = i0 + 1; if (o < cap_s) out_s[o] = i0 + 1; ++o; visit(v[u].y); }
          if (f & 4u) { out[o] = i0 + 2; if (o < cap_s) out_s[o] = i0 + 2; ++o; visit(v[u].z); }
          if (f & 8u) { out[o] = i0 + 3; if (o < cap_s) out_s[o] = i0 + 3; ++o; visit(v[u].w); }
        }
      }
    }
    running += s_scan[32];
    __syncthreads();
  }
  return running;
}

// The same compaction for a BOX query through the room's spatial index (DriverArgs::sp_*, lrg_spatial_index_kernel): out[]
// receives, ascending, every point i whose voxel lies in [lo, hi] and whose state word satisfies pred; visit(word) as above.
// (a) every thread tests one block box against the query and the blocks that meet it are collected (any order); (b) one
// warp per such block: the block's coordinates and point indices are static data in Morton order (coalesced 128-bit loads),
// only the points inside the box fetch their state word (a gather by point index), and the hits set their bit in a bitmap
// over the room's point indices in shared memory; (c) the bitmap is expanded in index order -- thread t owns a contiguous run
// of its words, one block scan of the popcounts gives the offsets.  Work follows the size of the box, not of the room: on the
// bench rooms (12 k points) a shell meets ~19 % of the blocks, on the 177 k-point outdoor scenes ~6 %.
// bitmap: (N+31)/32 words, blklist: (N+127)/128 ints, s_cnt: one int, s_scan: 33 ints.  Returns the count to every thread.
// BAR = 0: the group is the whole CTA (__syncthreads); else threads 0..NT-1 of the CTA meet at named barrier BAR, so that the
// rest of the CTA can do something else meanwhile (the median of the step, below).
template <int BAR, int G>
__device__ __forceinline__ void group_sync() {
  if constexpr (BAR == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(G) : "memory");
}

template <int NT, int BAR, class Pred, class Visit>
__device__ int scan_box_spatial(const DriverArgs& da, int room, const unsigned* pw, int N, const int (&lo)[3], const int (&hi)[3], Pred pred,
                                Visit visit, int* out, int* out_s, int cap_s, int* s_scan, unsigned* bitmap, int* blklist, int* s_cnt, bool l2) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long so = da.sp_off[room];
  const int nblk = (N + kSpBlock - 1) / kSpBlock, nwords = (N + 31) >> 5;
  for (int i = tid; i < nwords; i += NT) bitmap[i] = 0u;
  if (tid == 0) *s_cnt = 0;
  group_sync<BAR, NT>();
  const uint2* box = da.sp_box + so / kSpBlock;
  // (eight boxes per thread and round, loaded before any is tested: a large scene has more blocks than the group has threads,
  //  and one box per round made every round a round trip of its own)
  constexpr int kBoxU = 8;
  for (int b0 = 0; b0 < nblk; b0 += NT * kBoxU) {
    uint2 bb[kBoxU];
#pragma unroll
    for (int u = 0; u < kBoxU; ++u) {
      const int b = b0 + u * NT + tid;
      bb[u] = b < nblk ? __ldg(box + b) : make_uint2(PW_XYZ, 0u);
    }
#pragma unroll
    for (int u = 0; u < kBoxU; ++u) {
      if (b0 + u * NT >= nblk) break;                                       // (uniform)
      const int b = b0 + u * NT + tid;
      const bool hit = b < nblk && pw_x(bb[u].x) <= hi[0] && pw_x(bb[u].y) >= lo[0] && pw_y(bb[u].x) <= hi[1] && pw_y(bb[u].y) >= lo[1] &&
                       pw_z(bb[u].x) <= hi[2] && pw_z(bb[u].y) >= lo[2];
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(s_cnt, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) blklist[base + __popc(bal & ((1u << lane) - 1u))] = b;
      }
    }
  }
  group_sync<BAR, NT>();
  const int nhit = *s_cnt;
  const uint4* vox4 = reinterpret_cast<const uint4*>(da.sp_vox + so);
  const int4* perm4 = reinterpret_cast<const int4*>(da.sp_perm + so);
  auto inside = [&](unsigned v) {
    const int x = pw_x(v), y = pw_y(v), z = pw_z(v);
    return x >= lo[0] && x <= hi[0] && y >= lo[1] && y <= hi[1] && z >= lo[2] && z <= hi[2];
  };
  // kSpU blocks per warp and round, every load of the round in flight before any is used: a round is two dependent L2 round
  // trips (static coordinates + indices, then the state words of the points inside the box) whatever kSpU is, and the ~19
  // blocks a room's shell meets fit one round (one block per warp and round took three)
  constexpr int kSpU = 4;
  for (int i0 = warp; i0 < nhit; i0 += (NT / 32) * kSpU) {      // (block i0 + u * warps: the hits spread evenly over the warps)
    uint4 v[kSpU];
    int4 p[kSpU];
#pragma unroll
    for (int u = 0; u < kSpU; ++u) {
      v[u] = make_uint4(PW_XYZ, PW_XYZ, PW_XYZ, PW_XYZ);        // (coordinates 1023: outside every box)
      p[u] = make_int4(0, 0, 0, 0);
      if (i0 + u * (NT / 32) < nhit) {
        const int q = blklist[i0 + u * (NT / 32)] * (kSpBlock / 4) + lane;
        v[u] = __ldg(vox4 + q);
        p[u] = __ldg(perm4 + q);
      }
    }
    unsigned w[kSpU][4];
    unsigned f = 0;
#pragma unroll
    for (int u = 0; u < kSpU; ++u) {
      const bool f0 = inside(v[u].x), f1 = inside(v[u].y), f2 = inside(v[u].z), f3 = inside(v[u].w);
      f |= ((f0 ? 1u : 0u) | (f1 ? 2u : 0u) | (f2 ? 4u : 0u) | (f3 ? 8u : 0u)) << (4 * u);
      // (l2: speculative lanes update the words with atomics, which act in L2 -- read them there)
      w[u][0] = !f0 ? PW_VIS : l2 ? __ldcg(pw + p[u].x) : pw[p[u].x];
      w[u][1] = !f1 ? PW_VIS : l2 ? __ldcg(pw + p[u].y) : pw[p[u].y];
      w[u][2] = !f2 ? PW_VIS : l2 ? __ldcg(pw + p[u].z) : pw[p[u].z];
      w[u][3] = !f3 ? PW_VIS : l2 ? __ldcg(pw + p[u].w) : pw[p[u].w];
    }
    if (f) {
#pragma unroll
      for (int u = 0; u < kSpU; ++u) {
        const unsigned fu = (f >> (4 * u)) & 15u;
        if (fu) {
          if ((fu & 1u) && pred(w[u][0])) { atomicOr(bitmap + (p[u].x >> 5), 1u << (p[u].x & 31)); visit(w[u][0]); }
          if ((fu & 2u) && pred(w[u][1])) { atomicOr(bitmap + (p[u].y >> 5), 1u << (p[u].y & 31)); visit(w[u][1]); }
          if ((fu & 4u) && pred(w[u][2])) { atomicOr(bitmap + (p[u].z >> 5), 1u << (p[u].z & 31)); visit(w[u][2]); }
          if ((fu & 8u) && pred(w[u][3])) { atomicOr(bitmap + (p[u].w >> 5), 1u << (p[u].w & 31)); visit(w[u][3]); }
        }
      }
    }
  }
  group_sync<BAR, NT>();
  const int wpt = (nwords + NT - 1) / NT;
  const int w_lo = min(tid * wpt, nwords), w_hi = min(w_lo + wpt, nwords);
  int cnt = 0;
  for (int w = w_lo; w < w_hi; ++w) cnt += __popc(bitmap[w]);
  const int incl = warp_incl_scan(cnt, lane);
  if (lane == 31) s_scan[warp] = incl;
  group_sync<BAR, NT>();
  scan_warp_totals<NT>(s_scan, warp, lane);
  group_sync<BAR, NT>();
  int o = s_scan[warp] + incl - cnt;
  for (int w = w_lo; w < w_hi; ++w) {
    unsigned bits = bitmap[w];
    while (bits) {
      const int i = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      out[o] = i;
      if (o < cap_s) out_s[o] = i;
      ++o;
    }
  }
  const int total = s_scan[32];
  group_sync<BAR, NT>();
  return total;
}

// ----------------------------------------------------------------------------------------------------- step
constexpr int kMedianCap = 2048;   // inlier sets up to this size have their 9 median channels staged in shared memory
constexpr int kMedianCapSlots = kMedianCap / 32;
constexpr int kMedianSmall = 256;  // up to this size the channel's own warp transposes its keys by ballots (cheaper below ~256)
constexpr int kListCap = 1024;     // leading entries of the inlier / neighbour lists mirrored in shared memory
constexpr int kIncCap = 2048;      // regions up to this size get their inlier list updated from the previous one (no room scan)

struct StepShared {
  SlotState S;
  int scan[68];
  unsigned prefix[32];
  int rank[32];
  int sel[2][kMaxTilePts];      // sampled list positions: [0] inlier, [1] neighbor
  int4 odd[2 * kMaxTilePts];    // re-rounded voxels that do not match their source point (x, y, z, kind)
  int n_odd;
  int red[32 * 6];
  int flag;
  int all_done;                 // set when this call retired the last slot of the run
  float lp[2];                  // beam search, 'ml' scoring: addLogProb, rmvLogProb of the step being applied
  float lane_score;             // ... and the expansion's score (parent + both)
  int nsel[2];                  // ... selected rows per set
  int sp_cnt;                   // indexed shell scan: blocks that meet the shell
  unsigned long long est_sum;   // speculative lanes: sum of the other rooms' work estimates
  int sp_total;                 // ... and its result when only half of the CTA ran it
  unsigned wake;                // random restarts / beam search: lanes of this group (bit l) this call handed work to (they need a STEP)
  LaneGroup G;                  // random restarts: the group's room-level state while this CTA owns it
  unsigned nextkey[16];         // median: smallest key above the lower median, per channel
  int listI_s[kListCap], listJ_s[kListCap];
  unsigned malive[kMedianCapSlots];   // median: which keys of every 32-key slot exist
  union {                        // the histograms of the radix selects are never live together with the staged median bit planes
    unsigned planes[9][32 * (kMedianCapSlots + 1)];   // [channel][bit][slot]: bit planes of the inlier keys (sets above kMedianSmall)
    unsigned mkeys[9][kMedianSmall];                  // the keys themselves (sets up to kMedianSmall)
    int hist[18 * 256];
  };
};

enum { MODE_NEW_REGION = 0, MODE_SCAN = 1 };

__device__ __forceinline__ float confidence(float l0, float l1) {
  // scipy.special.softmax over the two logits, column 1 (test_region_grow.py:262-263)
  if (l1 >= l0) return 1.f / (expf(l0 - l1) + 1.f);
  float e = expf(l1 - l0);
  return e / (1.f + e);
}

// Median of one channel by ONE warp, keys in shared memory (n <= 1024 * G): bit-sliced radix select.  First the keys
// are transposed into bit planes with ballots -- lane s ends up owning slot s (the 32 keys s*32..s*32+31) as 32 words,
// P[b] = bit b of those 32 keys -- which are independent instructions that pipeline.  The select of rank (n-1)/2 then
// walks the bits from the top with one popc + one warp reduction per bit (not per key): zeros = sum over slots of
// popc(alive & ~P[b]).  For even n the upper median is the same key again if it has a further duplicate, else the
// smallest key above it.  Returns numpy.median's two middle keys in lo / hi (equal for odd n).
template <int G>
__device__ __forceinline__ void warp_median(const unsigned* __restrict__ keys, int n, unsigned& lo, unsigned& hi) {
  const int lane = threadIdx.x & 31;
  const int nslots = (n + 31) >> 5;
  unsigned P[G][32], alive[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    alive[g] = 0;
#pragma unroll
    for (int b = 0; b < 32; ++b) P[g][b] = 0;
    const int ns = min(32, nslots - 32 * g);
    for (int s = 0; s < ns; ++s) {
      const int j = (s + 32 * g) * 32 + lane;
      const unsigned key = j < n ? keys[j] : 0u;
      const unsigned av = __ballot_sync(0xffffffffu, j < n);
      if (lane == s) alive[g] = av;
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        const unsigned bal = __ballot_sync(0xffffffffu, (key >> b) & 1u);
        if (lane == s) P[g][b] = bal;
      }
    }
  }
  int rank = (n - 1) >> 1;
  unsigned prefix = 0;
#pragma unroll
  for (int b = 31; b >= 0; --b) {
    unsigned z = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) z += __popc(alive[g] & ~P[g][b]);
    const int zeros = (int)__reduce_add_sync(0xffffffffu, z);
    const bool take_ones = rank >= zeros;            // warp-uniform
    if (take_ones) { rank -= zeros; prefix |= 1u << b; }
#pragma unroll
    for (int g = 0; g < G; ++g) alive[g] &= take_ones ? P[g][b] : ~P[g][b];
  }
  lo = prefix;
  hi = prefix;
  if ((n & 1) == 0) {
    // candidates left = keys equal to prefix; `rank` = position of the lower median inside that run
    unsigned eq = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) eq += __popc(alive[g]);
    const int equal = (int)__reduce_add_sync(0xffffffffu, eq);
    if (rank + 1 >= equal) {
      unsigned above = 0xFFFFFFFFu;
      for (int j = lane; j < n; j += 32) {
        const unsigned k = keys[j];
        if (k > prefix) above = min(above, k);
      }
      hi = __reduce_min_sync(0xffffffffu, above);
    }
  }
}

// 32 x 32 bit transpose across a warp (five exchange steps): lane i passes key i; lane b receives the word whose bit i is
// bit b of key i -- the bit plane b of the warp's 32 keys.
__device__ __forceinline__ unsigned warp_transpose32(unsigned x, int lane) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const unsigned m = d == 16 ? 0x0000FFFFu : d == 8 ? 0x00FF00FFu : d == 4 ? 0x0F0F0F0Fu : d == 2 ? 0x33333333u : 0x55555555u;
    const unsigned y = __shfl_xor_sync(0xffffffffu, x, d);
    x = (lane & d) ? ((x & ~m) | ((y >> d) & m)) : ((x & m) | ((y << d) & ~m));
  }
  return x;
}

constexpr int kPlaneStride = kMedianCapSlots + 1;      // (+1: the 32 lanes of a warp write one column, bank-conflict free)

// numpy.median's two middle keys of one channel by ONE warp from bit planes staged in shared memory: planes[b *
// kPlaneStride + s] = bit b of the 32 keys of slot s, alive_s[s] = which of them exist.  Lane s (+32g) owns slot s as 32
// registers; the select walks the bits from the top with one popc + one warp reduction per bit (not per key), for the ranks
// (n-1)/2 and n/2 at once (independent chains that pipeline).  n <= 1024 * G.
template <int G>
__device__ __forceinline__ void warp_median_planes(const unsigned* __restrict__ planes, const unsigned* __restrict__ alive_s, int n,
                                                   unsigned& lo, unsigned& hi) {
  const int lane = threadIdx.x & 31;
  const int nslots = (n + 31) >> 5;
  unsigned P[G][32], a1[G], a2[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int sidx = lane + 32 * g;
    const bool ok = sidx < nslots;
    a1[g] = a2[g] = ok ? alive_s[sidx] : 0u;
#pragma unroll
    for (int b = 0; b < 32; ++b) P[g][b] = ok ? planes[b * kPlaneStride + sidx] : 0u;
  }
  int r1 = (n - 1) >> 1, r2 = n >> 1;
  unsigned p1 = 0, p2 = 0;
#pragma unroll
  for (int b = 31; b >= 0; --b) {
    unsigned z1 = 0, z2 = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) { z1 += __popc(a1[g] & ~P[g][b]); z2 += __popc(a2[g] & ~P[g][b]); }
    // one warp reduction per bit for both ranks: the two counts (<= 2048 each) travel in the halves of one word
    const unsigned zz = __reduce_add_sync(0xffffffffu, z1 | (z2 << 16));
    const int zeros1 = (int)(zz & 0xFFFFu), zeros2 = (int)(zz >> 16);
    const bool t1 = r1 >= zeros1, t2 = r2 >= zeros2;           // warp-uniform
    if (t1) { r1 -= zeros1; p1 |= 1u << b; }
    if (t2) { r2 -= zeros2; p2 |= 1u << b; }
#pragma unroll
    for (int g = 0; g < G; ++g) { a1[g] &= t1 ? P[g][b] : ~P[g][b]; a2[g] &= t2 ? P[g][b] : ~P[g][b]; }
  }
  lo = p1;
  hi = p2;
}

// numpy.median over n keys per channel (mean of the two middle values for even n, :241): block radix select of rank
// (n-1)/2 in every channel at once -- histogram updates are aggregated per warp with match.any, because the leading
// byte of a feature channel is nearly constant over a region -- then, for even n, the next key in sorted order is either
// the same value again (when the selected key has further duplicates) or the smallest key above it (one extra pass).
// keyfn(j, k[9]) yields the sortable keys of inlier j.  On return s_prefix[c] = lower median key, s_next[c] = upper.
template <int NT, class KeyFn>
__device__ void block_median9(int n, int nch, KeyFn keyfn, unsigned* s_prefix, int* s_rank, int* s_hist, unsigned* s_next) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 9) { s_prefix[tid] = 0; s_rank[tid] = (n - 1) / 2; }
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 9 * 256; i += NT) s_hist[i] = 0;
    __syncthreads();
    const unsigned himask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int j0 = 0; j0 < n; j0 += NT) {             // whole warps iterate together (match.any needs the full warp)
      const int j = j0 + tid;
      unsigned k[9];
      if (j < n) keyfn(j, k);
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const bool on = j < n && c < nch && ((k[c] ^ s_prefix[c]) & himask) == 0;
        const unsigned bin = on ? ((k[c] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (on && lane == __ffs(peers) - 1) atomicAdd(&s_hist[c * 256 + bin], __popc(peers));
      }
    }
    __syncthreads();
    for (int c = warp; c < nch; c += NT / 32) {
      int cnt[8], sum = 0;
#pragma unroll
      for (int t = 0; t < 8; ++t) { cnt[t] = s_hist[c * 256 + lane * 8 + t]; sum += cnt[t]; }
      const int incl = warp_incl_scan(sum, lane);
      int acc = incl - sum;
      const int rank = s_rank[c];
      if (acc <= rank && rank < incl) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (rank < acc + cnt[t]) {
            s_prefix[c] |= (unsigned)(lane * 8 + t) << shift;
            s_rank[c] = rank - acc;
            if (shift == 0) s_rank[9 + c] = cnt[t];      // how many keys equal the selected one
            break;
          }
          acc += cnt[t];
        }
      }
    }
    __syncthreads();
  }
  // upper median (rank n/2) for even n
  if (tid < 9) s_next[tid] = 0xFFFFFFFFu;
  __syncthreads();
  if ((n & 1) == 0) {
    unsigned best[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) best[c] = 0xFFFFFFFFu;
    for (int j = tid; j < n; j += NT) {
      unsigned k[9];
      keyfn(j, k);
#pragma unroll
      for (int c = 0; c < 9; ++c)
        if (k[c] > s_prefix[c] && k[c] < best[c]) best[c] = k[c];
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      unsigned b = best[c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) b = min(b, __shfl_xor_sync(0xffffffffu, b, d));
      if (lane == 0 && c < nch && b != 0xFFFFFFFFu) atomicMin(&s_next[c], b);
    }
    __syncthreads();
    if (tid < nch && s_rank[tid] + 1 < s_rank[9 + tid]) s_next[tid] = s_prefix[tid];   // a duplicate of the lower median follows it
  } else if (tid < 9) {
    s_next[tid] = s_prefix[tid];
  }
  __syncthreads();
}

// On return sh.S holds the slot's state (also written back): S.active != 0 means tiles are ready for a forward,
// S.finished != 0 means the slot has retired (no rooms left).
template <int NT>
__device__ void step_body(const DriverArgs& da, const int slot, StepShared& sh) {
  constexpr int VT = 2 * kMaxTilePts / NT;       // tile rows (512 inlier + 512 neighbor) handled per thread
  static_assert(VT * NT == 2 * kMaxTilePts, "NT must divide 1024");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long tstamp = clock64();
  auto stamp = [&](int stage) {
    if (da.dbg != nullptr && tid == 0) {
      const long long now = clock64();
      atomicAdd(da.dbg + stage, (unsigned long long)(now - tstamp));
      tstamp = now;
    }
  };
  auto mark = [&](int idx) {      // diagnostics: cycles since the last stamp, without restarting the interval
    if (da.dbg != nullptr && tid == 0) atomicAdd(da.dbg + idx, (unsigned long long)(clock64() - tstamp));
  };
  // metadata of the pending step's tile rows: addressed by the slot alone, so these loads go out together with the slot
  // state (one round trip); they are only meaningful -- and only used -- when a forward is pending
  int p_[VT], src_[VT];
#pragma unroll
  for (int k = 0; k < VT; ++k) {
    const int vt = tid + k * NT;                        // virtual thread: 0..511 inlier rows (remove), 512.. neighbor rows (add)
    const bool is_add = vt >= kMaxTilePts;
    const int r = is_add ? vt - kMaxTilePts : vt;
    const bool on = r < (is_add ? da.Nj : da.Ni);
    // padding rows duplicate a distinct row (:239-240,251-252): same input, same logits, own uniform draw
    src_[k] = on ? __ldcg(da.tilesrc[is_add ? 1 : 0] + (size_t)slot * kMaxTilePts + r) : 0;
    p_[k] = on ? __ldcg(da.tileidx[is_add ? 1 : 0] + (size_t)slot * kMaxTilePts + r) : -1;
  }
  // ... and so does the head of the slot's inlier list (four consecutive entries per thread): when a forward is pending and
  // the region is small, the new list is derived from it instead of a scan over the whole room
  static_assert(NT >= kMaxTilePts && NT * 4 >= kIncCap, "the list update assumes one neighbour row and four list entries per thread");
  int li_[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = tid * 4 + q;
    li_[q] = j < da.maxN ? __ldcg(da.listI + (size_t)slot * da.maxN + j) : 0;
  }
  SlotState* gS = da.slots + slot;
  for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
    reinterpret_cast<int*>(&sh.S)[i] = __ldcg(reinterpret_cast<const int*>(gS) + i);
  if (tid == 0) { sh.n_odd = 0; sh.flag = 0; sh.all_done = 0; sh.wake = 0; sh.nsel[0] = sh.nsel[1] = 0; sh.lp[0] = sh.lp[1] = 0.f; sh.lane_score = 0.f; }
  __syncthreads();
  SlotState& S = sh.S;
  if (S.finished) return;
  // random restarts (test_random_restart.py): slot = group * L + lane; a parked lane waits for the other restarts of its seed
  const int L = da.lanes > 1 ? da.lanes : 1;
  const int lane_id = L > 1 ? slot % L : 0;
  LaneGroup* const grp = (L > 1 || da.beam_width > 0) ? da.groups + slot / L : nullptr;
  // Philox coordinates of a draw: (room, seed point of the region, step within the region, stream, element) -- keyed by the
  // region, not by the room's running step count, so that a region's draws do not depend on what was grown before it;
  // restart lane l uses streams 8l + {0..5}
  // (speculative lanes: a region's draws must not depend on the lane that happens to grow it -- stream 0 like the plain driver)
  const unsigned lane_stream = (da.spec != 0) ? 0u : 8u * (unsigned)lane_id;
  // beam search (test_beam_search.py): lane q * SW + s expands candidate q of the seed's queue for the s-th time
  const bool beam = da.beam_width > 0;
  const int BW = da.beam_width, SW = da.search_width;
  // speculative lanes (DriverArgs::spec): the lanes of a group grow different regions of one room, commits in seed order
  const bool spec = da.spec != 0 && L > 1;
  if ((L > 1 || beam) && S.parked && !((beam || spec) && S.begin)) return;
  if ((beam || spec) && S.begin) {
    // handed a candidate by the lane that closed the previous round: everything it wrote precedes the flag, so read the
    // state again behind an acquire fence (the lock-step loop may have loaded a torn record)
    __threadfence();
    __syncthreads();
    for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
      reinterpret_cast<int*>(&sh.S)[i] = __ldcg(reinterpret_cast<const int*>(gS) + i);
    __syncthreads();
  }

  int* listI = da.listI + (size_t)slot * da.maxN;
  int* listJ = da.listJ + (size_t)slot * da.maxN;
  unsigned* keyI = da.keyI + (size_t)slot * da.maxN;
  unsigned* keyJ = da.keyJ + (size_t)slot * da.maxN;
  const float res = da.resolution;

  // room-dependent pointers (re-derived whenever the slot moves to another room)
  long long base = 0;
  int N = 0;
  const float* pts = nullptr;
  unsigned* pw = nullptr;
  int vmin0 = 0, vmin1 = 0, vmin2 = 0;            // voxel coordinates in the state words are relative to these
  auto bind_room = [&]() {
    base = da.room_off[S.room];
    N = (int)(da.room_off[S.room + 1] - base);
    pts = da.pts + base * 16;
    pw = da.pw + (long long)lane_id * da.pw_lane_stride + da.pw_off[S.room];
    const int4 vm = da.room_vmin[S.room];
    vmin0 = vm.x; vmin1 = vm.y; vmin2 = vm.z;
  };
  if (S.room >= 0) bind_room();
  // CURRENT flag updates of this lane's words.  Speculative lanes: another lane's commit may set VISITED in this copy at any
  // time (atomicOr), so the flag goes in and out with atomics too (a plain read-modify-write could lose the commit's bit).
  auto ldw = [&](int i) -> unsigned { return spec ? __ldcg(pw + i) : pw[i]; };   // (atomics act in L2: read them there)
  auto set_cur = [&](int i, unsigned w) { if (spec) atomicOr(pw + i, PW_CUR); else pw[i] = w | PW_CUR; };
  auto clr_cur = [&](int i, unsigned w) { if (spec) atomicAnd(pw + i, ~PW_CUR); else pw[i] = w & ~PW_CUR; };

  // next unvisited seed in curvature order from position `cursor` (:183-188); -1 when the room is exhausted
  auto find_seed = [&](int cursor) -> int {
    // (four NT-wide chunks per round with all eight loads in flight were measured slower: the seed is usually in the first chunk)
    const int* order = da.order + base;
    for (int start = cursor; start < N; start += NT) {
      const int pos = start + tid;
      const bool ok = pos < N && !(ldw(order[pos]) & PW_VIS);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) sh.red[warp] = bal ? (start + warp * 32 + __ffs(bal) - 1) : INT_MAX;
      __syncthreads();
      int best = INT_MAX;
      for (int w = 0; w < NT / 32; ++w) best = min(best, sh.red[w]);
      __syncthreads();
      if (best != INT_MAX) return best;
    }
    return -1;
  };

  // ---- random restarts (test_random_restart.py:170-197) ---------------------------------------------------------------
  // A restart that stops does not touch visited / labels: it records its score ('np': points in the region, :174), clears
  // its CURRENT flags and parks.  The lane that finishes a seed LAST commits it -- the first lane with the highest score
  // wins (numpy.argmax, :177), its region becomes visited in every lane's copy of the words and is labelled (:178-181) --
  // and then owns the group: it finds the next seed (or room), starts every lane on it and asks for their STEP items.
  auto stop_lane = [&](int reason, int n_cur) -> bool {
    for (int j = tid; j < n_cur; j += NT) {
      const int i = listI[j];
      pw[i] &= PW_XYZ;                                       // (a CURRENT point is never VISITED)
    }
    if (tid == 0) {
      const int slot_r = reason == STOP_NONEIGHBOR ? 0 : reason == STOP_NOEXPAND ? 1 : reason == STOP_STUCK ? 2 : 3;
      S.stops[slot_r] += 1;
      S.active = 0;
      S.parked = 1;
      S.n_in = n_cur;
    }
    __syncthreads();
    for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
      reinterpret_cast<int*>(gS)[i] = reinterpret_cast<const int*>(&sh.S)[i];
    __syncthreads();
    if (tid == 0) {
      *reinterpret_cast<volatile int*>(&grp->score[lane_id]) = n_cur;
      __threadfence();
      const int done = atomicAdd(&grp->done, 1) + 1;
      sh.flag = done;
      if (done == L) {
        __threadfence();
        for (int i = 0; i < (int)(sizeof(LaneGroup) / 4); ++i)
          reinterpret_cast<int*>(&sh.G)[i] = __ldcg(reinterpret_cast<const int*>(grp) + i);
        int best = 0;
        for (int l = 1; l < L; ++l)
          if (sh.G.score[l] > sh.G.score[best]) best = l;
        sh.red[0] = best;
        sh.G.done = 0;
      }
    }
    __syncthreads();
    const bool last = sh.flag == L;
    const int best = sh.red[0];
    __syncthreads();
    if (!last) return false;
    // commit the best restart of this seed
    const int n_best = sh.G.score[best];
    const bool labelled = n_best > da.cluster_threshold;
    const int cid = sh.G.cluster_id;
    const int* blist = da.listI + (size_t)(slot - lane_id + best) * da.maxN;
    int* label = da.label + base;
    unsigned* pw0 = da.pw + da.pw_off[S.room];
    for (int j = tid; j < n_best; j += NT) {
      const int i = best == lane_id ? blist[j] : __ldcg(blist + j);
      const unsigned w = (pw[i] & PW_XYZ) | PW_VIS;
      for (int l = 0; l < L; ++l) pw0[(long long)l * da.pw_lane_stride + i] = w;
      if (labelled) label[i] = cid;
    }
    __syncthreads();
    if (tid == 0) {
      if (labelled) sh.G.cluster_id += 1;
      sh.G.regions += 1;
      sh.G.visited += n_best;
      S.parked = 0;
    }
    __syncthreads();
    return true;
  };
  // Owner of the group: the room is exhausted (or this is the first call) -- publish its stats and fetch the next room from
  // the queue; with none left retire every lane.  Returns false when retired.
  auto group_next_room = [&]() -> bool {
    LaneGroup& G = sh.G;
    {
      // room exhausted (or first call): publish its stats and fetch the next room from the queue
      if (tid == 0) {
        if (G.room >= 0) {
          int steps = 0, stops[4] = {0, 0, 0, 0};
          for (int l = 0; l < L; ++l) {
            const SlotState* o = da.slots + (slot - lane_id + l);
            const int ts = l == lane_id ? S.total_steps : __ldcg(&o->total_steps);
            steps += ts;
            if (da.lane_steps != nullptr) da.lane_steps[(size_t)G.room * L + l] = ts;
            for (int k = 0; k < 4; ++k) stops[k] += l == lane_id ? S.stops[k] : __ldcg(&o->stops[k]);
          }
          LrgRoomStats& st = da.stats[G.room];
          st.n_points = N; st.grow_steps = steps; st.regions = G.regions; st.clusters = G.cluster_id - 1;
          st.stop_noneighbor = stops[0]; st.stop_noexpand = stops[1]; st.stop_stuck = stops[2]; st.stop_other = stops[3];
        }
        const int nr = atomicAdd(da.next_room, 1);
        G.room = nr < da.n_rooms ? (da.room_order != nullptr ? da.room_order[nr] : nr) : -1;
        G.cursor = 0; G.cluster_id = 1; G.regions = 0; G.visited = 0;
        for (int l = 0; l < L; ++l) {                      // per-room counters of every lane (the others are parked)
          SlotState* o = l == lane_id ? &S : da.slots + (slot - lane_id + l);
          o->total_steps = 0; o->stops[0] = o->stops[1] = o->stops[2] = o->stops[3] = 0;
        }
        S.room = G.room; S.active = 0;
      }
      __syncthreads();
      if (G.room < 0) {
        if (tid == 0) {
          for (int l = 0; l < L; ++l)
            if (l != lane_id) *reinterpret_cast<volatile int*>(&da.slots[slot - lane_id + l].finished) = 1;
          S.finished = 1;
          *grp = G;
          const int fin = atomicAdd(da.finished_slots, L) + L;
          if (fin == da.n_slots) {
            sh.all_done = 1;
            if (da.done_flag != nullptr) { *da.done_flag = 1; __threadfence_system(); }
          }
        }
        __syncthreads();
        return false;
      }
      bind_room();
      return true;
    }
  };
  // Owner of the group: next seed of the room, else the next room, else retire every lane.  Returns false when retired.
  auto advance_group = [&]() -> bool {
    LaneGroup& G = sh.G;
    while (true) {
      int found = -1;
      if (G.room >= 0) found = find_seed(G.cursor);
      if (found >= 0) {
        const int seed = da.order[base + found];
        const unsigned w = pw[seed];
        if (tid == 0) {
          G.cursor = found + 1;
          const int vx = pw_x(w), vy = pw_y(w), vz = pw_z(w);
          // order: the commit's word / label stores (all threads, before the barrier) and the seed's flags and list heads
          // become visible before any lane can observe its `begin` flag
          __threadfence();
          for (int l = 0; l < L; ++l) {
            (da.pw + (long long)l * da.pw_lane_stride + da.pw_off[G.room])[seed] = w | PW_CUR;
            (da.listI + (size_t)(slot - lane_id + l) * da.maxN)[0] = seed;
          }
          __threadfence();
          for (int l = 0; l < L; ++l) {
            SlotState* o = l == lane_id ? &S : da.slots + (slot - lane_id + l);
            o->active = 0; o->finished = 0; o->room = G.room; o->seed = seed;
            o->minD[0] = o->maxD[0] = o->seqMin[0] = o->seqMax[0] = vx;
            o->minD[1] = o->maxD[1] = o->seqMin[1] = o->seqMax[1] = vy;
            o->minD[2] = o->maxD[2] = o->seqMin[2] = o->seqMax[2] = vz;
            o->stuck = 0; o->steps = 0; o->n_in = 1; o->n_nb = 0;
            o->parked = 0;
            *reinterpret_cast<volatile int*>(&o->begin) = l == lane_id ? 0 : 1;
          }
          sh.listI_s[0] = seed;
          sh.wake = ((1u << L) - 1u) & ~(1u << lane_id);
          *grp = G;
        }
        __syncthreads();
        return true;
      }
      if (!group_next_room()) return false;
    }
  };

  // ---- beam search (test_beam_search.py:164-279) -----------------------------------------------------------------------
  // A lane that was handed candidate q of the queue: its region is the candidate's index list (this lane's copy of the words
  // carries no CURRENT flag between expansions), bounding box and size came with the slot state.
  auto beam_begin = [&]() {
    const int n = S.n_in;
    const int* par = da.parI + ((size_t)(slot / L) * BW + lane_id / SW) * da.maxN;
    for (int j = tid; j < n; j += NT) {
      const int i = __ldcg(par + j);
      pw[i] |= PW_CUR;
      listI[j] = i;
      if (j < kListCap) sh.listI_s[j] = i;
    }
    if (tid == 0) { S.begin = 0; S.parked = 0; }
    __syncthreads();
  };
  // The lane's expansion is over (listI[0..n_cur) = the expanded mask): report its score -- 'np': the size of the mask if the
  // expansion added a point (:262-267), else no candidate -- clear the CURRENT flags and park.  Returns true to the lane
  // that reports LAST in the round, which then owns the group (sh.G).
  auto beam_park = [&](bool candidate, int n_cur) -> bool {
    for (int j = tid; j < n_cur; j += NT) pw[listI[j]] &= ~PW_CUR;
    if (tid == 0) { S.active = 0; S.parked = 1; S.begin = 0; S.n_in = n_cur; }
    __syncthreads();
    for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
      reinterpret_cast<int*>(gS)[i] = reinterpret_cast<const int*>(&sh.S)[i];
    __syncthreads();
    if (tid == 0) {
      *reinterpret_cast<volatile int*>(&grp->score[lane_id]) = candidate ? n_cur : -1;
      if (da.score_ml) *reinterpret_cast<volatile float*>(&grp->fscore[lane_id]) = sh.lane_score;
      __threadfence();
      const int expect = __ldcg(&grp->expect);
      const int done = atomicAdd(&grp->done, 1) + 1;
      sh.flag = done == expect ? 1 : 0;
      if (done == expect) {
        __threadfence();
        for (int i = 0; i < (int)(sizeof(LaneGroup) / 4); ++i)
          reinterpret_cast<int*>(&sh.G)[i] = __ldcg(reinterpret_cast<const int*>(grp) + i);
      }
    }
    __syncthreads();
    const bool last = sh.flag != 0;
    __syncthreads();
    return last;
  };
  // Owner of the group (sh.G): close the round that just ended (if any), commit the seed's region when the search is over,
  // find the next seed / room, and start the next round.  Returns 0 when the group has retired, 1 when this lane expands a
  // candidate of the new round (its region is set up), 2 when it is not part of the new round (parked, state written back).
  auto beam_commit = [&]() -> int {
    LaneGroup& G = sh.G;
    const int slot0 = slot - lane_id;
    int* const par0 = da.parI + (size_t)(slot / L) * BW * da.maxN;
    if (G.nQ > 0) {
      // next Q = the BEAM_WIDTH best of newQ, stable (:273); newQ is in (candidate, search) order = lane order
      if (tid == 0) {
        unsigned used = 0;
        int nc = 0;
        for (int q = 0; q < BW; ++q) {
          int best = -1;
          for (int l = 0; l < G.expect; ++l)
            if (!((used >> l) & 1u) && G.score[l] >= 0 &&
                (best < 0 || (da.score_ml ? G.fscore[l] > G.fscore[best] : G.score[l] > G.score[best]))) best = l;
          if (best < 0) break;
          used |= 1u << best;
          sh.red[1 + nc++] = best;
        }
        sh.red[0] = nc;
      }
      __syncthreads();
      const int nc = sh.red[0];
      int fin = 0;                      // 0: the search goes on; else 1 + index into stops[] of why the seed's search ended
      if (nc == 0) {
        fin = 1 + 1;                    // no expansion added a point: Q is empty (:169), bestMask = the last Q[0] (:179)
      } else {
        for (int q = 0; q < nc; ++q) {
          const int l = sh.red[1 + q], n = G.score[l];
          const int* sl = da.listI + (size_t)(slot0 + l) * da.maxN;
          int* dst = par0 + (size_t)q * da.maxN;
          for (int j = tid; j < n; j += NT) __stcg(dst + j, l == lane_id ? sl[j] : __ldcg(sl + j));
        }
        if (tid == 0) {
          float ps[kMaxLanes];
          for (int q = 0; q < nc; ++q) ps[q] = G.fscore[sh.red[1 + q]];
          for (int q = 0; q < nc; ++q) G.par_score[q] = ps[q];
          for (int q = 0; q < nc; ++q) {
            const int l = sh.red[1 + q];
            const SlotState* o = da.slots + slot0 + l;
            G.par_n[q] = G.score[l];
            for (int a = 0; a < 3; ++a) {
              G.par_min[q][a] = l == lane_id ? S.minD[a] : __ldcg(&o->minD[a]);
              G.par_max[q][a] = l == lane_id ? S.maxD[a] : __ldcg(&o->maxD[a]);
            }
          }
          // head of the next round (qid == 0, :178-190): bestMask = Q[0]; stuck logic on its bounding box
          bool expanded = false;
          for (int a = 0; a < 3; ++a) expanded |= (G.par_min[0][a] < G.seqMin[a]) || (G.par_max[0][a] > G.seqMax[a]);
          int f = 0;
          if (!expanded) {
            if (G.stuck >= 1) f = 1 + 2; else G.stuck += 1;
          } else {
            G.stuck = 0;
          }
          for (int a = 0; a < 3; ++a) { G.seqMin[a] = min(G.seqMin[a], G.par_min[0][a]); G.seqMax[a] = max(G.seqMax[a], G.par_max[0][a]); }
          G.nQ = nc;
          G.round += 1;
          if (f == 0 && da.max_steps > 0 && G.round >= da.max_steps) f = 1 + 3;
          sh.flag = f;
        }
        __syncthreads();
        fin = sh.flag;
        __syncthreads();
      }
      if (fin != 0) {
        // visited[bestMask] = True; label it when it is larger than the threshold (:276-279)
        const int n_best = G.par_n[0];
        const bool labelled = n_best > da.cluster_threshold;
        const int cid = G.cluster_id;
        int* label = da.label + base;
        unsigned* pw0 = da.pw + da.pw_off[G.room];
        for (int j = tid; j < n_best; j += NT) {
          const int i = __ldcg(par0 + j);
          const unsigned w = (pw[i] & PW_XYZ) | PW_VIS;
          for (int l = 0; l < L; ++l) pw0[(long long)l * da.pw_lane_stride + i] = w;
          if (labelled) label[i] = cid;
        }
        __syncthreads();
        if (tid == 0) {
          if (labelled) G.cluster_id += 1;
          G.regions += 1;
          G.visited += n_best;
          G.nQ = 0;
          S.stops[fin - 1] += 1;
        }
        __syncthreads();
      }
    }
    // next seed of the room (:143-145), else the next room: Q = [(0, seed alone)] (:154-164)
    while (G.nQ == 0) {
      int found = -1;
      if (G.room >= 0) found = find_seed(G.cursor);
      if (found >= 0) {
        const int seed = da.order[base + found];
        const unsigned w = pw[seed];
        if (tid == 0) {
          G.cursor = found + 1;
          G.seed = seed;
          __stcg(par0, seed);
          G.par_n[0] = 1;
          G.par_score[0] = 0.f;
          const int v[3] = {pw_x(w), pw_y(w), pw_z(w)};
          for (int a = 0; a < 3; ++a) G.par_min[0][a] = G.par_max[0][a] = G.seqMin[a] = G.seqMax[a] = v[a];
          G.stuck = 1;                  // the head of the first round finds the seed's box inside itself (:180-184)
          G.round = 0;
          G.nQ = 1;
        }
        __syncthreads();
        break;
      }
      if (!group_next_room()) return 0;
    }
    // start the round: lane q * SW + s takes candidate q
    const int expect = G.nQ * SW;
    if (tid == 0) {
      G.done = 0;
      G.expect = expect;
      *grp = G;
      for (int l = 0; l < expect; ++l) {
        const int q = l / SW;
        SlotState* o = l == lane_id ? &S : da.slots + slot0 + l;
        o->active = 0; o->finished = 0; o->room = G.room; o->seed = G.seed;
        for (int a = 0; a < 3; ++a) { o->minD[a] = G.par_min[q][a]; o->maxD[a] = G.par_max[q][a]; }
        o->steps = G.round; o->n_in = G.par_n[q]; o->n_nb = 0; o->stuck = 0;
      }
      if (lane_id >= expect) { S.active = 0; S.parked = 1; S.begin = 0; }
    }
    __syncthreads();
    if (lane_id >= expect) {
      // not part of the round: this lane's record must be at rest before any other lane can close the round and write to it
      for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
        reinterpret_cast<int*>(gS)[i] = reinterpret_cast<const int*>(&sh.S)[i];
      __syncthreads();
    }
    if (tid == 0) {
      // order: the commit's word / label / list stores (all threads, before the barriers), the group record and the lanes'
      // states become visible before any lane can observe its `begin` flag
      __threadfence();
      unsigned wake = 0;
      for (int l = 0; l < expect; ++l)
        if (l != lane_id) {
          *reinterpret_cast<volatile int*>(&da.slots[slot0 + l].begin) = 2;
          wake |= 1u << l;
        }
      sh.wake = wake;
    }
    __syncthreads();
    if (lane_id >= expect) return 2;
    beam_begin();
    return 1;
  };

  // ---- speculative lanes (SpecSync, lrg_driver.cuh) ---------------------------------------------------------------------
  // The region of this lane has stopped (listI[0..n_cur) = the region): put the lane's state at rest, announce it, and find
  // out whether its ticket is the head of the commit order.  Returns true to the lane that may commit now (it owns the
  // group's critical section: sh.G is loaded), false when the lane waits -- whoever commits the ticket before it wakes it.
  SpecSync* const ssync = spec ? da.spec_sync + slot / L : nullptr;
  auto spec_load_group = [&]() {
    if (tid == 0) {
      __threadfence();                                       // acquire: everything the previous owners wrote
      for (int i = 0; i < (int)(sizeof(LaneGroup) / 4); ++i)
        reinterpret_cast<int*>(&sh.G)[i] = __ldcg(reinterpret_cast<const int*>(grp) + i);
    }
    __syncthreads();
  };
  auto spec_finish = [&](int reason, int n_cur) -> bool {
    if (tid == 0) { S.active = 0; S.parked = 1; S.begin = 0; S.n_in = n_cur; S.fin_reason = reason; }
    __syncthreads();
    for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
      reinterpret_cast<int*>(gS)[i] = reinterpret_cast<const int*>(&sh.S)[i];
    __syncthreads();
    if (tid == 0) {
      // Dekker with the committer of the ticket before mine: I raise `fin` then read commit_seq, it raises commit_seq then
      // reads (claims) `fin`; the fences in between are sequentially consistent, so at least one side sees the other
      __threadfence();
      atomicExch(&ssync->fin[lane_id], 1);
      __threadfence();
      const int cs = *reinterpret_cast<volatile int*>(&ssync->commit_seq);
      sh.flag = (cs == S.ticket && atomicCAS(&ssync->fin[lane_id], 1, 0) == 1) ? 1 : 0;
    }
    __syncthreads();
    const bool head = sh.flag != 0;
    __syncthreads();
    if (!head) return false;
    if (tid == 0) S.parked = 0;
    spec_load_group();
    return true;
  };
  // Owner: hand the next unvisited seed of the room (curvature order, :183-188) to lane l with the next ticket; false when the
  // room has no seed left.  The lane is idle (or is this one), so its state is at rest.
  auto spec_give_seed = [&](int l) -> bool {
    LaneGroup& G = sh.G;
    const int found = find_seed(G.cursor);
    if (found < 0) return false;
    const int seed = da.order[base + found];
    unsigned* pwl = da.pw + (long long)l * da.pw_lane_stride + da.pw_off[G.room];
    if (tid == 0) {
      const unsigned w = __ldcg(pwl + seed);
      const int vx = pw_x(w), vy = pw_y(w), vz = pw_z(w);
      G.cursor = found + 1;
      const int ticket = G.next_ticket++;
      G.lane_ticket[l] = ticket;
      atomicOr(pwl + seed, PW_CUR);
      (da.listI + (size_t)(slot - lane_id + l) * da.maxN)[0] = seed;
      SlotState* o = l == lane_id ? &S : da.slots + (slot - lane_id + l);
      o->active = 0; o->finished = 0; o->room = G.room; o->seed = seed;
      o->minD[0] = o->maxD[0] = o->seqMin[0] = o->seqMax[0] = vx;
      o->minD[1] = o->maxD[1] = o->seqMin[1] = o->seqMax[1] = vy;
      o->minD[2] = o->maxD[2] = o->seqMin[2] = o->seqMax[2] = vz;
      o->stuck = 0; o->steps = 0; o->n_in = 1; o->n_nb = 0;
      o->ticket = ticket; o->log_begin = G.log_n; o->fin_reason = 0;
      if (l == lane_id) {
        o->parked = 0; o->begin = 0;
        sh.listI_s[0] = seed;
      } else {
        // (the flag is raised after the fence in spec_release; here only the record)
        o->parked = 1;
        sh.wake |= 1u << l;
      }
    }
    __syncthreads();
    return true;
  };
  // Owner (sh.G loaded): validate and commit (or discard) this lane's stopped region if it has one, hand out seeds, pass the
  // head on.  Returns 0: the group has retired (no rooms left), 1: this lane grows a region (fresh seed, or the same seed
  // again), 2: this lane is idle (parked; its state is written back).
  auto spec_head = [&]() -> int {
    LaneGroup& G = sh.G;
    const int slot0 = slot - lane_id;
    bool bump = false;
    if (S.fin_reason != 0) {
      const int n_cur = S.n_in, reason = S.fin_reason;
      // what was committed since this region began: none of it may lie inside the boxes the region looked at
      // (seqMin - 1 .. seqMax + 1 covers every neighbour shell of its life, :222-229), and its seed must still be unvisited
      const int* clog = da.clog + (size_t)(slot / L) * da.maxN;
      const int lo0 = S.seqMin[0] - 1, lo1 = S.seqMin[1] - 1, lo2 = S.seqMin[2] - 1;
      const int hi0 = S.seqMax[0] + 1, hi1 = S.seqMax[1] + 1, hi2 = S.seqMax[2] + 1;
      int bad = 0;
      for (int j = S.log_begin + tid; j < G.log_n; j += NT) {
        const int i = __ldcg(clog + j);
        const unsigned w = __ldcg(pw + i);
        const int x = pw_x(w), y = pw_y(w), z = pw_z(w);
        if (i == S.seed) bad |= 2;
        if (x >= lo0 && x <= hi0 && y >= lo1 && y <= hi1 && z >= lo2 && z <= hi2) bad |= 1;
      }
      // (__syncthreads_or is a logical OR: one call per bit)
      bad = (__syncthreads_or(bad & 1) ? 1 : 0) | (__syncthreads_or(bad & 2) ? 2 : 0);
      if (bad != 0) {
        // grown on a stale visited set: forget the attempt
        for (int j = tid; j < n_cur; j += NT) atomicAnd(pw + listI[j], ~PW_CUR);
        __syncthreads();
        if (tid == 0) G.wasted_steps += S.steps;
        if (!(bad & 2)) {
          // ... and grow the seed again: every earlier region is committed now, so this attempt sees exactly the visited set
          // of the sequential driver (and nobody else can commit before it does)
          if (tid == 0) {
            const unsigned w = __ldcg(pw + S.seed);
            G.restarts += 1;
            atomicOr(pw + S.seed, PW_CUR);
            listI[0] = S.seed; sh.listI_s[0] = S.seed;
            S.minD[0] = S.maxD[0] = S.seqMin[0] = S.seqMax[0] = pw_x(w);
            S.minD[1] = S.maxD[1] = S.seqMin[1] = S.seqMax[1] = pw_y(w);
            S.minD[2] = S.maxD[2] = S.seqMin[2] = S.seqMax[2] = pw_z(w);
            S.stuck = 0; S.steps = 0; S.n_in = 1; S.n_nb = 0; S.log_begin = G.log_n; S.fin_reason = 0; S.parked = 0;
            grp->wasted_steps = G.wasted_steps; grp->restarts = G.restarts;      // (still the owner: nothing else changed)
          }
          __syncthreads();
          return 1;
        }
        if (tid == 0) G.dropped += 1;                          // the seed was swallowed by an earlier region (:187-188)
      } else {
        // commit (stop_growing, :210-217): visited in every lane's copy of the words, label, log
        const bool labelled = n_cur > da.cluster_threshold;
        const int cid = G.cluster_id;
        int* label = da.label + base;
        unsigned* pw0 = da.pw + da.pw_off[G.room];
        int* wlog = da.clog + (size_t)(slot / L) * da.maxN + G.log_n;
        for (int j = tid; j < n_cur; j += NT) {
          const int i = listI[j];
          for (int l = 0; l < L; ++l) atomicOr(pw0 + (long long)l * da.pw_lane_stride + i, PW_VIS);
          atomicAnd(pw + i, ~PW_CUR);
          if (labelled) label[i] = cid;
          __stcg(wlog + j, i);
        }
        __syncthreads();
        if (tid == 0) {
          if (labelled) G.cluster_id += 1;
          G.regions += 1;
          G.visited += n_cur;
          G.log_n += n_cur;
          G.useful_steps += S.steps;
          G.stops[reason == STOP_NONEIGHBOR ? 0 : reason == STOP_NOEXPAND ? 1 : reason == STOP_STUCK ? 2 : 3] += 1;
        }
      }
      if (tid == 0) { S.fin_reason = 0; G.lane_ticket[lane_id] = -1; G.commit_seq += 1; }
      __syncthreads();
      bump = true;
    }
    // hand out seeds: this lane first, then the idle lanes while the window has room; when nothing is in flight and the
    // room has no seed left, publish its statistics and move the group to the next room
    bool mine = false;
    while (true) {
      if (G.room >= 0) {
        // Window: speculation costs SM time (discarded attempts) and pays only on the run's critical path, so a seed goes to
        // a SECOND lane only (a) in the spec_top rooms with the most estimated work left -- the run ends with them -- or
        // (b) while many CTAs wait for work anyway (the tail).  Estimate = unvisited points x grow steps per visited point
        // of this room so far (prior: 0.2 steps per point).
        bool speculate = true;
        if (da.spec_est != nullptr) {
          const int gi = slot / L, ng = da.n_slots / L;
          const int mine = (int)(((long long)(N - G.visited) * (G.useful_steps + 20)) / (G.visited + 100));
          if (tid == 0) *reinterpret_cast<volatile int*>(da.spec_est + gi) = mine;
          int ahead = 0;
          long long part = 0;
          if (tid == 0) sh.est_sum = 0ull;
          for (int g0 = 0; g0 < ng; g0 += NT) {
            const int g = g0 + tid;
            const int other = (g < ng && g != gi) ? __ldcg(da.spec_est + g) : -1;
            if (other > 0) part += other;
            ahead += __syncthreads_count(other > mine || (other == mine && g < gi));
          }
          // the estimate of everything that is left: the other rooms in flight, this one, and the rooms nobody has started
          // (0.2 grow steps per point, the prior of the estimate above)
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
          if (lane == 0 && part != 0) atomicAdd(&sh.est_sum, (unsigned long long)part);
          __syncthreads();
          long long total_est = (long long)sh.est_sum + mine;
          {
            const int nr = min(*reinterpret_cast<volatile int*>(da.next_room), da.n_rooms);
            total_est += (da.pending_pts != nullptr ? da.pending_pts[nr] : da.total_pts - da.room_off[nr]) / 5;
          }
          const bool critical = da.spec_crit <= 0 || (long long)mine * da.spec_crit >= total_est;
          if (tid == 0) *reinterpret_cast<volatile int*>(da.spec_est + ng + gi) = (critical && ahead < da.spec_top) ? 1 : 0;
          // (the queue counters move while they are read: ONE thread looks, the verdict must be the same for the whole CTA)
          int idle_ok = 0;
          if (tid == 0 && da.q_ctr != nullptr) {
            const int b = (int)(*reinterpret_cast<const volatile unsigned*>(da.q_ctr + 1) - *reinterpret_cast<const volatile unsigned*>(da.q_ctr));
            idle_ok = (b < 0 ? -b : 0) >= da.spec_min_idle;
          }
          idle_ok = __syncthreads_or(idle_ok);
          speculate = (ahead < da.spec_top && critical) || idle_ok != 0;
        }
        for (int k = 0; k < L; ++k) {
          const int l = (lane_id + k) % L;
          if (G.lane_ticket[l] >= 0) continue;
          if (k > 0 && G.next_ticket > G.commit_seq && !speculate) continue;
          if (!spec_give_seed(l)) break;
          if (l == lane_id) mine = true;
        }
      }
      const bool in_flight = G.next_ticket > G.commit_seq;
      __syncthreads();                                       // (every thread has compared the two counters before thread 0 resets them below,
                                                             //  one after the other -- compute-sanitizer racecheck, profiles/r2aa_*)
      if (in_flight) break;                                  // something is in flight
      // the room is finished (or this is the first call of the run)
      if (tid == 0) {
        if (G.room >= 0) {
          LrgRoomStats& st = da.stats[G.room];
          st.n_points = N; st.grow_steps = G.useful_steps; st.regions = G.regions; st.clusters = G.cluster_id - 1;
          st.stop_noneighbor = G.stops[0]; st.stop_noexpand = G.stops[1]; st.stop_stuck = G.stops[2]; st.stop_other = G.stops[3];
          st.spec_wasted_steps = G.wasted_steps; st.spec_restarts = G.restarts; st.spec_dropped = G.dropped;
          if (da.lane_steps != nullptr)
            for (int l = 0; l < L; ++l) {
              SlotState* o = l == lane_id ? &S : da.slots + slot0 + l;
              da.lane_steps[(size_t)G.room * L + l] = l == lane_id ? S.total_steps : __ldcg(&o->total_steps);
            }
        }
        const int nr = atomicAdd(da.next_room, 1);
        G.room = nr < da.n_rooms ? (da.room_order != nullptr ? da.room_order[nr] : nr) : -1;
        G.cursor = 0; G.cluster_id = 1; G.regions = 0; G.visited = 0;
        G.commit_seq = 0; G.next_ticket = 0; G.log_n = 0;
        G.useful_steps = G.wasted_steps = G.restarts = G.dropped = 0;
        G.stops[0] = G.stops[1] = G.stops[2] = G.stops[3] = 0;
        for (int l = 0; l < L; ++l) {
          G.lane_ticket[l] = -1;
          SlotState* o = l == lane_id ? &S : da.slots + slot0 + l;
          o->total_steps = 0; o->room = G.room;
        }
        *reinterpret_cast<volatile int*>(&ssync->commit_seq) = 0;
        S.active = 0;
        bump = false;
      }
      __syncthreads();
      if (G.room < 0) {
        if (tid == 0) {
          if (da.spec_est != nullptr) *reinterpret_cast<volatile int*>(da.spec_est + slot / L) = 0;
          for (int l = 0; l < L; ++l)
            if (l != lane_id) *reinterpret_cast<volatile int*>(&da.slots[slot0 + l].finished) = 1;
          S.finished = 1;
          *grp = G;
          const int fin = atomicAdd(da.finished_slots, L) + L;
          if (fin == da.n_slots) {
            sh.all_done = 1;
            if (da.done_flag != nullptr) { *da.done_flag = 1; __threadfence_system(); }
          }
        }
        __syncthreads();
        return 0;
      }
      bind_room();
    }
    // release the critical section: group record, lane records and word flags first, then the ticket that may commit
    // next, then the lanes that were handed a seed / the lane that holds the new head ticket and already waits
    if (tid == 0) {
      *grp = G;
      __threadfence();
      const unsigned fresh = sh.wake;                           // lanes that were handed a seed in this call
      if (bump) {
        *reinterpret_cast<volatile int*>(&ssync->commit_seq) = G.commit_seq;
        __threadfence();
        for (int l = 0; l < L; ++l)
          if (l != lane_id && G.lane_ticket[l] == G.commit_seq && !((fresh >> l) & 1u) && atomicCAS(&ssync->fin[l], 1, 0) == 1) {
            *reinterpret_cast<volatile int*>(&da.slots[slot0 + l].begin) = 2;
            sh.wake |= 1u << l;
          }
      }
      for (int l = 0; l < L; ++l)
        if ((fresh >> l) & 1u) *reinterpret_cast<volatile int*>(&da.slots[slot0 + l].begin) = 1;
      if (!mine) { S.active = 0; S.parked = 1; S.begin = 0; }
    }
    __syncthreads();
    if (!mine) {
      for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
        reinterpret_cast<int*>(gS)[i] = reinterpret_cast<const int*>(&sh.S)[i];
      __syncthreads();
      return 2;
    }
    return 1;
  };

  // stop_growing (:210-217): visited |= current; label when the region is larger than the threshold.
  // listI[0..n_cur) holds the current region.
  auto stop_region = [&](int reason, int n_cur) -> bool {
    if (spec) return spec_finish(reason, n_cur);
    if (L > 1) return stop_lane(reason, n_cur);
    const bool labelled = n_cur > da.cluster_threshold;
    int* label = da.label + base;
    const int cid = S.cluster_id;
    for (int j = tid; j < n_cur; j += NT) {
      const int i = listI[j];
      pw[i] = (pw[i] & PW_XYZ) | PW_VIS;
      if (labelled) label[i] = cid;
    }
    __syncthreads();
    if (tid == 0) {
      if (labelled) S.cluster_id += 1;
      S.regions += 1;
      S.visited += n_cur;
      int slot_r = reason == STOP_NONEIGHBOR ? 0 : reason == STOP_NOEXPAND ? 1 : reason == STOP_STUCK ? 2 : 3;
      S.stops[slot_r] += 1;
      S.active = 0;
    }
    __syncthreads();
    return true;
  };

  int mode = MODE_NEW_REGION;
  if (beam && !S.active) {
    if (S.begin) {                      // a candidate of the new round handed over by the lane that closed the previous one
      beam_begin();
      mode = MODE_SCAN;
    } else {                            // first call of the run (lane 0): take the group (no seed in flight: nQ == 0)
      if (tid == 0)
        for (int i = 0; i < (int)(sizeof(LaneGroup) / 4); ++i)
          reinterpret_cast<int*>(&sh.G)[i] = __ldcg(reinterpret_cast<const int*>(grp) + i);
      __syncthreads();
    }
  } else if (spec && !S.active) {
    const int begin_flag = S.begin;
    __syncthreads();                    // (every thread has read the flag before thread 0 clears it)
    if (begin_flag == 1) {              // a fresh seed handed over by the lane that held the head ticket (state re-read above)
      if (tid == 0) { S.begin = 0; S.parked = 0; sh.listI_s[0] = S.seed; }
      __syncthreads();
      mode = MODE_SCAN;
    } else {                            // woken as the new head with a stopped region (begin == 2), or the first call of the run
      if (tid == 0) { S.begin = 0; S.parked = 0; }
      spec_load_group();
    }
  } else if (L > 1 && !S.active) {
    if (S.begin) {                      // a fresh seed handed over by the lane that committed the previous one
      __threadfence();                  // (acquire: the committing CTA's stores to this lane's words and list)
      __syncthreads();
      if (tid == 0) { S.begin = 0; sh.listI_s[0] = S.seed; }      // (the region is the seed alone: list mirror of listI[0])
      mode = MODE_SCAN;
    } else {                            // first call of the run (lane 0): take the group
      if (tid == 0)
        for (int i = 0; i < (int)(sizeof(LaneGroup) / 4); ++i)
          reinterpret_cast<int*>(&sh.G)[i] = __ldcg(reinterpret_cast<const int*>(grp) + i);
    }
    __syncthreads();
  }
  stamp(0);

  // ------------------------------------------------------------------ apply the pending step (:262-306)
  if (S.active) {
    const int room_rng = da.room_id_base + S.room;
    const unsigned step_rng = (unsigned)S.steps;
    const unsigned rng_lane = ((unsigned)S.seed << 8) | lane_stream;
    LrgStepTrace* tr = nullptr;
    if (da.trace != nullptr && S.total_steps < da.trace_capacity)
      tr = da.trace + ((size_t)S.room * L + lane_id) * da.trace_capacity + S.total_steps;
    float2 lg_[VT], xy_[VT];
    unsigned w_[VT];
#pragma unroll
    for (int k = 0; k < VT; ++k) {
      const int vt = tid + k * NT;
      const bool is_add = vt >= kMaxTilePts;
      const int nrows = is_add ? da.Nj : da.Ni;
      if (p_[k] >= 0) {
        lg_[k] = __ldcg(reinterpret_cast<const float2*>(da.logits[is_add ? 1 : 0] + ((size_t)slot * nrows + src_[k]) * 2));
        xy_[k] = *reinterpret_cast<const float2*>(pts + (size_t)p_[k] * 16);
        w_[k] = ldw(p_[k]);
      }
    }
    // the uniform draws depend on nothing that was loaded: their ~80 integer instructions per row run under the loads' latency
    float u_[VT];
#pragma unroll
    for (int k = 0; k < VT; ++k) {
      const int vt = tid + k * NT;
      const bool is_add = vt >= kMaxTilePts;
      const unsigned draw = philox_draw(da.seed, room_rng, step_rng, rng_lane + (is_add ? kStreamAddUniform : kStreamRemoveUniform),
                                        is_add ? vt - kMaxTilePts : vt);
      u_[k] = (float)(draw >> 8) * (1.0f / 16777216.0f);
    }
    bool m_[VT], normal_[VT];
    int upd = 0;
    if (da.dbg != nullptr && tid == 0 && __float_as_uint(lg_[0].x) + w_[0] == 0x12345u) mark(14);   // (waits for thread 0's loads)
    mark(10);
#pragma unroll
    for (int k = 0; k < VT; ++k) {
      const int vt = tid + k * NT;
      const bool is_add = vt >= kMaxTilePts;
      const int r = is_add ? vt - kMaxTilePts : vt;
      const int p = p_[k];
      bool m = false;
      if (p >= 0) {
        const float conf = confidence(lg_[k].x, lg_[k].y);
        m = u_[k] < conf;                                      // :266-267
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
        const float x = __fadd_rn(__fsub_rn(xy_[k].x, cx), cx);
        const float y = __fadd_rn(__fsub_rn(xy_[k].y, cy), cy);
        const int vx = voxel_of(x, res) - vmin0, vy = voxel_of(y, res) - vmin1;
        normal = (vx == pw_x(w_[k]) && vy == pw_y(w_[k]));
        if (!normal) {
          int o = atomicAdd(&sh.n_odd, 1);
          sh.odd[o] = make_int4(vx, vy, pw_z(w_[k]), is_add ? 1 : 0);
        }
      }
      // adds first, removes second (:283-286)
      if (m && normal && is_add) { set_cur(p, w_[k]); upd = 1; }      // (a neighbour row is never CURRENT before this)
      m_[k] = m; normal_[k] = normal;
    }
    if (beam && da.score_ml) {
      // 'ml' scoring (test_beam_search.py:238-256): every padded tile row asks whether ITS re-rounded voxel is in the set of the
      // re-rounded voxels of the rows that sampled True (rows of one source point agree; a row that re-rounds into another
      // point's voxel is handled by comparing the voxels themselves), takes log(conf) or log(1 - conf), / NUM_NEIGHBOR_POINT
      unsigned* selkey = reinterpret_cast<unsigned*>(sh.planes);          // [2][kMaxTilePts]   (the median planes are not live here)
      float* terms = reinterpret_cast<float*>(selkey + 2 * kMaxTilePts);  // [2][kMaxTilePts]
      unsigned key_[VT];
#pragma unroll
      for (int k = 0; k < VT; ++k) {
        const int vt = tid + k * NT;
        const bool is_add = vt >= kMaxTilePts;
        key_[k] = 0;
        if (p_[k] >= 0) {
          const float cx = S.center[0], cy = S.center[1];
          const float x = __fadd_rn(__fsub_rn(xy_[k].x, cx), cx);
          const float y = __fadd_rn(__fsub_rn(xy_[k].y, cy), cy);
          const int vx = voxel_of(x, res) - vmin0, vy = voxel_of(y, res) - vmin1;
          key_[k] = ((unsigned)(vx + 1) & 0x7ffu) | (((unsigned)(vy + 1) & 0x7ffu) << 11) | ((unsigned)pw_z(w_[k]) << 22);
          if (m_[k]) selkey[(is_add ? 1 : 0) * kMaxTilePts + atomicAdd(&sh.nsel[is_add ? 1 : 0], 1)] = key_[k];
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < VT; ++k) {
        const int vt = tid + k * NT;
        const bool is_add = vt >= kMaxTilePts;
        const int r = is_add ? vt - kMaxTilePts : vt;
        float t = 0.f;
        if (p_[k] >= 0) {
          const unsigned* sk = selkey + (is_add ? 1 : 0) * kMaxTilePts;
          const int ns = sh.nsel[is_add ? 1 : 0];
          bool hit = false;
          for (int j = 0; j < ns; ++j) hit |= sk[j] == key_[k];
          const float conf = confidence(lg_[k].x, lg_[k].y);
          t = __fdiv_rn(logf(hit ? conf : __fsub_rn(1.f, conf)), (float)da.Nj);
        }
        terms[(is_add ? 1 : 0) * kMaxTilePts + r] = t;
      }
      __syncthreads();
      if (tid == 0 || tid == 32) {                             // row-by-row float32 accumulation, one thread per set (:243-245,255-257)
        const int set = tid == 0 ? 1 : 0;
        const int nrows = set ? da.Nj : da.Ni;
        float acc = 0.f;
        for (int r = 0; r < nrows; ++r) acc = __fadd_rn(acc, terms[set * kMaxTilePts + r]);
        sh.lp[set ? 0 : 1] = acc;                               // lp[0] = addLogProb, lp[1] = rmvLogProb
      }
    }
    mark(11);
    __syncthreads();
    mark(12);
    const int n_odd = sh.n_odd;
    if (n_odd > 0) {
      for (int i = tid; i < N; i += NT) {
        const unsigned w = ldw(i);
        bool hit = false;
        for (int o = 0; o < n_odd; ++o) hit |= (sh.odd[o].w == 1 && sh.odd[o].x == pw_x(w) && sh.odd[o].y == pw_y(w) && sh.odd[o].z == pw_z(w));
        if (hit && !(w & PW_CUR)) { set_cur(i, w); upd = 1; }
      }
    }
    const int updated = __syncthreads_or(upd);
#pragma unroll
    for (int k = 0; k < VT; ++k) {
      const bool is_add = (tid + k * NT) >= kMaxTilePts;
      if (m_[k] && normal_[k] && !is_add) clr_cur(p_[k], w_[k]);      // (an inlier row is CURRENT and not VISITED)
    }
    if (n_odd > 0) {
      __syncthreads();
      for (int i = tid; i < N; i += NT) {
        const unsigned w = ldw(i);
        bool hit = false;
        for (int o = 0; o < n_odd; ++o) hit |= (sh.odd[o].w == 0 && sh.odd[o].x == pw_x(w) && sh.odd[o].y == pw_y(w) && sh.odd[o].z == pw_z(w));
        if (hit) clr_cur(i, w);
      }
    }
    mark(13);
    __syncthreads();
    if (tid == 0) { S.steps += 1; S.total_steps += 1; }     // :288
    stamp(1);

    // inlier list + bounding box of the updated region (:292-293); also what stop_growing marks when the region ends
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    auto grow_box = [&](unsigned w) {
      const int x = pw_x(w), y = pw_y(w), z = pw_z(w);
      mn[0] = min(mn[0], x); mn[1] = min(mn[1], y); mn[2] = min(mn[2], z);
      mx[0] = max(mx[0], x); mx[1] = max(mx[1], y); mx[2] = max(mx[2], z);
    };
    int n_in;
    const int n_old = S.n_in;                                 // length of listI = the region before this step's masks
    if (n_odd == 0 && n_old <= kIncCap) {
      // Small region, every selected row kept its voxel: the new list is (old list minus the removed points) merged with the
      // added neighbour rows -- both ascending and disjoint -- instead of a scan over the room's N state words.
      int* keepI = reinterpret_cast<int*>(sh.odd);            // [kIncCap]      (the re-rounding table is empty on this path)
      int* addA = keepI + kIncCap;                            // [kMaxTilePts]
      int* addflag = addA + kMaxTilePts;                      // [kMaxTilePts]
      static_assert(sizeof(sh.odd) >= sizeof(int) * (kIncCap + 2 * kMaxTilePts), "scratch of the list update");
      for (int r = tid; r < kMaxTilePts; r += NT) addflag[r] = 0;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < VT; ++k)                            // a padding row that sampled True adds the row it duplicates
        if ((tid + k * NT) >= kMaxTilePts && m_[k]) addflag[src_[k]] = 1;
      // kept inliers: the old list filtered by the CURRENT flag the removals just cleared (ordered compaction)
      unsigned kw[4], kf = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = tid * 4 + q;
        kw[q] = j < n_old ? ldw(li_[q]) : 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (kw[q] & PW_CUR) { kf |= 1u << q; grow_box(kw[q]); }
      const int kc = __popc(kf);
      const int kincl = warp_incl_scan(kc, lane);
      if (lane == 31) sh.scan[warp] = kincl;
      __syncthreads();                                        // (also publishes addflag)
      scan_warp_totals<NT>(sh.scan, warp, lane);
      __syncthreads();
      const int nK = sh.scan[32];
      {
        int o = sh.scan[warp] + kincl - kc;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if ((kf >> q) & 1u) keepI[o++] = li_[q];
      }
      __syncthreads();
      // added points: the distinct neighbour rows (ascending in point index) whose flag is set
      const int nsetJ = min(S.n_nb, da.Nj);
      int arow = -1;                                          // this thread's neighbour row (one per thread)
      unsigned aw = 0;
      int ap = 0;
#pragma unroll
      for (int k = 0; k < VT; ++k) {
        const int vt = tid + k * NT;
        if (vt >= kMaxTilePts) { arow = vt - kMaxTilePts; aw = w_[k]; ap = p_[k]; }
      }
      const bool af = arow >= 0 && arow < nsetJ && addflag[arow] != 0;
      if (af) grow_box(aw);
      const int aincl = warp_incl_scan(af ? 1 : 0, lane);
      if (lane == 31) sh.scan[warp] = aincl;
      __syncthreads();
      scan_warp_totals<NT>(sh.scan, warp, lane);
      __syncthreads();
      const int nA = sh.scan[32];
      if (af) addA[sh.scan[warp] + aincl - 1] = ap;
      __syncthreads();
      // merge by rank: position = own rank + number of smaller elements of the other list
      auto lower_bound = [](const int* a, int n, int v) {
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
        return lo;
      };
      for (int j = tid; j < nK; j += NT) {
        const int v = keepI[j], pos = j + lower_bound(addA, nA, v);
        listI[pos] = v;
        if (pos < kListCap) sh.listI_s[pos] = v;
      }
      for (int t = tid; t < nA; t += NT) {
        const int v = addA[t], pos = t + lower_bound(keepI, nK, v);
        listI[pos] = v;
        if (pos < kListCap) sh.listI_s[pos] = v;
      }
      n_in = nK + nA;
      __syncthreads();
    } else if (da.sp_perm != nullptr && N >= kSpMinN && N <= kSpMaxN) {
      // every CURRENT point lies within two voxels of the region's box before this step (the adds come from its shell; a
      // re-rounded row may land one voxel further): a box query through the spatial index instead of a scan of the room
      unsigned* bitmap = reinterpret_cast<unsigned*>(sh.planes) + 9 * kMedianSmall;
      int* blklist = reinterpret_cast<int*>(bitmap + kSpMaxN / 32);
      const int lo[3] = {S.minD[0] - 2, S.minD[1] - 2, S.minD[2] - 2}, hi[3] = {S.maxD[0] + 2, S.maxD[1] + 2, S.maxD[2] + 2};
      n_in = scan_box_spatial<NT, 0>(da, S.room, pw, N, lo, hi, [](unsigned w) { return (w & PW_CUR) != 0u; }, grow_box, listI, sh.listI_s,
                                  kListCap, sh.scan, bitmap, blklist, &sh.sp_cnt, spec);
    } else {
      n_in = scan_words<NT>(pw, N, [](unsigned w) { return (w & PW_CUR) != 0; }, grow_box, listI, sh.listI_s, kListCap, sh.scan, spec);
    }
    int reason = STOP_NONE;
    if (!updated) {
      reason = STOP_NOEXPAND;                                // :304-306 (removals alone do not count)
    } else {
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
      if (tid == 0) {
        S.n_in = n_in;
        if (n_in > 0) {
          for (int a = 0; a < 3; ++a) {
            int lo = INT_MAX, hi = INT_MIN;
            for (int w = 0; w < NT / 32; ++w) { lo = min(lo, sh.red[w * 6 + a]); hi = max(hi, sh.red[w * 6 + 3 + a]); }
            S.minD[a] = lo; S.maxD[a] = hi;
          }
          int rsn = STOP_NONE;
          if (!beam) {                                       // (beam search runs this on Q[0] at the head of a round)
            bool expanded = false;
            for (int a = 0; a < 3; ++a) expanded |= (S.minD[a] < S.seqMin[a]) || (S.maxD[a] > S.seqMax[a]);
            if (!expanded) {                                 // :294-299
              if (S.stuck >= 1) rsn = STOP_STUCK; else S.stuck += 1;
            } else {
              S.stuck = 0;                                   // :300-301
            }
            for (int a = 0; a < 3; ++a) { S.seqMin[a] = min(S.seqMin[a], S.minD[a]); S.seqMax[a] = max(S.seqMax[a], S.maxD[a]); }
            if (rsn == STOP_NONE && da.max_steps > 0 && S.steps >= da.max_steps) rsn = STOP_MAXSTEPS;
          }
          sh.flag = rsn;
        } else {
          sh.flag = STOP_EMPTY;
        }
      }
      __syncthreads();
      reason = sh.flag;
      __syncthreads();
    }
    if (tr != nullptr && tid == 0) { tr->stop_reason = reason; tr->size_after = n_in; }
    if (beam && da.score_ml && tid == 0) {                     // newScore = currentScore + addLogProb + rmvLogProb (:264)
      const float parent = __ldcg(&grp->par_score[lane_id / SW]);
      sh.lane_score = __fadd_rn(__fadd_rn(parent, sh.lp[0]), sh.lp[1]);
      if (tr != nullptr) { tr->log_prob[0] = sh.lp[0]; tr->log_prob[1] = sh.lp[1]; tr->score = sh.lane_score; }
    }
    stamp(2);
    if (beam) {
      // one expansion per lane and round: the mask joins newQ if it added a point (:262-267)
      if (!beam_park(reason == STOP_NONE, n_in)) return;
      mode = MODE_NEW_REGION;                              // this lane reported last: it closes the round
    } else if (reason != STOP_NONE) {
      if (!stop_region(reason, n_in)) return;              // (random restarts: parked until the seed's other restarts end)
      mode = MODE_NEW_REGION;
    } else {
      mode = MODE_SCAN;
    }
  }

  bool median_done = false;
  // median of every centred channel over ALL current points (:241): channels 0,1 and 6..F-1
  const int nch = 2 + (da.F > 6 ? da.F - 6 : 0);
  auto row_keys = [&](int j, unsigned (&k)[9]) {
    const float* row = pts + (size_t)(j < kListCap ? sh.listI_s[j] : listI[j]) * 16;
    const float4 a = *reinterpret_cast<const float4*>(row);
    const float4 b = *reinterpret_cast<const float4*>(row + 4);
    const float4 c = *reinterpret_cast<const float4*>(row + 8);
    const float4 d = *reinterpret_cast<const float4*>(row + 12);
    const float vals[9] = {a.x, a.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x};
#pragma unroll
    for (int s = 0; s < 9; ++s) k[s] = sortable(vals[s]);
  };
  stamp(3);
  // ------------------------------------------------------------------ find the next region that needs a forward
  while (true) {
    if (mode == MODE_NEW_REGION && beam) {
      const int r = beam_commit();
      if (r == 0) break;
      if (r == 2) return;
      mode = MODE_SCAN;
    } else if (mode == MODE_NEW_REGION && spec) {
      const int r = spec_head();
      if (r == 0) break;
      if (r == 2) return;
      mode = MODE_SCAN;
    } else if (mode == MODE_NEW_REGION && L > 1) {
      if (!advance_group()) break;
      mode = MODE_SCAN;
    } else if (mode == MODE_NEW_REGION) {
      // next unvisited seed in curvature order (:183-188)
      int found = -1;
      if (S.room >= 0) found = find_seed(S.cursor);
      if (found < 0) {
        // room exhausted (or first launch): publish its stats and fetch the next room from the queue
        if (tid == 0) {
          if (S.room >= 0) {
            LrgRoomStats& st = da.stats[S.room];
            st.n_points = N; st.grow_steps = S.total_steps; st.regions = S.regions; st.clusters = S.cluster_id - 1;
            st.stop_noneighbor = S.stops[0]; st.stop_noexpand = S.stops[1]; st.stop_stuck = S.stops[2]; st.stop_other = S.stops[3];
          }
          const int nr = atomicAdd(da.next_room, 1);
          S.room = nr < da.n_rooms ? (da.room_order != nullptr ? da.room_order[nr] : nr) : -1;
          S.cursor = 0; S.cluster_id = 1; S.total_steps = 0; S.regions = 0; S.visited = 0;
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
        const unsigned w = pw[seed];
        S.cursor = found + 1;
        S.seed = seed;
        S.minD[0] = S.maxD[0] = S.seqMin[0] = S.seqMax[0] = pw_x(w);
        S.minD[1] = S.maxD[1] = S.seqMin[1] = S.seqMax[1] = pw_y(w);
        S.minD[2] = S.maxD[2] = S.seqMin[2] = S.seqMax[2] = pw_z(w);
        S.stuck = 0; S.steps = 0; S.n_in = 1;
        pw[seed] = w | PW_CUR;
        listI[0] = seed;
        sh.listI_s[0] = seed;
      }
      __syncthreads();
      mode = MODE_SCAN;
    }
    stamp(4);
    // neighbour shell: bbox +- 1 voxel, not current, not visited (:222-229)
    const int lo0 = S.minD[0] - 1, lo1 = S.minD[1] - 1, lo2 = S.minD[2] - 1;
    const int hi0 = S.maxD[0] + 1, hi1 = S.maxD[1] + 1, hi2 = S.maxD[2] + 1;
    int n_nb;
    median_done = false;
    if (da.sp_perm != nullptr && N >= kSpMinN && N <= kSpMaxN) {
      // through the room's spatial index: only the blocks of Morton-ordered points whose box meets the shell are read
      unsigned* bitmap = reinterpret_cast<unsigned*>(sh.planes) + 9 * kMedianSmall;   // (behind mkeys; planes / histograms are not live)
      int* blklist = reinterpret_cast<int*>(bitmap + kSpMaxN / 32);
      static_assert(sizeof(sh.planes) >= sizeof(unsigned) * (9 * kMedianSmall + kSpMaxN / 32) + sizeof(int) * (kSpMaxN / kSpBlock), "scratch of the indexed shell scan");
      const int lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
      auto free_word = [](unsigned w) { return (w & (PW_CUR | PW_VIS)) == 0u; };
      if (S.n_in <= kMedianSmall && !(da.tune_step & 1)) {
        // The indexed scan is a chain of L2 round trips with little arithmetic, and the 9-channel median of a small region
        // (:241; needs only the inlier list) is independent of it: part of the CTA scans, the rest selects.
        // (NT = 512: 7 warps scan, 9 warps take one median channel each -- with 8 + 8 one warp selected two channels in a row)
        constexpr int GA = NT == 512 ? 224 : NT / 2, GB = NT - GA;
        if (tid < GA) {
          const int r = scan_box_spatial<GA, 1>(da, S.room, pw, N, lo, hi, free_word, [](unsigned) {}, listJ, sh.listJ_s, kListCap, sh.scan,
                                                bitmap, blklist, &sh.sp_cnt, spec);
          if (tid == 0) sh.sp_total = r;
        } else {
          const int t2 = tid - GA, n_cur = S.n_in;
          for (int j = t2; j < n_cur; j += GB) {
            unsigned k[9];
            row_keys(j, k);
#pragma unroll
            for (int s = 0; s < 9; ++s) sh.mkeys[s][j] = k[s];
          }
          group_sync<2, GB>();
          for (int c = t2 >> 5; c < nch; c += GB / 32) {
            unsigned lo_k, hi_k;
            warp_median<1>(sh.mkeys[c], n_cur, lo_k, hi_k);
            if (lane == 0) { sh.prefix[c] = lo_k; sh.nextkey[c] = hi_k; }
          }
        }
        __syncthreads();
        n_nb = sh.sp_total;
        median_done = true;
      } else {
        n_nb = scan_box_spatial<NT, 0>(da, S.room, pw, N, lo, hi, free_word, [](unsigned) {}, listJ, sh.listJ_s, kListCap, sh.scan, bitmap,
                                       blklist, &sh.sp_cnt, spec);
      }
    } else {
      n_nb = scan_words<NT>(pw, N, [&](unsigned w) {
        if (w & (PW_CUR | PW_VIS)) return false;
        const int x = pw_x(w), y = pw_y(w), z = pw_z(w);
        return x >= lo0 && x <= hi0 && y >= lo1 && y <= hi1 && z >= lo2 && z <= hi2;
      }, [](unsigned) {}, listJ, sh.listJ_s, kListCap, sh.scan, spec);
    }
    if (n_nb == 0 && beam) {                                  // empty shell: the candidate is not expanded (:206)
      if (!beam_park(false, S.n_in)) return;
      mode = MODE_NEW_REGION;
      continue;
    }
    if (n_nb == 0) {                                          // :233-235
      if (!stop_region(STOP_NONEIGHBOR, S.n_in)) return;
      mode = MODE_NEW_REGION;
      continue;
    }
    if (tid == 0) S.n_nb = n_nb;
    __syncthreads();
    stamp(5);
    break;
  }

  if (!S.finished) {
    // ---------------------------------------------------------------- tiles for the next forward (:237-254)
    const int n_in = S.n_in, n_nb = S.n_nb;
    const int room_rng = da.room_id_base + S.room;
    const unsigned step_rng = (unsigned)S.steps;
    const unsigned rng_lane = ((unsigned)S.seed << 8) | lane_stream;
    if (median_done) {
      // (selected beside the shell scan, above)
    } else if (n_in <= kMedianSmall) {
      for (int j = tid; j < n_in; j += NT) {
        unsigned k[9];
        row_keys(j, k);
#pragma unroll
        for (int s = 0; s < 9; ++s) sh.mkeys[s][j] = k[s];
      }
      __syncthreads();
      // one warp per channel: ballots transpose the keys into bit planes, then a bit-sliced rank select
      for (int c = warp; c < nch; c += NT / 32) {
        unsigned lo, hi;
        warp_median<1>(sh.mkeys[c], n_in, lo, hi);
        if (lane == 0) { sh.prefix[c] = lo; sh.nextkey[c] = hi; }
      }
      __syncthreads();
    } else if (n_in <= kMedianCap) {
      // every thread loads the row of one inlier; each warp turns the 32 keys of a channel into 32 bit-plane words with a
      // five-step shuffle transpose (no per-key work in the select below)
      const int nslots = (n_in + 31) >> 5;
      for (int j0 = 0; j0 < nslots * 32; j0 += NT) {
        const int j = j0 + tid, sl = j >> 5;                  // sl is warp-uniform
        if (sl < nslots) {
          unsigned k[9];
          if (j < n_in) row_keys(j, k);
          else {
#pragma unroll
            for (int c = 0; c < 9; ++c) k[c] = 0u;
          }
          const unsigned av = __ballot_sync(0xffffffffu, j < n_in);
          if (lane == 0) sh.malive[sl] = av;
#pragma unroll
          for (int c = 0; c < 9; ++c)
            if (c < nch) sh.planes[c][lane * kPlaneStride + sl] = warp_transpose32(k[c], lane);
        }
      }
      __syncthreads();
      // one warp per channel, no atomics and no block barriers (measured per step, sets of 257-512 / 513-1024 / 1025-2048
      // points: 11k / 15k / 23k cycles; ballot transposition by the channel's own warp 18k / 31k / 55k; a histogram select
      // with shared-memory atomics 2x, a CTA-wide bitonic sort 3-4x slower than that)
      for (int c = warp; c < nch; c += NT / 32) {
        unsigned lo, hi;
        if (n_in <= 1024) warp_median_planes<1>(sh.planes[c], sh.malive, n_in, lo, hi);
        else warp_median_planes<2>(sh.planes[c], sh.malive, n_in, lo, hi);
        if (lane == 0) { sh.prefix[c] = lo; sh.nextkey[c] = hi; }
      }
      __syncthreads();
    } else {
      __syncthreads();
      block_median9<NT>(n_in, nch, row_keys, sh.prefix, sh.rank, sh.hist, sh.nextkey);
    }
    if (tid < 16) {
      float cval = 0.f;
      const int ch = tid < 2 ? tid : tid - 4;                 // feature column -> median channel
      if ((tid < 2 || tid >= 6) && tid < da.F) {
        const float lo = unsortable(sh.prefix[ch]), hi = unsortable(sh.nextkey[ch]);
        cval = (n_in & 1) ? lo : __fmul_rn(__fadd_rn(lo, hi), 0.5f);   // numpy.median: mean of the two middle values
      }
      S.center[tid] = cval;
    }
    __syncthreads();
    if (da.dbg != nullptr && tid == 0) {   // median cost by set size: [16 + 2b] cycles, [17 + 2b] steps; b = size bucket
      const int bkt = n_in <= 64 ? 0 : n_in <= 256 ? 1 : n_in <= 512 ? 2 : n_in <= 1024 ? 3 : n_in <= 2048 ? 4 : 5;
      atomicAdd(da.dbg + 16 + 2 * bkt, (unsigned long long)(clock64() - tstamp));
      atomicAdd(da.dbg + 17 + 2 * bkt, 1ull);
    }
    stamp(6);

    // sampling (:237-240, :249-252) with the Philox stream of oracle/lrg_driver.py PhiloxRng
    const bool fullI = n_in >= da.Ni, fullJ = n_nb >= da.Nj;
    // (sets of up to kKeyCap points keep their sampling keys in shared memory -- behind the histograms of the radix select, in the
    //  scratch of the median planes, which are dead by now -- instead of in global memory: the select reads every key five times)
    constexpr int kKeyCap = 4096;
    static_assert(sizeof(sh.planes) >= 32768 + 2 * kKeyCap * sizeof(unsigned) && sizeof(sh.hist) <= 32768, "sampling keys behind the histograms");
    unsigned* const s_keys = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(sh.planes) + 32768);
    if (n_in <= kKeyCap) keyI = s_keys;
    if (n_nb <= kKeyCap) keyJ = s_keys + kKeyCap;
    if (fullI) for (int j = tid; j < n_in; j += NT) keyI[j] = philox_draw(da.seed, room_rng, step_rng, rng_lane + kStreamInlierKey, j);
    if (fullJ) for (int j = tid; j < n_nb; j += NT) keyJ[j] = philox_draw(da.seed, room_rng, step_rng, rng_lane + kStreamNeighborKey, j);
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
        sh.sel[0][r] = r < n_in ? r : (int)__umulhi(philox_draw(da.seed, room_rng, step_rng, rng_lane + kStreamInlierPad, r - n_in), (unsigned)n_in);
    if (fullJ) block_select_smallest<NT>(n_nb, keyJ, TJ, EJ, sh.sel[1], sh.scan);
    else
      for (int r = tid; r < da.Nj; r += NT)
        sh.sel[1][r] = r < n_nb ? r : (int)__umulhi(philox_draw(da.seed, room_rng, step_rng, rng_lane + kStreamNeighborPad, r - n_nb), (unsigned)n_nb);
    __syncthreads();

    stamp(7);
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
          const int p = pos < kListCap ? (is_nb ? sh.listJ_s : sh.listI_s)[pos] : (is_nb ? listJ : listI)[pos];
          if (r < nset) {                                    // a distinct point: materialise its tile row (16 floats, zero padded)
            const float4* row = reinterpret_cast<const float4*>(pts + (size_t)p * 16);
            float4* out = reinterpret_cast<float4*>(da.tile[is_nb ? 1 : 0] + ((size_t)slot * nrows + r) * 16);
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 t = row[q]; v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w; }
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if ((c < 2 || c >= 6) && c < da.F) v[c] = __fsub_rn(v[c], S.center[c]);
#pragma unroll
            for (int q = 0; q < 4; ++q) out[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
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
          LrgStepTrace* tr = da.trace + ((size_t)S.room * L + lane_id) * da.trace_capacity + S.total_steps;
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
    stamp(8);
  }
  __syncthreads();
  for (int i = tid; i < (int)(sizeof(SlotState) / 4); i += NT)
    reinterpret_cast<int*>(gS)[i] = reinterpret_cast<const int*>(&sh.S)[i];
  stamp(9);
  if (da.dbg != nullptr && tid == 0) atomicAdd(da.dbg + 15, 1ull);
}

}  // namespace lrg
