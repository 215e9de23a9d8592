// The persistent grow kernel: the whole region-growing run of every uploaded room as ONE launch, one CTA per SM.
//
// The lock-step loop (lrg_engine.cu) advances all rooms together, one {step, branch, gproj, head} kernel quartet per
// iteration, so every iteration costs the slowest room's step plus four launch boundaries, and the run lasts as many
// iterations as the longest room has steps.  Here each room slot instead walks its own dependency chain
//
//     STEP(slot) -> BRANCH(slot, branch, tile)... -> { pooled projection, HEAD(slot, head, tile)... } -> STEP(slot) ...
//
// through a device-side work queue: CTAs pop items, run the same device bodies the stand-alone kernels use
// (lrg_step_body.cuh, lrg_tc_tiles.cuh) and the CTA that retires the last item of a stage publishes the next stage.
// Items are only published when they are runnable, so a CTA never waits on another item -- with one exception: head tiles
// go out together with the projection and spin on its counter after their prologue, which is safe because whoever
// computes the projection never waits on a work item (the projection servers below; without them the 8 GPROJ items sit in
// front of the head tiles in the same FIFO).  Rooms progress independently (a slow step of one room no longer stalls the
// others) and nothing is launched per step.
//
// The pooled projection (head layer 0 applied to the pooled 1024-vector, a GEMV) is answered by 16 SERVER CTAs that keep
// the weights in shared memory for the whole run (proj_server) instead of 8 work items that stream them from L2.
//
// Memory ordering: producers finish their global writes, __syncthreads(), then one thread does __threadfence() and
// the atomic / queue store that publishes; the consumer's popping thread spins on the volatile queue entry, does
// __threadfence() (acquire: drops stale L1 lines) and the CTA barrier hands the ordering to the other threads.
// Data produced inside the launch is additionally read with ld.global.cg where it is consumed (never via ld.global.nc).
#include "lrg_persistent.cuh"
#include "lrg_step_body.cuh"
#include "lrg_tc_tiles.cuh"

namespace lrg {

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Multi-producer / multi-consumer ticket ring.  Entry = (generation << 32) | item; generation = ticket / capacity + 1.
__device__ __forceinline__ void queue_push(const GrowQueue& q, const unsigned* items, int n) {
  const unsigned t = atomicAdd(q.tail, (unsigned)n);
  for (int i = 0; i < n; ++i) {
    const unsigned idx = t + (unsigned)i;
    const unsigned long long gen = (unsigned long long)(idx / (q.cap_mask + 1u)) + 1ull;
    *reinterpret_cast<volatile unsigned long long*>(q.ring + (idx & q.cap_mask)) = (gen << 32) | items[i];
  }
}

// Watchdog of the spinning CTAs.  Idle CTAs are normal (one large room keeps ~10 of 148 busy), so the test is GLOBAL progress:
// thread 0 of every CTA bumps GrowArgs::progress after each item; a spinner that has seen that counter stand still for
// kStallNs (20 s of globaltimer -- independent of the SM clock, profilers and throttling slow the run but items keep retiring)
// raises GrowArgs::abort, and every spinner that sees the flag leaves.  The host then returns LRG_E_STATE; nothing traps, the
// context stays usable.
constexpr unsigned long long kStallNs = 20000000000ull;
struct StallWatch {
  unsigned last;
  unsigned long long t_last;
  unsigned spins;
  __device__ __forceinline__ void start(const GrowArgs& ga) { last = *reinterpret_cast<volatile unsigned*>(ga.progress); t_last = global_ns(); spins = 0; }
  // true: give up (the run is wedged, or somebody else found it to be)
  __device__ __forceinline__ bool stalled(const GrowArgs& ga) {
    if ((++spins & 1023u) != 0) return false;
    if (*reinterpret_cast<volatile int*>(ga.abort) != 0) return true;
    const unsigned p = *reinterpret_cast<volatile unsigned*>(ga.progress);
    const unsigned long long now = global_ns();
    if (p != last) { last = p; t_last = now; return false; }
    if (now - t_last > kStallNs) {
      *reinterpret_cast<volatile int*>(ga.abort) = 1;
      __threadfence();
      return true;
    }
    return false;
  }
};

// Blocking pop of the normal ring: take a ticket, spin on that ticket's own entry (idle CTAs poll distinct addresses).
__device__ __forceinline__ unsigned queue_pop(const GrowArgs& ga, const GrowQueue& q) {
  const unsigned h = atomicAdd(q.head, 1u);
  const unsigned long long gen = (unsigned long long)(h / (q.cap_mask + 1u)) + 1ull;
  const volatile unsigned long long* e = q.ring + (h & q.cap_mask);
  unsigned long long v = *e;
  if ((v >> 32) != gen) {
    StallWatch w;
    w.start(ga);
    while (((v = *e) >> 32) != gen) {
      __nanosleep(32);
      if (w.stalled(ga)) return make_item(ITEM_EXIT, 0, 0, 0);
    }
  }
  return (unsigned)v;                                // (the caller issues the acquire fence where the item needs one)
}

// Pooled-projection server: gproj[slot][h][c0 + c] = bias0_h[c0 + c] + sum_k pooled[slot][k] * W0g_h[k][c0 + c] for the 32
// columns of this CTA.  The CTA's 1024 x 32 weights live in REGISTERS for the whole run -- lane c of warp j keeps the 64
// weights of column c0 + c in K-groups 2j and 2j+1 (rows 64j .. 64j+63) -- so nothing is streamed from L2 per grow step
// (256 KB per 64-column block item otherwise) and the weights are not re-read from shared memory either (a first version
// kept them there: 128 KB of shared-memory reads per request, ~2k cycles of the port, 64 % utilisation at the bench's request
// rate -- the head tiles waited 7.8 us for their projection under load against 2.4 us on an idle machine).
// The 16 warps run free of one another (no CTA barrier in the loop).  Requests are numbered by ticket; for ticket t
//   * warp t mod 16 is its LOADER: it polls the broadcast request ring in global memory (one polling warp per server and
//     ticket -- 256 warps polling one L2 line made the producers' stores queue behind them), copies the slot's pooled row
//     (4 KB) into a ring of row buffers in shared memory and then publishes the request in a shared-memory ring; it does this
//     up to kProjLook tickets AHEAD of its own arithmetic, so under load the two L2 round trips are off the request path;
//   * EVERY warp waits for the request in the shared-memory ring, reads its 64 values of the row (broadcast 128-bit loads),
//     runs its two 32-long fmaf chains and drops the two partial sums per column into a ring of partial-sum buffers;
//   * warp t mod 16 is also its REDUCER: it waits for the 16 arrivals on the buffer's mbarrier, combines the 32 K-groups in
//     order on top of the bias and stores every value together with the tag of the slot's forward as one 8-byte word -- the
//     head tiles poll the tags, so the server needs no fence and no counter (a fence + atomic per request kept the reducer
//     ~800 cycles from the next request's arithmetic, which every request needs from every warp: a serial chain).
// Summation order = tc_gproj_block's (32 K-groups of 32 rows, each a sequential fmaf chain from 0, combined in order on top of
// the bias), so both paths give the same bits.
constexpr int kProjBufs = 8;                         // row / partial-sum buffers in flight per server
constexpr int kProjLook = 4;                         // tickets a loader runs ahead of its own arithmetic (< kProjBufs)
constexpr int kProjReqRing = 32;                     // shared-memory request ring (> kProjBufs + kProjLook)
__device__ void proj_server(const GrowArgs& ga, unsigned char* smem) {
  const int tid = threadIdx.x, c = tid & 31, j = tid >> 5;           // j = warp: K-groups 2j, 2j + 1
  const int h = (int)blockIdx.x / 8, c0 = ((int)blockIdx.x % 8) * 32;
  float* const sP = reinterpret_cast<float*>(smem);                  // [kProjBufs][1024] pooled rows
  float* const sR = sP + kProjBufs * 1024;                           // [kProjBufs][32 K-groups][32 columns]
  volatile unsigned long long* const s_req = reinterpret_cast<volatile unsigned long long*>(sR + kProjBufs * 1024);   // [kProjReqRing]
  uint64_t* const bars = reinterpret_cast<uint64_t*>(const_cast<unsigned long long*>(s_req) + kProjReqRing);        // full[], empty[]
  if (tid < kProjReqRing) s_req[tid] = 0ull;
  if (tid == 0) {
    for (int i = 0; i < kProjBufs; ++i) { mbar_init(smem_u32(&bars[i]), 16); mbar_init(smem_u32(&bars[kProjBufs + i]), 1); }
    fence_barrier_init();
  }
  float w[64];
  {
    const float* W = ga.net.W0g[h] + (size_t)(64 * j) * 256 + c0 + c;
#pragma unroll
    for (int i = 0; i < 64; ++i) w[i] = __ldg(W + (size_t)i * 256);
  }
  const float bias = __ldg(ga.net.head_bias0[h] + c0 + c);
  __syncthreads();                                                   // (ring and barriers initialised: the only CTA barrier)
  auto gen_of = [&](unsigned ticket) { return (unsigned long long)(ticket / (ga.greq_mask + 1u)) + 1ull; };
  // Loader duty for ticket t2 (whole warp): 1 = done, 0 = not possible yet (non-blocking only), -1 = the run was abandoned
  auto load_request = [&](unsigned t2, bool blocking) -> int {
    const int buf = (int)(t2 % kProjBufs);
    if (t2 >= (unsigned)kProjBufs) {
      // the row buffer is free once all 16 warps have arrived for ticket t2 - kProjBufs (they read the row before they arrive)
      const uint32_t par = ((t2 - kProjBufs) / kProjBufs) & 1u;
      if (blocking) mbar_wait(smem_u32(&bars[buf]), par);
      else {
        int ok = 0;                                    // (one lane asks: the answer must be warp-uniform)
        if (c == 0) ok = (int)mbar_try_wait(smem_u32(&bars[buf]), par);
        if (!__shfl_sync(0xffffffffu, ok, 0)) return 0;
      }
    }
    const volatile unsigned long long* e = ga.greq_ring + (t2 & ga.greq_mask);
    unsigned long long v = 0;
    if (c == 0) v = *e;
    v = __shfl_sync(0xffffffffu, v, 0);
    if ((v >> 32) != gen_of(t2)) {
      if (!blocking) return 0;
      int state = 0;                                   // lane 0 polls; 1 = there, -1 = give up
      if (c == 0) {
        StallWatch sw;
        sw.start(ga);
        while (((v = *e) >> 32) != gen_of(t2)) {
          __nanosleep(20);
          if (sw.stalled(ga)) { state = -1; break; }
        }
      }
      state = __shfl_sync(0xffffffffu, state, 0);
      if (state < 0) return -1;
      v = __shfl_sync(0xffffffffu, v, 0);
    }
    const unsigned slot = (unsigned)v & 0x1FFFu;
    if (slot != kProjExit) {
      const float4* src = reinterpret_cast<const float4*>(ga.fa.pooled + (size_t)slot * 1024);
      float4 r[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) r[q] = __ldcg(src + q * 32 + c);
      float4* dst = reinterpret_cast<float4*>(sP + buf * 1024);
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q * 32 + c] = r[q];
    }
    __syncwarp();
    if (c == 0) {
      __threadfence_block();
      s_req[t2 % kProjReqRing] = ((unsigned long long)(t2 + 1u) << 32) | (unsigned)v;   // (tagged by ticket: the ring is shorter than a generation)
    }
    return 1;
  };
  unsigned next_duty = (unsigned)j;                    // my next ticket as loader
  for (unsigned ticket = 0;; ++ticket) {
    if (next_duty <= ticket + (unsigned)kProjLook) {
      const int r = load_request(next_duty, next_duty == ticket);
      if (r < 0) break;
      if (r > 0) next_duty += 16u;
    }
    // the request, once its loader has staged the row
    unsigned long long v = s_req[ticket % kProjReqRing];
    if ((v >> 32) != (unsigned long long)(ticket + 1u)) {
      unsigned spins = 0;
      bool give_up = false;
      while (((v = s_req[ticket % kProjReqRing]) >> 32) != (unsigned long long)(ticket + 1u)) {
        __nanosleep(20);
        if ((++spins & 0xFFFFu) == 0 && *reinterpret_cast<volatile int*>(ga.abort) != 0) { give_up = true; break; }
        // (my own loader duty may have become possible meanwhile: rows are staged kProjLook tickets ahead)
        if ((spins & 63u) == 0 && next_duty <= ticket + (unsigned)kProjLook && next_duty != ticket) {
          const int r = load_request(next_duty, false);
          if (r > 0) next_duty += 16u;
        }
      }
      if (give_up) break;
    }
    __threadfence_block();
    const unsigned slot = (unsigned)v & 0x1FFFu, tag = proj_tag((unsigned)v >> 13);
    if (slot == kProjExit) break;
    const int buf = (int)(ticket % kProjBufs);
    const uint32_t phase = (ticket / kProjBufs) & 1u;
    float a0 = 0.f, a1 = 0.f;
    {
      const float4* p4 = reinterpret_cast<const float4*>(sP + buf * 1024 + 64 * j);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x = p4[q], y = p4[8 + q];
        a0 = fmaf(x.x, w[q * 4 + 0], a0); a1 = fmaf(y.x, w[32 + q * 4 + 0], a1);
        a0 = fmaf(x.y, w[q * 4 + 1], a0); a1 = fmaf(y.y, w[32 + q * 4 + 1], a1);
        a0 = fmaf(x.z, w[q * 4 + 2], a0); a1 = fmaf(y.z, w[32 + q * 4 + 2], a1);
        a0 = fmaf(x.w, w[q * 4 + 3], a0); a1 = fmaf(y.w, w[32 + q * 4 + 3], a1);
      }
    }
    mbar_wait(smem_u32(&bars[kProjBufs + buf]), phase ^ 1u);         // the buffer's previous request has been reduced
    float* R = sR + buf * 1024;
    R[(2 * j) * 32 + c] = a0;
    R[(2 * j + 1) * 32 + c] = a1;
    __syncwarp();
    if (c == 0) mbar_arrive(smem_u32(&bars[buf]));
    if (j == (int)(ticket & 15u)) {
      mbar_wait(smem_u32(&bars[buf]), phase);
      float s2 = bias;
#pragma unroll
      for (int g2 = 0; g2 < 32; ++g2) s2 += R[g2 * 32 + c];
      // value and tag in ONE 8-byte store: whoever reads the tag of this forward has the value -- no fence, no counter
      __stcg(ga.gproj_tagged + ((size_t)slot * 2 + h) * 256 + c0 + c, make_uint2(__float_as_uint(s2), tag));
      __syncwarp();
      if (c == 0) mbar_arrive(smem_u32(&bars[kProjBufs + buf]));
    }
  }
}

template <bool F16>
__global__ void __launch_bounds__(kGrowThreads, 1) lrg_grow_kernel(const __grid_constant__ GrowArgs ga) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(16) TcStatic st;
  __shared__ uint32_t tmem_base;
  __shared__ unsigned s_item;
  const int tid = threadIdx.x, warp = tid >> 5;
  if ((int)blockIdx.x < ga.n_servers) {               // (no tensor memory, no work items: requests only)
    proj_server(ga, smem);
    return;
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base), kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base;
  // the driver scratch aliases the weight ring of the tensor tiles: a CTA runs one item at a time
  StepShared& sh = *reinterpret_cast<StepShared*>(smem);
  float* const sP = reinterpret_cast<float*>(smem);
  float* const sR = sP + 1024;

  // Scheduling (optional, hi_ctas > 0): the run ends with the rooms that have the most work left, so the few slots with the
  // most unvisited points are served by RESERVED CTAs -- CTAs below hi_ctas pop ring 0 only, the others ring 1 only; a
  // producer sends a high-priority slot's items to ring 0 only when that many reserved CTAs are waiting there right now
  // (all or nothing, so the FIFO argument for the head tiles holds within a ring), else to ring 1 like everybody else's.
  const int my_ring = (ga.hi_ctas > 0 && (int)blockIdx.x - ga.n_servers < ga.hi_ctas) ? 0 : 1;
  unsigned chained = 0;                               // item this CTA hands to itself (the STEP that follows the last head tile)
  while (true) {
    if (tid == 0) {
      unsigned it = chained;
      chained = 0;
      if (it == 0) it = queue_pop(ga, ga.q[my_ring]);
      // acquire: the driver step reads state other CTAs wrote with plain stores (drop stale L1 lines); the tensor tiles and
      // the projection read everything produced in this launch with ld.global.cg and need no fence
      if ((it & 7u) == ITEM_STEP) __threadfence();
      s_item = it;
    }
    __syncthreads();
    const unsigned item = s_item;
    const int type = (int)(item & 7u), slot = (int)((item >> 3) & 0x1FFFu), a = (int)((item >> 16) & 15u), t = (int)((item >> 20) & 15u);
    if (type == ITEM_EXIT) break;
    const unsigned long long t0 = (tid == 0) ? global_ns() : 0ull;
    SlotSync* sy = ga.sync + slot;
    if (tid == 0 && ga.busy_ns != nullptr) {
      // diagnostics: how long the item sat between being published and being picked up
      const unsigned long long tp = *reinterpret_cast<volatile unsigned long long*>(&sy->t_pub);
      if (tp != 0 && t0 > tp) atomicAdd(ga.busy_ns + 16 + type, t0 - tp);
    }
    unsigned next[32 + kMaxLanes];
    int n_next = 0;
    if (type == ITEM_STEP) {
      step_body<kGrowThreads>(ga.da, slot, sh);
      __syncthreads();
      if (tid == 0) {
        if (sh.S.finished) *reinterpret_cast<volatile int*>(ga.remaining + slot) = 0;
        if (sh.all_done) {
          // the last slot has retired: nothing is in flight any more, release every CTA
          if (ga.n_servers > 0) {                          // one closing request: every server reads every entry
            __threadfence();
            const unsigned idx = atomicAdd(ga.greq_tail, 1u);
            *reinterpret_cast<volatile unsigned long long*>(ga.greq_ring + (idx & ga.greq_mask)) =
                (((unsigned long long)(idx / (ga.greq_mask + 1u)) + 1ull) << 32) | kProjExit;
          }
          for (int ring = 0; ring < 2; ++ring)
            for (unsigned left = ring == 0 ? (unsigned)ga.hi_ctas : gridDim.x - (unsigned)(ga.hi_ctas + ga.n_servers); left > 0;) {
              const int n = left > 16u ? 16 : (int)left;
              for (int i = 0; i < n; ++i) next[i] = make_item(ITEM_EXIT, 0, 0, 0);
              __threadfence();
              queue_push(ga.q[ring], next, n);
              left -= (unsigned)n;
            }
        } else if (sh.S.active && !sh.S.finished) {
          // scheduling (optional): rank this slot by the unvisited points of its room; the bar for the high-priority queue
          // (the hi_slots-th largest count over all slots) is refreshed by every 64th step of a slot
          sy->prio = 1;
          if (ga.hi_ctas > 0 && ga.hi_crit) {
            const int L2 = ga.da.lanes > 1 ? ga.da.lanes : 1;
            if (*reinterpret_cast<volatile int*>(ga.da.spec_est + ga.da.n_slots / L2 + slot / L2) != 0) sy->prio = 0;
          } else if (ga.hi_ctas > 0) {
            const int mine = (int)(ga.da.room_off[sh.S.room + 1] - ga.da.room_off[sh.S.room]) - sh.S.visited;
            *reinterpret_cast<volatile int*>(ga.remaining + slot) = mine;
            volatile int* bar = ga.remaining + ga.da.n_slots;
            if ((sh.S.total_steps & 63) == 0) {
              int top[8];
              for (int k = 0; k < ga.hi_slots; ++k) top[k] = 0;
              for (int i = 0; i < ga.da.n_slots; ++i) {
                int v = *reinterpret_cast<volatile int*>(ga.remaining + i);
                for (int k = 0; k < ga.hi_slots; ++k)
                  if (v > top[k]) { const int t2 = top[k]; top[k] = v; v = t2; }
              }
              *bar = top[ga.hi_slots - 1];
            }
            if (mine >= *bar) sy->prio = 0;
          }
          // only rows that carry distinct points are evaluated (the rest are padding duplicates of them)
          const int tilesI = (min(sh.S.n_in, ga.fa.n_pts[0]) + 127) / 128, tilesJ = (min(sh.S.n_nb, ga.fa.n_pts[1]) + 127) / 128;
          // When enough CTAs are idle every branch tile is split over 2 or 4 CTAs by column block of the last
          // layer -- each recomputes the cheap first four layers -- which shortens the critical path of the run's tail;
          // under load tiles stay whole (splitting costs SM time).  a = branch | log2(parts) << 1, t = tile | part << 2.
          // (idle CTAs hold tickets ahead of the tail, so tail - head is negative when the machine has spare SMs)
          const int b1 = (int)(*reinterpret_cast<volatile unsigned*>(ga.q[1].tail) - *reinterpret_cast<volatile unsigned*>(ga.q[1].head));
          const int idle = b1 < 0 ? -b1 : 0;               // CTAs waiting for work right now
          const int lg = !(ga.tune & 1) ? 0 : idle >= 8 * (tilesI + tilesJ) ? 2 : idle >= 4 * (tilesI + tilesJ) ? 1 : 0;
          const int parts = 1 << lg;
          sy->branch_left = (tilesI + tilesJ) * parts;
          sy->seq += 1u;                                 // (this CTA owns the slot between a STEP and its publication)
          sy->gproj_left = 8;
          sy->head_left = tilesI + tilesJ;
          sy->tiles[0] = tilesI;
          sy->tiles[1] = tilesJ;
          for (int part = 0; part < parts; ++part) {
            for (int i = 0; i < tilesI; ++i) next[n_next++] = make_item(ITEM_BRANCH, slot, 0 | (lg << 1), i | (part << 2));
            for (int i = 0; i < tilesJ; ++i) next[n_next++] = make_item(ITEM_BRANCH, slot, 1 | (lg << 1), i | (part << 2));
          }
        }
        if (sh.wake && !sh.all_done) {
          // random restarts: this step committed a seed and started every lane of its group on the next one; beam search:
          // it closed a round and handed the candidates of the next one to these lanes
          const int L = ga.da.lanes, first_slot = slot - slot % L;
          for (int l = 0; l < L; ++l)
            if ((sh.wake >> l) & 1u) next[n_next++] = make_item(ITEM_STEP, first_slot + l, 0, 0);
        }
      }
    } else if (type == ITEM_BRANCH) {
      {
        const int br = a & 1, part = t >> 2, nbs = 4 >> (a >> 1);
        tc_branch_tile<F16>(ga.net, ga.fa, slot, br, t & 3, forward_valid_rows(ga.fa, slot, br), part * nbs, (part + 1) * nbs, smem, st, tmem);
      }
      if (tid == 0) {
        __threadfence();
        if (atomicSub(&sy->branch_left, 1) == 1) {
          // the pooled row is complete: publish the projection blocks and, behind them in the FIFO, the head tiles -- a CTA
          // that pops a head tile knows every projection block of its slot is already running (or done), so the head's
          // wait on gproj_left cannot deadlock, and its prologue overlaps the projection
          if (ga.n_servers > 0) {
            // one request to the projection servers (they hold the weights in shared memory); the fence above ordered the
            // pooled row before it.  The head tiles go out right away: prologue and first MMAs overlap the servers, the tiles
            // spin on gproj_left (servers never wait on a work item, so this cannot deadlock).  (Letting the server that
            // answers last publish them instead -- no spinning CTAs -- was measured slower in every regime: 342 vs 297 ms
            // plain, 1350 vs 1180 ms with 10 restarts.)
            const unsigned idx = atomicAdd(ga.greq_tail, 1u);
            *reinterpret_cast<volatile unsigned long long*>(ga.greq_ring + (idx & ga.greq_mask)) =
                (((unsigned long long)(idx / (ga.greq_mask + 1u)) + 1ull) << 32) | (unsigned)slot | ((__ldcg(&sy->seq) & kProjSeqMask) << 13);
          } else {
            for (int h = 0; h < 2; ++h)
              for (int cb = 0; cb < 4; ++cb) next[n_next++] = make_item(ITEM_GPROJ, slot, h, cb);
          }
          if ((ga.tune & 2) || ga.n_servers > 0) {
            const int tilesI = __ldcg(&sy->tiles[0]), tilesJ = __ldcg(&sy->tiles[1]);
            for (int i = 0; i < tilesI; ++i) next[n_next++] = make_item(ITEM_HEAD, slot, 0, i);
            for (int i = 0; i < tilesJ; ++i) next[n_next++] = make_item(ITEM_HEAD, slot, 1, i);
          }
        }
      }
    } else if (type == ITEM_GPROJ) {
      tc_gproj_block(ga.net, ga.fa, slot, a, t, sP, sR);
      if (tid == 0) {
        __threadfence();
        if (atomicSub(&sy->gproj_left, 1) == 1 && !(ga.tune & 2)) {   // (with tune bit 1 the head tiles are already out, spinning on this)
          const int tilesI = __ldcg(&sy->tiles[0]), tilesJ = __ldcg(&sy->tiles[1]);
          for (int i = 0; i < tilesI; ++i) next[n_next++] = make_item(ITEM_HEAD, slot, 0, i);
          for (int i = 0; i < tilesJ; ++i) next[n_next++] = make_item(ITEM_HEAD, slot, 1, i);
        }
      }
    } else if (type == ITEM_HEAD) {
      if (ga.n_servers > 0)
        tc_head_tile<F16>(ga.net, ga.fa, slot, a, t, forward_valid_rows(ga.fa, slot, a), nullptr, ga.gproj_tagged + ((size_t)slot * 2 + a) * 256,
                          proj_tag(__ldcg(&sy->seq)), smem, st, tmem);
      else
        tc_head_tile<F16>(ga.net, ga.fa, slot, a, t, forward_valid_rows(ga.fa, slot, a), &sy->gproj_left, nullptr, 0u, smem, st, tmem);
      if (tid == 0) {
        __threadfence();
        // the CTA that retires the slot's last head tile runs the slot's next driver step itself: no queue hop, and under
        // load the step does not wait behind other slots' tiles
        if (atomicSub(&sy->head_left, 1) == 1) chained = make_item(ITEM_STEP, slot, 0, 0);
      }
    }
    if (tid == 0) {
      // keep the first successor for this CTA (a branch tile after a STEP, a projection block after the last branch tile:
      // neither ever waits on another item), publish the rest
      int first = 0;
      if (n_next > 0 && (type == ITEM_STEP || type == ITEM_BRANCH)) { chained = next[0]; first = 1; }
      if (n_next > first) {
        *reinterpret_cast<volatile unsigned long long*>(&sy->t_pub) = global_ns();
        // release: a STEP publishes the counters it just wrote; the tile / projection finishers already fenced before the
        // atomic that made them last (the push is control-dependent on that atomic's result)
        if (type == ITEM_STEP) __threadfence();
        int ring = 1;
        if (ga.hi_ctas > 0 && (*reinterpret_cast<volatile int*>(&sy->prio) & 1) == 0) {
          const int waiting = (int)(*reinterpret_cast<volatile unsigned*>(ga.q[0].head) - *reinterpret_cast<volatile unsigned*>(ga.q[0].tail));
          if (waiting >= n_next - first) ring = 0;
        }
        queue_push(ga.q[ring], next + first, n_next - first);
      }
      if (ga.busy_ns != nullptr) {
        atomicAdd(ga.busy_ns + type, global_ns() - t0);
        atomicAdd(ga.busy_ns + 8 + type, 1ull);
      }
      atomicAdd(ga.progress, 1u);                      // (the watchdog of the spinning CTAs looks at this)
    }
    __syncthreads();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

int grow_configure() {
  LRG_CUDA(cudaFuncSetAttribute(lrg_grow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LRG_CUDA(cudaFuncSetAttribute(lrg_grow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  return LRG_OK;
}

// Cooperative launch: the kernel's CTAs wait for one another (work queue, projection servers), so they must all be resident
// -- one per SM, each with all 512 TMEM columns.  cudaLaunchCooperativeKernel guarantees co-residency (the launch is held
// back until the grid fits) and refuses a grid that can never fit, which is reported as LRG_E_STATE instead of a spin.
int launch_grow(const GrowArgs& ga, int n_ctas, bool f16, cudaStream_t stream) {
  static_assert(sizeof(StepShared) <= kTcSmem, "driver scratch must fit the shared memory it aliases");
  const void* fn = f16 ? (const void*)lrg_grow_kernel<true> : (const void*)lrg_grow_kernel<false>;
  int per_sm = 0, dev = 0, sms = 0, coop = 0;
  LRG_CUDA(cudaGetDevice(&dev));
  LRG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  LRG_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  LRG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kGrowThreads, kTcSmem));
  if (per_sm < 1 || n_ctas > per_sm * sms || !coop) {
    set_error("persistent grow kernel: %d CTAs cannot be co-resident on this device (%d SMs x %d CTAs per SM, cooperative launch %s)",
              n_ctas, sms, per_sm, coop ? "supported" : "unsupported");
    return LRG_E_STATE;
  }
  void* args[] = {const_cast<GrowArgs*>(&ga)};
  const cudaError_t err = cudaLaunchCooperativeKernel(fn, dim3((unsigned)n_ctas), dim3(kGrowThreads), args, kTcSmem, stream);
  if (err == cudaErrorCooperativeLaunchTooLarge) {
    cudaGetLastError();
    set_error("persistent grow kernel: the device cannot hold %d co-resident CTAs right now", n_ctas);
    return LRG_E_STATE;
  }
  LRG_CUDA(err);
  return LRG_OK;
}

}  // namespace lrg
