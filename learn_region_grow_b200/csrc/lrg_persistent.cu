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

// Blocking pop of the normal ring: take a ticket, spin on that ticket's own entry (idle CTAs poll distinct addresses).
__device__ __forceinline__ unsigned queue_pop(const GrowQueue& q) {
  const unsigned h = atomicAdd(q.head, 1u);
  const unsigned long long gen = (unsigned long long)(h / (q.cap_mask + 1u)) + 1ull;
  const volatile unsigned long long* e = q.ring + (h & q.cap_mask);
  unsigned long long v = *e;
  if ((v >> 32) != gen) {
    const long long t0 = clock64();
    while (((v = *e) >> 32) != gen) {
      __nanosleep(32);
      if (clock64() - t0 > 60000000000ll) asm volatile("trap;");   // ~30 s without work: the run is wedged, fail loudly
    }
  }
  return (unsigned)v;                                // (the caller issues the acquire fence where the item needs one)
}

// Pooled-projection server: gproj[slot][h][c0 + c] = bias0_h[c0 + c] + sum_k pooled[slot][k] * W0g_h[k][c0 + c] for the 32
// columns of this CTA, the weights resident in shared memory (128 KB) instead of being streamed from L2 by every grow step
// (256 KB per 64-column block item).  Four request groups of 128 threads work on different requests at once (group g takes
// the tickets = g mod 4 of the broadcast ring); summation order = tc_gproj_block's (32 K-groups of 32 rows, each a
// sequential fmaf chain from 0, combined in order on top of the bias), so both paths give the same bits.
__device__ void proj_server(const GrowArgs& ga, unsigned char* smem) {
  const int tid = threadIdx.x, grp = tid >> 7, t = tid & 127, c = t & 31, kq = t >> 5;
  const int h = (int)blockIdx.x / 8, c0 = ((int)blockIdx.x % 8) * 32;
  float* const sW = reinterpret_cast<float*>(smem);                 // [1024][32]
  float* const sP = sW + 1024 * 32 + grp * 2048;                    // [1024] pooled row of this group's request
  float* const sR = sP + 1024;                                      // [32][32] K-group partial sums
  {
    const float4* W = reinterpret_cast<const float4*>(ga.net.W0g[h] + c0);
    for (int i = tid; i < 1024 * 8; i += kGrowThreads) {
      const int k = i >> 3, q = i & 7;
      reinterpret_cast<float4*>(sW)[i] = __ldg(W + (size_t)k * 64 + q);
    }
  }
  const float bias = __ldg(ga.net.head_bias0[h] + c0 + c);
  __syncthreads();
  for (unsigned ticket = (unsigned)grp;; ticket += 4u) {
    const unsigned long long gen = (unsigned long long)(ticket / (ga.greq_mask + 1u)) + 1ull;
    const volatile unsigned long long* e = ga.greq_ring + (ticket & ga.greq_mask);
    unsigned long long v = *e;
    if ((v >> 32) != gen) {
      const long long t0 = clock64();
      while (((v = *e) >> 32) != gen) {
        __nanosleep(20);
        if (clock64() - t0 > 60000000000ll) asm volatile("trap;");
      }
    }
    const unsigned slot = (unsigned)v;
    if (slot == kProjExit) break;
    // (every thread polled the entry itself: no broadcast through shared memory, no barrier before the loads)
    const float4* prow = reinterpret_cast<const float4*>(ga.fa.pooled + (size_t)slot * 1024);
    reinterpret_cast<float4*>(sP)[t] = __ldcg(prow + t);
    reinterpret_cast<float4*>(sP)[t + 128] = __ldcg(prow + t + 128);
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1));
    {
      // eight K-groups per thread as eight independent fmaf chains (each chain sequential in k like tc_gproj_block)
      float acc[8];
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) acc[g8] = 0.f;
      const float* w = sW + (kq * 8 * 32) * 32 + c;
      const float* p = sP + kq * 8 * 32;
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) acc[g8] = fmaf(p[g8 * 32 + i], w[(g8 * 32 + i) * 32], acc[g8]);
      }
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) sR[(kq * 8 + g8) * 32 + c] = acc[g8];
    }
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1));
    if (t < 32) {
      float s2 = bias;
#pragma unroll
      for (int g2 = 0; g2 < 32; ++g2) s2 += sR[g2 * 32 + t];
      ga.fa.gproj[((size_t)slot * 2 + h) * 256 + c0 + t] = s2;
      __syncwarp();
      if (t == 0) {
        __threadfence();
        atomicSub(&ga.sync[slot].gproj_left, 1);
      }
    }
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1));     // sP / sR are reused by the group's next request
  }
}

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
      if (it == 0) it = queue_pop(ga.q[my_ring]);
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
          if (ga.n_servers > 0) {                          // one closing request per request group of the servers
            __threadfence();
            const unsigned t4 = atomicAdd(ga.greq_tail, 4u);
            for (unsigned i = 0; i < 4u; ++i) {
              const unsigned idx = t4 + i;
              *reinterpret_cast<volatile unsigned long long*>(ga.greq_ring + (idx & ga.greq_mask)) =
                  (((unsigned long long)(idx / (ga.greq_mask + 1u)) + 1ull) << 32) | kProjExit;
            }
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
          if (ga.hi_ctas > 0) {
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
          sy->gproj_left = ga.n_servers > 0 ? ga.n_servers : 8;
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
        tc_branch_tile(ga.net, ga.fa, slot, br, t & 3, forward_valid_rows(ga.fa, slot, br), part * nbs, (part + 1) * nbs, smem, st, tmem);
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
                (((unsigned long long)(idx / (ga.greq_mask + 1u)) + 1ull) << 32) | (unsigned)slot;
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
      tc_head_tile(ga.net, ga.fa, slot, a, t, forward_valid_rows(ga.fa, slot, a), &sy->gproj_left, smem, st, tmem);
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
    }
    __syncthreads();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

int grow_configure() {
  LRG_CUDA(cudaFuncSetAttribute(lrg_grow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  return LRG_OK;
}

int launch_grow(const GrowArgs& ga, int n_ctas, cudaStream_t stream) {
  static_assert(sizeof(StepShared) <= kTcSmem, "driver scratch must fit the shared memory it aliases");
  lrg_grow_kernel<<<n_ctas, kGrowThreads, kTcSmem, stream>>>(ga);
  LRG_CUDA(cudaGetLastError());
  return LRG_OK;
}

}  // namespace lrg
