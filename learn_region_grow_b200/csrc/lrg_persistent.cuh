// Persistent grow kernel: work items, queue and launch arguments (internal).
#pragma once
#include "lrg_driver.cuh"
#include "lrg_tc.cuh"

namespace lrg {

constexpr int kGrowThreads = 512;
enum { ITEM_STEP = 1, ITEM_BRANCH = 2, ITEM_GPROJ = 3, ITEM_HEAD = 4, ITEM_EXIT = 7 };

// type: bits [0,3); slot: bits [3,16); a (branch / head index): bits [16,20); t (tile / column block): bits [20,24)
__host__ __device__ inline unsigned make_item(int type, int slot, int a, int t) {
  return (unsigned)type | ((unsigned)slot << 3) | ((unsigned)a << 16) | ((unsigned)t << 20);
}
constexpr int kMaxGrowSlots = 8190;       // (slot ids travel in 13 bits; 0x1FFF is the servers' exit request)
constexpr int kProjServers = 16;         // 2 heads x 8 slices of 32 columns
constexpr unsigned kProjExit = 0x1FFFu;  // request that ends a server's request group

struct GrowQueue {
  unsigned long long* ring;     // capacity entries, zero-initialised; entry = (generation << 32) | item
  unsigned cap_mask;            // capacity - 1 (capacity is a power of two > the items that can be outstanding)
  unsigned* head;               // next ticket to pop
  unsigned* tail;               // next ticket to push
};

// seq: forwards this slot has started in this run (the tag of its pooled projection when the servers answer it)
struct SlotSync { int branch_left, gproj_left, head_left, prio; int tiles[2]; unsigned long long t_pub; unsigned seq; unsigned pad; };
// Request of the projection servers: bits [0,13) slot, bits [13,32) the low bits of the slot's forward count
constexpr unsigned kProjSeqMask = 0x7FFFFu;
__host__ __device__ inline unsigned proj_tag(unsigned seq) { return (seq & kProjSeqMask) + 1u; }   // never 0 (the buffer starts zeroed)

struct GrowArgs {
  DriverArgs da;
  ForwardArgs fa;
  TcNet net;
  GrowQueue q[2];               // [0] served by the reserved CTAs (items of the slots with the most unvisited points left), [1] everybody else
  int* remaining;               // (n_slots + 1) unvisited points of the slot's room, refreshed by every STEP; [n_slots] = the
                                // count a slot needs to be served from the high-priority queue (refreshed every 64 steps of a slot)
  int hi_slots;                 // how many slots are served from the high-priority queue
  int hi_ctas;                  // CTAs reserved for the high-priority queue (blockIdx < hi_ctas pop ring 0 only, the rest ring 1 only)
  int hi_crit;                  // 1: the reserved CTAs serve the rooms flagged critical by the speculative window instead of the largest rooms
  int tune;                     // bit 0: split branch tiles over CTAs when the backlog is short; bit 1: publish head tiles with the projection blocks
  // pooled-projection servers (n_servers = 16 or 0): the first n_servers CTAs keep one 32-column slice of a head's pooled
  // weights W0[:1024] in shared memory for the whole run and answer one request per (slot, grow step)
  int n_servers;
  unsigned long long* greq_ring;  // broadcast ring: every server reads every entry; entry = (generation << 32) | seq << 13 | slot
  uint2* gproj_tagged;            // (n_slots, 2 heads, 256) {value bits, tag}: what the servers write and the head tiles poll -- a
                                  // value is valid when its tag = proj_tag(seq of the slot's current forward); one 8-byte store
                                  // per value, so the servers need neither a fence nor a counter to publish
  unsigned greq_mask;             // capacity - 1
  unsigned* greq_tail;
  SlotSync* sync;               // (n_slots)
  unsigned* progress;           // items retired so far by all CTAs (what the watchdog of the spinning CTAs looks at)
  int* abort;                   // raised by a spinner that saw no item retire anywhere for 20 s: every spinner leaves, the host reports LRG_E_STATE
  unsigned long long* busy_ns;  // [24]: [type] = summed handler time in ns, [8 + type] = items handled, [16 + type] = summed queue delay; may be NULL
};

int grow_configure();
int launch_grow(const GrowArgs& ga, int n_ctas, bool f16, cudaStream_t stream);   // f16: 3xFP16 tensor tiles, else 3xTF32

}  // namespace lrg
