// On-device region-grow driver: data structures shared by lrg_driver.cu and lrg_engine.cu (internal).
#pragma once
#include "lrg_common.cuh"

namespace lrg {

constexpr int kStepThreads = 1024;
constexpr int kMaxTilePts = 512;       // NUM_INLIER_POINT / NUM_NEIGHBOR_POINT upper bound (test_region_grow.py:22-23)
enum { STOP_NONE = 0, STOP_NONEIGHBOR = 1, STOP_NOEXPAND = 2, STOP_STUCK = 3, STOP_MAXSTEPS = 4, STOP_EMPTY = 5 };

// Per-slot grow state (one slot = one room in flight).  `active` is read by the forward kernels.
struct SlotState {
  int active;          // a forward is pending: tiles are valid, logits will be applied by the next step kernel
  int finished;        // no more rooms for this slot
  int room;            // room index or -1
  int cursor;          // next position in the room's seed order (test_region_grow.py:186)
  int seed;
  int minD[3], maxD[3], seqMin[3], seqMax[3];   // :200-203
  int stuck, steps, total_steps;
  int n_in, n_nb;
  int cluster_id;      // :177
  int regions, stops[4];
  int visited;         // points of the current room already assigned to a finished region (scheduling hint)
  float center[16];    // :241
  // random-restart driver only (lanes > 1): this lane has finished the current seed and waits for the other restarts /
  // the lane was handed a fresh seed by the lane that committed the previous one (skip the seed search)
  int parked, begin;
  // speculative lanes only (DriverArgs::spec): the order ticket of the region this lane grows, the length of the group's commit
  // log when the region began (what was committed later must stay outside the boxes it looked at), and -- once the region
  // has stopped and waits for its turn to commit -- why it stopped (0: nothing waits)
  int ticket, log_begin, fin_reason;
};

// Random-restart driver (test_random_restart.py): the NUM_RESTARTS restarts of a seed are `lanes` consecutive slots (a
// group) that grow side by side from the same visited state, each on its own copy of the room's state words.  What the
// reference keeps per room lives here, owned by whichever lane commits a seed (the last one to finish it).
//
// Beam-search driver (test_beam_search.py): a group is BEAM_WIDTH x SEARCH_WIDTH lanes; lane q * SEARCH_WIDTH + s expands
// candidate q of the seed's queue Q for the s-th time.  A ROUND = every candidate expanded SEARCH_WIDTH times (one grow step
// per lane); the lane that reports last builds the next Q (the BEAM_WIDTH largest updated masks, stable) as index lists in
// the group's parent buffers, runs the stuck logic on Q[0] and starts the next round or commits Q[0] / the previous Q[0].
constexpr int kMaxLanes = 16;
struct LaneGroup {
  int done;                // lanes that have finished the current seed (beam search: reported this round)
  int score[kMaxLanes];    // 'np' score of every lane: points in its final region (test_random_restart.py:174); beam search:
                           // size of the expanded mask, -1 if the expansion added nothing (test_beam_search.py:262-267)
  int room, cursor, cluster_id, regions, visited;
  // beam search only
  int expect;              // lanes that report this round = candidates in Q x SEARCH_WIDTH
  int nQ;                  // candidates in Q (0: no seed in flight)
  int round;               // rounds completed for this seed = the Philox step coordinate of the round's expansions
  int stuck, seed;         // test_beam_search.py:161,180-186
  int seqMin[3], seqMax[3];
  int par_n[kMaxLanes];    // candidate q of Q: points, bounding box (its index list lives in DriverArgs::parI)
  int par_min[kMaxLanes][3], par_max[kMaxLanes][3];
  float fscore[kMaxLanes];    // 'ml' scoring (DriverArgs::score_ml): log-probability score of every lane's expansion (:264)
  float par_score[kMaxLanes]; // ... and of candidate q of Q (0 for the seed alone, :164)
  // speculative lanes only: written by the lane that holds the head ticket (the commit critical section)
  int commit_seq;          // ticket that commits next (published to the other lanes through SpecSync::commit_seq)
  int next_ticket;         // tickets handed out so far = seeds issued in curvature order
  int lane_ticket[kMaxLanes];   // ticket every lane holds, -1: the lane is idle
  int log_n;               // entries in the group's commit log (DriverArgs::clog)
  int useful_steps, wasted_steps, restarts, dropped;   // grow steps of committed regions / of discarded attempts
  int stops[4];
};

// Speculative lanes (test_region_grow.py:183-217 with intra-room parallelism): the `lanes` slots of a group grow
// the next unvisited seeds of ONE room side by side, each on its own copy of the state words (private CURRENT flags, VISITED
// set by every commit in every copy).  Seeds are handed out in curvature order with increasing tickets; regions COMMIT
// strictly in ticket order.  A region that stops waits until its ticket is the head, then validates itself against the log
// of points committed since it began: a committed point inside the envelope it looked at (seqMin-1 .. seqMax+1) means it grew
// on a stale visited set -> it is grown again, now as the head (exactly the sequential state); its own seed among them ->
// dropped (the sequential driver would have skipped it).  What the other lanes of a group poll lives here.
struct SpecSync {
  int commit_seq;          // ticket that may commit now
  int fin[kMaxLanes];      // 1: the lane's region has stopped and its state is at rest (claimed with a CAS by whoever makes it the head)
};

struct DriverArgs {
  int n_rooms;
  const long long* room_off;    // (n_rooms+1)
  const float* pts;             // (T,16) padded feature rows
  unsigned* pw;                 // per-point state words: 10+10+10 bits of room-relative voxel coordinates (:175), bit 30 CURRENT, bit 31 VISITED
  const long long* pw_off;      // (n_rooms+1) word offset of every room in pw (rooms padded to a multiple of 4 words)
  const int4* room_vmin;        // (n_rooms) voxel coordinates the words are relative to
  // spatial index of every room (static; built at upload by lrg_spatial_index_kernel): the room's points in Morton order of
  // their voxels, cut into blocks of kSpBlock; NULL = off (every shell scan reads all N state words)
  const long long* sp_off;      // (n_rooms+1) offset of every room in sp_perm / sp_vox (rooms padded to whole blocks)
  const int* sp_perm;           // room-local point index of the m-th point in Morton order (padding: 0)
  const unsigned* sp_vox;       // its packed voxel coordinates (padding: 0x3FFFFFFF, outside every shell)
  const uint2* sp_box;          // (sp_off / kSpBlock) per block: packed minimum / maximum voxel coordinates
  int tune_step;                // A/B switches of the step: bit 0 = never run the median beside the indexed shell scan
  int* label;                   // (T) cluster_label (:176)
  const int* order;             // (T) room-local seed order (:183)
  SlotState* slots;
  int n_slots;
  int maxN;
  int* listI;                   // (n_slots, maxN) ascending indices of the current region
  int* listJ;                   // (n_slots, maxN) ascending indices of the neighbour shell
  unsigned* keyI;               // (n_slots, maxN) sampling keys
  unsigned* keyJ;
  float* tile[2];               // [0] inlier (n_slots, Ni, 16), [1] neighbor (n_slots, Nj, 16): rows padded to 16 floats
  int* tileidx[2];              // (n_slots, 512) source point of every tile row
  int* tilesrc[2];              // (n_slots, 512) tile row whose logits row r uses: r itself, or the row it duplicates
  const float* logits[2];       // [0] remove_output (n_slots, Ni, 2), [1] add_output (n_slots, Nj, 2)
  float* pooled;                // (n_slots, pooled_per_slot) zeroed here for the next forward
  int pooled_per_slot;
  int Ni, Nj, F;
  float resolution;
  int cluster_threshold;
  unsigned long long seed;
  int max_steps;
  int room_id_base;
  int* next_room;
  // The order in which rooms are started (NULL: by index): largest first, so that the run does not end with a long room that was
  // started late -- a room's label set does not depend on when it runs.  pending_pts[k] = points of the rooms order[k..]).
  const int* room_order;        // (n_rooms)
  const long long* pending_pts; // (n_rooms + 1)
  int* finished_slots;
  volatile int* done_flag;      // mapped pinned host memory
  LrgRoomStats* stats;          // (n_rooms)
  LrgStepTrace* trace;          // (n_rooms, trace_capacity) or NULL
  int trace_capacity;
  unsigned long long* dbg;      // diagnostics (NULL = off): summed clock64 cycles per step stage [0..14], steps in [15]
  // random restarts (lanes > 1): slot = group * lanes + lane; lane l works on pw + l * pw_lane_stride
  int lanes;
  LaneGroup* groups;            // (n_slots / lanes)
  long long pw_lane_stride;
  int* lane_steps;              // (n_rooms, lanes) out: grow steps every lane took in the room (trace lengths)
  // beam search (beam_width > 0): lanes = beam_width * search_width
  int beam_width, search_width;
  int score_ml;                 // beam search: rank candidates by accumulated log-probability instead of size (LRG_FLAG_SCORE_ML)
  int* parI;                    // (n_slots / lanes, beam_width, maxN) ascending index lists of the candidates in Q
  // speculative lanes (spec != 0, lanes > 1, no restarts / beam)
  int spec;
  SpecSync* spec_sync;          // (n_slots / lanes)
  int* clog;                    // (n_slots / lanes, maxN) commit log: room-local indices of committed points, in commit order
  const unsigned* q_ctr;        // persistent kernel: [0] head, [1] tail of the work queue (load hint for the window), or NULL
  // which rooms speculate (a seed goes to a SECOND lane of a room only then): the spec_top groups with the largest estimate
  // in spec_est (grow steps the group's room still needs; every owner refreshes its own entry), or anybody while at least
  // spec_min_idle CTAs of the persistent kernel wait for work
  int spec_top;
  int spec_min_idle;
  int spec_crit;                // ... and only while its estimate x spec_crit >= the estimate of all remaining work (0: always)
  long long total_pts;          // points of all rooms (the rooms not yet started enter the estimate with 0.2 grow steps per point)
  int* spec_est;                // (n_slots / lanes)
};

struct FillArgs {
  int n_rooms;
  const long long* room_off;
  const float* pts;
  const int* label;
  int* label_filled;
  int* lab_list;                // (T) per room: ascending indices of labelled points
  int* unl_list;                // (T) per room: ascending indices of unlabeled points
  int* n_lab;                   // (n_rooms)
  int* n_unl;
  int F;
  // spatial index of the rooms (DriverArgs::sp_*; sp_perm == NULL: every unlabeled point scans every labelled point)
  const long long* sp_off;
  const int* sp_perm;
  const uint2* sp_box;
  const int4* room_vmin;
  float resolution;
};

constexpr int kSpBlock = 128;          // points per block of the spatial index (one 128-bit load per lane of a warp)
constexpr int kSpMaxN = 262144;        // rooms up to this many points use the index (bitmap of the hits in shared memory)
constexpr int kSpMinN = 2048;          // ... and smaller rooms are not worth the extra round trip
// Morton order + block boxes of every room; d_keys: (sum of P_r) sort scratch, P_r = power of two >= N_r at d_key_off[r]
int launch_spatial_index(int n_rooms, const long long* d_room_off, const long long* d_pw_off, const unsigned* d_pw, const long long* d_sp_off,
                         const long long* d_key_off, unsigned long long* d_keys, long long max_keys, int* d_sp_perm, unsigned* d_sp_vox,
                         uint2* d_sp_box, cudaStream_t stream, int* n_launches = nullptr);
// feature rows -> padded 16-float rows + state words; *d_err is set to a room index + 1 if a room spans > 1022 voxels
int launch_pack(const float* d_points, int F, int n_rooms, const long long* d_room_off, const long long* d_pw_off, float resolution,
                float* d_pts16, unsigned* d_pw, int4* d_room_vmin, int* d_err, cudaStream_t stream);
// caller-prepared rooms: seed_order a permutation per room, one point per voxel; d_scratch = 3 * total_words zeroed words;
// *d_err receives (room + 1) | kind << 28 (1 = bad seed order, 2 = two points in one voxel)
int launch_validate_rooms(int n_rooms, const long long* d_room_off, const long long* d_pw_off, const unsigned* d_pw, const int* d_order,
                          unsigned* d_scratch, int* d_err, cudaStream_t stream);
// clears the CURRENT / VISITED flags of every room (start of a run)
// (lanes > 1: also re-creates the copies 1..lanes-1 of the words, lane_stride words apart)
int launch_reset_words(int n_rooms, const long long* d_room_off, const long long* d_pw_off, unsigned* d_pw, int lanes, long long lane_stride,
                       cudaStream_t stream);
size_t step_smem_bytes();
int launch_step(const DriverArgs& da, cudaStream_t stream);
int launch_fill(const FillArgs& fa, cudaStream_t stream);

}  // namespace lrg
