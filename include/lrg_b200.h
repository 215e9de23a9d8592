/*
 * lrg_b200.h -- C ABI of the B200-native LRGNet grow engine (liblrg_b200.so).
 *
 * Every entry point returns 0 on success and a negative LRG_E_* code on failure (never throws across the
 * ABI); lrg_last_error() returns a human-readable message for the calling thread.  Pointers named d_* are
 * device pointers, everything else is host memory.  Kernels are launched on the engine's own stream unless a
 * cudaStream_t argument is given.
 *
 * Each declaration cites the reference interface it replaces (paths relative to jingdao/learn_region_grow).
 */
#ifndef LRG_B200_H_
#define LRG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LrgEngine LrgEngine;
typedef void* lrg_stream_t;            /* cudaStream_t */

enum {
  LRG_OK = 0,
  LRG_E_INVALID = -1,                  /* bad argument (the reference raises InvalidArgument via OP_REQUIRES) */
  LRG_E_CUDA = -2,                     /* a CUDA runtime call or kernel launch failed */
  LRG_E_STATE = -3,                    /* call order violated (e.g. forward before load_weights) */
  LRG_E_NOMEM = -4,
  LRG_E_RANGE = -5                     /* LRG_FORWARD_TENSOR_F16 only: an activation left the fp16 range */
};

const char* lrg_last_error(void);
int lrg_version(void);
int lrg_device_count(void);

/* ------------------------------------------------------------------------------------------------------
 * LrgNet forward (learn_region_grow_util.py:76-162)
 * ------------------------------------------------------------------------------------------------------ */

/* Replaces LrgNet.__init__(batch_size, seq_len, num_inlier_points, num_neighbor_points, feature_size, lite)
 * (learn_region_grow_util.py:76-103): lite 0 -> conv 64,64,64,128,512 / head 256,128; 1 -> 64,64 / 64;
 * 2 -> 64,64,256 / 64,64 (:77-85).  max_batch = batch_size*seq_len tiles per forward call. */
int lrg_engine_create(LrgEngine** out, int device, int feature_size, int num_inlier_points,
                      int num_neighbor_points, int lite, int max_batch);
int lrg_engine_destroy(LrgEngine* e);

/* Number of float32 parameters the engine expects, in graph-construction order (util.py:106-162):
 * for prefix in (lrg_, lrg_neighbor_): kernel_i [Cin,Cout] row-major, bias_i ...; then for prefix in
 * (lrg_add_, lrg_remove_): kernel_i, bias_i.  This is the order tf.train.Saver enumerates them. */
size_t lrg_engine_weight_count(const LrgEngine* e);
/* Replaces tf.train.Saver().restore(sess, path) (test_region_grow.py:92-93): the host side reads the
 * checkpoint-V2 files and hands over one flat float32 blob in the order above. */
int lrg_engine_load_weights(LrgEngine* e, const float* blob, size_t n_floats);

/* Which kernels evaluate the network.  AUTO = tensor cores when the model is the full one (lite=0), FMA otherwise.
 *   LRG_FORWARD_TENSOR_F16  tcgen05 tensor cores, every contraction as a three-term split product of fp16 halves
 *                           (x = hi + lo, D += hi.hi + lo.hi + hi.lo, fp32 accumulation; weights pre-scaled per layer by a
 *                           power of two): ~22 mantissa bits per product at half the MMAs and half the weight bytes of
 *                           3xTF32.  Valid while every activation stays below 65,000 (the shipped model: < 500); a
 *                           synchronous call that sees a larger one fails with LRG_E_RANGE.
 *   LRG_FORWARD_TENSOR      the same with tf32 halves (3xTF32): no range limit
 *   LRG_FORWARD_AUTO        3xFP16; a synchronous call (lrg_forward_host, lrg_segment_resident, lrg_segment_rooms_host) that
 *                           sees an activation beyond the fp16 range is repeated with 3xTF32 before it returns
 *   LRG_FORWARD_FMA         fp32 FMA pipe (all lite variants; A/B reference for the tensor paths)
 * No reference equivalent (TF picks its own conv kernels). */
enum { LRG_FORWARD_AUTO = 0, LRG_FORWARD_FMA = 1, LRG_FORWARD_TENSOR = 2, LRG_FORWARD_TENSOR_F16 = 3 };
int lrg_engine_set_forward_mode(LrgEngine* e, int mode);
/* The mode forward calls currently take (LRG_FORWARD_FMA, LRG_FORWARD_TENSOR or LRG_FORWARD_TENSOR_F16; valid after load_weights). */
int lrg_engine_forward_mode(const LrgEngine* e);
/* For callers of the asynchronous lrg_forward_device in 3xFP16: synchronises the device, reports (and clears) whether a tile
 * saw an activation beyond the fp16 range since the last check, and how many synchronous calls were repeated with 3xTF32 so
 * far.  Either pointer may be NULL. */
int lrg_engine_range_overflow(LrgEngine* e, int* overflow, int* fallbacks);

/* Replaces sess.run([net.add_output, net.remove_output], {inlier_pl, neighbor_pl}) (test_region_grow.py:257-258).
 * inlier (B, Ni, F), neighbor (B, Nj, F) float32 row-major; add_out (B, Nj, 2), remove_out (B, Ni, 2).
 * _host: synchronous, host buffers, H2D/D2H inside.  _device: asynchronous on `stream` (0 = engine stream). */
int lrg_forward_host(LrgEngine* e, int B, const float* inlier, const float* neighbor, float* add_out,
                     float* remove_out);
int lrg_forward_device(LrgEngine* e, int B, const float* d_inlier, const float* d_neighbor, float* d_add_out,
                       float* d_remove_out, lrg_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * On-device region-grow driver (test_region_grow.py:175-316)
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LrgGrowParams {
  float resolution;            /* test_region_grow.py:30   (0.1; 0.3 for Semantic-KITTI) */
  int cluster_threshold;       /* :33  regions with more points than this get a label (10) */
  uint64_t seed;               /* Philox key; the reference's single MT19937 stream (:21) cannot be sharded */
  int max_slots;               /* rooms in flight per GPU (0 = default) */
  int max_steps_per_region;    /* 0 = unbounded like the reference */
  int room_id_base;            /* added to the local room index in the RNG counter (multi-GPU sharding) */
  int trace_capacity;          /* >0: record up to this many grow steps per room (tests) */
  int flags;                   /* LRG_FLAG_* */
  int num_restarts;            /* 0/1: test_region_grow.py.  N > 1: test_random_restart.py with NUM_RESTARTS = N (:24) and
                                  'np' scoring (:40,174): every seed is grown N times from the same visited state, the N
                                  restarts side by side on the device, and the largest result is kept (max 16) */
  int beam_width;              /* > 0: test_beam_search.py with BEAM_WIDTH (:24) = beam_width, SEARCH_WIDTH (:25) = search_width */
  int search_width;            /* and 'np' scoring (:41,266): per seed a queue of at most beam_width candidate masks, each expanded
                                  search_width times per round by one sampled grow step, the beam_width largest updated masks
                                  kept (:273); the beam_width * search_width expansions of a round run side by side on the
                                  device (product max 16).  Excludes num_restarts > 1. */
  int spec_lanes;              /* test_region_grow.py only (no restarts / beam): > 1 grows up to spec_lanes regions of ONE room side
                                  by side, speculatively, and commits them strictly in seed order (:183-217) -- a region grown on a
                                  stale visited set is detected when it reaches the head of the order and grown again there, so the
                                  labels are bit-identical to spec_lanes = 1.  0 = engine default (in the
                                  persistent kernel when no trace is recorded: 4, or 8 when the rooms average >= 65,536 points; else
                                  1), 1 = off, max 16 */
  int spec_top;                /* speculative lanes: only the spec_top rooms in flight with the most estimated work left (unvisited
                                  points x grow steps per visited point so far) hand seeds to more than one lane -- the run ends
                                  with its longest rooms, speculation elsewhere only costs SM time.  A cap on top of spec_crit below.
                                  0 = engine default (8), < 0 = no cap */
  int spec_min_idle;           /* speculative lanes: a room outside the spec_top still speculates while at least this many CTAs of
                                  the persistent kernel wait for work (the tail of a run).  0 = engine default (96), < 0 = never */
  int spec_crit;               /* speculative lanes: a room among the spec_top speculates only while it is CRITICAL -- its estimated
                                  remaining grow steps x spec_crit >= the estimated remaining grow steps of the whole run (rooms in
                                  flight + rooms not started): a chain link costs ~50 us of latency but a grow step only ~1 us of the
                                  machine's time (130 us of SM time over 132 SMs), so a room whose share of the remaining work is
                                  below ~1/50 finishes inside the throughput bound anyway and speculation would only add discarded
                                  steps to it.  0 = engine default (40), < 0 = every room counts as critical */
} LrgGrowParams;

enum {
  LRG_FLAG_KERNEL_TIMING = 1,  /* time the forward kernels separately with CUDA events (no graph; slower) */
  LRG_FLAG_NO_GRAPH = 2,       /* lock-step loop with direct kernel launches instead of a CUDA graph */
  LRG_FLAG_PRIORITY = 8,       /* persistent kernel: reserve CTAs for the rooms the run ends with -- with speculative lanes 16 CTAs for
                                  the rooms the window calls critical (spec_crit), with one lane 24 CTAs for the two slots with the
                                  most unvisited points.  Off by default: it shortens a run only when a GPU holds few enough rooms to
                                  be chain-bound yet enough of them to queue (34 rooms: 7 % faster) and costs throughput otherwise */
  LRG_FLAG_LOCKSTEP = 4,       /* lock-step loop (one {step, branch, gproj, head} kernel quartet per iteration, CUDA graph)
                                  instead of the persistent grow kernel; implied by the two flags above and by FMA mode */
  /* A/B switches of the persistent kernel (the defaults are the measured best, profiles/README.md; results do not depend on
   * any of them, tests/test_driver_gpu.py): */
  LRG_FLAG_NO_PROJ_SERVERS = 16,   /* no pooled-projection server CTAs: 8 projection work items per grow step instead */
  LRG_FLAG_NO_TILE_SPLIT = 32,     /* never split a branch tile over idle CTAs */
  LRG_FLAG_NO_SPATIAL_INDEX = 256, /* every neighbour-shell scan reads all N state words of the room instead of the blocks of the
                                      room's Morton-ordered spatial index that meet the shell, and the fill compares every unlabeled
                                      point with every labelled point instead of pruning blocks by their distance bound */
  LRG_FLAG_ROOMS_IN_ORDER = 1024,  /* rooms are started in index order instead of largest first (a run that holds more rooms than
                                      slots then tends to end with a long room that was started late) */
  LRG_FLAG_NO_STEP_OVERLAP = 512,  /* the median of a small region is selected after the indexed shell scan instead of beside it */
  LRG_FLAG_HEADS_AFTER_PROJ = 64,  /* publish the head tiles when the projection is complete instead of together with it
                                      (only without projection servers) */
  /* beam search only: 'ml' scoring (test_beam_search.py:46-47,238-256,263-264) -- a candidate's score is its parent's score plus
   * the log-probability of the step under the network's confidences: over all padded tile rows of each set log(conf) where
   * the row's re-rounded voxel is in the sampled add / remove set, log(1 - conf) elsewhere, each / NUM_NEIGHBOR_POINT,
   * accumulated row by row in float32 like the reference's numpy scalars.  Without the flag: 'np' (:41,266). */
  LRG_FLAG_SCORE_ML = 128
};

typedef struct LrgRoomStats {
  int32_t n_points;
  int32_t grow_steps;          /* Session.run calls the reference would have made */
  int32_t regions;             /* seeds grown (labelled or not) */
  int32_t clusters;            /* labelled regions (cluster_id - 1) */
  int32_t stop_noneighbor, stop_noexpand, stop_stuck, stop_other;
  /* speculative lanes (spec_lanes > 1) only, else 0: grow steps of discarded attempts (not part of grow_steps), regions grown
   * again at the head of the order, attempts dropped because an earlier region swallowed their seed */
  int32_t spec_wasted_steps, spec_restarts, spec_dropped, reserved;
} LrgRoomStats;

/* One record per grow step when tracing (parity tests): everything needed to re-drive the CPU oracle. */
typedef struct LrgStepTrace {
  int32_t seed_point, step_in_region, n_inlier, n_neighbor;
  int32_t stop_reason;         /* 0 continue, 1 noneighbor(unused here), 2 noexpand, 3 stuck, 4 maxsteps, 5 empty */
  int32_t size_after;
  float center[16];
  uint32_t add_mask[16];       /* bit r of word r/32: tile row r sampled True */
  uint32_t remove_mask[16];
  uint32_t inlier_idx_crc, neighbor_idx_crc;   /* sum-of-products checksums of the sampled point indices */
  float log_prob[2];           /* LRG_FLAG_SCORE_ML: addLogProb, rmvLogProb of this expansion (test_beam_search.py:238-256) */
  float score;                 /* LRG_FLAG_SCORE_ML: parent's score + addLogProb + rmvLogProb (:264) */
  int32_t reserved;            /* zero */
} LrgStepTrace;

/* Upload rooms: points (sum N, F) float32 rows = the 13-D features of test_region_grow.py:165-172,
 * room_offsets (n_rooms+1) prefix sums of N_r, seed_order (sum N) = argsort(curvatures) per room (:183),
 * room-local indices.  Replaces the numpy arrays the driver keeps per room (:175-183).
 * Preconditions, checked on the device (LRG_E_INVALID): every room holds at most ONE point per voxel at `resolution` -- the
 * reference applies its masks by voxel (:282-287), the device by point, which is the same thing only after the reference's
 * equalisation (:125-136; lrg_rooms_upload_raw does it) -- and seed_order is a permutation of 0..N_r-1 per room. */
int lrg_rooms_upload(LrgEngine* e, int n_rooms, const int64_t* room_offsets, const float* points,
                     const int32_t* seed_order, float resolution);
/* Upload rooms as RAW points and prepare the features on the device: replaces test_region_grow.py:119-173 (equalise to one
 * point per voxel in first-seen order, room-normalised coordinates, normal + curvature from the covariance of the raw
 * points in the 27 surrounding voxels, curvature / max, seed order = argsort(curvatures)).  raw_points: (sum Nr, n_cols)
 * float32 rows x y z r g b [...] as stored in the reference's H5 files (learn_region_grow_util.py:13-20); the engine's
 * feature_size selects the columns like the driver's ablation switches (6: xyz + room xyz, 9: + rgb, 12: + normal, 13: +
 * curvature).  A room may hold at most 1,048,575 raw points. */
int lrg_rooms_upload_raw(LrgEngine* e, int n_rooms, const int64_t* raw_offsets, const float* raw_points, int n_cols,
                         float resolution);
/* Same with the raw points already in device memory (caller-owned, read only; raw_offsets is host memory). */
int lrg_rooms_upload_raw_device(LrgEngine* e, int n_rooms, const int64_t* raw_offsets, const float* d_raw_points, int n_cols,
                                float resolution);
/* Device time (ms, CUDA events on the engine stream) of the last lrg_rooms_upload_raw / _device call: host-to-device copy (host
 * variant), feature preparation and packing, including the one host round trip that sizes the equalised rooms. */
int lrg_last_prepare_ms(LrgEngine* e, float* ms);
/* Kernels launched by the last upload (feature preparation with its sort passes, packing, spatial index). */
int lrg_last_prepare_launches(LrgEngine* e, int* n);
/* After an upload: (n_rooms+1) prefix sums of the equalised room sizes. */
int lrg_rooms_equalized_offsets(LrgEngine* e, int64_t* eq_offsets);
/* After lrg_rooms_upload_raw: the prepared features (sum Neq, F), the seed order, equalized_idx (sum Neq: raw index of every
 * equalised point, room-local) and unequalized_idx (sum Nr: equalised index of every raw point, :130); any pointer may be NULL. */
int lrg_rooms_features_download(LrgEngine* e, float* points, int32_t* seed_order, int32_t* equalized_idx,
                                int32_t* unequalized_idx);
/* cluster_label[unequalized_idx] (:366): the labels of lrg_labels_download mapped back to the raw points (sum Nr). */
int lrg_labels_download_raw(LrgEngine* e, int32_t* labels_raw, int filled);

/* Grow every uploaded room to completion on the device (no host round trip per step), then fill unlabeled
 * points (:308-316).  stats may be NULL or n_rooms entries.  params->resolution must be 0 or the resolution of the upload.
 * One call at a time per engine handle: a second call that arrives while one is running returns LRG_E_STATE (use one
 * engine per thread / stream); the persistent kernel needs the whole device -- a launch that cannot be co-resident
 * returns LRG_E_STATE instead of spinning. */
int lrg_segment_resident(LrgEngine* e, const LrgGrowParams* params, LrgRoomStats* stats);
/* labels (sum N) int32; filled != 0 returns the labels after the nearest-neighbour fill (:308-316),
 * otherwise the raw cluster_label with 0 = unlabeled (:176,214). */
int lrg_labels_download(LrgEngine* e, int32_t* labels, int filled);
int lrg_trace_download(LrgEngine* e, int room, LrgStepTrace* out, int capacity, int* n_steps);
/* After a run with num_restarts > 1: the steps restart lane `lane` took in `room` (all seeds, in order). */
int lrg_trace_download_lane(LrgEngine* e, int room, int lane, LrgStepTrace* out, int capacity, int* n_steps);
/* Host-buffer convenience: upload + segment + download (the end-to-end call bench.py times). */
int lrg_segment_rooms_host(LrgEngine* e, int n_rooms, const int64_t* room_offsets, const float* points,
                           const int32_t* seed_order, const LrgGrowParams* params, int32_t* labels_filled,
                           LrgRoomStats* stats);
/* Device time (ms, CUDA events on the engine stream) of the last lrg_segment_resident call, the number of
 * lock-step iterations it ran and the kernels it launched. */
int lrg_last_segment_profile(LrgEngine* e, float* grow_ms, float* fill_ms, int64_t* iterations,
                             int64_t* kernel_launches, float* forward_ms);

/* Persistent grow kernel (default when the tensor-core forward is in use): whether the last lrg_segment_resident call
 * ran as one persistent launch, and per work-item type (step, branch tile, pooled-projection block, head tile) the summed
 * handler time over all CTAs in ms and the number of items handled. */
int lrg_last_grow_profile(LrgEngine* e, int* persistent, double busy_ms[4], int64_t items[4]);
/* (With the projection servers -- the default -- no projection items exist: items[2] = 0 and the servers' time is not counted.) */
/* Summed time (ms) the items of each type waited in the device queue between publication and pick-up (same order). */
int lrg_last_grow_queue_delay(LrgEngine* e, double delay_ms[4]);

/* Diagnostics: after lrg_engine_set_tile_timing(e, 1) the tensor tiles and the driver step add the SM cycles of each
 * of their stages to counters (out[0..13] branch tile stages, out[15] branch tiles; out[16..23] head tile stages,
 * out[31] head tiles; out[32..41] driver step stages, out[47] steps, out[48..59] median cycles / steps by set-size bucket);
 * tools/grow_profile.py prints them. */
int lrg_engine_set_tile_timing(LrgEngine* e, int on);
int lrg_tile_timing(LrgEngine* e, uint64_t out[64], int reset);

/* With LRG_FLAG_KERNEL_TIMING: summed CUDA-event durations (ms) of the four kernels of the lock-step loop over the
 * last lrg_segment_resident call: out[0] step (driver), out[1] branch MLPs, out[2] pooled projection, out[3] heads. */
int lrg_last_kernel_times(LrgEngine* e, float out_ms[4]);
/* Device pointer of the (sum N) int32 label array (filled != 0: after the fill), valid until the next upload;
 * lets the host hand the labels to NCCL without a copy. */
int lrg_labels_device_ptr(LrgEngine* e, int filled, void** d_ptr);

/* ------------------------------------------------------------------------------------------------------
 * Per-room segmentation statistics (test_region_grow.py:319-349)
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LrgRoomMetrics {
  double nmi;                  /* sklearn normalized_mutual_info_score(obj_id, cluster_label)   (:346) */
  double ami;                  /* sklearn adjusted_mutual_info_score                            (:347) */
  double ars;                  /* sklearn adjusted_rand_score                                   (:348) */
  double prc;                  /* numpy.mean(dt_match): matched clusters / cluster_label.max()  (:342) */
  double rcl;                  /* gt_match / len(set(obj_id))                                   (:343) */
  double iou;                  /* mean over objects of the best IoU with an unmatched cluster   (:344) */
  int32_t n_points;
  int32_t n_classes;           /* len(set(obj_id)) */
  int32_t n_clusters;          /* cluster_label.max() */
  int32_t gt_match;            /* objects matched to a cluster with IoU > 0.5 (:333-335) */
} LrgRoomMetrics;

/* Replaces the statistics block of test_region_grow.py:319-349 for n_rooms rooms stored back to back: d_obj_id and
 * d_cluster_label are DEVICE arrays of room_offsets[n_rooms] int32 (room_offsets is host memory and starts at 0).  One
 * contingency table per room is built on the device, the expected mutual information is evaluated on the device, the
 * O(classes x clusters) closed forms on the host.  Objects are matched in descending point count; equal counts in the order
 * of a stable ascending sort reversed (numpy.argsort's tie order at :328 is unspecified).  d_cluster_label2 (optional,
 * device, same length) receives cluster_label2 (:323,335,339-341).  An empty room yields NaNs. */
int lrg_segmentation_metrics(int n_rooms, const int64_t* room_offsets, const int32_t* d_obj_id, const int32_t* d_cluster_label,
                             LrgRoomMetrics* out, int32_t* d_cluster_label2, lrg_stream_t s);
/* Same on the rooms held by the engine after lrg_segment_resident.  obj_id is HOST memory: the ground-truth object id of
 * every equalised point (raw == 0, sum Neq values) or of every raw point (raw != 0, sum Nr values, rooms uploaded with
 * lrg_rooms_upload_raw; gathered with equalized_idx on the device like :136).  filled selects the labels after the
 * nearest-neighbour fill (what the reference scores) or before it.  cluster_label2 (optional, host, sum Neq). */
int lrg_room_metrics(LrgEngine* e, const int32_t* obj_id, int raw, int filled, LrgRoomMetrics* out, int32_t* cluster_label2);

/* ------------------------------------------------------------------------------------------------------
 * tf_ops primitives.  Same argument lists as the reference's C++ launchers plus a trailing stream; all
 * pointers are DEVICE pointers (the reference receives TF-allocated device buffers).
 * ------------------------------------------------------------------------------------------------------ */
/* tf_ops/sampling/tf_sampling_g.cu:203  farthestpointsamplingLauncher(b,n,m,inp,temp,out); temp (32,n) workspace
 * is accepted for signature parity and may be NULL (min-distances live on chip when n fits). */
int lrg_farthest_point_sampling(int b, int n, int m, const float* d_inp, float* d_temp, int* d_out, lrg_stream_t s);
/* tf_sampling_g.cu:206 gatherpointLauncher(b,n,m,inp,idx,out) */
int lrg_gather_point(int b, int n, int m, const float* d_inp, const int* d_idx, float* d_out, lrg_stream_t s);
/* train_pointnet.py:113-123 sample_and_group(npoint, radius, nsample, xyz, points) as two launches instead of four custom ops
 * and four graph ops: farthest point sampling, then gather_point of the centres (new_xyz) + ball query + group_point of xyz with
 * the translation normalisation (:117) + group_point of the features + concat (:119-120) in one kernel (a store per round or an
 * epilogue inside the sampling kernel was measured to slow its serial loop by 27 %, so the gather rides with the queries).
 * xyz (b,n,3), points (b,n,c) or NULL with c = 0; out: fps_idx (b,npoint), new_xyz (b,npoint,3), new_points
 * (b,npoint,nsample,3+c), idx (b,npoint,nsample), pts_cnt (b,npoint), grouped_xyz (b,npoint,nsample,3) or NULL.  temp: the
 * (32,n) workspace of farthestpointsamplingLauncher, only read for n > 65,536.  Results equal the separate ops bit for bit. */
int lrg_sample_and_group(int b, int n, int npoint, float radius, int nsample, int c, const float* d_xyz, const float* d_points, float* d_temp,
                         int* d_fps_idx, float* d_new_xyz, float* d_new_points, int* d_idx, int* d_pts_cnt, float* d_grouped_xyz, lrg_stream_t s);
/* tests: clouds larger than n use the thread-block-cluster FPS kernel (default 8,192 = where one CTA runs out of registers) */
int lrg_fps_set_cluster_min(int n);
/* tf_sampling_g.cu:209 scatteraddpointLauncher(b,n,m,out_g,idx,inp_g); inp_g must be zeroed (tf_sampling.cpp:174) */
int lrg_scatter_add_point(int b, int n, int m, const float* d_out_g, const int* d_idx, float* d_inp_g, lrg_stream_t s);
/* tf_sampling_g.cu:198 probsampleLauncher(b,n,m,inp_p,inp_r,temp,out); temp (b,n) workspace */
int lrg_prob_sample(int b, int n, int m, const float* d_inp_p, const float* d_inp_r, float* d_temp, int* d_out, lrg_stream_t s);
/* tf_ops/grouping/tf_grouping_g.cu:125 queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt).
 * Rows with no point in the ball are left untouched like the reference (uninitialised there). */
int lrg_query_ball_point(int b, int n, int m, float radius, int nsample, const float* d_xyz1, const float* d_xyz2,
                         int* d_idx, int* d_pts_cnt, lrg_stream_t s);
/* tf_grouping_g.cu:129 selectionSortLauncher(b,n,m,k,dist,outi,out) */
int lrg_selection_sort(int b, int n, int m, int k, const float* d_dist, int* d_outi, float* d_out, lrg_stream_t s);
/* tf_grouping.py:66-68 (knn_point's graph ops tile / subtract / square / reduce_sum): dist (b,m,n) = squared distances between
 * every xyz2 (b,m,c) query and every xyz1 (b,n,c) point; the input of selectionSortLauncher in knn_point (:70) */
int lrg_pairwise_sqdist(int b, int n, int m, int c, const float* d_xyz1, const float* d_xyz2, float* d_dist, lrg_stream_t s);
/* tf_grouping_g.cu:133 groupPointLauncher(b,n,c,m,nsample,points,idx,out) */
int lrg_group_point(int b, int n, int c, int m, int nsample, const float* d_points, const int* d_idx, float* d_out, lrg_stream_t s);
/* tf_grouping_g.cu:137 groupPointGradLauncher(b,n,c,m,nsample,grad_out,idx,grad_points); grad_points zeroed (tf_grouping.cpp:204) */
int lrg_group_point_grad(int b, int n, int c, int m, int nsample, const float* d_grad_out, const int* d_idx, float* d_grad_points, lrg_stream_t s);
/* tf_ops/3d_interpolation/tf_interpolate.cpp:60 threenn_cpu(b,n,m,xyz1,xyz2,dist,idx) -- CPU-only in the reference */
int lrg_three_nn(int b, int n, int m, const float* d_xyz1, const float* d_xyz2, float* d_dist, int* d_idx, lrg_stream_t s);
/* tf_interpolate.cpp:107 threeinterpolate_cpu(b,m,c,n,points,idx,weight,out) */
int lrg_three_interpolate(int b, int m, int c, int n, const float* d_points, const int* d_idx, const float* d_weight, float* d_out, lrg_stream_t s);
/* tf_interpolate.cpp:131 threeinterpolate_grad_cpu(b,n,c,m,grad_out,idx,weight,grad_points); grad_points zeroed (:244) */
int lrg_three_interpolate_grad(int b, int n, int c, int m, const float* d_grad_out, const int* d_idx, const float* d_weight, float* d_grad_points, lrg_stream_t s);

/* Plain device-memory helpers so a ctypes host needs no second CUDA binding. */
int lrg_malloc(void** d_ptr, size_t bytes);
int lrg_free(void* d_ptr);
int lrg_memcpy_h2d(void* d_dst, const void* src, size_t bytes);
int lrg_memcpy_d2h(void* dst, const void* d_src, size_t bytes);
int lrg_memset(void* d_ptr, int value, size_t bytes);
int lrg_device_synchronize(void);
int lrg_set_device(int device);
/* Pinned host buffers for the end-to-end path. */
int lrg_host_alloc(void** ptr, size_t bytes);
int lrg_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* LRG_B200_H_ */
