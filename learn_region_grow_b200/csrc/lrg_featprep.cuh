// Device feature preparation (test_region_grow.py:119-173): launch arguments (internal).
#pragma once
#include "lrg_common.cuh"

namespace lrg {

struct FeatPrepArgs {
  int n_rooms;
  int C;                          // floats per raw point (x y z r g b [...]), >= 6
  int F;                          // feature columns written: 6 (xyz, room xyz), 9 (+rgb), 12 (+normal), 13 (+curvature)
  float res;
  const long long* raw_off;       // (R+1) raw point offsets
  const float* raw;               // (sum Nr, C)
  const long long* sort_off;      // (R+1) offsets of the per-room sort buffers (power-of-two sizes >= Nr)
  long long max_sort;             // host: the largest sort buffer (phase 1 sorts whole buffers)
  long long max_order_sort;       // host: the largest seed-order sort of phase 2 = max over rooms of min(buffer, pow2ceil(max(2, Neq)))
  unsigned long long* keys;       // (sum P) voxel sort keys; later the seed-order keys
  unsigned long long* keys2;      // (sum P) first-seen sort keys; later the seed-order indices
  int4* raw_vmin;                 // (R) voxel origin of every room
  int* n_eq;                      // (R) out: voxels (= equalised points) per room
  int* err;                       // out: room + 1 whose voxel span exceeds 10 bits
  // per voxel in key order, indexed raw_off[r] + u (a room has at most Nr voxels)
  unsigned* uniq_vox;             // packed 10+10+10-bit voxel coordinates
  int* uniq_start;                // start of the voxel's run in the sorted keys
  int* eq_of_uniq;                // first-seen position of the voxel = index of its equalised point
  double* sums;                   // (.., 10): n, sum x y z, sum xx xy xz yy yz zz
  int* raw_rank;                  // (sum Nr) key-order voxel rank of every raw point
  // phase 2
  const long long* eq_off;        // (R+1) equalised point offsets
  float* feat;                    // (sum Neq, F) dense feature rows
  double* curv;                   // (sum Neq) curvature (normalised by the room maximum on return)
  int* order;                     // (sum Neq) seed order
  int* equalized_idx;             // (sum Neq) raw index of every equalised point (room-local)
  int* unequalized_idx;           // (sum Nr) equalised index of every raw point (room-local)
  float* extent;                  // (R, 6) min / max xyz of the equalised points
  unsigned long long* cmax;       // (R) bit pattern of the room's largest curvature
  int* has_nan;                   // (R) a curvature of the room is NaN
};

// n_launches (optional) is incremented by the kernels launched
int launch_featprep_phase1(const FeatPrepArgs& a, cudaStream_t stream, int* n_launches = nullptr);
int launch_featprep_phase2(const FeatPrepArgs& a, cudaStream_t stream, int* n_launches = nullptr);
int launch_labels_raw(int n_rooms, const long long* raw_off, const long long* eq_off, const int* unequalized_idx, const int* label,
                      int* out, cudaStream_t stream);

}  // namespace lrg
