// Per-room segmentation statistics (test_region_grow.py:319-349): internal launch interface of lrg_metrics.cu.
#pragma once
#include "lrg_common.cuh"

namespace lrg {

int segmentation_metrics(int n_rooms, const int64_t* room_offsets, const int32_t* d_obj_id, const int32_t* d_label,
                         LrgRoomMetrics* out, int32_t* d_label2, cudaStream_t st);
int launch_gather_equalized(int n_rooms, const long long* d_raw_off, const long long* d_eq_off, const int* d_equalized_idx, const int* d_obj_raw,
                            int* d_obj_eq, cudaStream_t st);

}  // namespace lrg
