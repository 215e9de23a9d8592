"""Beam-search driver (test_beam_search.py) on the bench workload: time per pass and steps/s for a few beam shapes, next to
the plain driver, with the segmentation statistics (the reference's reason for the local search)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from learn_region_grow_b200.engine import Engine

rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 68
raw_off, raw = bench.make_workload(rooms, 1000)
eng = Engine(1, 1, 512, 512, 13, 0); eng.load_weights(bench.load_weights())
eng.upload_raw_concatenated(raw_off, raw, 0.1)
obj_raw = [raw[raw_off[i]:raw_off[i + 1], 6].astype(np.int32) for i in range(rooms)]
for B, W in ((0, 0), (1, 1), (2, 2), (3, 3), (4, 4)):
    best = None
    for it in range(2):
        t0 = time.perf_counter()
        st = eng.segment_resident(resolution=0.1, seed=0, beam_width=B, search_width=W)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    m = eng.room_metrics(obj_raw, raw=True)
    steps = int(st['grow_steps'].sum())
    print('beam %d x %d: %7.1f ms/pass  %8d grow steps  %7.0f k steps/s  %6.2f M raw points/s | NMI %.3f AMI %.3f ARS %.3f PRC %.3f RCL %.3f IOU %.3f' %
          (B, W, 1e3 * best, steps, steps / best / 1e3, raw_off[-1] / best / 1e6, m['nmi'].mean(), m['ami'].mean(), m['ars'].mean(),
           np.nanmean(m['prc']), m['rcl'].mean(), m['iou'].mean()), flush=True)
