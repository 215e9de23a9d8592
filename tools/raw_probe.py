"""Time the raw-points path: device feature preparation (upload_raw) vs growing, per pass."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from learn_region_grow_b200 import _lib
from learn_region_grow_b200.engine import Engine
raw_off, raw = bench.make_workload(68, 1000)
eng = Engine(1, 1, 512, 512, 13, 0); eng.load_weights(bench.load_weights())
h_raw = bench.pinned_array(_lib, raw.shape, np.float32); h_raw[...] = raw
for it in range(4):
    t0 = time.perf_counter(); eng.upload_raw_concatenated(raw_off, h_raw, 0.1); t1 = time.perf_counter()
    st = eng.segment_resident(resolution=0.1, seed=0); t2 = time.perf_counter()
    lab = eng.raw_labels(True); t3 = time.perf_counter()
    print('upload_raw (feature prep) %.1f ms   segment %.1f ms   raw labels %.1f ms' % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
# statistics on the device (test_region_grow.py:319-349) against the ground-truth object ids of the raw points (column 6)
obj_raw = [raw[raw_off[i]:raw_off[i + 1], 6].astype(np.int32) for i in range(68)]
for it in range(3):
    t0 = time.perf_counter(); m = eng.room_metrics(obj_raw, raw=True); t1 = time.perf_counter()
    print('room metrics %.1f ms   mean NMI %.3f AMI %.3f ARS %.3f PRC %.3f RCL %.3f IOU %.3f' %
          (1e3 * (t1 - t0), m['nmi'].mean(), m['ami'].mean(), m['ars'].mean(), m['prc'].mean(), m['rcl'].mean(), m['iou'].mean()))
