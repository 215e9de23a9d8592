"""Workload calibration probe (the reference's S3DIS logs show ~50 regions and ~950 grow steps per room): regions,
clusters and grow steps per room of the synthetic generator as a function of its knobs.  python tools/calib_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tools import rooms
from learn_region_grow_b200.engine import Engine

e = Engine(1, 1, 512, 512, 13, 0); e.load_weights(bench.load_weights())
print('n_boxes color xyz_noise | raw  eq | regions clusters steps (mean over 12 rooms; max steps) | stops nn/ne/st | NMI PRC RCL')
for n_boxes in (None, 20, 10, 5):
    for color in (0.5, 0.25, 0.1):
        for noise in (0.01, 0.004):
            rs = [rooms.generate_room(1000 + i, n_boxes=n_boxes, color_jitter=color, xyz_noise=noise) for i in range(12)]
            eq = e.upload_raw_rooms(rs, 0.1)
            st = e.segment_resident(resolution=0.1, seed=0)
            m = e.room_metrics([r[:, 6].astype(np.int32) for r in rs], raw=True)
            print('%6s %5.2f %6.3f | %5d %5d | %6.1f %6.1f %7.1f (%5d) | %5.1f %5.1f %5.1f | %.3f %.3f %.3f' % (
                n_boxes, color, noise, np.mean([len(r) for r in rs]), eq[-1] / 12, st['regions'].mean(), st['clusters'].mean(), st['grow_steps'].mean(),
                st['grow_steps'].max(), st['stop_noneighbor'].mean(), st['stop_noexpand'].mean(), st['stop_stuck'].mean(),
                m['nmi'].mean(), m['prc'].mean(), m['rcl'].mean()), flush=True)
