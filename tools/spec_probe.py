"""Speculative-lane probe: grow time, committed / discarded steps by spec_lanes and by which rooms speculate (spec_top rooms
with the most work left, anybody while spec_min_idle CTAs idle; -1 = all / never), on one long room and on the 68 bench rooms.  python tools/spec_probe.py [rooms]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                    # noqa: E402
from learn_region_grow_b200.engine import Engine     # noqa: E402

n_rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 68
e = Engine(1, 1, 512, 512, 13, 0)
e.load_weights(bench.load_weights())
raw_off, raw = bench.make_workload(n_rooms, 1000)
ref = None
for sel, name in ((slice(26, 27) if n_rooms > 26 else slice(0, 1), 'longest room'), (slice(0, n_rooms), '%d rooms' % n_rooms)):
    rooms_raw = [raw[raw_off[i]:raw_off[i + 1]] for i in range(n_rooms)][sel]
    base = sel.start
    e.upload_raw_rooms(rooms_raw, 0.1)
    ref = None
    for lanes, top, min_idle in ((1, 0, 0), (2, -1, -1), (4, -1, -1), (4, 4, -1), (4, 2, -1), (3, 4, -1), (4, 6, -1), (4, 8, -1), (0, 0, 0), (4, 4, 120),
                                 (4, 4, 64), (6, 4, 96), (8, 2, -1)):
        if name == 'longest room' and (top, min_idle) not in ((0, 0), (-1, -1)):
            continue
        ms = []
        for rep in range(3):
            st = e.segment_resident(resolution=0.1, seed=0, room_id_base=base, spec_lanes=lanes, spec_top=top, spec_min_idle=min_idle)
            ms.append(e.profile()['grow_ms'])
        lab = np.concatenate(e.labels(True))
        if ref is None:
            ref = lab
        pr = e.profile()
        busy = sum(pr['busy_ms'].values())
        print('%-13s lanes %d top %2d min_idle %3d: grow %7.2f ms (min of 3; %s)  steps %7d  discarded %6d (regrown %4d dropped %4d)  SM busy %.2f  same labels %s'
              % (name, lanes, top, min_idle, min(ms), ' '.join('%.1f' % m for m in ms), st['grow_steps'].sum(), st['spec_wasted_steps'].sum(),
                 st['spec_restarts'].sum(), st['spec_dropped'].sum(), busy / (148 * pr['grow_ms']), np.array_equal(lab, ref)), flush=True)
