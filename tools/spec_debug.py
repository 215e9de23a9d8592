import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import bench
from learn_region_grow_b200.engine import Engine
from util_rooms import golden_room
e = Engine(1, 1, 512, 512, 13, 0)
e.load_weights(bench.load_weights())
points, order = golden_room(1001)
for lanes in (1, 2, 2, 4, 8):
    for flags in (0, 4):
        labels, st = e.segment_rooms([points], [order], resolution=0.1, seed=12345, spec_lanes=lanes, flags=flags)
        print(lanes, flags, {k: int(st[k][0]) for k in st.dtype.names}, flush=True)
