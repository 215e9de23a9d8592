import sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from tools import rooms
from learn_region_grow_b200.engine import Engine
e = Engine(1, 1, 512, 512, 13, 0); e.load_weights(bench.load_weights())
raw = rooms.generate_room(1000, n_raw=2500, n_boxes=4, dims=np.array([3.0, 2.5, 2.2]))
lab, st = e.segment_raw_rooms([raw], resolution=0.1, seed=0)
print('ok', st['grow_steps'])
