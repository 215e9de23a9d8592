import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, ctypes as C
import bench
from learn_region_grow_b200 import _lib
from learn_region_grow_b200.engine import Engine
offsets, points, order, raw = bench.make_workload(68, 1000)
eng = Engine(1,1,512,512,13,0); eng.load_weights(bench.load_weights())
hp = bench.pinned_array(_lib, points.shape, np.float32); hp[...] = points
ho = bench.pinned_array(_lib, order.shape, np.int32); ho[...] = order
for it in range(3):
    t0=time.perf_counter(); eng.upload_concatenated(offsets, hp, ho, 0.1); t1=time.perf_counter()
    st = eng.segment_resident(resolution=0.1, seed=0); t2=time.perf_counter()
    lab = eng.labels(); t3=time.perf_counter()
    print('upload %.1f ms  segment %.1f ms  labels %.1f ms   (device grow %.1f fill %.1f)' % (1e3*(t1-t0),1e3*(t2-t1),1e3*(t3-t2), eng.profile()['grow_ms'], eng.profile()['fill_ms']))
