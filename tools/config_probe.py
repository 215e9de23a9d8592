"""BASELINE.json configs 3 and 5 on the device through the raw-points path: variable-N ScanNet-shaped rooms and one
Semantic-KITTI-shaped scene at 0.3 m.  Prints sizes, steps and times (never a bench number)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tools import rooms
from learn_region_grow_b200.engine import Engine

eng = Engine(1, 1, 512, 512, 13, 0); eng.load_weights(bench.load_weights())
which = sys.argv[1] if len(sys.argv) > 1 else 'both'
if which in ('scannet', 'both'):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    area = rooms.generate_area(n, seed_base=2000, log_uniform=(5000, 60000))
    for it in range(2):
        t0 = time.perf_counter(); eq = eng.upload_raw_rooms(area, 0.1); t1 = time.perf_counter()
        st = eng.segment_resident(resolution=0.1, seed=0); t2 = time.perf_counter()
        lab = eng.raw_labels(); t3 = time.perf_counter()
    m = eng.room_metrics([r[:, 6].astype(np.int32) for r in area], raw=True)
    print('scannet-shaped: %d rooms, raw %d..%d (sum %d), equalised sum %d | prep %.1f ms grow+fill %.1f ms labels %.1f ms | %d steps, longest %d | %.2f M raw points/s | NMI %.3f' %
          (n, min(map(len, area)), max(map(len, area)), sum(map(len, area)), int(eq[-1]), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2),
           int(st['grow_steps'].sum()), int(st['grow_steps'].max()), sum(map(len, area)) / (t3 - t0) / 1e6, m['nmi'].mean()), flush=True)
if which in ('kitti', 'both'):
    scene = rooms.generate_outdoor_scene(3000)
    from learn_region_grow_b200 import _lib
    for flags in (0, 0, _lib.FLAG_NO_STEP_OVERLAP, _lib.FLAG_NO_SPATIAL_INDEX):
        t0 = time.perf_counter(); eq = eng.upload_raw_rooms([scene], 0.3); t1 = time.perf_counter()
        st = eng.segment_resident(resolution=0.3, seed=0, flags=flags); t2 = time.perf_counter()
        lab = eng.raw_labels(); t3 = time.perf_counter()
        print('flags %3d grow %.1f ms ' % (flags, eng.profile()['grow_ms']), end='')
        print('kitti-shaped: raw %d, equalised %d | prep %.1f ms grow+fill %.1f ms labels %.1f ms | %d steps %d regions %d clusters | %.2f M raw points/s' %
              (len(scene), int(eq[-1]), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), int(st['grow_steps'][0]), int(st['regions'][0]),
               int(st['clusters'][0]), len(scene) / (t3 - t0) / 1e6), flush=True)
