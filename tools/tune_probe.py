"""A/B switches of the persistent kernel on the 68 bench rooms: grow time (min of 3) by flags x spec_top.  python tools/tune_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from learn_region_grow_b200.engine import Engine
from learn_region_grow_b200 import _lib

e = Engine(1, 1, 512, 512, 13, 0); e.load_weights(bench.load_weights())
raw_off, raw = bench.make_workload(68, 1000)
e.upload_raw_concatenated(raw_off, raw, 0.1)
ref = None
for name, flags in (('default', 0), ('no step overlap', _lib.FLAG_NO_STEP_OVERLAP), ('no spatial index', _lib.FLAG_NO_SPATIAL_INDEX)):
    for lanes, top in ((1, 0), (4, 4), (4, 2)):
        ms = []
        for rep in range(3):
            st = e.segment_resident(resolution=0.1, seed=0, flags=flags, spec_lanes=lanes, spec_top=top)
            ms.append(e.profile()['grow_ms'])
        lab = np.concatenate(e.labels(True))
        ref = lab if ref is None else ref
        pr = e.profile()
        print('%-16s lanes %d top %d: grow %.1f ms (%s) | items %s | us/item %s | busy %.2f | labels %s' % (
            name, lanes, top, min(ms), ' '.join('%.1f' % m for m in ms), ' '.join('%s %d' % (k, v) for k, v in pr['items'].items() if v),
            ' '.join('%s %.1f' % (k, 1e3 * pr['busy_ms'][k] / max(pr['items'][k], 1)) for k in pr['items'] if pr['items'][k]),
            sum(pr['busy_ms'].values()) / (148 * pr['grow_ms']), 'same' if np.array_equal(lab, ref) else 'DIFFERENT'), flush=True)
