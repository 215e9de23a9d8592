"""Pooled-projection servers on/off (LRG_FLAG_NO_PROJ_SERVERS): grow time of the bench workload, one room alone, restarts and beam; labels
checked against the run without servers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from learn_region_grow_b200.engine import Engine
from learn_region_grow_b200 import _lib

rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 68
raw_off, raw = bench.make_workload(rooms, 1000)
eng = Engine(1, 1, 512, 512, 13, 0); eng.load_weights(bench.load_weights())
eng.upload_raw_concatenated(raw_off, raw, 0.1)
for name, kw in (('plain', {}), ('restarts 10', dict(num_restarts=10)), ('beam 3x3', dict(beam_width=3, search_width=3))):
    ref = None
    for s in ('0', '1', '0', '1'):
        ms = []
        for it in range(3):
            st = eng.segment_resident(resolution=0.1, seed=0, flags=0 if s == '1' else _lib.FLAG_NO_PROJ_SERVERS, **kw)
            ms.append(eng.profile()['grow_ms'])
        lab = np.concatenate(eng.labels(True))
        if ref is None:
            ref = lab
        pr = eng.profile()
        print('%-12s servers %s: grow ms %s  steps %d  labels %s | busy ms %s | delay us/item %s' % (
            name, s, ' '.join('%.1f' % m for m in ms), int(st['grow_steps'].sum()), 'same' if np.array_equal(lab, ref) else 'DIFFERENT',
            ' '.join('%s %.0f' % (k, pr['busy_ms'][k]) for k in ('step', 'branch', 'gproj', 'head')),
            ' '.join('%s %.1f' % (k, 1e3 * pr['queue_delay_ms'][k] / max(pr['items'][k], 1)) for k in ('branch', 'gproj', 'head'))), flush=True)
