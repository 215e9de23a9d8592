"""Reserved-CTA scheduling sweep (LRG_HI="slots,ctas"): grow time of the bench workload per setting, labels checked against
the plain FIFO run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from learn_region_grow_b200.engine import Engine

rooms = int(sys.argv[1]) if len(sys.argv) > 1 else 68
settings = sys.argv[2:] or ['0,0', '2,16', '2,24', '3,24', '4,32', '1,12', '0,0']
raw_off, raw = bench.make_workload(rooms, 1000)
eng = Engine(1, 1, 512, 512, 13, 0); eng.load_weights(bench.load_weights())
eng.upload_raw_concatenated(raw_off, raw, 0.1)
ref = None
for s in settings:
    os.environ['LRG_HI'] = s
    ms = []
    for it in range(4):
        st = eng.segment_resident(resolution=0.1, seed=0)
        ms.append(eng.profile()['grow_ms'])
    lab = np.concatenate(eng.labels(True))
    if ref is None:
        ref = lab
    pr = eng.profile()
    print('LRG_HI=%-6s grow ms %s  best %.1f  labels %s  | delay us/item: %s' % (
        s, ' '.join('%.1f' % m for m in ms), min(ms[1:]), 'same' if np.array_equal(lab, ref) else 'DIFFERENT',
        ' '.join('%s %.1f' % (k, 1e3 * pr['queue_delay_ms'][k] / max(pr['items'][k], 1)) for k in ('branch', 'gproj', 'head'))), flush=True)
