"""Engine defaults of the speculative lanes across the BASELINE configurations: grow time (min of 2) by spec_top / scheduling flag.
python tools/spec_top_probe.py [configs]  (default 2,3,4,5; config 4 simulates one of 8 ranks with --shard)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from learn_region_grow_b200.engine import Engine
from learn_region_grow_b200 import _lib, parallel

e = Engine(1, 1, 512, 512, 13, 0); e.load_weights(bench.load_weights())
configs = [int(c) for c in (sys.argv[1].split(',') if len(sys.argv) > 1 else '2,3,4,5'.split(','))]
for cfg in configs:
    name, kind, total, res, _ = bench.WORKLOADS[cfg]
    total = total or 68
    rows = [bench._room(kind, g) for g in range(total)]
    cases = [('all rooms', rows)]
    if cfg == 4:     # what one of 8 ranks gets (LPT shard 0): the strong-scaling case is bound by the rank's longest chain
        for nr in (8, 4, 16):
            shards = parallel.shard_rooms(np.array([len(r) for r in rows], np.int64), nr)
            cases.append(('rank 0 of %d' % nr, [rows[int(g)] for g in shards[0]]))
        cases = cases[:2]
    for cname, rr in cases:
        raw_off, raw = bench.concat_rooms([r[:, :6] for r in rr])
        e.upload_raw_concatenated(raw_off, raw, res)
        ref = None
        for label, kw in (('default', dict()), ('reserved CTAs for critical rooms', dict(flags=_lib.FLAG_PRIORITY)), ('rooms in index order', dict(flags=_lib.FLAG_ROOMS_IN_ORDER)),
                          ('top 4 crit off (first rule)', dict(spec_top=4, spec_crit=-1)), ('8 lanes', dict(spec_lanes=8)), ('1 lane', dict(spec_lanes=1))):
            ms = []
            for rep in range(2):
                st = e.segment_resident(resolution=res, seed=0, spec_lanes=kw.get('spec_lanes', 0), spec_top=kw.get('spec_top', 0), flags=kw.get('flags', 0), spec_crit=kw.get('spec_crit', 0), spec_min_idle=kw.get('spec_min_idle', 0), max_steps_per_region=kw.get('msr', 0))
                ms.append(e.profile()['grow_ms'])
            lab = np.concatenate(e.labels(True))
            ref = lab if ref is None else ref
            print('config %d %-13s %-32s grow %8.1f ms | steps %d longest %d | labels %s' % (cfg, cname, label, min(ms), int(st['grow_steps'].sum()),
                  int(st['grow_steps'].max()), 'same' if np.array_equal(lab, ref) else 'DIFFERENT'), flush=True)
