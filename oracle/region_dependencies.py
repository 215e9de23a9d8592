"""How much of a room's region growing is inherently sequential?  ANALYSIS TOOL on top of the CPU oracle -- TEST
INFRASTRUCTURE (oracle/__init__.py), not part of the product.

    python -m oracle.region_dependencies [room_seed ...]

Runs the plain driver restatement (oracle/lrg_driver.py RoomGrower, Philox streams keyed by (room, seed point, step in
region) like the device) on bench-shaped rooms and records every region's points and the envelope it ever looked at
(seqMin - 1 .. seqMax + 1, /root/reference/test_region_grow.py:222-229,302-303).  Region j (later in seed order) would have
grown exactly the same with region i (earlier) still uncommitted iff no point of i lies inside j's envelope: then j never
sees the difference between "visited by i" and "not yet visited".  The regions with these edges form a DAG; its longest
path in grow steps is what a driver that grows independent regions of one room side by side (committing in seed order, so
the labels stay the reference's) cannot go below, whatever the number of lanes.
"""
import sys
import time

import numpy as np

from . import feature_prep, lrg_driver, lrg_forward


class _Recorder(lrg_driver.RoomGrower):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.rec = []

    def stop_growing(self, reason):
        self.rec.append(dict(seed=self.seed_id, steps=max(self.steps, 1), pts=np.nonzero(self.currentMask)[0],
                             lo=np.asarray(self.seqMinDims) - 1, hi=np.asarray(self.seqMaxDims) + 1))
        super().stop_growing(reason)


def analyse(room_seed, weights, n_raw=None):
    from tools import rooms
    raw = rooms.generate_room(room_seed) if n_raw is None else rooms.generate_room(room_seed, n_raw=n_raw)
    f = feature_prep.prepare_features(raw)
    fwd = lambda a, b: lrg_forward.forward(weights, a, b)
    g = _Recorder(f['points'], f['order'], fwd, lrg_driver.PhiloxRng(0))
    t0 = time.time()
    g.run()
    vox = g.point_voxels
    R = len(g.rec)
    steps = np.array([r['steps'] for r in g.rec])
    # finish[j] = earliest time (in grow steps) region j can be complete when every region starts as soon as the regions it
    # depends on are complete (unbounded lanes); lanes-limited schedules are simulated below
    deps = []
    for j, rj in enumerate(g.rec):
        d = []
        for i in range(j):
            p = vox[g.rec[i]['pts']]
            if np.any(np.all((p >= rj['lo']) & (p <= rj['hi']), axis=1)):
                d.append(i)
        deps.append(d)
    finish = np.zeros(R)
    for j in range(R):
        finish[j] = (max(finish[d] for d in deps[j]) if deps[j] else 0.0) + steps[j]
    out = dict(room=room_seed, n_points=len(f['points']), regions=R, total_steps=int(steps.sum()), critical_path=int(finish.max()),
               oracle_s=time.time() - t0)
    # Speculative lanes as a device could run them: seeds are handed out in order to L lanes; a region starts the moment a
    # lane is free; it is valid only if every region it depends on was COMMITTED (commits happen in seed order) when it
    # started, otherwise it starts over when the last of them commits (the lane is busy either way); `wasted` = grow steps of
    # discarded attempts, counted up to the restart.
    for L in (2, 3, 4, 8):
        lane_free = np.zeros(L)
        commit = np.zeros(R)
        last_start, wasted, restarts = 0.0, 0.0, 0
        for j in range(R):
            l = int(np.argmin(lane_free))
            start = max(lane_free[l], last_start)
            ready = max([commit[d] for d in deps[j]], default=0.0)
            if ready > start:
                restarts += 1
                wasted += min(ready - start, steps[j])
            fin = max(start, ready) + steps[j]
            commit[j] = max(fin, commit[j - 1] if j else 0.0)
            lane_free[l] = fin
            last_start = start
        out['lanes_%d' % L] = dict(makespan=int(commit[-1]), restarts=restarts, wasted_steps=int(wasted))
    return out


if __name__ == '__main__':
    import os
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'lrgnet_model5.npz')
    with np.load(gold) as z:
        w = {k: z[k] for k in z.files}
    for s in [int(x) for x in sys.argv[1:]] or [1000, 1026]:
        r = analyse(s, w)
        print(r, flush=True)
