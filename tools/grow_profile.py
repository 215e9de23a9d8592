#!/usr/bin/env python
"""Print the persistent grow kernel's per-item busy-time breakdown for one pass over N synthetic rooms."""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rooms', type=int, default=68)
    ap.add_argument('--repeat', type=int, default=2)
    ap.add_argument('--slots', type=int, default=0)
    ap.add_argument('--flags', type=int, default=0)
    ap.add_argument('--lanes', type=int, default=0)
    ap.add_argument('--kitti', type=int, default=0, help='N Semantic-KITTI-shaped scenes at 0.3 m instead of rooms')
    args = ap.parse_args()
    import bench
    from learn_region_grow_b200.engine import Engine
    res = 0.3 if args.kitti else 0.1
    if args.kitti:
        import numpy as np
        from tools import rooms as Rm
        scenes = [Rm.generate_outdoor_scene(3000 + i)[:, :6] for i in range(args.kitti)]
        raw_off = np.concatenate([[0], np.cumsum([len(s) for s in scenes])]).astype(np.int64)
        raw = np.ascontiguousarray(np.vstack(scenes), dtype=np.float32)
    else:
        raw_off, raw = bench.make_workload(args.rooms, 1000)
    eng = Engine(1, 1, 512, 512, 13, 0)
    eng.load_weights(bench.load_weights())
    if os.environ.get('LRG_TILE_TIMING'):          # (this tool's own switch; the library has an explicit call for it)
        from learn_region_grow_b200 import _lib
        _lib.check(eng.lib.lrg_engine_set_tile_timing(eng._h, 1))
    eng.upload_raw_concatenated(raw_off, raw, res)
    for _ in range(args.repeat):
        stats = eng.segment_resident(resolution=res, seed=0, max_slots=args.slots, flags=args.flags, spec_lanes=args.lanes)
        pr = eng.profile()
    steps = int(stats['grow_steps'].sum())
    print('rooms %d  grow steps %d (max/room %d)  grow %.1f ms  fill %.1f ms  persistent %s' %
          (args.rooms, steps, int(stats['grow_steps'].max()), pr['grow_ms'], pr['fill_ms'], pr['persistent']))
    for k in ('step', 'branch', 'gproj', 'head'):
        n = max(pr['items'][k], 1)
        print('  %-7s items %8d  busy %9.1f ms  avg %7.2f us/item   queue delay avg %6.2f us/item' %
              (k, pr['items'][k], pr['busy_ms'][k], 1e3 * pr['busy_ms'][k] / n, 1e3 * pr['queue_delay_ms'][k] / n))
    tot = sum(pr['busy_ms'].values())
    print('  busy total %.1f ms = %.1f %% of %d SMs x %.1f ms' % (tot, 100 * tot / (148 * pr['grow_ms']), 148, pr['grow_ms']))
    print('  longest room: %.1f us per step end to end' % (1e3 * pr['grow_ms'] / int(stats['grow_steps'].max())))
    if os.environ.get('LRG_TILE_TIMING'):
        import ctypes as C
        from learn_region_grow_b200 import _lib
        out = (C.c_uint64 * 64)()
        _lib.check(eng.lib.lrg_tile_timing(eng._h, C.byref(out), 0))
        names_b = ['init+bias', 'x tile', 'epi L0', 'epi L1', 'epi L2', 'epi L3', 'L4 nb0', 'L4 nb1', 'L4 nb2', 'L4 nb3', 'teardown']
        names_h = ['init', 'h1 tile+vec', 'epi nb0', 'epi nb1', 'epi nb2', 'epi nb3', 'final', 'teardown']
        nb, nh = max(out[15], 1), max(out[31], 1)
        print('  branch tile stages (cycles/tile, thread 0):', ', '.join('%s %d' % (n, out[i] // nb) for i, n in enumerate(names_b)),
              '| total', sum(out[i] for i in range(11)) // nb,
              '| MMA thread waits: weights %d, accumulator %d, activations %d' % (out[11] // nb, out[12] // nb, out[13] // nb))
        names_s = ['load state', 'apply', 'scan inliers+stop logic', 'stop/mark', 'seed search', 'scan neighbours', 'median', 'sampling', 'gather', 'write back']
        ns = max(out[47], 1)
        print('  step stages        (cycles/step, thread 0):', ', '.join('%s %d' % (n, out[32 + i] // ns) for i, n in enumerate(names_s)),
              '| total', sum(out[32 + i] for i in range(10)) // ns)
        print('  apply, cumulative cycles (thread 0): loads returned %d, masks + add stores %d, after barrier %d, removals stored %d' %
              tuple(out[32 + i] // ns for i in (10, 11, 12, 13)))
        bn = ['<=64', '<=256', '<=512', '<=1024', '<=2048', '>2048']
        print('  median by inlier-set size: ' + ', '.join('%s: %d steps x %d cyc' % (bn[b], out[49 + 2 * b], out[48 + 2 * b] // max(out[49 + 2 * b], 1)) for b in range(6)))
        print('  head tile stages   (cycles/tile, thread 0):', ', '.join('%s %d' % (n, out[16 + i] // nh) for i, n in enumerate(names_h)),
              '| total', sum(out[16 + i] for i in range(8)) // nh)


if __name__ == '__main__':
    main()
