#!/usr/bin/env python
"""One resident pass of the grow engine over N synthetic rooms with direct kernel launches (no CUDA graph), for use under
`ncu` (profiles/README.md has the commands).  Never a bench number: anything printed here ran under a profiler."""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rooms', type=int, default=68)
    ap.add_argument('--graph', action='store_true')
    args = ap.parse_args()
    import bench
    from learn_region_grow_b200 import _lib
    from learn_region_grow_b200.engine import Engine
    raw_off, raw = bench.make_workload(args.rooms, 1000)
    eng = Engine(1, 1, 512, 512, 13, 0)
    eng.load_weights(bench.load_weights())
    eng.upload_raw_concatenated(raw_off, raw, 0.1)
    flags = 0 if args.graph else _lib.FLAG_NO_GRAPH
    stats = eng.segment_resident(resolution=0.1, seed=0, flags=flags)
    pr = eng.profile()
    print('rooms %d  grow steps %d  iterations %d  grow %.1f ms  fill %.1f ms' %
          (args.rooms, int(stats['grow_steps'].sum()), pr['iterations'], pr['grow_ms'], pr['fill_ms']))


if __name__ == '__main__':
    main()
