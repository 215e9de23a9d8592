"""Generate the committed golden fixtures under tests/golden/ -- run in the BUILD container only
(needs /root/reference).  TEST INFRASTRUCTURE (oracle/__init__.py).

    python -m oracle.make_golden

1. ``lrgnet_model5.npz``      - the 32 trainable tensors of /root/reference/models/lrgnet_model5.ckpt (the only
                                LrgNet weight set the reference ships), read with learn_region_grow_b200/ckpt.py.
2. ``driver_trace_<seed>.npz`` - the UNMODIFIED /root/reference/test_region_grow.py executed on one synthetic room
                                (oracle/run_reference.py; forward = oracle/lrg_forward.py): the 13-D features and seed
                                order it computed, a CRC of every tile it fed to / logits it got from Session.run, and
                                its final (filled) cluster labels.  tests/test_oracle_driver.py replays these with
                                oracle/lrg_driver.py.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = '/root/reference'
GOLD = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, REPO)

from learn_region_grow_b200 import ckpt              # noqa: E402
from tools import rooms                               # noqa: E402
from oracle import run_reference                      # noqa: E402


def golden_weights():
    t = ckpt.load_checkpoint(os.path.join(REF, 'models', 'lrgnet_model5.ckpt'))
    keep = {k: v for k, v in t.items() if k.startswith('lrg_') and 'Adam' not in k}
    assert len(keep) == 32 and sum(v.size for v in keep.values()) == 791044
    np.savez(os.path.join(GOLD, 'lrgnet_model5.npz'), **keep)
    print('weights: %d tensors' % len(keep))


def golden_trace(seed, n_raw, n_boxes):
    room = rooms.generate_room(seed, n_raw=n_raw, n_boxes=n_boxes, dims=np.array([3.0, 2.5, 2.2]))
    scratch = '/tmp/lrg_golden_%d' % seed
    os.makedirs(os.path.join(scratch, 'data'), exist_ok=True)
    os.makedirs(os.path.join(scratch, 'models'), exist_ok=True)
    for ext in ('index', 'data-00000-of-00001'):
        dst = os.path.join(scratch, 'models', 'lrgnet_model5.ckpt.' + ext)
        if not os.path.exists(dst):
            os.symlink(os.path.join(REF, 'models', 'lrgnet_model5.ckpt.' + ext), dst)
    open(os.path.join(scratch, 'data', 's3dis_sampled.txt'), 'w').close()
    head, tail = run_reference.shim_paths(REF)
    sys.path[:0] = head
    sys.path.extend(tail)
    from learn_region_grow_b200 import io_util
    cwd = os.getcwd()
    os.chdir(scratch)
    try:
        io_util.saveToH5('data/s3dis_area5.h5', [room])
        import learn_region_grow_util as shim_util
        shim_util.TRACE = []
        shim_util.FORWARD_DTYPE = np.float64
        buf = io.StringIO()
        t0 = time.time()
        with contextlib.redirect_stdout(buf):
            g = run_reference.run(os.path.join(REF, 'test_region_grow.py'), ['--area', '5'])
        wall = time.time() - t0
        trace = shim_util.TRACE
        shim_util.TRACE = None
        shim_util.FORWARD_DTYPE = np.float32
    finally:
        os.chdir(cwd)
    log = buf.getvalue()
    out = dict(room=room, points=g['points'].astype(np.float32), order=np.asarray(g['order']),
               curvatures=np.asarray(g['curvatures']), cluster_label=np.asarray(g['cluster_label']),
               inlier_crc=np.array([t['inlier_crc'] for t in trace], dtype=np.uint32),
               neighbor_crc=np.array([t['neighbor_crc'] for t in trace], dtype=np.uint32),
               add_crc=np.array([t['add_crc'] for t in trace], dtype=np.uint32),
               remove_crc=np.array([t['remove_crc'] for t in trace], dtype=np.uint32),
               log=np.array(log), reference_wall_s=np.array(wall),
               buckets=np.array([g['comp_time_analysis'][k] for k in ('feature', 'net', 'neighbor', 'inlier')]))
    np.savez_compressed(os.path.join(GOLD, 'driver_trace_%d.npz' % seed), **out)
    print('trace %d: N_raw %d N_eq %d steps %d wall %.1fs' % (seed, len(room), len(out['points']), len(trace), wall))
    print(log[-400:])


def golden_restart_trace(seed, n_raw, n_boxes):
    """3. ``restart_trace_<seed>.npz`` - the UNMODIFIED /root/reference/test_random_restart.py (10 restarts per seed, 'np'
    scoring) on the room of driver_trace_<seed>.npz: tile / logits CRCs of every Session.run call, final labels, log."""
    room = rooms.generate_room(seed, n_raw=n_raw, n_boxes=n_boxes, dims=np.array([3.0, 2.5, 2.2]))
    scratch = '/tmp/lrg_golden_rr_%d' % seed
    os.makedirs(os.path.join(scratch, 'data'), exist_ok=True)
    os.makedirs(os.path.join(scratch, 'models'), exist_ok=True)
    for ext in ('index', 'data-00000-of-00001'):
        dst = os.path.join(scratch, 'models', 'lrgnet_model5.ckpt.' + ext)
        if not os.path.exists(dst):
            os.symlink(os.path.join(REF, 'models', 'lrgnet_model5.ckpt.' + ext), dst)
    head, tail = run_reference.shim_paths(REF)
    sys.path[:0] = head
    sys.path.extend(tail)
    from learn_region_grow_b200 import io_util
    cwd = os.getcwd()
    os.chdir(scratch)
    try:
        io_util.saveToH5('data/s3dis_area5.h5', [room])
        import learn_region_grow_util as shim_util
        shim_util.TRACE = []
        shim_util.FORWARD_DTYPE = np.float64
        buf = io.StringIO()
        t0 = time.time()
        with contextlib.redirect_stdout(buf):
            g = run_reference.run(os.path.join(REF, 'test_random_restart.py'), ['--area', '5'])
        wall = time.time() - t0
        trace = shim_util.TRACE
        shim_util.TRACE = None
        shim_util.FORWARD_DTYPE = np.float32
    finally:
        os.chdir(cwd)
    log = buf.getvalue()
    out = dict(cluster_label=np.asarray(g['cluster_label']),
               inlier_crc=np.array([t['inlier_crc'] for t in trace], dtype=np.uint32),
               neighbor_crc=np.array([t['neighbor_crc'] for t in trace], dtype=np.uint32),
               log=np.array(log), reference_wall_s=np.array(wall), num_restarts=np.array(g['NUM_RESTARTS']))
    np.savez_compressed(os.path.join(GOLD, 'restart_trace_%d.npz' % seed), **out)
    print('restart trace %d: steps %d wall %.1fs' % (seed, len(trace), wall))
    print(log[-300:])


def golden_beam_trace(seed, n_raw, n_boxes, scoring='np'):
    """4. ``beam_trace_<seed>.npz`` - the UNMODIFIED /root/reference/test_beam_search.py (BEAM_WIDTH 3, SEARCH_WIDTH 3, 'np'
    scoring) on the room of driver_trace_<seed>.npz.  The script concatenates ``range(n) + list(...)`` (:212,224), which is
    Python 2; it is executed as it is with a list-returning ``range`` among its module globals (run_reference.py2_range).
    ``scoring='ml'`` passes ``--scoring ml`` (:46-47,263-264: candidates ranked by accumulated log-probability) and writes
    ``beam_ml_trace_<seed>.npz``."""
    room = rooms.generate_room(seed, n_raw=n_raw, n_boxes=n_boxes, dims=np.array([3.0, 2.5, 2.2]))
    scratch = '/tmp/lrg_golden_bs_%d' % seed
    os.makedirs(os.path.join(scratch, 'data'), exist_ok=True)
    os.makedirs(os.path.join(scratch, 'models'), exist_ok=True)
    for ext in ('index', 'data-00000-of-00001'):
        dst = os.path.join(scratch, 'models', 'lrgnet_model5.ckpt.' + ext)
        if not os.path.exists(dst):
            os.symlink(os.path.join(REF, 'models', 'lrgnet_model5.ckpt.' + ext), dst)
    head, tail = run_reference.shim_paths(REF)
    sys.path[:0] = head
    sys.path.extend(tail)
    from learn_region_grow_b200 import io_util
    cwd = os.getcwd()
    os.chdir(scratch)
    try:
        io_util.saveToH5('data/s3dis_area5.h5', [room])
        import learn_region_grow_util as shim_util
        shim_util.TRACE = []
        shim_util.FORWARD_DTYPE = np.float64
        buf = io.StringIO()
        t0 = time.time()
        with contextlib.redirect_stdout(buf):
            g = run_reference.run(os.path.join(REF, 'test_beam_search.py'), ['--area', '5'] + (['--scoring', scoring] if scoring != 'np' else []),
                                  init_globals={'range': run_reference.py2_range})
        assert g['scoring'] == scoring
        wall = time.time() - t0
        trace = shim_util.TRACE
        shim_util.TRACE = None
        shim_util.FORWARD_DTYPE = np.float32
    finally:
        os.chdir(cwd)
    log = buf.getvalue()
    out = dict(cluster_label=np.asarray(g['cluster_label']),
               inlier_crc=np.array([t['inlier_crc'] for t in trace], dtype=np.uint32),
               neighbor_crc=np.array([t['neighbor_crc'] for t in trace], dtype=np.uint32),
               log=np.array(log), reference_wall_s=np.array(wall),
               beam_width=np.array(g['BEAM_WIDTH']), search_width=np.array(g['SEARCH_WIDTH']))
    np.savez_compressed(os.path.join(GOLD, 'beam%s_trace_%d.npz' % ('' if scoring == 'np' else '_' + scoring, seed)), **out)
    print('beam trace %d (%s): steps %d wall %.1fs' % (seed, scoring, len(trace), wall))
    print(log[-300:])


def golden_forward_pin():
    """5. ``forward_graphdef.npz`` - the forward PINNED TO THE SHIPPED GRAPH: tile pairs the oracle driver feeds the network
    on the golden room 1000 (steps 3, 10 and 25 of the first regions that get that far -- SURVEY 8d config 1 -- plus two
    random pairs) evaluated by interpreting /root/reference/models/lrgnet_model5.ckpt.meta op by op
    (oracle/graphdef_forward.py) on the checkpoint's own tensors, in float64 and in float32."""
    from oracle import graphdef_forward, lrg_driver, lrg_forward
    with np.load(os.path.join(GOLD, 'driver_trace_1000.npz')) as z:
        points, order = z['points'], z['order']
    with np.load(os.path.join(GOLD, 'lrgnet_model5.npz')) as z:
        weights = {k: z[k] for k in z.files}
    g = lrg_driver.RoomGrower(points, order, lambda a, b: lrg_forward.forward(weights, a, b), lrg_driver.PhiloxRng(0))
    g.trace = []
    g.run()
    tiles, where = [], []
    per_region = {}
    for t in g.trace:
        per_region.setdefault(t['seed'], []).append(t)
    for seed, steps in per_region.items():
        for want in (3, 10, 25):
            if len(steps) > want and len(tiles) < 6:
                tiles.append((steps[want]['inlier'][0], steps[want]['neighbor'][0]))
                where.append((seed, want, steps[want]['n_inlier'], steps[want]['n_neighbor']))
    rng = np.random.RandomState(5)
    for _ in range(2):
        tiles.append((rng.randn(512, 13).astype(np.float32), rng.randn(512, 13).astype(np.float32)))
        where.append((-1, -1, 512, 512))
    inlier = np.stack([t[0] for t in tiles]).astype(np.float32)
    neighbor = np.stack([t[1] for t in tiles]).astype(np.float32)
    out = dict(inlier=inlier, neighbor=neighbor, where=np.array(where))
    for name, dt in (('f64', np.float64), ('f32', np.float32)):
        sg = graphdef_forward.ShippedGraphForward(os.path.join(REF, 'models', 'lrgnet_model5.ckpt'), dt)
        add, rmv = sg.forward(inlier, neighbor)
        out['add_' + name], out['remove_' + name] = add, rmv
    out['ops'] = np.array(sorted('%s x %d' % kv for kv in sg.interp.ops_seen.items()))
    out['endpoints'] = np.array([sg.inlier_pl, sg.neighbor_pl, sg.add_out, sg.remove_out])
    np.savez_compressed(os.path.join(GOLD, 'forward_graphdef.npz'), **out)
    print('forward pin: %d tile pairs %s; ops %s' % (len(tiles), where, list(out['ops'])))


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ['beam']:
        golden_beam_trace(1000, 2500, 4)
        sys.exit(0)
    if sys.argv[1:] == ['beam_ml']:
        golden_beam_trace(1000, 2500, 4, scoring='ml')
        sys.exit(0)
    if sys.argv[1:] == ['forward']:
        golden_forward_pin()
        sys.exit(0)
    golden_weights()
    golden_trace(1000, 2500, 4)
    golden_trace(1001, 6000, 8)
    golden_restart_trace(1000, 2500, 4)
    golden_beam_trace(1000, 2500, 4)
    golden_beam_trace(1000, 2500, 4, scoring='ml')
    golden_forward_pin()
