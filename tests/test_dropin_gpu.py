"""The drop-in end to end on the GPU box: learn_region_grow_b200/run_reference.py executes a driver script
that imports ``tensorflow`` / ``h5py`` / ``learn_region_grow_util`` and calls ``Session.run`` once per grow step exactly like
/root/reference/test_region_grow.py:86-99,257-258 (tests/dropin_driver_script.py -- repo-authored, because the reference tree
is not on the GPU box; where it IS present the unmodified script itself runs through the same entry point,
``python -m learn_region_grow_b200.run_reference /root/reference/test_region_grow.py --area 5``).

The script's grow loop runs on the host (the oracle's restatement of the driver) with the network evaluated by the sm_100a
engine through the drop-in; its labels must equal what the on-device driver (``Engine.segment_rooms``, persistent kernel and
lock-step loop) produces for the same rooms and Philox seed: the same forward bits per row whether a tile is evaluated alone
through ``lrg_forward_host`` or inside the grow kernel."""
import os
import sys

import numpy as np
import pytest

from learn_region_grow_b200 import _lib, ckpt, io_util, run_reference

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_driver_script_through_the_dropin_equals_the_device_driver(golden_weights, tmp_path):
    from tools import rooms as Rm
    from learn_region_grow_b200.engine import Engine
    raw = [Rm.generate_room(1000, n_raw=2500, n_boxes=4, dims=np.array([3.0, 2.5, 2.2])), Rm.generate_room(1001, n_raw=3000, n_boxes=3)]
    standins = os.path.join(REPO, 'learn_region_grow_b200', 'dropin', 'standins')
    sys.path.append(standins)
    try:
        io_util.saveToH5(str(tmp_path / 'rooms.h5'), raw)
    finally:
        sys.path.remove(standins)
        sys.modules.pop('h5py', None)
    ckpt.save_checkpoint(str(tmp_path / 'lrgnet_model5.ckpt'), golden_weights)
    for m in ('tensorflow', 'learn_region_grow_util', 'h5py'):
        sys.modules.pop(m, None)
    try:
        g = run_reference.run(os.path.join(REPO, 'tests', 'dropin_driver_script.py'),
                              ['--h5', str(tmp_path / 'rooms.h5'), '--ckpt', str(tmp_path / 'lrgnet_model5.ckpt'), '--seed', '3'])
    finally:
        for m in ('tensorflow', 'learn_region_grow_util', 'h5py'):
            sys.modules.pop(m, None)
    assert g['session_runs'] > 300 and len(g['cluster_labels']) == 2
    assert type(g['net']).__module__ == 'learn_region_grow_util' and g['net'].engine.lib is not None
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    pts = [f[0] for f in g['features']]
    orders = [f[1] for f in g['features']]
    for flags in (0, _lib.FLAG_LOCKSTEP):
        e.upload_rooms(pts, orders, resolution=0.1)
        stats = e.segment_resident(resolution=0.1, seed=3, flags=flags, spec_lanes=1)
        assert int(stats['grow_steps'].sum()) == g['session_runs']
        for dev, dev_filled, host, host_filled in zip(e.labels(filled=False), e.labels(filled=True), g['cluster_labels'], g['filled_labels']):
            np.testing.assert_array_equal(dev, host)
            np.testing.assert_array_equal(dev_filled, host_filled)
    # raw points in: the device's own feature preparation instead of the script's host-side one
    raw_labels, _ = e.segment_raw_rooms([r[:, :6] for r in raw], resolution=0.1, seed=3)
    assert [len(l) for l in raw_labels] == [len(r) for r in raw]
    e.close()
