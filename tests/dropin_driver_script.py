"""A repo-authored driver with the reference driver's import and call sequence (/root/reference/test_region_grow.py:1-19,
68,86-99,110,257-258), for tests/test_dropin_gpu.py: the real script does not travel to the GPU box, this one does.

It is run by learn_region_grow_b200/run_reference.py exactly like the real one -- ``import tensorflow``, ``import h5py`` and
``from learn_region_grow_util import *`` resolve to the drop-in -- opens the rooms with ``loadFromH5``, restores the
checkpoint with ``tf.compat.v1.train.Saver``, and grows every room on the HOST with the oracle's restatement of the driver
loop, one ``sess.run`` per grow step.  The grow loop's random draws come from the Philox streams the device driver uses, so
that its labels can be compared with ``Engine.segment_rooms``.

    dropin_driver_script.py --h5 rooms.h5 --ckpt model.ckpt --seed 3 [--max-rooms 1]
"""
import sys

import numpy
import h5py                                                  # noqa: F401  (the reference imports it, :2)
import tensorflow as tf
from learn_region_grow_util import *                         # noqa: F401,F403  (:18)

from oracle import feature_prep, lrg_driver

NUM_INLIER_POINT = 512                                       # :22-24
NUM_NEIGHBOR_POINT = 512
FEATURE_SIZE = 13
resolution = 0.1
args = dict(zip(sys.argv[1::2], sys.argv[2::2]))
seed = int(args.get('--seed', 0))

tf.compat.v1.reset_default_graph()                           # :68
config = tf.compat.v1.ConfigProto()                          # :86-93
config.gpu_options.allow_growth = True
config.allow_soft_placement = True
config.log_device_placement = False
sess = tf.compat.v1.Session(config=config)
net = LrgNet(1, 1, NUM_INLIER_POINT, NUM_NEIGHBOR_POINT, FEATURE_SIZE)          # noqa: F405
saver = tf.compat.v1.train.Saver()
saver.restore(sess, args['--ckpt'])

all_points, all_obj_id, all_cls_id = loadFromH5(args['--h5'])                   # noqa: F405  (:96-99)
session_runs = 0


def forward(inlier_points, neighbor_points):
    """One grow step of the network, fetched like the driver does (:257-258)."""
    global session_runs
    session_runs += 1
    input_add = numpy.zeros((1, NUM_NEIGHBOR_POINT), dtype=numpy.int32)
    input_remove = numpy.zeros((1, NUM_INLIER_POINT), dtype=numpy.int32)
    ls, add, add_acc, rmv, rmv_acc = sess.run([net.loss, net.add_output, net.add_acc, net.remove_output, net.remove_acc],
                                              {net.inlier_pl: inlier_points, net.neighbor_pl: neighbor_points,
                                               net.add_mask_pl: input_add, net.remove_mask_pl: input_remove})
    return add, rmv


cluster_labels, filled_labels, features = [], [], []
for room_id in range(min(len(all_points), int(args.get('--max-rooms', len(all_points))))):      # :110
    f = feature_prep.prepare_features(all_points[room_id], resolution)                  # :119-173
    grower = lrg_driver.RoomGrower(f['points'], f['order'], forward, lrg_driver.PhiloxRng(seed), resolution=resolution, room_id=room_id)
    grower.run()
    features.append((f['points'], f['order']))
    cluster_labels.append(grower.cluster_label.copy())
    filled_labels.append(grower.fill())
    print('room %d: %d points, %d grow steps, %d clusters' % (room_id, len(f['points']), grower.total_steps, grower.cluster_id - 1))
