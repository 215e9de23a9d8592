"""Drop-in for /root/reference/learn_region_grow_util.py: the names the drivers pull in with
``from learn_region_grow_util import *`` (test_region_grow.py:18), with ``LrgNet`` evaluated by the sm_100a engine.

    net = LrgNet(batch_size, seq_len, num_inlier_points, num_neighbor_points, feature_size, lite)   # util.py:76
    sess.run([net.loss, net.add_output, net.add_acc, net.remove_output, net.remove_acc],
             {net.inlier_pl: ..., net.neighbor_pl: ..., net.add_mask_pl: ..., net.remove_mask_pl: ...})

There is no CPU path: constructing ``LrgNet`` without a CUDA device or without liblrg_b200.so raises.
"""
import os                                                                      # noqa: F401  (re-exported like the reference)

import numpy
import h5py                                                                    # noqa: F401
import tensorflow as tf

from learn_region_grow_b200.engine import Engine
from learn_region_grow_b200.io_util import loadFromH5, savePCD, savePLY, saveToH5   # noqa: F401


def _log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    return x - m - numpy.log(numpy.exp(x - m).sum(axis=-1, keepdims=True))


class LrgNet:
    def __init__(self, batch_size, seq_len, num_inlier_points, num_neighbor_points, feature_size, lite=0):
        B = batch_size * seq_len
        self.engine = Engine(batch_size, seq_len, num_inlier_points, num_neighbor_points, feature_size, lite)
        H = tf.Handle
        self.inlier_pl = H(self, 'inlier_pl', (B, num_inlier_points, feature_size), 'float32')        # util.py:100
        self.neighbor_pl = H(self, 'neighbor_pl', (B, num_neighbor_points, feature_size), 'float32')  # :101
        self.add_mask_pl = H(self, 'add_mask_pl', (B, num_neighbor_points), 'int32')                  # :102
        self.remove_mask_pl = H(self, 'remove_mask_pl', (B, num_inlier_points), 'int32')              # :103
        for name in ('loss', 'add_output', 'add_acc', 'remove_output', 'remove_acc', 'add_loss', 'remove_loss',
                     'add_prc', 'add_rcl', 'remove_prc', 'remove_rcl'):
            setattr(self, name, H(self, name))
        self.train_op = H(self, 'train_op')
        tf.register_net(self)

    # -- hooks used by the session stand-in -----------------------------------------------------------------------
    def _load_variables(self, tensors):
        self.engine.load_weights(tensors)

    def _variables(self):
        return self.engine.variables()

    def _evaluate(self, names, feeds):
        if 'train_op' in names:
            raise NotImplementedError('training is outside the inference hot path of this engine (SURVEY.md 8 a8)')
        for req in ('inlier_pl', 'neighbor_pl'):
            if req not in feeds:
                raise ValueError('You must feed a value for placeholder tensor %r' % req)
        add, rmv = self.engine.forward(feeds['inlier_pl'], feeds['neighbor_pl'])
        out = {'add_output': add, 'remove_output': rmv}
        scalars = [n for n in names if n not in out]
        if scalars:
            for req in ('add_mask_pl', 'remove_mask_pl'):
                if req not in feeds:
                    raise ValueError('You must feed a value for placeholder tensor %r' % req)
            out.update(_fetch_scalars(add, rmv, numpy.asarray(feeds['add_mask_pl']), numpy.asarray(feeds['remove_mask_pl'])))
        return out


def _fetch_scalars(add, rmv, add_mask, rmv_mask):
    """The fetch-only scalars of util.py:165-186, from the logits (host side: 2x512x2 values)."""
    add_mask = add_mask.astype(numpy.int64)
    rmv_mask = rmv_mask.astype(numpy.int64)
    ce_add = -numpy.take_along_axis(_log_softmax(add), add_mask[..., None], -1)[..., 0]
    ce_rmv = -numpy.take_along_axis(_log_softmax(rmv), rmv_mask[..., None], -1)[..., 0]
    pos = rmv_mask.astype(bool)
    add_loss = numpy.float32(ce_add.mean())
    pos_loss = numpy.float32(ce_rmv[pos].mean()) if pos.any() else numpy.float32(0)      # NaN -> 0 (:170-171)
    neg_loss = numpy.float32(ce_rmv[~pos].mean()) if (~pos).any() else numpy.float32(0)
    add_pred = add.argmax(-1)
    rmv_pred = rmv.argmax(-1)
    rmv_hard = _log_softmax(rmv)[..., 1] > numpy.log(0.5)                                # softmax[...,1] > 0.5 (:181)
    tp_add = numpy.float32(numpy.sum((add_pred == 1) & (add_mask == 1)))
    tp_rmv = numpy.float32(numpy.sum(rmv_hard & (rmv_mask == 1)))
    return {
        'add_loss': add_loss, 'remove_loss': numpy.float32(pos_loss + neg_loss),
        'loss': numpy.float32(add_loss + pos_loss + neg_loss),                           # :186
        'add_acc': numpy.float32(numpy.mean(add_pred == add_mask)),                      # :175
        'remove_acc': numpy.float32(numpy.mean(rmv_pred == rmv_mask)),                   # :180
        'add_prc': tp_add / (numpy.float32(add_pred.sum()) + 1), 'add_rcl': tp_add / (numpy.float32(add_mask.sum()) + 1),
        'remove_prc': tp_rmv / (numpy.float32(rmv_hard.sum()) + 1), 'remove_rcl': tp_rmv / (numpy.float32(rmv_mask.sum()) + 1),
    }
