"""CPU stand-in for /root/reference/learn_region_grow_util.py used ONLY to run the *unmodified* reference
driver on the host (golden-vector generation and the CPU baseline) -- TEST INFRASTRUCTURE (oracle/__init__.py).

``LrgNet`` has the reference constructor (util.py:76) and handle attributes (:100-103,149,162,175,180,186);
``Session.run`` evaluates them with the numpy oracle forward.  The product drop-in of the same name lives in
learn_region_grow_b200/dropin/ and never imports this file.
"""
import os
import zlib

import numpy
import h5py
import tensorflow as tf

from learn_region_grow_b200.io_util import loadFromH5, savePCD, savePLY     # noqa: F401 (drivers import *)
from oracle import lrg_forward

TRACE = None        # set to a list to record one dict per Session.run call (oracle/make_golden.py)
FORWARD_DTYPE = numpy.float32   # make_golden uses float64 (rounded to float32) so the masks do not depend on the BLAS build


def _crc(a):
    return zlib.crc32(numpy.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


class LrgNet:
    def __init__(self, batch_size, seq_len, num_inlier_points, num_neighbor_points, feature_size, lite=0):
        B = batch_size * seq_len
        self.lite = lite
        self.feature_size = feature_size
        H = tf.Handle
        self.inlier_pl = H(self, 'inlier_pl', (B, num_inlier_points, feature_size), 'float32')
        self.neighbor_pl = H(self, 'neighbor_pl', (B, num_neighbor_points, feature_size), 'float32')
        self.add_mask_pl = H(self, 'add_mask_pl', (B, num_neighbor_points), 'int32')
        self.remove_mask_pl = H(self, 'remove_mask_pl', (B, num_inlier_points), 'int32')
        for name in ('loss', 'add_output', 'add_acc', 'remove_output', 'remove_acc'):
            setattr(self, name, H(self, name))
        self.weights = None
        tf.register_net(self)

    def _load_variables(self, tensors):
        self.weights = {n: numpy.asarray(tensors[n], numpy.float32).reshape(s)
                        for n, s in lrg_forward.variable_shapes(self.feature_size, self.lite)}

    def _variables(self):
        return dict(self.weights)

    def _evaluate(self, names, feeds):
        add, rmv = lrg_forward.forward(self.weights, feeds['inlier_pl'], feeds['neighbor_pl'], self.lite, FORWARD_DTYPE)
        add, rmv = add.astype(numpy.float32), rmv.astype(numpy.float32)
        out = {'add_output': add, 'remove_output': rmv}
        if any(n in names for n in ('loss', 'add_acc', 'remove_acc')):
            loss, aacc, racc = lrg_forward.fetch_scalars(add, rmv, feeds['add_mask_pl'], feeds['remove_mask_pl'])
            out.update(loss=loss, add_acc=aacc, remove_acc=racc)
        if TRACE is not None:
            TRACE.append(dict(inlier_crc=_crc(feeds['inlier_pl']), neighbor_crc=_crc(feeds['neighbor_pl']),
                              add_crc=_crc(add), remove_crc=_crc(rmv)))
        return out
