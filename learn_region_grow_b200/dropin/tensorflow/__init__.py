"""Minimal stand-in for the TensorFlow session API used by the reference's inference driver
(/root/reference/test_region_grow.py:68,86-93,257-258).  Only importable when the real TensorFlow is absent.

A "graph" is just the list of network objects built since the last ``reset_default_graph``; ``Session.run``
groups the requested fetch handles by owner and asks the owner to evaluate them with the fed arrays; the
``Saver`` reads checkpoint-V2 files directly (learn_region_grow_b200/ckpt.py).
"""
import types

float32 = 'float32'
int32 = 'int32'
__version__ = '0.0-lrg-b200-shim'


class _Graph:
    def __init__(self):
        self.nets = []


_default_graph = _Graph()


class Handle:
    """A fetchable / feedable tensor handle owned by a network object."""

    def __init__(self, owner, name, shape=None, dtype=None):
        self.owner, self.name, self.shape, self.dtype = owner, name, shape, dtype

    def __repr__(self):
        return '<Handle %s shape=%s>' % (self.name, self.shape)

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


def register_net(net):
    _default_graph.nets.append(net)


def _reset_default_graph():
    global _default_graph
    _default_graph = _Graph()


class _GpuOptions:
    allow_growth = False


class _ConfigProto:
    def __init__(self, **kw):
        self.gpu_options = _GpuOptions()
        self.allow_soft_placement = False
        self.log_device_placement = False
        for k, v in kw.items():
            setattr(self, k, v)


class _Session:
    def __init__(self, target='', graph=None, config=None):
        self.graph = _default_graph
        self.config = config

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        feed_dict = feed_dict or {}
        owners = []
        for h in flist:
            if not isinstance(h, Handle):
                raise TypeError('Session.run can only fetch handles of the drop-in networks, got %r' % (h,))
            if h.owner not in owners:
                owners.append(h.owner)
        results = {}
        for owner in owners:
            feeds = {h.name: v for h, v in feed_dict.items() if isinstance(h, Handle) and h.owner is owner}
            names = [h.name for h in flist if h.owner is owner]
            results[id(owner)] = owner._evaluate(names, feeds)
        out = [results[id(h.owner)][h.name] for h in flist]
        return out[0] if single else out

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Saver:
    def __init__(self, *a, **kw):
        self.graph = _default_graph

    def restore(self, sess, save_path):
        from learn_region_grow_b200 import ckpt
        tensors = ckpt.load_checkpoint(save_path)
        for net in self.graph.nets:
            net._load_variables(tensors)

    def save(self, sess, save_path, **kw):
        from learn_region_grow_b200 import ckpt
        tensors = {}
        for net in self.graph.nets:
            tensors.update(net._variables())
        ckpt.save_checkpoint(save_path, tensors)
        return save_path


def _placeholder(dtype, shape=None, name=None):
    return Handle(None, name or 'placeholder', shape, dtype)


v1 = types.ModuleType('tensorflow.compat.v1')
v1.reset_default_graph = _reset_default_graph
v1.ConfigProto = _ConfigProto
v1.Session = _Session
v1.placeholder = _placeholder
v1.disable_eager_execution = lambda: None
v1.train = types.SimpleNamespace(Saver=_Saver)
compat = types.ModuleType('tensorflow.compat')
compat.v1 = v1
train = v1.train

import sys as _sys
_sys.modules.setdefault('tensorflow.compat', compat)
_sys.modules.setdefault('tensorflow.compat.v1', v1)
