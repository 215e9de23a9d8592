"""The LrgNet forward evaluated by INTERPRETING the reference's own shipped graph, op by op -- TEST INFRASTRUCTURE
(see oracle/__init__.py).

/root/reference/models/lrgnet_model5.ckpt.meta is the MetaGraphDef TensorFlow 1.14 wrote when the reference's
``LrgNet(100, 1, 512, 512, 13)`` (learn_region_grow_util.py:76-189) was trained: a reference-held artefact that records
the wiring of the forward -- which tensor feeds which op, the concat order ``[pooled | conv[1]]`` (:128-135), which layer
feeds the heads -- independently of anybody's reading of the Python.  This module walks that GraphDef backwards from the
two output tensors (``BiasAdd_12`` = add_output, ``BiasAdd_15`` = remove_output; found by name of their bias variable, not
hard-coded) and evaluates the dozen op types it meets in numpy:

    Placeholder  VariableV2 / Identity  Const  ExpandDims  Conv2D (1 x 1, NHWC, VALID, stride 1)  Squeeze  BiasAdd  Relu
    Max  ConcatV2  Reshape  Tile

Variables come from the checkpoint-V2 files next to the .meta (read with learn_region_grow_b200/ckpt.py).  The graph is
fixed at batch 100 (the Reshape / Tile shape constants hold 100), so it is evaluated at batch 100.

``oracle/make_golden.py`` uses it to write ``tests/golden/forward_graphdef.npz`` (driver tiles in, logits out);
``tests/test_forward_pin.py`` holds ``oracle/lrg_forward.py`` to those vectors, and -- where /root/reference exists --
re-runs the interpreter live.  tensorboard's ``meta_graph_pb2`` (in the image) parses the protobuf; TensorFlow itself
is not needed.
"""
import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def load_meta_graph(meta_path):
    from tensorboard.compat.proto import meta_graph_pb2
    m = meta_graph_pb2.MetaGraphDef()
    with open(meta_path, 'rb') as f:
        m.ParseFromString(f.read())
    return m


def _const_value(node):
    t = node.attr['value'].tensor
    dt = _DTYPES[t.dtype]
    shape = [d.size for d in t.tensor_shape.dim]
    if t.tensor_content:
        return np.frombuffer(t.tensor_content, dtype=dt).reshape(shape).copy()
    vals = list(t.float_val) or list(t.int_val) or list(t.int64_val) or list(t.double_val) or list(t.bool_val)
    n = int(np.prod(shape)) if shape else 1
    if len(vals) == 1 and n > 1:
        vals = vals * n
    return np.asarray(vals, dtype=dt).reshape(shape)


class GraphInterpreter:
    """Evaluates tensors of a TF1 GraphDef in numpy.  ``variables``: name -> array; ``dtype``: the float type to compute in."""

    def __init__(self, graph_def, variables, dtype=np.float32):
        self.nodes = {n.name: n for n in graph_def.node}
        self.variables = variables
        self.dtype = dtype
        self.ops_seen = {}

    def _f(self, a):
        a = np.asarray(a)
        return a.astype(self.dtype) if a.dtype.kind == 'f' else a

    def run(self, fetches, feed):
        memo = {k.split(':')[0]: self._f(v) for k, v in feed.items()}
        return [self._eval(f.split(':')[0], memo) for f in fetches]

    def _eval(self, name, memo):
        if name in memo:
            return memo[name]
        node = self.nodes[name]
        op = node.op
        self.ops_seen[op] = self.ops_seen.get(op, 0) + 1
        inp = [self._eval(i.split(':')[0].lstrip('^'), memo) for i in node.input if not i.startswith('^')]
        if op == 'Placeholder':
            raise KeyError('placeholder %s is not fed' % name)
        elif op == 'VariableV2':
            out = self._f(self.variables[name])
            shape = [d.size for d in node.attr['shape'].shape.dim]
            assert list(out.shape) == shape, (name, out.shape, shape)
        elif op == 'Const':
            out = self._f(_const_value(node))
        elif op == 'Identity':
            out = inp[0]
        elif op == 'ExpandDims':
            out = np.expand_dims(inp[0], int(inp[1]))
        elif op == 'Conv2D':
            x, w = inp
            assert node.attr['data_format'].s in (b'NHWC', b''), node.attr['data_format'].s
            assert list(node.attr['strides'].list.i) == [1, 1, 1, 1] and node.attr['padding'].s == b'VALID'
            assert all(d == 1 for d in (list(node.attr['dilations'].list.i) or [1]))
            kh, kw, cin, cout = w.shape
            assert kh == 1 and kw == 1 and x.shape[-1] == cin, (x.shape, w.shape)
            # a 1 x 1 VALID convolution in NHWC is a matrix product over the channel axis at every (n, h, w)
            out = x @ w[0, 0]
        elif op == 'Squeeze':
            dims = tuple(int(d) for d in node.attr['squeeze_dims'].list.i)
            out = np.squeeze(inp[0], axis=dims if dims else None)
        elif op == 'BiasAdd':
            assert node.attr['data_format'].s in (b'NHWC', b'')
            out = inp[0] + inp[1]
        elif op == 'Relu':
            out = np.maximum(inp[0], 0)
        elif op == 'Max':
            out = np.max(inp[0], axis=tuple(np.atleast_1d(inp[1]).tolist()), keepdims=bool(node.attr['keep_dims'].b))
        elif op == 'ConcatV2':
            out = np.concatenate(inp[:-1], axis=int(inp[-1]))
        elif op == 'Reshape':
            out = np.reshape(inp[0], [int(d) for d in inp[1]])
        elif op == 'Tile':
            out = np.tile(inp[0], [int(d) for d in inp[1]])
        else:
            raise NotImplementedError('op %s (%s) is not part of the forward' % (op, name))
        memo[name] = out
        return out


def find_forward_endpoints(graph_def):
    """(inlier placeholder, neighbor placeholder, add_output, remove_output) node names, located structurally: the outputs
    are the BiasAdd ops that read lrg_add_bias2 / lrg_remove_bias2 (learn_region_grow_util.py:145-149,158-162); the
    placeholders are found by walking back from the first layer of each branch (lrg_kernel0 / lrg_neighbor_kernel0)."""
    nodes = {n.name: n for n in graph_def.node}

    def consumer(var, op):
        read = var + '/read'
        hits = [n.name for n in graph_def.node if n.op == op and not n.name.startswith('gradients') and
                any(i.split(':')[0] == read for i in n.input)]
        assert len(hits) == 1, (var, op, hits)
        return hits[0]

    def placeholder_of(kernel_var):
        # kernel/read -> ExpandDims_1 -> Conv2D; the Conv2D's other input is ExpandDims(placeholder)
        exp = consumer(kernel_var, 'ExpandDims')
        conv = [n for n in graph_def.node if n.op == 'Conv2D' and exp in [i.split(':')[0] for i in n.input]]
        assert len(conv) == 1
        other = [i.split(':')[0] for i in conv[0].input if i.split(':')[0] != exp][0]
        src = nodes[other]
        while src.op != 'Placeholder':
            src = nodes[src.input[0].split(':')[0]]
        return src.name

    return (placeholder_of('lrg_kernel0'), placeholder_of('lrg_neighbor_kernel0'),
            consumer('lrg_add_bias2', 'BiasAdd'), consumer('lrg_remove_bias2', 'BiasAdd'))


class ShippedGraphForward:
    """``forward(inlier (B,512,13), neighbor (B,512,13)) -> (add (B,512,2), remove (B,512,2))`` through the shipped graph;
    B <= 100 (inputs are zero-padded to the graph's batch of 100 -- no op of the forward mixes batch rows)."""

    def __init__(self, ckpt_prefix, dtype=np.float64):
        from learn_region_grow_b200 import ckpt
        self.meta = load_meta_graph(ckpt_prefix + '.meta')
        tensors = ckpt.load_checkpoint(ckpt_prefix)
        self.interp = GraphInterpreter(self.meta.graph_def, tensors, dtype)
        self.inlier_pl, self.neighbor_pl, self.add_out, self.remove_out = find_forward_endpoints(self.meta.graph_def)
        shp = [d.size for d in self.interp.nodes[self.inlier_pl].attr['shape'].shape.dim]
        self.batch, self.n_points, self.feature_size = shp

    def forward(self, inlier, neighbor):
        B = len(inlier)
        assert B <= self.batch
        xi = np.zeros((self.batch, self.n_points, self.feature_size), np.float32)
        xj = np.zeros_like(xi)
        xi[:B] = inlier
        xj[:B] = neighbor
        add, rmv = self.interp.run([self.add_out, self.remove_out], {self.inlier_pl: xi, self.neighbor_pl: xj})
        return add[:B], rmv[:B]
