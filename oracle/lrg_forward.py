"""CPU restatement of the LrgNet forward graph -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/learn_region_grow_util.py:76-162 *as written* (tile + concat of the pooled vector in
front of every per-point row, then the 1088-wide first head layer), so it shares no algebra with the CUDA
kernels, which factor the head as g.W[:1024] + h1.W[1024:].

Parity: weights pinned to the shipped checkpoint; activations "parity unpinned" (TensorFlow is absent and the
reference holds no activation vectors) -- anchored instead on the float64 evaluation of the same graph.
"""
import numpy as np


def channel_lists(lite):
    """learn_region_grow_util.py:77-85"""
    if lite == 0 or lite is None:
        return [64, 64, 64, 128, 512], [256, 128]
    if lite == 1:
        return [64, 64], [64]
    if lite == 2:
        return [64, 64, 256], [64, 64]
    raise ValueError('lite must be 0/None, 1 or 2')


def variable_shapes(feature_size=13, lite=0):
    """Names and shapes of the trainable variables in graph-construction order (util.py:106-162)."""
    conv, conv2 = channel_lists(lite)
    out = []
    for prefix in ('lrg_', 'lrg_neighbor_'):
        for i, c in enumerate(conv):
            cin = feature_size if i == 0 else conv[i - 1]
            out.append((prefix + 'kernel%d' % i, (1, cin, c)))
            out.append((prefix + 'bias%d' % i, (c,)))
    for prefix in ('lrg_add_', 'lrg_remove_'):
        for i, c in enumerate(conv2 + [2]):
            cin = conv[-1] * 2 + conv[1] if i == 0 else conv2[i - 1]
            out.append((prefix + 'kernel%d' % i, (1, cin, c)))
            out.append((prefix + 'bias%d' % i, (c,)))
    return out


def random_weights(feature_size=13, lite=0, seed=0):
    """Glorot-uniform kernels (util.py:108) with small random biases, for lite variants that ship no checkpoint."""
    rng = np.random.RandomState(seed)
    w = {}
    for name, shape in variable_shapes(feature_size, lite):
        if 'kernel' in name:
            lim = np.sqrt(6.0 / (shape[1] + shape[2]))
            w[name] = rng.uniform(-lim, lim, shape).astype(np.float32)
        else:
            w[name] = (rng.randn(*shape) * 0.1).astype(np.float32)
    return w


def forward(weights, inlier, neighbor, lite=0, dtype=np.float32):
    """inlier (B,Ni,F), neighbor (B,Nj,F) -> add_output (B,Nj,2), remove_output (B,Ni,2)."""
    conv_ch, conv2_ch = channel_lists(lite)
    W = {k: np.asarray(v, dtype=dtype) for k, v in weights.items() if k.startswith('lrg_')}
    x_i = np.asarray(inlier, dtype=dtype)
    x_j = np.asarray(neighbor, dtype=dtype)
    B, Ni, _ = x_i.shape
    Nj = x_j.shape[1]

    def branch(x, prefix):                                   # util.py:106-111 / :114-119
        acts = []
        for i in range(len(conv_ch)):
            x = np.maximum(x @ W[prefix + 'kernel%d' % i][0] + W[prefix + 'bias%d' % i], 0)
            acts.append(x)
        return acts

    conv = branch(x_i, 'lrg_')
    nconv = branch(x_j, 'lrg_neighbor_')
    pooled = np.concatenate([conv[-1].max(axis=1), nconv[-1].max(axis=1)], axis=1)   # util.py:122-125

    def head(local, prefix, n):                              # util.py:128-162
        z = np.concatenate([np.broadcast_to(pooled[:, None, :], (B, n, pooled.shape[1])), local], axis=2)
        for i in range(len(conv2_ch)):
            z = np.maximum(z @ W[prefix + 'kernel%d' % i][0] + W[prefix + 'bias%d' % i], 0)
        i = len(conv2_ch)
        return z @ W[prefix + 'kernel%d' % i][0] + W[prefix + 'bias%d' % i]

    remove_output = head(conv[1], 'lrg_remove_', Ni)
    add_output = head(nconv[1], 'lrg_add_', Nj)
    return add_output, remove_output


def _log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def fetch_scalars(add_output, remove_output, add_mask, remove_mask):
    """loss, add_acc, remove_acc as the graph defines them (util.py:165-186)."""
    add_output = np.asarray(add_output, np.float32)
    remove_output = np.asarray(remove_output, np.float32)
    add_mask = np.asarray(add_mask).astype(np.int64)
    remove_mask = np.asarray(remove_mask).astype(np.int64)
    ce_add = -np.take_along_axis(_log_softmax(add_output), add_mask[..., None], -1)[..., 0]
    ce_rmv = -np.take_along_axis(_log_softmax(remove_output), remove_mask[..., None], -1)[..., 0]
    add_loss = ce_add.mean(dtype=np.float32)
    pos = remove_mask.astype(bool)
    pos_loss = ce_rmv[pos].mean(dtype=np.float32) if pos.any() else np.float32(0)   # NaN -> 0 (:170-171)
    neg_loss = ce_rmv[~pos].mean(dtype=np.float32) if (~pos).any() else np.float32(0)
    add_acc = np.mean(add_output.argmax(-1) == add_mask, dtype=np.float32)
    remove_acc = np.mean(remove_output.argmax(-1) == remove_mask, dtype=np.float32)
    return np.float32(add_loss + pos_loss + neg_loss), np.float32(add_acc), np.float32(remove_acc)
