"""LrgNet forward: CUDA kernels (through the C ABI) vs the CPU oracle on the golden weights."""
import numpy as np
import pytest

from oracle import lrg_driver, lrg_forward
from util_rooms import golden_room

pytestmark = pytest.mark.gpu

# Stated tolerance: the kernels and the oracle are both fp32 but sum in different orders (and the kernels factor head
# layer 0); against the float64 evaluation of the reference graph both sit at ~1e-5 on logits of magnitude <= ~15.
ATOL_F64 = 2e-4
ATOL_F32 = 2e-4


def _driver_tiles(weights, n_tiles=6):
    """Real tiles: run the oracle driver on the golden room and keep a few steps' inputs."""
    points, order = golden_room(1000)
    tiles = []

    def fwd(inlier, neighbor):
        add, rmv = lrg_forward.forward(weights, inlier, neighbor)
        if len(tiles) < 64:
            tiles.append((inlier.copy(), neighbor.copy()))
        return add, rmv

    g = lrg_driver.RoomGrower(points, order, fwd, lrg_driver.PhiloxRng(3))
    for seed_id in np.arange(len(points))[order]:
        if g.visited[seed_id]:
            continue
        g.grow_region(seed_id)
        if len(tiles) >= 64:
            break
    pick = np.linspace(0, len(tiles) - 1, n_tiles).astype(int)
    return np.concatenate([tiles[i][0] for i in pick]), np.concatenate([tiles[i][1] for i in pick])


@pytest.fixture(scope='module')
def engine(golden_weights):
    from learn_region_grow_b200.engine import Engine
    e = Engine(1, 1, 512, 512, 13, 0)
    e.load_weights(golden_weights)
    yield e
    e.close()


def test_forward_matches_oracle_on_driver_tiles(engine, golden_weights):
    inlier, neighbor = _driver_tiles(golden_weights)
    add, rmv = engine.forward(inlier, neighbor)
    add64, rmv64 = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
    add32, rmv32 = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float32)
    assert np.abs(add - add64).max() < ATOL_F64 and np.abs(rmv - rmv64).max() < ATOL_F64
    assert np.abs(add - add32).max() < ATOL_F32 and np.abs(rmv - rmv32).max() < ATOL_F32
    # same add/remove masks for the same uniforms (test_region_grow.py:262-267), except draws that fall inside the tolerance band
    for b in range(len(inlier)):
        for got, ref, s in ((add[b], add64[b], 7), (rmv[b], rmv64[b], 8)):
            u = np.random.RandomState(s).random_sample(512)
            c_got, c_ref = lrg_driver.confidence(got), lrg_driver.confidence(ref.astype(np.float32))
            differ = (u < c_got) != (u < c_ref)
            assert np.all(np.abs(u[differ] - c_ref[differ]) < 1e-4)
            assert differ.sum() <= 1


def test_forward_batch_and_single_agree(engine, golden_weights):
    rng = np.random.RandomState(0)
    inlier = rng.randn(5, 512, 13).astype(np.float32)
    neighbor = rng.randn(5, 512, 13).astype(np.float32)
    add, rmv = engine.forward(inlier, neighbor)
    for b in range(5):
        a1, r1 = engine.forward(inlier[b:b + 1], neighbor[b:b + 1])
        np.testing.assert_array_equal(a1[0], add[b])      # tiles are independent: bit-identical regardless of batching
        np.testing.assert_array_equal(r1[0], rmv[b])
    add64, rmv64 = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
    scale = max(np.abs(add64).max(), np.abs(rmv64).max())
    assert np.abs(add - add64).max() < 1e-5 * scale + ATOL_F64 and np.abs(rmv - rmv64).max() < 1e-5 * scale + ATOL_F64


@pytest.mark.parametrize('lite,F,Ni,Nj', [(1, 13, 512, 512), (2, 13, 512, 512), (0, 9, 256, 256), (0, 6, 128, 128), (0, 12, 512, 256)])
def test_forward_variants(lite, F, Ni, Nj):
    """lite / feature_size / set-size variants of the constructor (util.py:77-85, test_region_grow.py:70-83)."""
    from learn_region_grow_b200.engine import Engine
    w = lrg_forward.random_weights(F, lite, seed=lite * 10 + F)
    e = Engine(1, 1, Ni, Nj, F, lite)
    e.load_weights(w)
    rng = np.random.RandomState(1)
    inlier = rng.randn(3, Ni, F).astype(np.float32)
    neighbor = rng.randn(3, Nj, F).astype(np.float32)
    add, rmv = e.forward(inlier, neighbor)
    add64, rmv64 = lrg_forward.forward(w, inlier, neighbor, lite=lite, dtype=np.float64)
    assert add.shape == (3, Nj, 2) and rmv.shape == (3, Ni, 2)
    assert np.abs(add - add64).max() < ATOL_F64 and np.abs(rmv - rmv64).max() < ATOL_F64
    e.close()


def test_forward_errors(golden_weights):
    from learn_region_grow_b200.engine import Engine
    from learn_region_grow_b200._lib import LrgError
    e = Engine(1, 1, 512, 512, 13, 0)
    with pytest.raises(LrgError):          # restore() must come first, like an uninitialised TF variable
        e.forward(np.zeros((1, 512, 13), np.float32), np.zeros((1, 512, 13), np.float32))
    e.load_weights(golden_weights)
    with pytest.raises(ValueError):
        e.forward(np.zeros((1, 256, 13), np.float32), np.zeros((1, 512, 13), np.float32))
    with pytest.raises(KeyError):
        e.load_weights({'lrg_kernel0': golden_weights['lrg_kernel0']})
    e.close()
    with pytest.raises(LrgError):
        Engine(1, 1, 1024, 512, 13, 0)


def test_dropin_session_run(golden_weights, tmp_path):
    """The reference's call sequence (test_region_grow.py:86-94,257-258) through the drop-in modules."""
    import os
    import sys
    from learn_region_grow_b200 import ckpt
    dropin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'learn_region_grow_b200', 'dropin')
    sys.path[:0] = [dropin]
    sys.path.append(os.path.join(dropin, 'standins'))
    try:
        for m in ('tensorflow', 'learn_region_grow_util', 'h5py'):
            sys.modules.pop(m, None)
        import tensorflow as tf
        from learn_region_grow_util import LrgNet
        ckpt.save_checkpoint(str(tmp_path / 'lrgnet_model5.ckpt'), golden_weights)
        tf.compat.v1.reset_default_graph()
        config = tf.compat.v1.ConfigProto()
        config.gpu_options.allow_growth = True
        sess = tf.compat.v1.Session(config=config)
        net = LrgNet(1, 1, 512, 512, 13, None)
        tf.compat.v1.train.Saver().restore(sess, str(tmp_path / 'lrgnet_model5.ckpt'))
        rng = np.random.RandomState(5)
        feed = {net.inlier_pl: rng.randn(1, 512, 13).astype(np.float32), net.neighbor_pl: rng.randn(1, 512, 13).astype(np.float32),
                net.add_mask_pl: rng.randint(0, 2, (1, 512)).astype(np.int32), net.remove_mask_pl: rng.randint(0, 2, (1, 512)).astype(np.int32)}
        ls, add, add_acc, rmv, rmv_acc = sess.run([net.loss, net.add_output, net.add_acc, net.remove_output, net.remove_acc], feed)
        a32, r32 = lrg_forward.forward(golden_weights, feed[net.inlier_pl], feed[net.neighbor_pl])
        ls_o, aacc_o, racc_o = lrg_forward.fetch_scalars(a32, r32, feed[net.add_mask_pl], feed[net.remove_mask_pl])
        assert np.abs(add - a32).max() < 1e-5 * np.abs(a32).max() + ATOL_F32
        assert abs(float(ls) - float(ls_o)) < 1e-3 * max(1.0, abs(float(ls_o)))
        assert abs(float(add_acc) - float(aacc_o)) <= 2 / 512 and abs(float(rmv_acc) - float(racc_o)) <= 2 / 512
    finally:
        sys.path.remove(dropin)
        sys.path.remove(os.path.join(dropin, 'standins'))
        for m in ('tensorflow', 'learn_region_grow_util', 'h5py'):
            sys.modules.pop(m, None)
