"""Tensor-core (tcgen05; 3xFP16 = the default, 3xTF32 = its fallback) forward vs the fp32-FMA kernels and the CPU oracle,
through the C ABI.

Stated tolerance: 2e-4 absolute on logits of magnitude <= ~15 against the float64 evaluation of the reference graph
(learn_region_grow_util.py:106-162) -- the same bar the fp32-FMA kernels are held to (tests/test_forward_gpu.py)."""
import numpy as np
import pytest

from oracle import lrg_driver, lrg_forward
from test_forward_gpu import _driver_tiles

pytestmark = pytest.mark.gpu
ATOL = 2e-4
TENSOR_MODES = pytest.mark.parametrize('mode', [2, 3], ids=['3xTF32', '3xFP16'])      # _lib.FORWARD_TENSOR, _lib.FORWARD_TENSOR_F16


def _engine(weights, mode, Ni=512, Nj=512, F=13):
    from learn_region_grow_b200.engine import Engine
    e = Engine(1, 1, Ni, Nj, F, 0, forward_mode=mode)
    e.load_weights(weights)
    return e


def test_auto_mode_is_tensor_for_the_full_model(golden_weights):
    from learn_region_grow_b200 import _lib
    e = _engine(golden_weights, 0)
    assert e.forward_mode() == _lib.FORWARD_TENSOR_F16
    e.set_forward_mode(_lib.FORWARD_TENSOR)
    assert e.forward_mode() == _lib.FORWARD_TENSOR
    e.close()
    from learn_region_grow_b200.engine import Engine
    lite = Engine(1, 1, 512, 512, 13, 1)
    lite.load_weights(lrg_forward.random_weights(13, 1, seed=1))
    assert lite.forward_mode() == _lib.FORWARD_FMA
    with pytest.raises(_lib.LrgError):
        lite.set_forward_mode(_lib.FORWARD_TENSOR)
    lite.close()


@TENSOR_MODES
def test_tensor_forward_matches_oracle_and_fma(golden_weights, mode):
    from learn_region_grow_b200 import _lib
    inlier, neighbor = _driver_tiles(golden_weights, n_tiles=8)
    tc, fma = _engine(golden_weights, mode), _engine(golden_weights, _lib.FORWARD_FMA)
    add_t, rmv_t = tc.forward(inlier, neighbor)
    add_f, rmv_f = fma.forward(inlier, neighbor)
    add64, rmv64 = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
    err_t = max(np.abs(add_t - add64).max(), np.abs(rmv_t - rmv64).max())
    err_f = max(np.abs(add_f - add64).max(), np.abs(rmv_f - rmv64).max())
    print('max |logit - f64|: tensor %s %.3e, fp32 FMA %.3e (|logit| max %.2f)' % ('3xTF32' if mode == 2 else '3xFP16', err_t, err_f, np.abs(add64).max()))
    assert err_t < ATOL and err_f < ATOL
    flips = 0
    for b in range(len(inlier)):
        for got, ref, s in ((add_t[b], add64[b], 7), (rmv_t[b], rmv64[b], 8)):
            u = np.random.RandomState(s).random_sample(512)
            c_got, c_ref = lrg_driver.confidence(got), lrg_driver.confidence(ref.astype(np.float32))
            differ = (u < c_got) != (u < c_ref)
            assert np.all(np.abs(u[differ] - c_ref[differ]) < 1e-4)
            flips += int(differ.sum())
    assert flips <= 2
    tc.close()
    fma.close()


@TENSOR_MODES
def test_tensor_forward_random_inputs_and_batching(golden_weights, mode):
    rng = np.random.RandomState(11)
    inlier = rng.randn(7, 512, 13).astype(np.float32)
    neighbor = rng.randn(7, 512, 13).astype(np.float32)
    tc = _engine(golden_weights, mode)
    add, rmv = tc.forward(inlier, neighbor)
    add64, rmv64 = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
    scale = max(np.abs(add64).max(), np.abs(rmv64).max())
    assert np.abs(add - add64).max() < 1e-5 * scale + ATOL and np.abs(rmv - rmv64).max() < 1e-5 * scale + ATOL
    for b in (0, 3, 6):
        a1, r1 = tc.forward(inlier[b:b + 1], neighbor[b:b + 1])
        np.testing.assert_array_equal(a1[0], add[b])
        np.testing.assert_array_equal(r1[0], rmv[b])
    tc.close()


@TENSOR_MODES
@pytest.mark.parametrize('F,Ni,Nj', [(9, 256, 256), (6, 128, 128), (12, 512, 256), (13, 200, 77), (16, 512, 512)])
def test_tensor_forward_shapes(F, Ni, Nj, mode):
    """feature_size / set-size variants of the full model (test_region_grow.py:70-83), including ragged last tiles."""
    w = lrg_forward.random_weights(F, 0, seed=100 + F)
    e = _engine(w, mode, Ni, Nj, F)
    rng = np.random.RandomState(2)
    inlier = rng.randn(3, Ni, F).astype(np.float32)
    neighbor = rng.randn(3, Nj, F).astype(np.float32)
    add, rmv = e.forward(inlier, neighbor)
    add64, rmv64 = lrg_forward.forward(w, inlier, neighbor, lite=0, dtype=np.float64)
    assert add.shape == (3, Nj, 2) and rmv.shape == (3, Ni, 2)
    assert np.abs(add - add64).max() < ATOL and np.abs(rmv - rmv64).max() < ATOL
    e.close()


def test_fp16_range_overflow_falls_back_to_tf32(golden_weights):
    """3xFP16 is valid while every activation stays below 65,000 (the shipped model on real features: < 500).  Inputs scaled far
    outside the feature distribution push activations beyond that: LRG_FORWARD_TENSOR_F16 reports LRG_E_RANGE, the default
    mode repeats the call with 3xTF32 and returns its result."""
    from learn_region_grow_b200 import _lib
    inlier, neighbor = _driver_tiles(golden_weights, n_tiles=2)
    inlier, neighbor = (inlier * 20000.0).astype(np.float32), (neighbor * 20000.0).astype(np.float32)   # inputs < 4e4, layer 1 > 1e5
    add64, rmv64 = lrg_forward.forward(golden_weights, inlier, neighbor, dtype=np.float64)
    f16 = _engine(golden_weights, _lib.FORWARD_TENSOR_F16)
    with pytest.raises(_lib.LrgError, match='fp16 range'):
        f16.forward(inlier, neighbor)
    f16.close()
    auto, tf32 = _engine(golden_weights, _lib.FORWARD_AUTO), _engine(golden_weights, _lib.FORWARD_TENSOR)
    add_a, rmv_a = auto.forward(inlier, neighbor)
    add_t, rmv_t = tf32.forward(inlier, neighbor)
    np.testing.assert_array_equal(add_a, add_t)
    np.testing.assert_array_equal(rmv_a, rmv_t)
    over, fallbacks = auto.range_overflow()
    assert fallbacks == 1 and not over
    scale = max(np.abs(add64).max(), np.abs(rmv64).max())
    assert np.abs(add_a - add64).max() < 1e-5 * scale and np.abs(rmv_a - rmv64).max() < 1e-5 * scale
    # in range again: no further fallback, and the 3xFP16 result
    small_i, small_n = _driver_tiles(golden_weights, n_tiles=2)
    auto.forward(small_i, small_n)
    assert auto.range_overflow() == (False, 1)
    auto.close()
    tf32.close()
