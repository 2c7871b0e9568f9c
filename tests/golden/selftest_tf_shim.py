#!/usr/bin/env python
"""Self-test of the numpy TensorFlow stand-in (oracle/tf_shim) against the worked examples of the TensorFlow 1.x API
documentation (the examples are quoted from the docstrings of the ops named in each block).  The goldens are only as
good as the stand-in's reading of those semantics; this pins that reading.  Run by tests/test_tf_shim.py in a fresh
interpreter (the stand-in must never be importable as `tensorflow` inside the test process itself).
    python tests/golden/selftest_tf_shim.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), ROOT]

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the stand-in)
from tensorflow import _layers  # noqa: E402


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(a, b), (a, b)


def main():
    assert "tf_shim" in tf.__file__
    # ---- dtype rules of convert_to_tensor: python floats -> float32, python ints -> int32 -----------------------
    assert np.asarray(tf.constant(1.5)).dtype == np.float32 and np.asarray(tf.constant([1, 2])).dtype == np.int32
    assert np.asarray(tf.constant(np.float64(1.5))).dtype == np.float64            # tensors keep their dtype
    assert np.asarray(tf.cast(tf.constant([1.8, -1.8]), tf.int32)).tolist() == [1, -1]   # truncation
    # ---- tf.pad ---------------------------------------------------------------------------------------------------
    t = tf.constant([[1, 2, 3], [4, 5, 6]])
    eq(tf.pad(t, tf.constant([[1, 1], [2, 2]]), "CONSTANT"),
       [[0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 2, 3, 0, 0], [0, 0, 4, 5, 6, 0, 0], [0, 0, 0, 0, 0, 0, 0]])
    # ---- tf.tile / tf.stack / tf.unstack / tf.concat / tf.split ----------------------------------------------------
    eq(tf.tile(tf.constant([1, 2, 3]), [2]), [1, 2, 3, 1, 2, 3])
    x, y, z = tf.constant([1, 4]), tf.constant([2, 5]), tf.constant([3, 6])
    eq(tf.stack([x, y, z]), [[1, 4], [2, 5], [3, 6]])
    eq(tf.stack([x, y, z], axis=1), [[1, 2, 3], [4, 5, 6]])
    assert [np.asarray(u).tolist() for u in tf.unstack(tf.constant([[1, 2], [3, 4]]), axis=1)] == [[1, 3], [2, 4]]
    t1, t2 = tf.constant([[1, 2, 3], [4, 5, 6]]), tf.constant([[7, 8, 9], [10, 11, 12]])
    eq(tf.concat([t1, t2], 0), [[1, 2, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]])
    eq(tf.concat([t1, t2], 1), [[1, 2, 3, 7, 8, 9], [4, 5, 6, 10, 11, 12]])
    parts = tf.split(tf.constant(np.arange(30).reshape(5, 6)), [1, 2, 3], axis=1)
    assert [np.asarray(p).shape for p in parts] == [(5, 1), (5, 2), (5, 3)]
    # ---- tf.gather / gather_nd / boolean_mask / where --------------------------------------------------------------
    p = tf.constant([[0, 1], [2, 3], [4, 5]])
    eq(tf.gather(p, tf.constant([2, 0])), [[4, 5], [0, 1]])
    eq(tf.gather(p, tf.constant([1]), axis=1), [[1], [3], [5]])
    eq(tf.gather(tf.constant(np.arange(24).reshape(2, 3, 4)), tf.constant([[0, 2], [1, 1]]), axis=1).shape, (2, 2, 2, 4))
    eq(tf.gather_nd(tf.constant([['a', 'b'], ['c', 'd']]), tf.constant([[0, 0], [1, 1]])), ['a', 'd'])
    eq(tf.boolean_mask(tf.constant([0, 1, 2, 3]), np.array([True, False, True, False])), [0, 2])
    eq(tf.boolean_mask(tf.constant([[1, 2], [3, 4], [5, 6]]), np.array([True, False, True])), [[1, 2], [5, 6]])
    eq(tf.where(tf.constant([True, False, True])), [[0], [2]])                       # coordinates, int64, [n, rank]
    assert np.asarray(tf.where(tf.constant([True, False]))).dtype == np.int64
    eq(tf.where(tf.constant([True, False]), tf.constant([1, 2]), tf.constant([9, 8])), [1, 8])
    # ---- tf.one_hot -------------------------------------------------------------------------------------------------
    eq(tf.one_hot(tf.constant([0, 1, 2]), 3), np.eye(3, dtype=np.float32))
    eq(tf.one_hot(tf.constant([0, 2, -1, 1]), 3, on_value=5.0, off_value=0.0),
       [[5, 0, 0], [0, 0, 5], [0, 0, 0], [0, 5, 0]])
    eq(tf.one_hot(1, 3, on_value=False, off_value=True, dtype=tf.bool), [True, False, True])   # scalar index
    # ---- reductions / scans / elementwise ---------------------------------------------------------------------------
    x = tf.constant([[1., 1., 1.], [1., 1., 1.]])
    assert float(tf.reduce_sum(x)) == 6 and np.asarray(tf.reduce_sum(x, 0)).tolist() == [2, 2, 2]
    eq(tf.reduce_sum(x, 1, keepdims=True), [[3.], [3.]])
    eq(tf.cumsum(tf.constant([1, 2, 3])), [1, 3, 6])
    eq(tf.scan(lambda a, b: a + b, tf.constant([1, 2, 3, 4, 5, 6])), [1, 3, 6, 10, 15, 21])
    eq(tf.reverse(tf.constant([[1, 2], [3, 4]]), axis=[0]), [[3, 4], [1, 2]])
    eq(tf.round(tf.constant([0.9, 2.5, 2.3, 1.5, -4.5])), [1.0, 2.0, 2.0, 2.0, -4.0])   # half to even
    eq(tf.clip_by_value(tf.constant([-1., 0.5, 3.]), 0., 1.), [0., 0.5, 1.])
    eq(tf.floormod(tf.constant([7, -7]), 3), [1, 2])                                 # sign of the divisor
    eq(tf.floor_div(tf.constant([7, -7]), 2), [3, -4])
    eq(tf.argmax(tf.constant([[1, 9, 3], [8, 2, 7]]), axis=1), [1, 0])
    # ---- tf.nn.top_k: 'If two elements are equal, the lower-index element appears first' ----------------------------
    v, i = tf.nn.top_k(tf.constant([1., 3., 3., 2., 3.]), k=4)
    assert np.asarray(v).tolist() == [3., 3., 3., 2.] and np.asarray(i).tolist() == [1, 2, 4, 3]
    v, i = tf.nn.top_k(tf.constant([[1., 2.], [4., 3.]]), k=1)
    assert np.asarray(i).tolist() == [[1], [0]]
    # ---- control flow ------------------------------------------------------------------------------------------------
    i, = tf.while_loop(lambda i: tf.less(i, 10), lambda i: [tf.add(i, 1)], [tf.constant(0)])
    assert int(i) == 10
    assert int(tf.cond(tf.constant(True), lambda: tf.constant(1), lambda: tf.constant(2))) == 1
    eq(tf.map_fn(lambda e: e * e, tf.constant([1, 2, 3, 4, 5, 6])), [1, 4, 9, 16, 25, 36])
    a, b = tf.map_fn(lambda e: (e[0] + e[1], e[0] - e[1]), [tf.constant([1, 2]), tf.constant([5, 7])])
    assert np.asarray(a).tolist() == [6, 9] and np.asarray(b).tolist() == [-4, -5]
    ta = tf.TensorArray(tf.float32, size=0, dynamic_size=True)
    ta = ta.write(0, tf.constant([1., 2.])).write(1, tf.constant([3., 4.]))
    eq(ta.stack(), [[1., 2.], [3., 4.]])
    # ---- variable scopes: explicit names as given; default names uniquified per parent ('conv2d', 'conv2d_1') --------
    _layers.reset_variables()
    img = tf.constant(np.ones((1, 5, 5, 2), np.float32))
    with tf.variable_scope("m"):
        tf.layers.conv2d(img, 3, 3)
        tf.layers.conv2d(img, 3, 3)
        with tf.variable_scope("head"):
            tf.layers.conv2d(img, 3, 1, use_bias=False)
        tf.layers.conv2d(img, 4, 1, name="conv2d_7")
        tf.layers.batch_normalization(img)
    assert list(_layers.VARIABLES) == [
        "m/conv2d/kernel", "m/conv2d/bias", "m/conv2d_1/kernel", "m/conv2d_1/bias", "m/head/conv2d/kernel",
        "m/conv2d_7/kernel", "m/conv2d_7/bias", "m/batch_normalization/gamma", "m/batch_normalization/beta",
        "m/batch_normalization/moving_mean", "m/batch_normalization/moving_variance"], list(_layers.VARIABLES)
    assert _layers.VARIABLES["m/conv2d/kernel"].shape == (3, 3, 2, 3)
    # ---- 'SAME' / 'VALID' output sizes: ceil(n / s) and ceil((n - k + 1) / s); 'SAME' pads the extra cell at the end ---
    _layers.reset_variables()
    _layers.VARIABLE_FACTORY = lambda name, shape: np.ones(shape, np.float32)
    x = tf.constant(np.ones((1, 7, 8, 1), np.float32))
    assert np.asarray(tf.layers.conv2d(x, 1, 3, strides=2, padding="same", use_bias=False)).shape == (1, 4, 4, 1)
    assert np.asarray(tf.layers.conv2d(x, 1, 3, strides=2, padding="valid", use_bias=False)).shape == (1, 3, 3, 1)
    y = np.asarray(tf.layers.conv2d(tf.constant(np.ones((1, 4, 4, 1), np.float32)), 1, 3, strides=2, padding="same",
                                    use_bias=False, name="c"))[0, :, :, 0]
    eq(y, [[9., 6.], [6., 4.]])            # 4 -> 2 with pad (0 before, 1 after): only the last window is clipped
    mp = np.asarray(tf.layers.max_pooling2d(tf.constant(-np.ones((1, 4, 4, 1), np.float32)), 3, 2, padding="same"))
    assert mp.shape == (1, 2, 2, 1) and (mp == -1).all()                            # padding never wins a max
    d = np.asarray(tf.layers.dense(tf.constant(np.ones((2, 3, 4), np.float32)), 5, name="d"))
    assert d.shape == (2, 3, 5) and (d == 5.0).all()                                # kernel ones [4,5] + bias ones
    _layers.VARIABLE_FACTORY = None
    print("tf_shim self-test ok")


if __name__ == "__main__":
    main()
