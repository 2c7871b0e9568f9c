"""tf.variable_scope + tf.layers.{conv2d, separable_conv2d, batch_normalization, max_pooling2d, dense} for the numpy
TensorFlow stand-in: enough of TF 1.6's layer semantics that the reference's NETWORK BUILDERS (net/xception_body.py
XceptionBody / get_rpn / large_sep_kernel / get_head, net/resnet_v2.py, net/xdet_body.py) run unmodified and produce
(a) the variable names / shapes TF would create and (b) forward values.  TEST INFRASTRUCTURE ONLY.

What is restated here rather than executed from TensorFlow (it is not installable offline), from the TF 1.6 docs:
  * variable scopes: an explicit name is used as is; ``default_name`` is uniquified among the scopes opened so far
    under the same parent ("conv2d", "conv2d_1", ...), counts of sub-scopes are forgotten when the parent closes;
  * 'SAME' padding: out = ceil(n / s), total = max((out - 1) * s + (k - 1) * d + 1 - n, 0), floor(total / 2) in front;
    max-pooling ignores padded cells;
  * kernels are [kh, kw, in, out]; depthwise [kh, kw, in, multiplier]; dense [in, out];
  * inference batch-norm y = gamma * (x - mean) / sqrt(var + eps) + beta; training=True uses the batch mean and the
    biased batch variance.
Arithmetic is numpy float64 over sliding-window views, rounded to float32 per layer -- deliberately NOT torch, so the
goldens are independent of the library oracle/net.py is written in.

Variables are created on first use by ``VARIABLE_FACTORY(full_name, shape)`` (default: oracle.net.seeded_variable,
a deterministic function of the name) and recorded, in creation order, in ``VARIABLES``.
"""
import collections
import contextlib

import numpy as np

VARIABLES = collections.OrderedDict()
VARIABLE_FACTORY = None
AUTO_REUSE = "auto_reuse"

_scope = []          # names of the open variable scopes
_scope_counts = {}   # full scope name -> times opened (the store's variable_scope_count)


def reset_variables():
    VARIABLES.clear()
    _scope_counts.clear()
    del _scope[:]


def _full(name):
    return "/".join(_scope + [name]) if name else "/".join(_scope)


def _unique(default_name):
    if _scope_counts.get(_full(default_name), 0) == 0:
        return default_name
    idx = 1
    while _scope_counts.get(_full("%s_%d" % (default_name, idx)), 0) > 0:
        idx += 1
    return "%s_%d" % (default_name, idx)


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None, **kw):
    name = name_or_scope if name_or_scope is not None else _unique(default_name)
    full = _full(name)
    _scope_counts[full] = _scope_counts.get(full, 0) + 1
    _scope.append(name)
    try:
        yield full
    finally:
        _scope.pop()
        for k in list(_scope_counts):
            if k.startswith(full + "/"):
                _scope_counts[k] = 0


def _variable(leaf, shape):
    name = _full(leaf)
    shape = tuple(int(s) for s in shape)
    if name not in VARIABLES:
        factory = VARIABLE_FACTORY
        if factory is None:
            from oracle.net import seeded_variable as factory
        VARIABLES[name] = np.asarray(factory(name, shape), np.float32).reshape(shape)
    assert VARIABLES[name].shape == shape, (name, VARIABLES[name].shape, shape)
    return VARIABLES[name].astype(np.float64)


def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def _same(n, k, s, d):
    total = max((-(-n // s) - 1) * s + (k - 1) * d + 1 - n, 0)
    return total // 2, total - total // 2


def _to_nhwc(x, data_format):
    x = np.asarray(x, np.float64)
    return x.transpose(0, 2, 3, 1) if data_format == "channels_first" else x


def _from_nhwc(y, data_format):
    from . import _t
    y = y.astype(np.float32)
    return _t(y.transpose(0, 3, 1, 2) if data_format == "channels_first" else y)


def _windows(x, k, s, d, padding, fill):
    """x [N,H,W,C] -> view [N,Ho,Wo,C,kh,kw] of the (strided, dilated) k windows under TF padding."""
    (kh, kw), (sh, sw), (dh, dw) = k, s, d
    if padding.upper() == "SAME":
        ph, pw = _same(x.shape[1], kh, sh, dh), _same(x.shape[2], kw, sw, dw)
        x = np.pad(x, ((0, 0), ph, pw, (0, 0)), mode="constant", constant_values=fill)
    else:
        assert padding.upper() == "VALID", padding
    win = np.lib.stride_tricks.sliding_window_view(x, ((kh - 1) * dh + 1, (kw - 1) * dw + 1), axis=(1, 2))
    return win[:, ::sh, ::sw, :, ::dh, ::dw]


def _finish(y, activation):
    return y if activation is None else np.asarray(activation(y.astype(np.float32)), np.float64)


def conv2d(inputs, filters, kernel_size, strides=(1, 1), padding="valid", data_format="channels_last",
           dilation_rate=(1, 1), activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
           name=None, reuse=None, **kw):
    x = _to_nhwc(inputs, data_format)
    k, s, d = _pair(kernel_size), _pair(strides), _pair(dilation_rate)
    with variable_scope(name, default_name="conv2d"):
        kernel = _variable("kernel", k + (x.shape[3], filters))
        bias = _variable("bias", (filters,)) if use_bias else None
    y = np.tensordot(_windows(x, k, s, d, padding, 0.0), kernel, axes=([3, 4, 5], [2, 0, 1]))
    if bias is not None:
        y = y + bias
    return _from_nhwc(_finish(y, activation), data_format)


def separable_conv2d(inputs, filters, kernel_size, strides=(1, 1), padding="valid", data_format="channels_last",
                     dilation_rate=(1, 1), depth_multiplier=1, activation=None, use_bias=True, name=None, reuse=None,
                     **kw):
    assert depth_multiplier == 1
    x = _to_nhwc(inputs, data_format)
    k, s, d = _pair(kernel_size), _pair(strides), _pair(dilation_rate)
    with variable_scope(name, default_name="separable_conv2d"):
        depthwise = _variable("depthwise_kernel", k + (x.shape[3], 1))
        pointwise = _variable("pointwise_kernel", (1, 1, x.shape[3], filters))
        bias = _variable("bias", (filters,)) if use_bias else None
    y = np.einsum("nhwcij,ijc->nhwc", _windows(x, k, s, d, padding, 0.0), depthwise[..., 0])
    y = y @ pointwise[0, 0]
    if bias is not None:
        y = y + bias
    return _from_nhwc(_finish(y, activation), data_format)


def batch_normalization(inputs, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, training=False,
                        name=None, reuse=None, fused=None, **kw):
    from . import _t
    x = np.asarray(inputs, np.float64)
    axis = axis % x.ndim
    c = x.shape[axis]
    with variable_scope(name, default_name="batch_normalization"):
        gamma = _variable("gamma", (c,)) if scale else np.ones(c)
        beta = _variable("beta", (c,)) if center else np.zeros(c)
        mean = _variable("moving_mean", (c,))
        var = _variable("moving_variance", (c,))
    if training:
        red = tuple(a for a in range(x.ndim) if a != axis)
        mean, var = x.mean(axis=red), x.var(axis=red)
    sh = [1] * x.ndim
    sh[axis] = c
    y = (x - mean.reshape(sh)) / np.sqrt(var.reshape(sh) + epsilon) * gamma.reshape(sh) + beta.reshape(sh)
    return _t(y.astype(np.float32))


def max_pooling2d(inputs, pool_size, strides, padding="valid", data_format="channels_last", name=None):
    x = _to_nhwc(inputs, data_format)
    y = _windows(x, _pair(pool_size), _pair(strides), (1, 1), padding, -np.inf).max(axis=(4, 5))
    return _from_nhwc(y, data_format)


def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name=None,
          reuse=None, **kw):
    from . import _t
    x = np.asarray(inputs, np.float64)
    with variable_scope(name, default_name="dense"):
        kernel = _variable("kernel", (x.shape[-1], units))
        bias = _variable("bias", (units,)) if use_bias else None
    y = x @ kernel
    if bias is not None:
        y = y + bias
    return _t(_finish(y, activation).astype(np.float32))
