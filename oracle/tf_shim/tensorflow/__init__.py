"""A numpy-backed stand-in for the few dozen TensorFlow 1.x graph ops that the reference's *Python* code on the
Light-Head R-CNN path uses (preprocessing/anchor_manipulator.py, net/xception_body.py, net/resnet_v2.py,
net/xdet_body.py, utility/eval_helper.py), evaluated eagerly; tf.variable_scope and tf.layers live in _layers.py.  TEST INFRASTRUCTURE ONLY (like oracle/ref_shim for the C++ op): with this
directory first on sys.path, ``import tensorflow as tf`` inside the UNMODIFIED reference modules resolves here, so
their own code can be run in this container to mint golden vectors (tests/golden/make_tfpath_golden.py).

Semantics kept: float32 arithmetic (python scalars are weak, python float lists become float32, ints int32, as TF's
convert_to_tensor does), tf.nn.top_k = descending with ties to the lower index, tf.boolean_mask / gather / pad / tile /
while_loop / TensorArray as documented for TF 1.6.  NOT provided by TensorFlow's own code here: tf.nn.top_k,
tf.image.non_max_suppression (TF r1.6 NonMaxSuppressionV2, restated in oracle/proposals.py) and tf.random_shuffle
(replaced by the injected-key order the product uses) -- the goldens pin everything AROUND those three."""
import contextlib
import sys
import types

import numpy as np

class DType(object):
    """tf.float32 & co.: usable wherever TF takes a dtype, with .max / .min / .as_numpy_dtype."""

    def __init__(self, np_type):
        self.as_numpy_dtype = np_type
        self.name = np.dtype(np_type).name
        if np.issubdtype(np_type, np.integer):
            self.max, self.min = int(np.iinfo(np_type).max), int(np.iinfo(np_type).min)
        elif np.issubdtype(np_type, np.floating):
            self.max, self.min = float(np.finfo(np_type).max), float(np.finfo(np_type).min)

    def __eq__(self, other):
        return _np(other) == self.as_numpy_dtype if other is not None else False

    def __hash__(self):
        return hash(self.name)


def _np(dtype):
    """numpy type of a DType / numpy dtype / None."""
    if dtype is None:
        return None
    return dtype.as_numpy_dtype if isinstance(dtype, DType) else np.dtype(dtype).type


float32, float64, int32, int64, bool = DType(np.float32), DType(np.float64), DType(np.int32), DType(np.int64), DType(np.bool_)
uint8 = DType(np.uint8)


class _Shape(object):
    def __init__(self, dims):
        self._d = [int(d) for d in dims]

    def as_list(self):
        return list(self._d)

    def is_fully_defined(self):
        return True

    def with_rank(self, rank):
        assert len(self._d) == rank
        return self

    @property
    def ndims(self):
        return len(self._d)

    def __len__(self):
        return len(self._d)

    def __getitem__(self, i):
        return self._d[i]


class Tensor(np.ndarray):
    def get_shape(self):
        return _Shape(self.shape)

    def set_shape(self, shape):
        return None


def _t(x, dtype=None):
    """convert_to_tensor: python floats -> float32, python ints -> int32, arrays keep their dtype."""
    dtype = _np(dtype)
    if isinstance(x, (np.ndarray, np.generic)):  # tensors (and the numpy scalars reductions return) keep their dtype
        x = np.asarray(x)
        a = x if dtype is None else x.astype(dtype)
    else:
        a = np.asarray(x)
        if dtype is not None:
            a = a.astype(dtype)
        elif a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    return np.ascontiguousarray(a).view(Tensor) if a.ndim else np.asarray(a).reshape(()).view(Tensor)


def convert_to_tensor(x, dtype=None, name=None):
    return _t(x, dtype)


def constant(value, dtype=None, shape=None, name=None):
    return _t(value, dtype)


def cast(x, dtype, name=None):
    return _t(np.asarray(_t(x)).astype(_np(dtype)))


def to_float(x):
    return cast(x, float32)


def to_double(x):
    return cast(x, float64)


NAMED = {}        # tf.identity(x, name=...) results by name (how the reference labels its loss tensors)
GATHER_LOG = None  # a list -> every tf.gather call appends (params.shape, indices, axis)


def identity(x, name=None):
    if name is not None:
        NAMED[name] = x
    return x


def stop_gradient(x, name=None):
    return x


def Print(x, data=None, message=None, summarize=None, name=None):
    return x


@contextlib.contextmanager
def name_scope(*a, **k):
    yield


@contextlib.contextmanager
def device(*a, **k):
    yield


def shape(x, name=None, out_type=None):
    return _t(np.array(np.shape(x), dtype=np.int32))


def size(x, name=None, out_type=None):
    return np.int32(np.size(x))


def reshape(x, shp, name=None):
    return _t(np.reshape(_t(x), [int(s) for s in np.asarray(shp).reshape(-1)]))


def transpose(x, perm=None, name=None):
    return _t(np.transpose(_t(x), perm))


def expand_dims(x, axis=None, name=None, dim=None):
    return _t(np.expand_dims(_t(x), axis if axis is not None else dim))


def stack(values, axis=0, name=None):
    vals = [_t(v) for v in values]
    return _t(np.stack(vals, axis=axis))


def unstack(x, num=None, axis=0, name=None):
    x = _t(x)
    return [_t(np.take(x, i, axis=axis)) for i in range(x.shape[axis])]


def concat(values, axis, name=None):
    if len(values) == 1:  # TF returns identity(values[0]) for a single input, whatever the axis (array_ops.concat)
        return _t(values[0])
    return _t(np.concatenate([_t(v) for v in values], axis=axis))


def tile(x, multiples, name=None):
    return _t(np.tile(_t(x), [int(m) for m in np.asarray(multiples).reshape(-1)]))


def pad(x, paddings, mode='CONSTANT', name=None):
    assert mode == 'CONSTANT'
    p = np.asarray(paddings).astype(np.int64)
    return _t(np.pad(_t(x), [(int(a), int(b)) for a, b in p], mode='constant'))


def slice(x, begin, size, name=None):  # noqa: A001
    x = _t(x)
    idx = tuple(np.s_[int(b):(None if int(s) < 0 else int(b) + int(s))] for b, s in zip(begin, size))
    return _t(x[idx])


def gather(params, indices, name=None, axis=0):
    if GATHER_LOG is not None:
        GATHER_LOG.append((np.shape(params), np.array(indices), axis))
    return _t(np.take(_t(params), np.asarray(indices).astype(np.int64), axis=axis))


def boolean_mask(tensor, mask, name=None):
    return _t(_t(tensor)[np.asarray(mask).astype(np.bool_)])


def where(condition, x=None, y=None, name=None):
    if x is None and y is None:  # coordinates of the true elements, [n, rank] int64
        return _t(np.argwhere(np.asarray(condition)).astype(np.int64))
    return _t(np.where(np.asarray(condition), _t(x), _t(y)))


def range(*args, **kw):  # noqa: A001
    dtype = _np(kw.get('dtype', None))
    if dtype is None:  # dtype of the limits (python ints -> int32, tensors keep theirs), as tf.range infers it
        dtype = np.result_type(*[np.asarray(_t(a)).dtype for a in args]).type
    return _t(np.arange(*[int(a) for a in args]).astype(dtype))


def meshgrid(*args, **kw):
    return [_t(m) for m in np.meshgrid(*[_t(a) for a in args], indexing=kw.get('indexing', 'xy'))]


def zeros(shp, dtype=float32, name=None):
    return _t(np.zeros([int(s) for s in np.asarray(shp).reshape(-1)], dtype=_np(dtype)))


def zeros_like(x, dtype=None, name=None):
    return _t(np.zeros_like(_t(x), dtype=_np(dtype)))


def ones_like(x, dtype=None, name=None):
    return _t(np.ones_like(_t(x), dtype=_np(dtype)))


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
    """Negative / out-of-range indices give an all-`off` slice, as in TF."""
    idx = np.asarray(_t(indices)).astype(np.int64)
    depth = int(np.asarray(depth))
    dt = _np(dtype) or (np.asarray(on_value).dtype.type if on_value is not None else np.float32)
    on = np.asarray(1 if on_value is None else on_value).astype(dt)
    off = np.asarray(0 if off_value is None else off_value).astype(dt)
    hot = (idx[..., None] == np.arange(depth)).astype(np.bool_)          # new axis last
    out = np.where(hot, on, off).astype(dt)
    ax = -1 if axis is None else axis
    return _t(np.moveaxis(out, -1, ax if ax >= 0 else out.ndim + ax))


def clip_by_value(x, lo, hi, name=None):
    x = _t(x)
    return _t(np.clip(x, np.asarray(lo).astype(x.dtype), np.asarray(hi).astype(x.dtype)))


def gather_nd(params, indices, name=None):
    idx = np.asarray(_t(indices)).astype(np.int64)
    return _t(_t(params)[tuple(idx[..., i] for i in builtins_range(idx.shape[-1]))])


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    if not np.isscalar(num_or_size_splits):   # a list gives the SIZES of the parts (np.split wants the cut points)
        num_or_size_splits = np.cumsum(np.asarray(num_or_size_splits))[:-1]
    return [_t(p) for p in np.split(_t(value), num_or_size_splits, axis=axis)]


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    return _t(np.squeeze(_t(x), axis=axis if axis is not None else squeeze_dims))


def _binary(fn):
    def op(a, b, name=None):
        a, b = _t(a), _t(b)
        with np.errstate(divide="ignore", invalid="ignore"):  # e.g. safe_divide's 0/0, masked afterwards by tf.where
            return _t(fn(a, b))
    return op


maximum, minimum = _binary(np.maximum), _binary(np.minimum)
greater, less, equal = _binary(np.greater), _binary(np.less), _binary(np.equal)
greater_equal, less_equal = _binary(np.greater_equal), _binary(np.less_equal)
logical_and, logical_or = _binary(np.logical_and), _binary(np.logical_or)
divide = _binary(np.true_divide)
truediv = _binary(np.true_divide)
floor_div = _binary(np.floor_divide)
floormod = _binary(np.mod)


def logical_not(x, name=None):
    return _t(np.logical_not(_t(x)))


def exp(x, name=None):
    return _t(np.exp(_t(x)))


def sqrt(x, name=None):
    return _t(np.sqrt(_t(x)))


def log(x, name=None):
    with np.errstate(divide="ignore", invalid="ignore"):
        return _t(np.log(_t(x)))


def round(x, name=None):  # noqa: A001
    return _t(np.rint(_t(x)))  # half to even, as tf.round


def reduce_sum(x, axis=None, keepdims=False, name=None):
    return _t(np.sum(_t(x), axis=axis, keepdims=keepdims))


def reduce_max(x, axis=None, keepdims=False, name=None, keep_dims=None):
    return _t(np.max(_t(x), axis=axis, keepdims=keepdims if keep_dims is None else keep_dims))


def count_nonzero(x, axis=None, name=None):
    return np.int64(np.count_nonzero(_t(x), axis=axis))


def argmax(x, axis=None, name=None, output_type=None):
    return _t(np.asarray(np.argmax(_t(x), axis=axis)).astype(_np(output_type) or np.int64))  # first maximum, as TF


def cond(pred, true_fn=None, false_fn=None, name=None, **kw):
    return true_fn() if builtins_bool(pred) else false_fn()


def builtins_bool(x):
    return True if np.asarray(x).reshape(-1)[0] else False


MAP_INDEX = 0  # element index of the innermost running tf.map_fn (read by random_shuffle)


def map_fn(fn, elems, dtype=None, back_prop=True, infer_shape=True, parallel_iterations=None, name=None, **kw):
    global MAP_INDEX
    multi = isinstance(elems, (list, tuple))
    n = len(elems[0]) if multi else len(elems)
    outs = []
    saved = MAP_INDEX
    for i in builtins_range(n):
        MAP_INDEX = i
        outs.append(fn(type(elems)(_t(e[i]) for e in elems) if multi else _t(elems[i])))
    MAP_INDEX = saved
    if isinstance(outs[0], (list, tuple)):
        return type(outs[0])(_t(np.stack([np.asarray(o[j]) for o in outs])) for j in builtins_range(len(outs[0])))
    return _t(np.stack([np.asarray(o) for o in outs]))


import builtins as _b  # noqa: E402

builtins_range = _b.range


def while_loop(cond_fn, body, loop_vars, parallel_iterations=None, back_prop=False, **kw):
    v = list(loop_vars)
    while builtins_bool(cond_fn(*v)):
        v = list(body(*v))
    return v


class TensorArray(object):
    def __init__(self, dtype, size=0, dynamic_size=False, infer_shape=True, **kw):
        self.dtype, self.items, self.dynamic = _np(dtype), [None] * int(size), dynamic_size

    def write(self, i, value):
        if self.dynamic and int(i) >= len(self.items):
            self.items.extend([None] * (int(i) + 1 - len(self.items)))
        self.items[int(i)] = np.asarray(value).astype(self.dtype)
        return self

    def stack(self):
        return _t(np.stack(self.items)) if self.items else _t(np.zeros((0,), self.dtype))


# tf.random_shuffle(range(n)) := stable argsort of keys[:n] (the injected-key order of the product and the oracles).
# SHUFFLE_KEYS: [batch, n] float32 used for every call, or a callable (map_index, n, caller_lineno) -> keys [>= n]
# when a function shuffles for several purposes (ext_encode_rois: fg / bg down-sampling, up-sampling).
SHUFFLE_KEYS = None


def random_shuffle(value, seed=None, name=None):
    v = _t(value)
    if SHUFFLE_KEYS is None:
        keys = np.arange(len(v), dtype=np.float32)
    elif callable(SHUFFLE_KEYS):
        frames, f = [], sys._getframe(1)
        while f is not None and len(frames) < 8:
            frames.append((f.f_code.co_name, f.f_lineno, f.f_code.co_filename))
            f = f.f_back
        keys = np.asarray(SHUFFLE_KEYS(MAP_INDEX, len(v), frames), np.float32)[:len(v)]
    else:
        keys = np.asarray(SHUFFLE_KEYS[MAP_INDEX][:len(v)], np.float32)
    return _t(v[np.argsort(keys, kind='stable')])


class _Any(object):
    """Attribute sink for module-level references the golden scripts never execute (initialisers, tf.layers, ...)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    if name == "tuple":
        return _tf_tuple
    return _Any()


# ---- submodules ------------------------------------------------------------------------------------------------
nn = types.ModuleType("tensorflow.nn")
image = types.ModuleType("tensorflow.image")


def _top_k(x, k=1, sorted=True, name=None):  # noqa: A002
    x = _t(x)
    k = int(np.asarray(k))
    if x.ndim > 1:  # row-wise
        rows = [_top_k(r, k) for r in x.reshape(-1, x.shape[-1])]
        return (_t(np.stack([r[0] for r in rows]).reshape(x.shape[:-1] + (k,))),
                _t(np.stack([r[1] for r in rows]).reshape(x.shape[:-1] + (k,))))
    order = np.lexsort((np.arange(x.shape[0]), -x.astype(np.float64)))[:k]   # descending, ties -> lower index
    return _t(x[order]), _t(order.astype(np.int32))


def _softmax(x, axis=-1, name=None, dim=None):
    x = _t(x)
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return _t((e / np.sum(e, axis=axis, keepdims=True)).astype(x.dtype))


def _nms(boxes, scores, max_output_size, iou_threshold=0.5, name=None):
    from oracle import proposals as P  # the restated TF r1.6 NonMaxSuppressionV2
    idx = P.non_max_suppression_fast(np.asarray(boxes, np.float32), np.asarray(scores, np.float32), int(np.asarray(max_output_size)),
                                     float(iou_threshold))
    return _t(idx.astype(np.int32))


def _convert_image_dtype(x, dtype, saturate=False, name=None):
    """integer -> float: cast, then multiply by 1 / dtype.max (the documented rule)."""
    x = _t(x)
    if np.issubdtype(x.dtype, np.integer) and np.issubdtype(_np(dtype), np.floating):
        return _t((x.astype(np.float64) * (1.0 / np.iinfo(x.dtype).max)).astype(_np(dtype)))
    return _t(x.astype(_np(dtype)))


class _ResizeMethod(object):
    BILINEAR, NEAREST_NEIGHBOR, BICUBIC, AREA = 0, 1, 2, 3


def _lerp_matrix(n_out, n_in):
    """[n_out, n_in] float64 interpolation matrix of TF r1.6 ResizeBilinear, align_corners=False: output i samples
    the input at i * (n_in / n_out) (no half-pixel offset), between floor() and min(floor() + 1, n_in - 1)."""
    m = np.zeros((n_out, n_in))
    for i in builtins_range(n_out):
        pos = i * (np.float32(n_in) / np.float32(n_out))
        lo = int(np.floor(pos))
        hi = min(lo + 1, n_in - 1)
        m[i, lo] += 1.0 - (pos - lo)
        m[i, hi] += pos - lo
    return m


def _resize_images(images, size, method=_ResizeMethod.BILINEAR, align_corners=False):
    assert method == _ResizeMethod.BILINEAR and not align_corners
    x = np.asarray(_t(images), np.float64)  # [N,H,W,C]
    ho, wo = int(np.asarray(size[0])), int(np.asarray(size[1]))
    ry, rx = _lerp_matrix(ho, x.shape[1]), _lerp_matrix(wo, x.shape[2])
    return _t(np.einsum("ih,nhwc,jw->nijc", ry, x, rx).astype(np.float32))


image.convert_image_dtype, image.resize_images, image.ResizeMethod = _convert_image_dtype, _resize_images, _ResizeMethod
nn.top_k, nn.softmax = _top_k, _softmax
image.non_max_suppression = _nms
sys.modules.setdefault("tensorflow.nn", nn)
sys.modules.setdefault("tensorflow.image", image)


# ---- variable scopes + tf.layers (network builders; see _layers.py) ----------------------------------------------
import importlib  # noqa: E402

_layers = importlib.import_module(__name__ + "._layers")  # ('from . import' would hit the __getattr__ sink)

layers = types.ModuleType("tensorflow.layers")
for _n in ("conv2d", "separable_conv2d", "batch_normalization", "max_pooling2d", "dense"):
    setattr(layers, _n, getattr(_layers, _n))
sys.modules.setdefault("tensorflow.layers", layers)
variable_scope, AUTO_REUSE = _layers.variable_scope, _layers.AUTO_REUSE
add, subtract, multiply = _binary(np.add), _binary(np.subtract), _binary(np.multiply)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    return _t(np.mean(_t(x), axis=axis, keepdims=keepdims, dtype=np.float64).astype(np.float32))


def _relu(x, name=None):
    return _t(np.maximum(_t(x), 0))


nn.relu = _relu


# ---- what the reference's TRAINING model_fn needs beyond the graph builders (light_head_rfcn_train.py:277-451) -----
control_dependencies = name_scope


def abs(x, name=None):  # noqa: A001
    return _t(np.abs(_t(x)))


def add_n(values, name=None):
    total = _t(values[0])
    for v in values[1:]:
        total = _t(total + _t(v))
    return total


class _Variable(object):
    def __init__(self, name, value):
        self.name, self.value = name + ":0", value
        self.op = types.SimpleNamespace(name=name)


def trainable_variables():
    return [_Variable(k, v) for k, v in _layers.VARIABLES.items() if not k.rsplit("/", 1)[-1].startswith("moving_")]


class GraphKeys(object):
    TRAINABLE_VARIABLES, UPDATE_OPS, GLOBAL_VARIABLES = "trainable_variables", "update_ops", "variables"


def get_collection(key, scope=None):
    return trainable_variables() if key == GraphKeys.TRAINABLE_VARIABLES else []


def _log_softmax(logits):
    z = np.asarray(logits, np.float64)
    z = z - z.max(axis=-1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def _sparse_xent(labels=None, logits=None, name=None, _sentinel=None):
    lp = _log_softmax(logits)
    lab = np.asarray(labels).astype(np.int64)
    return _t((-np.take_along_axis(lp, lab[..., None], axis=-1)[..., 0]).astype(np.float32))


def _l2_loss(v, name=None):
    a = np.asarray(v.value if isinstance(v, _Variable) else v, np.float64)
    return _t(np.float32((a * a).sum() / 2))


nn.sparse_softmax_cross_entropy_with_logits, nn.l2_loss = _sparse_xent, _l2_loss

losses = types.ModuleType("tensorflow.losses")
# weights = 1, reduction SUM_BY_NONZERO_WEIGHTS: the mean over all elements
losses.sparse_softmax_cross_entropy = lambda labels, logits, **kw: _t(
    np.float32(np.asarray(_sparse_xent(labels, logits), np.float64).mean()))
losses.add_loss = lambda *a, **k: None

estimator = types.ModuleType("tensorflow.estimator")


class _ModeKeys(object):
    TRAIN, EVAL, PREDICT = "train", "eval", "infer"


class _EstimatorSpec(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


estimator.ModeKeys, estimator.EstimatorSpec = _ModeKeys, _EstimatorSpec


class _Flags(object):
    """tf.app.flags: DEFINE_* record their defaults; OVERRIDES plays the command line."""
    OVERRIDES = {}

    class _Values(object):
        pass

    FLAGS = _Values()

    @classmethod
    def _define(cls, name, default, doc=None, **kw):
        setattr(cls.FLAGS, name, cls.OVERRIDES.get(name, default))

    DEFINE_integer = DEFINE_float = DEFINE_string = DEFINE_boolean = DEFINE_bool = _define


app = types.ModuleType("tensorflow.app")
app.flags = _Flags
train = types.ModuleType("tensorflow.train")
train.get_or_create_global_step = lambda: _t(np.int64(0))
CHECKPOINT_TENSORS = None  # a set of names -> what tf.train.NewCheckpointReader(...).has_tensor answers from
SAVERS = []                # var_list of every tf.train.Saver constructed


class _CheckpointReader(object):
    def __init__(self, path):
        self.path = path

    def has_tensor(self, name):
        return CHECKPOINT_TENSORS is None or name in CHECKPOINT_TENSORS


class _Saver(object):
    def __init__(self, var_list=None, reshape=False, **kw):
        SAVERS.append(var_list)

    def build(self):
        return None


train.latest_checkpoint = lambda checkpoint_dir, latest_filename=None: None
train.NewCheckpointReader, train.Saver = _CheckpointReader, _Saver
gfile = types.ModuleType("tensorflow.gfile")
gfile.IsDirectory = lambda path: False
train.piecewise_constant = lambda x, boundaries, values, name=None: _t(
    np.float32(values[int(np.searchsorted(np.asarray(boundaries), np.asarray(x), side="left"))]))


def _train_sink(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Any()


for _m in (train, nn, image, layers, losses, estimator, app, gfile):  # PEP 562: names not defined above resolve to sinks
    _m.__getattr__ = _train_sink
OP_LIBRARIES = {}  # path suffix -> object standing for the loaded custom-op module (tf.load_op_library)


def load_op_library(path):
    for k, v in OP_LIBRARIES.items():
        if path.endswith(k):
            return v
    return _Any()


# ---- tensorflow.python.* internals that utility/metrics.py reaches for: local variables (keyed by scope/name so
#      that calling streaming_tp_fp_arrays once per batch accumulates, as running its update ops would), assigns,
#      cumsum / scan / reverse ------------------------------------------------------------------------------------
LOCAL_VARIABLES = {}


class Variable(object):
    def __new__(cls, initial_value=None, name=None, trainable=True, collections=None, validate_shape=True, **kw):
        key = _layers._full(name)
        if key not in LOCAL_VARIABLES:
            self = object.__new__(cls)
            self.name, self.value = key, _t(initial_value)
            LOCAL_VARIABLES[key] = self
        return LOCAL_VARIABLES[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value) if dtype is None else np.asarray(self.value).astype(dtype)


def _assign(ref, value, validate_shape=None, name=None):
    ref.value = _t(value).astype(ref.value.dtype)
    return ref.value


def _assign_add(ref, value, name=None):
    ref.value = _t(ref.value + _t(value).astype(ref.value.dtype))
    return ref.value


def cumsum(x, axis=0, name=None):
    return _t(np.cumsum(_t(x), axis=axis))


def reverse(x, axis, name=None):
    return _t(np.flip(_t(x), axis=tuple(axis)))


def scan(fn, elems, initializer=None, **kw):
    e = _t(elems)
    acc, outs = (e[0], [e[0]]) if initializer is None else (initializer, [])
    for v in e[(1 if initializer is None else 0):]:
        acc = fn(acc, v)
        outs.append(acc)
    return _t(np.stack([np.asarray(o) for o in outs]))


def _tf_tuple(tensors, name=None):  # tf.tuple: served by the module __getattr__ (keeps the builtin usable here)
    return list(tensors)


def _module(fullname, **attrs):
    m = types.ModuleType(fullname)
    m.__dict__.update(attrs)
    m.__getattr__ = _train_sink
    m.__path__ = []
    sys.modules[fullname] = m
    parent, _, leaf = fullname.rpartition(".")
    if parent != "tensorflow":
        setattr(sys.modules[parent], leaf, m)  # 'from parent import leaf' must not fall into the parent's sink
    return m


_this = sys.modules[__name__]
_module("tensorflow.python")
_module("tensorflow.python.framework")
_module("tensorflow.python.ops")
_module("tensorflow.python.framework.dtypes", float32=float32, float64=float64, int32=int32, int64=int64, bool=bool)
_module("tensorflow.python.framework.ops", name_scope=name_scope, convert_to_tensor=convert_to_tensor,
        control_dependencies=control_dependencies, GraphKeys=GraphKeys, add_to_collections=lambda *a, **k: None)
_module("tensorflow.python.ops.array_ops", zeros=zeros, shape=shape, unstack=unstack)
_module("tensorflow.python.ops.math_ops", greater=greater, divide=divide,
        to_int64=lambda x, name=None: cast(x, int64), to_float=lambda x, name=None: cast(x, float32))
_module("tensorflow.python.ops.state_ops", assign=_assign, assign_add=_assign_add)
_module("tensorflow.python.ops.variable_scope", variable_scope=variable_scope)
_module("tensorflow.python.ops.variables", Variable=Variable)
GraphKeys.LOCAL_VARIABLES = "local_variables"


# ---- any other tensorflow.* import (tensorflow.contrib.framework..., tensorflow.python.ops...) resolves to an empty
#      package of attribute sinks: modules that merely IMPORT such names at the top (net/depth_conv2d.py) load --------
import importlib.abc  # noqa: E402
import importlib.machinery  # noqa: E402


class _SinkFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith("tensorflow."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        def sink(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Any()
        module.__getattr__ = sink


sys.meta_path.append(_SinkFinder())
