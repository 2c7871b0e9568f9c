"""Variable store: the stand-in for TF's variable scopes on this path.

Variables are fp32 torch tensors keyed by the reference's TF names (``rpn_head/conv2d/kernel``,
``large_sep_feature/Branch_0/conv2d_1/bias``, ``final_head/subnet_fc/kernel``,
``batch_normalization_7/moving_variance`` ...), so that a converted TF checkpoint of the reference can be
dropped in as a state dict.  Kernels keep TF's layouts: conv [KH,KW,Cin,Cout], dense [in,out].
Missing variables are created with the reference's initialisers (seeded), since no weights exist offline:
``tf.glorot_normal_initializer`` (net/xception_body.py:25-26) and ``tf.variance_scaling_initializer``
(net/resnet_v2.py:89) -- both TRUNCATED normals in TF 1.6 (resampled beyond two standard deviations) --, zeros for
biases, and TensorFlow's batch-norm defaults (gamma = 1, beta = 0, moving_mean = 0, moving_variance = 1): what a
training or fine-tuning run of the reference starts from (its restore map, utility/train_helper.py, never restores
moving statistics).  ``randomize_bn=True`` (tests only; tests/conftest.py turns it on through ``RANDOMIZE_BN``) draws
the batch-norm variables NON-trivially instead (gamma~U(.5,1.5), beta,mean~N(0,.1), var~U(.5,1.5)) so that folding
errors would show in parity tests.  ``created`` lists the variables that did not come from ``state_dict``.

Derived tensors (packed bf16 GEMM weights, folded batch-norm scale/bias) are cached per name.
"""
import math

import torch


RANDOMIZE_BN = False  # default of VariableStore(randomize_bn=None); the test suite sets it


class VariableStore(object):
    def __init__(self, device="cuda", seed=0, state_dict=None, randomize_bn=None):
        self.device = torch.device(device)
        self.randomize_bn = RANDOMIZE_BN if randomize_bn is None else bool(randomize_bn)
        self.created = []
        # state_dict: {TF variable name: tensor or numpy array} (e.g. utility.train_helper.load_state_dict of a TF
        # checkpoint); values become fp32 tensors on the device
        self.vars = {} if state_dict is None else {
            k: torch.as_tensor(v, dtype=torch.float32).to(self.device).contiguous() for k, v in state_dict.items()}
        self.derived = {}
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self._scope = []
        self._counters = [{}]

    # ---- scopes and TF-style automatic layer names -------------------------------------------
    class _Scope(object):
        def __init__(self, store, name):
            self.store, self.name = store, name

        def __enter__(self):
            self.store._scope.append(self.name)
            self.store._counters.append({})
            return self

        def __exit__(self, *a):
            self.store._scope.pop()
            self.store._counters.pop()

    def scope(self, name):
        return VariableStore._Scope(self, name)

    def auto_name(self, base):
        """tf.layers naming: ``conv2d``, ``conv2d_1``, ... numbered per enclosing variable scope."""
        c = self._counters[-1]
        i = c.get(base, 0)
        c[base] = i + 1
        return base if i == 0 else "%s_%d" % (base, i)

    def full(self, name):
        return "/".join(self._scope + [name])

    # ---- creation -----------------------------------------------------------------------------
    def get(self, name, shape, init):
        key = self.full(name)
        if key not in self.vars:
            self.vars[key] = init(shape).to(self.device)
            self.created.append(key)
        v = self.vars[key]
        if tuple(v.shape) != tuple(shape):
            raise ValueError("variable %s has shape %s, expected %s" % (key, tuple(v.shape), tuple(shape)))
        return key, v

    def _normal(self, shape, std):
        return torch.randn(shape, generator=self._gen, dtype=torch.float32) * std

    def _truncated_normal(self, shape, std):
        """tf.truncated_normal: values beyond two standard deviations are redrawn."""
        t = torch.empty(shape, dtype=torch.float32)
        return torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0 * std, b=2.0 * std, generator=self._gen)

    def glorot_normal(self, shape):
        rf = 1
        for s in shape[:-2]:
            rf *= s
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        return self._truncated_normal(shape, math.sqrt(2.0 / (fan_in + fan_out)))

    def variance_scaling(self, shape):
        rf = 1
        for s in shape[:-2]:
            rf *= s
        return self._truncated_normal(shape, math.sqrt(1.0 / (shape[-2] * rf)))

    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.float32)

    def batch_norm(self, name, channels):
        """-> dict(gamma, beta, moving_mean, moving_variance) keys under ``name``."""
        with self.scope(name):
            if self.randomize_bn:
                g = self.get("gamma", (channels,), lambda s: torch.rand(s, generator=self._gen) + 0.5)
                b = self.get("beta", (channels,), lambda s: self._normal(s, 0.1))
                m = self.get("moving_mean", (channels,), lambda s: self._normal(s, 0.1))
                v = self.get("moving_variance", (channels,), lambda s: torch.rand(s, generator=self._gen) + 0.5)
            else:  # tf.layers.batch_normalization defaults
                g = self.get("gamma", (channels,), lambda s: torch.ones(s, dtype=torch.float32))
                b = self.get("beta", (channels,), self.zeros)
                m = self.get("moving_mean", (channels,), self.zeros)
                v = self.get("moving_variance", (channels,), lambda s: torch.ones(s, dtype=torch.float32))
        return {"gamma": g, "beta": b, "mean": m, "var": v}

    def folded_bn(self, bn, eps):
        """Inference batch-norm as y = x*scale + bias (fp32 CUDA tensors, cached)."""
        key = ("bn", bn["gamma"][0], eps)
        if key not in self.derived:
            g, b, m, v = (bn[k][1].double() for k in ("gamma", "beta", "mean", "var"))
            scale = g / torch.sqrt(v + eps)
            self.derived[key] = (scale.float().contiguous(), (b - m * scale).float().contiguous())
        return self.derived[key]

    def state_dict(self):
        return dict(self.vars)
