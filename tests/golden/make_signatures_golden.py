#!/usr/bin/env python
"""Record the call signatures (parameter names, order, defaults) of the reference's public functions on the
Light-Head R-CNN path, read with inspect from the reference modules imported unmodified under the numpy TensorFlow
stand-in.  tests/test_host_logic.py holds the product's same-named functions to them (same names,
argument order and defaults).
    python tests/golden/make_signatures_golden.py"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference", ROOT]

import tensorflow as tf  # noqa: E402,F401  (the stand-in)

MODULES = ["net.xception_body", "net.resnet_v2", "preprocessing.anchor_manipulator",
           "preprocessing.common_preprocessing", "utility.eval_helper", "utility.metrics", "utility.train_helper"]


def default(p):
    if p.default is inspect.Parameter.empty:
        return None
    d = p.default
    if isinstance(d, (bool, int, float, str, type(None))):
        return {"value": d}
    if isinstance(d, (tuple, list)):
        return {"value": list(d)}
    import enum
    return {"repr": d.name if isinstance(d, enum.Enum) else type(d).__name__}   # enums by member name, else type name


def signature(fn):
    return [[n, default(p)] for n, p in inspect.signature(fn).parameters.items()]


def main():
    import importlib
    out = {}
    for name in MODULES:
        mod = importlib.import_module(name)
        sigs = {}
        for k, obj in vars(mod).items():
            if getattr(obj, "__module__", None) != mod.__name__:
                continue
            if inspect.isfunction(obj):
                sigs[k] = signature(obj)
            elif inspect.isclass(obj):
                for mk, m in vars(obj).items():
                    if inspect.isfunction(m):
                        sigs["%s.%s" % (k, mk)] = signature(m)
        out[name] = sigs
    path = os.path.join(HERE, "signatures_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
