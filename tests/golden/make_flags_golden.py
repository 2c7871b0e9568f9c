#!/usr/bin/env python
"""Record the command-line flags (names and defaults) that the reference's light_head_rfcn_train.py (:38-168) and
light_head_rfcn_eval.py (:41-139) define, by importing both scripts unmodified under the numpy TensorFlow stand-in
(tf.app.flags.DEFINE_* record their defaults).  --run_on_cloud=False for the train script only (its default would
shell out to cmake at import); that flag is recorded with the reference's own default.
    python tests/golden/make_flags_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference", ROOT]

import tensorflow as tf  # noqa: E402  (the stand-in)


def main():
    import light_head_rfcn_eval as le
    flags = {"eval": dict(vars(le.FLAGS))}
    for k in list(vars(tf.app.flags.FLAGS)):
        delattr(tf.app.flags.FLAGS, k)
    tf.app.flags.OVERRIDES.update(run_on_cloud=False)
    import light_head_rfcn_train as lt
    flags["train"] = dict(vars(lt.FLAGS), run_on_cloud=True)
    path = os.path.join(HERE, "flags_golden.json")
    with open(path, "w") as f:
        json.dump(flags, f, indent=1, sort_keys=True)
    print("wrote", path, {k: len(v) for k, v in flags.items()})


if __name__ == "__main__":
    main()
