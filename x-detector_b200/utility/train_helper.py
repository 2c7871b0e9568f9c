"""Checkpoint restore helpers -- the API of the reference's ``utility/train_helper.py`` (:5-93) without TensorFlow:
which variables are restored under which checkpoint names (``--checkpoint_exclude_scopes``, ``--model_scope`` ->
``--checkpoint_model_scope`` renaming, ``--ignore_missing_vars``), reading a V2 checkpoint (``tensor_bundle.py``) and
producing the ``state_dict`` that ``LightHeadRFCN(state_dict=...)`` / ``LightHeadTrainer(state_dict=...)`` take (their
variables are keyed by the reference's TF names)."""
import os
import re

import numpy as np

from .tensor_bundle import TensorBundleReader


def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: the prefix named by the ``checkpoint`` state file of a directory, or None."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.isfile(state):
        return None
    with open(state) as f:
        m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', f.read(), re.M)
    if not m:
        return None
    path = m.group(1)
    return path if os.path.isabs(path) else os.path.join(checkpoint_dir, path)


def variables_to_restore(var_names, model_scope, checkpoint_model_scope=None, checkpoint_exclude_scopes=None):
    """utility/train_helper.py:13-30 -> {name in the checkpoint: name of the model variable}."""
    exclusions = [s.strip() for s in checkpoint_exclude_scopes.split(',')] if checkpoint_exclude_scopes else []
    kept = [v for v in var_names if not any(v.startswith(e) for e in exclusions)]
    if checkpoint_model_scope is None:
        return {v: v for v in kept}
    if checkpoint_model_scope.strip() == '':
        return {v.replace(model_scope + '/', checkpoint_model_scope): v for v in kept}
    return {v.replace(model_scope, checkpoint_model_scope): v for v in kept}


def resolve_checkpoint_path(checkpoint_path):
    """A directory means its latest checkpoint (:32-35)."""
    return latest_checkpoint(checkpoint_path) if os.path.isdir(checkpoint_path) else checkpoint_path


def load_state_dict(checkpoint_path, var_names, model_scope, checkpoint_model_scope=None, checkpoint_exclude_scopes=None,
                    ignore_missing_vars=False, shapes=None, verify_checksums=True):
    """The restore of ``get_init_fn_for_scaffold`` (:5-72) as data: {model variable name: float32 numpy array} read from
    the V2 checkpoint at ``checkpoint_path`` (file prefix, or a directory holding a ``checkpoint`` state file).
    Missing variables raise KeyError unless ``ignore_missing_vars``; ``shapes`` ({name: shape}) enables the
    ``reshape=False`` check of ``tf.train.Saver`` (:65)."""
    path = resolve_checkpoint_path(checkpoint_path)
    if path is None:
        raise FileNotFoundError("no checkpoint found under %s" % checkpoint_path)
    mapping = variables_to_restore(var_names, model_scope, checkpoint_model_scope, checkpoint_exclude_scopes)
    if not mapping:
        raise ValueError('variables_to_restore cannot be empty')
    reader = TensorBundleReader(path, verify_checksums=verify_checksums)
    out = {}
    for ckpt_name, var in mapping.items():
        if not reader.has_tensor(ckpt_name):
            if ignore_missing_vars:
                continue
            raise KeyError("Variable %s missing in checkpoint %s" % (ckpt_name, path))
        t = np.asarray(reader.get_tensor(ckpt_name), dtype=np.float32)
        if shapes is not None and var in shapes and tuple(shapes[var]) != tuple(t.shape):
            raise ValueError("shape mismatch for %s: checkpoint %s vs variable %s" % (var, t.shape, tuple(shapes[var])))
        out[var] = t
    return out


def checkpoint_to_state_dict(checkpoint_path, model_scope, checkpoint_model_scope=None, verify_checksums=True):
    """Everything a checkpoint holds for the model, keyed by MODEL variable names: the inverse of the renaming above,
    applied to every float tensor of the checkpoint except optimizer slots (``.../Momentum``) and counters.  What the
    Estimator does when it restores a model_dir checkpoint for evaluation (``get_latest_checkpoint_for_evaluate``,
    :74-93): variables absent from the checkpoint keep their initial values."""
    path = resolve_checkpoint_path(checkpoint_path)
    if path is None:
        raise FileNotFoundError("no checkpoint found under %s" % checkpoint_path)
    reader = TensorBundleReader(path, verify_checksums=verify_checksums)
    src = model_scope if checkpoint_model_scope is None else checkpoint_model_scope
    out = {}
    for name in reader.get_variable_to_shape_map():
        if name.endswith("/Momentum") or name == "global_step" or "ExponentialMovingAverage" in name:
            continue
        if checkpoint_model_scope is not None and checkpoint_model_scope.strip() == '':
            model_name = model_scope + '/' + name
        elif name.startswith(src):
            model_name = model_scope + name[len(src):]
        else:
            continue
        t = reader.get_tensor(name)
        if t.dtype.kind == 'f':
            out[model_name] = np.asarray(t, dtype=np.float32)
    return out


def get_init_fn_for_scaffold(flags, var_names, shapes=None):
    """utility/train_helper.py:5-72 under the reference's name: None when ``flags.model_dir`` already holds a
    checkpoint (the Estimator resumes from it and ``--checkpoint_path`` is ignored, :10-12); otherwise a callable that
    performs the fine-tuning restore and returns ``{model variable name: float32 array}`` -- pass it as
    ``VariableStore(state_dict=...)`` / ``LightHeadRFCN.from_checkpoint``.  ``var_names`` stands for TF's
    TRAINABLE_VARIABLES collection (:17): the model's trainable variable names (no moving statistics)."""
    checkpoint_path = flags.checkpoint_path
    if flags.run_on_cloud:
        checkpoint_path = os.path.join(flags.data_dir, flags.cloud_checkpoint_path)
    if latest_checkpoint(flags.model_dir):
        return None
    if not variables_to_restore(var_names, flags.model_scope, flags.checkpoint_model_scope,
                                flags.checkpoint_exclude_scopes):
        raise ValueError('variables_to_restore cannot be empty')

    def callback(scaffold=None, session=None):
        return load_state_dict(checkpoint_path, var_names, flags.model_scope, flags.checkpoint_model_scope,
                               flags.checkpoint_exclude_scopes, flags.ignore_missing_vars, shapes)
    return callback


def get_latest_checkpoint_for_evaluate(flags):
    """utility/train_helper.py:74-93: the checkpoint an evaluation should read -- None when ``flags.model_dir`` holds
    one (the Estimator picks that up by itself), else ``--checkpoint_path`` (``model_dir`` on the cloud) resolved to
    its latest checkpoint if it is a directory."""
    checkpoint_path = flags.checkpoint_path
    if flags.run_on_cloud:
        checkpoint_path = flags.model_dir
    if latest_checkpoint(flags.model_dir):
        return None
    return resolve_checkpoint_path(checkpoint_path)
