"""CPU: TF V2 checkpoint import -- the tensor-bundle reader against bundles written by the independent
fixture writer (tests/bundle_writer.py), the restore-name rules of utility/train_helper.py:13-30, and a model built
from the imported state dict.  No TensorFlow offline: the format is restated, parity unpinned."""
import os

import numpy as np
import pytest
import torch

import xdet_b200  # noqa: F401
from tests import bundle_writer as bw
from xdet_b200.net.variables import VariableStore
from xdet_b200.utility import tensor_bundle as tb
from xdet_b200.utility import train_helper as th


def test_crc32c_known_answers():
    assert tb.crc32c(b"123456789") == 0xE3069283      # the CRC-32C check value
    assert tb.crc32c(b"") == 0
    assert tb.crc32c(bytes(32)) == 0x8A9136AA          # RFC 3720 B.4: 32 bytes of zeros


def _tensors(rng):
    t = {"xception_lighthead/conv2d/kernel": rng.standard_normal((7, 7, 3, 64)).astype(np.float32),
         "xception_lighthead/batch_normalization/gamma": rng.random(64).astype(np.float32),
         "xception_lighthead/batch_normalization/moving_mean": rng.standard_normal(64).astype(np.float32),
         "xception_lighthead/final_head/fc_cls/bias": rng.standard_normal(21).astype(np.float32),
         "global_step": np.array(12345, np.int64)}
    for i in range(1, 30):  # enough keys for several table blocks and shared key prefixes
        t["xception_lighthead/conv2d_%d/kernel" % i] = rng.standard_normal((1, 1, 8, 4 + i)).astype(np.float32)
    return t


def test_reader_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = _tensors(rng)
    prefix = str(tmp_path / "model.ckpt-7")
    bw.write_bundle(prefix, tensors)
    r = tb.TensorBundleReader(prefix)
    assert set(r.get_variable_to_shape_map()) == set(tensors)
    for k, v in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and list(got.shape) == list(v.shape) and np.array_equal(got, v), k
    assert r.has_tensor("global_step") and not r.has_tensor("nope")
    # corruption is detected
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(100)
        b = f.read(1)
        f.seek(100)
        f.write(bytes([b[0] ^ 0xFF]))
    bad = [k for k in tensors if _raises(lambda: tb.TensorBundleReader(prefix).get_tensor(k))]
    assert len(bad) == 1
    with open(prefix + ".index", "r+b") as f:
        f.seek(-3, os.SEEK_END)
        f.write(b"\x00")
    with pytest.raises(ValueError):
        tb.TensorBundleReader(prefix)


def _raises(fn):
    try:
        fn()
        return False
    except ValueError:
        return True


def test_restore_name_rules():
    names = ["xception_lighthead/conv2d/kernel", "xception_lighthead/rpn_head/conv2d/kernel",
             "xception_lighthead/final_head/fc_cls/bias"]
    assert th.variables_to_restore(names, "xception_lighthead") == {n: n for n in names}
    m = th.variables_to_restore(names, "xception_lighthead", "xception", "xception_lighthead/final_head, other")
    assert m == {"xception/conv2d/kernel": names[0], "xception/rpn_head/conv2d/kernel": names[1]}
    m = th.variables_to_restore(names[:1], "xception_lighthead", " ")   # blank: strip the scope (train_helper.py:27-28)
    assert m == {" conv2d/kernel": names[0]}


def test_load_state_dict_into_store(tmp_path):
    rng = np.random.default_rng(1)
    tensors = _tensors(rng)
    ckpt_dir = tmp_path / "logs"
    ckpt_dir.mkdir()
    # the checkpoint uses another scope name, as the published Xception backbone does (--checkpoint_model_scope)
    renamed = {k.replace("xception_lighthead", "xception"): v for k, v in tensors.items()}
    bw.write_bundle(str(ckpt_dir / "model.ckpt-9"), renamed)
    (ckpt_dir / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-9"\nall_model_checkpoint_paths: "model.ckpt-9"\n')
    assert th.latest_checkpoint(str(ckpt_dir)) == str(ckpt_dir / "model.ckpt-9")
    want = [k for k in tensors if k != "global_step"] + ["xception_lighthead/not_in_checkpoint/kernel"]
    with pytest.raises(KeyError):
        th.load_state_dict(str(ckpt_dir), want, "xception_lighthead", "xception")
    sd = th.load_state_dict(str(ckpt_dir), want, "xception_lighthead", "xception", ignore_missing_vars=True,
                            shapes={"xception_lighthead/conv2d/kernel": (7, 7, 3, 64)})
    assert set(sd) == set(want[:-1])
    with pytest.raises(ValueError):
        th.load_state_dict(str(ckpt_dir), want, "xception_lighthead", "xception", ignore_missing_vars=True,
                           shapes={"xception_lighthead/conv2d/kernel": (3, 3, 3, 64)})
    store = VariableStore(device="cpu", seed=0, state_dict=sd)
    with store.scope("xception_lighthead"):
        with store.scope("conv2d"):
            key, k = store.get("kernel", (7, 7, 3, 64), store.glorot_normal)   # found, not re-initialised
    assert key == "xception_lighthead/conv2d/kernel" and torch.equal(k, torch.from_numpy(tensors[key]))
    assert k.dtype == torch.float32


def test_checkpoint_to_state_dict(tmp_path):
    rng = np.random.default_rng(2)
    tensors = _tensors(rng)
    tensors["xception_lighthead/conv2d/kernel/Momentum"] = np.zeros((7, 7, 3, 64), np.float32)
    renamed = {k.replace("xception_lighthead", "xception"): v for k, v in tensors.items()}
    renamed["unrelated/weights"] = np.ones(3, np.float32)
    prefix = str(tmp_path / "m.ckpt")
    bw.write_bundle(prefix, renamed)
    sd = th.checkpoint_to_state_dict(prefix, "xception_lighthead", "xception")
    assert set(sd) == {k for k in tensors if not k.endswith("Momentum") and k != "global_step"}
    assert np.array_equal(sd["xception_lighthead/final_head/fc_cls/bias"], tensors["xception_lighthead/final_head/fc_cls/bias"])


def test_restore_map_matches_reference_init_fn(tmp_path):
    """The {checkpoint name: model variable} map that the reference's own get_init_fn_for_scaffold
    (utility/train_helper.py:5-72) hands to tf.train.Saver for the real Xception graph under the train script's
    default flags -- captured by running the reference unmodified under the numpy TensorFlow stand-in
    (tests/golden/make_trainstep_golden.py) -- against the product's rules, including ignore_missing_vars."""
    import json
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "trainstep_golden.npz"))
    r = json.loads(str(G["restore"]))
    meta = json.loads(str(G["meta"]))
    # the reference restores TRAINABLE variables only (tf.GraphKeys.TRAINABLE_VARIABLES): no moving statistics
    trainable = [n for n, _ in meta["variables"] if not n.rsplit("/", 1)[-1].startswith("moving_")]
    m = th.variables_to_restore(trainable, r["model_scope"], r["checkpoint_model_scope"], r["checkpoint_exclude_scopes"])
    assert m == r["all"] and len(m) == 154
    assert not any("/rpn_head/" in v or "/large_sep_feature/" in v or "/final_head/" in v for v in m.values())
    # a checkpoint holding every other tensor: the missing ones are skipped, the rest restored under model names
    shapes = {n: tuple(s) for n, s in meta["variables"]}
    rng = np.random.default_rng(4)
    tensors = {k: rng.standard_normal(shapes[r["all"][k]]).astype(np.float32) for k in r["checkpoint_tensors"]}
    prefix = str(tmp_path / "xception_model.ckpt")
    bw.write_bundle(prefix, tensors)
    sd = th.load_state_dict(prefix, trainable, r["model_scope"], r["checkpoint_model_scope"],
                            r["checkpoint_exclude_scopes"], ignore_missing_vars=r["ignore_missing_vars"], shapes=shapes)
    assert set(sd) == set(r["half"].values()) and len(sd) == 77
    for ck, name in r["half"].items():
        assert np.array_equal(np.asarray(sd[name]), tensors[ck])


def test_reference_named_entry_points(tmp_path):
    """get_init_fn_for_scaffold / get_latest_checkpoint_for_evaluate under the reference's names and flag object."""
    import types
    rng = np.random.default_rng(2)
    names = ["xception_lighthead/block1_conv1/kernel", "xception_lighthead/rpn_head/conv2d/kernel"]
    tensors = {"block1_conv1/kernel": rng.standard_normal((3, 3, 3, 32)).astype(np.float32)}
    pre = tmp_path / "pretrained"
    pre.mkdir()
    bw.write_bundle(str(pre / "xception_model.ckpt"), tensors)
    logs = tmp_path / "logs"
    logs.mkdir()
    flags = types.SimpleNamespace(checkpoint_path=str(pre / "xception_model.ckpt"), run_on_cloud=False, data_dir="",
                                  cloud_checkpoint_path="", model_dir=str(logs), model_scope="xception_lighthead",
                                  checkpoint_model_scope="", checkpoint_exclude_scopes="xception_lighthead/rpn_head",
                                  ignore_missing_vars=True)
    init_fn = th.get_init_fn_for_scaffold(flags, names, shapes={names[0]: (3, 3, 3, 32)})
    sd = init_fn(None, None)
    assert list(sd) == names[:1] and np.array_equal(sd[names[0]], tensors["block1_conv1/kernel"])
    assert th.get_latest_checkpoint_for_evaluate(flags) == flags.checkpoint_path
    # the cloud layout: data_dir/cloud_checkpoint_path
    cloud = types.SimpleNamespace(**dict(vars(flags), run_on_cloud=True, data_dir=str(pre),
                                         cloud_checkpoint_path="xception_model.ckpt", checkpoint_path="unused"))
    assert list(th.get_init_fn_for_scaffold(cloud, names)(None, None)) == names[:1]
    # a checkpoint in model_dir wins: no init_fn, and evaluation leaves the choice to the model_dir
    bw.write_bundle(str(logs / "model.ckpt-5"), tensors)
    (logs / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-5"\n')
    assert th.get_init_fn_for_scaffold(flags, names) is None and th.get_latest_checkpoint_for_evaluate(flags) is None
    with pytest.raises(ValueError):
        flags.model_dir = str(tmp_path / "empty")
        flags.checkpoint_exclude_scopes = "xception_lighthead"
        th.get_init_fn_for_scaffold(flags, names)


# ---- second opinions that do not come from this repository -------------------------------------------------------
def test_crc32c_matches_tensorboards_implementation():
    """CRC-32C and its TensorFlow mask against the implementation Google ships in TensorBoard's TensorFlow stub (the
    only TensorFlow-authored code installed here), for the byte-loop and the vectorised large-input paths."""
    tbs = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    from xdet_b200.utility import tensor_bundle as tb
    rng = np.random.default_rng(0)
    for n in (0, 1, 9, 4095, 4096, 70000, 16 * 4096 + 5, 3 * 1024 * 1024 + 77):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tb.crc32c(d) == (tbs.crc32c(d) & 0xFFFFFFFF), n
        assert tb.masked_crc32c(d) == (tbs.masked_crc32c(d) & 0xFFFFFFFF), n


def test_entry_protos_parse_with_tensorboards_tensorflow_protos(tmp_path):
    """The dtype enum and the TensorShapeProto bytes inside every BundleEntryProto written here, parsed by the compiled
    TensorFlow protobufs TensorBoard ships (types.proto, tensor_shape.proto)."""
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    from xdet_b200.utility import tensor_bundle as tb
    w = tb.TensorBundleWriter(str(tmp_path / "m.ckpt-1"))
    tensors = {"a/kernel": np.zeros((3, 3, 8, 16), np.float32), "global_step": np.array(7, np.int64),
               "b/counts": np.arange(5, dtype=np.int32), "c/d": np.ones((2, 0, 4), np.float64)}
    for k, v in tensors.items():
        w.add(k, v)
    w.finish()
    want = {np.dtype(np.float32): types_pb2.DT_FLOAT, np.dtype(np.int64): types_pb2.DT_INT64,
            np.dtype(np.int32): types_pb2.DT_INT32, np.dtype(np.float64): types_pb2.DT_DOUBLE}
    r = tb.TensorBundleReader(str(tmp_path / "m.ckpt-1"))
    with open(str(tmp_path / "m.ckpt-1.index"), "rb") as f:
        raw_index = f.read()
    for name, arr in tensors.items():
        e = r.entries[name]
        assert e.dtype == want[arr.dtype] and list(e.shape) == list(arr.shape)
        assert np.array_equal(r.get_tensor(name), arr)
        # the shape submessage, byte for byte, is what TensorFlow's own message class serialises
        proto = shape_pb2.TensorShapeProto()
        for s in arr.shape:
            proto.dim.add().size = int(s)
        assert proto.SerializeToString() in raw_index


def test_trainer_checkpoint_roundtrip_through_the_product_writer(tmp_path):
    """TensorBundleWriter -> TensorBundleReader, the ``checkpoint`` state file and its keep-N pruning."""
    from xdet_b200.utility import tensor_bundle as tb
    from xdet_b200.utility import train_helper as th
    rng = np.random.default_rng(1)
    for step in (1, 2, 3):
        w = tb.TensorBundleWriter(str(tmp_path / ("model.ckpt-%d" % step)))
        ts = {"s/v%03d/kernel" % i: rng.standard_normal((3, 3, i + 1, 2)).astype(np.float32) for i in range(200)}
        ts["global_step"] = np.array(step, np.int64)
        for k, v in ts.items():
            w.add(k, v)
        w.finish()
        tb.update_checkpoint_state(str(tmp_path), str(tmp_path / ("model.ckpt-%d" % step)), keep=2)
    assert th.latest_checkpoint(str(tmp_path)).endswith("model.ckpt-3")
    assert not (tmp_path / "model.ckpt-1.index").exists() and (tmp_path / "model.ckpt-2.index").exists()
    r = tb.TensorBundleReader(th.latest_checkpoint(str(tmp_path)))
    assert set(r.entries) == set(ts) and int(r.get_tensor("global_step")) == 3
    for k, v in ts.items():
        assert np.array_equal(r.get_tensor(k), v), k
