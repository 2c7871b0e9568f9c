"""CPU: host-side logic that needs no GPU -- the fp32x3 operand split / weight packing (ops/conv.py), the
post-processing constants (light_head_rfcn_eval.det_min_size) and the oracle chain they are checked against."""
import numpy as np
import torch

import xdet_b200  # noqa: F401
from oracle import detections as od
from xdet_b200.ops import conv as conv_ops


def test_split3_values_is_exact_to_fp32():
    g = torch.Generator().manual_seed(0)
    w = torch.randn((64, 37), generator=g) * 5
    hi, mid, lo = conv_ops.split3_values(w)
    for part in (hi, mid, lo):  # every piece is representable in bf16
        assert torch.equal(part, part.to(torch.bfloat16).float())
    err = (hi.double() + mid.double() + lo.double() - w.double()).abs()
    assert bool((err <= w.double().abs() * 2.0 ** -23).all())


def test_fp32x3_packing_pairs_the_six_significant_products():
    """Emulate one output of the split-operand convolution on the CPU: activation blocks [mid|lo|hi|mid|hi|hi] against
    the packed weight blocks must equal x.w to fp32 level (dropped products are < 2^-24 of the result)."""
    g = torch.Generator().manual_seed(1)
    cin, cout = 40, 8
    x = torch.randn((cin,), generator=g)
    w = torch.randn((cout, cin, 1, 1), generator=g)
    with conv_ops.precision("fp32x3"):
        wp = conv_ops.pack_conv_weight(w)            # [cout, pad64(6*cin)] bf16
    assert wp.shape == (cout, 256) and wp.dtype == torch.bfloat16
    hi, mid, lo = conv_ops.split3_values(x)
    xs = torch.cat([mid, lo, hi, mid, hi, hi]).double()          # what xdet_split3_bf16 writes
    got = wp[:, :6 * cin].double() @ xs
    ref = w.reshape(cout, cin).double() @ x.double()
    assert float((got - ref).abs().max()) < 1e-6 * float(ref.abs().max() + 1)
    assert not wp[:, 6 * cin:].float().any()                     # zero padding up to the 64-channel chunk
    # default precision: plain bf16 pack, one block
    assert conv_ops.pack_conv_weight(w).shape == (cout, 64)


def test_weight_chunk_selects_the_same_channels_in_all_six_blocks():
    g = torch.Generator().manual_seed(2)
    cin, cout, taps = 24, 4, 3
    w = torch.randn((cout, cin, 1, taps), generator=g)
    with conv_ops.precision("fp32x3"):
        wp = conv_ops.pack_conv_weight(w)
        full = wp.view(cout, taps, -1)[:, :, :6 * cin].float().view(cout, taps, 6, cin)
        chunk = conv_ops._weight_chunk(wp, cout, taps, cin, 8, 16)
    cv = chunk.view(cout, taps, -1)
    assert cv.shape[2] == 64
    assert torch.equal(cv[:, :, :48].float().view(cout, taps, 6, 8), full[:, :, :, 8:16])


def test_det_min_size_follows_eval_helper():
    from xdet_b200 import light_head_rfcn_eval as lh
    shapes = [(375, 500), (480, 480), (1, 1), (1200, 1600)]
    ms = lh.det_min_size(shapes, 480, "cpu").numpy()
    for s, m in zip(shapes, ms):
        assert m == od.min_size_of(0.03, s, (480, 480))
    assert ms[2] == np.float32(0.0001)  # the floor of utility/eval_helper.py:295


def test_oracle_detection_chain_small_case():
    """Hand-checkable case: two overlapping boxes of one class (the weaker one is suppressed at IoU 0.3), one box of
    another class, one below the score threshold."""
    probs = np.zeros((4, 3), np.float32)
    probs[:, 0] = 1.0
    probs[0, 1], probs[1, 1], probs[2, 2], probs[3, 1] = 0.9, 0.8, 0.7, 0.005
    boxes = np.array([[0.1, 0.1, 0.5, 0.5], [0.12, 0.1, 0.5, 0.52], [0.6, 0.6, 0.9, 0.9], [0.2, 0.2, 0.4, 0.4]], np.float32)
    s, b = od.bboxes_eval_select(probs, boxes, np.array([0, 0, 1, 1], np.float32), (480, 480), 3)
    assert s[1][0] == np.float32(0.9) and not s[1][1:].any() and s[1].shape == (200,)
    assert np.array_equal(b[1][0], boxes[0]) and not b[1][1:].any()
    assert s[2][0] == np.float32(0.7) and np.array_equal(b[2][0], boxes[2])


def test_flag_names_and_defaults_match_reference_scripts():
    """tests/golden/flags_golden.json: what tf.app.flags.DEFINE_* record when the reference's train / eval scripts are
    imported unmodified under the numpy TensorFlow stand-in.  The drop-in contract: same flag names, same
    defaults.  Deviations, both deliberate: data_format (NHWC is the kernels' native layout; the NCHW contract is kept
    at the op boundary) and the extra flags backbone / precision / resnet_layers."""
    import json
    import os
    from xdet_b200 import light_head_rfcn_eval as pe
    from xdet_b200 import light_head_rfcn_train as pt
    with open(os.path.join(os.path.dirname(__file__), "golden", "flags_golden.json")) as f:
        gold = json.load(f)
    for which, mod, extras in (("train", pt, {"backbone", "resnet_layers", "precision"}), ("eval", pe, {"backbone", "precision"})):
        mine, ref = mod.make_params(), gold[which]
        assert set(mine) - set(ref) == extras, (which, sorted(set(mine) - set(ref)))
        assert set(ref) - set(mine) == set(), (which, sorted(set(ref) - set(mine)))
        diff = {k for k in ref if mine[k] != ref[k]}
        assert diff == {"data_format"}, (which, {k: (mine[k], ref[k]) for k in diff})
        assert ref["data_format"] == "channels_first" and mine["data_format"] == "channels_last"
        # ... and the launchers accept every one of them on the command line
        ap = mod.arg_parser()
        args = ap.parse_args(["--rpn_nms_thres", "0.6", "--run_on_cloud", "false", "--model_scope", "m"])
        assert args.rpn_nms_thres == 0.6 and args.run_on_cloud is False and args.model_scope == "m"
        assert all(hasattr(args, k) for k in ref)
    assert pt.arg_parser().parse_args(["--resnet_layers", "1,1,1,1", "--train_epochs", "3"]).resnet_layers == (1, 1, 1, 1)


def test_public_signatures_are_supersets_of_the_reference():
    """tests/golden/signatures_golden.json: inspect.signature of the reference's functions (imported unmodified under
    the numpy TensorFlow stand-in).  Every same-named product function must take the reference's parameters under the
    same names, required ones at the same positions, with the same defaults -- so that reference call sites keep
    working; the product may add trailing keyword parameters (store=, out=, state=, device=...).  Deliberate default
    deviations are listed here, nowhere else."""
    import importlib
    import inspect
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "signatures_golden.json")) as f:
        gold = json.load(f)
    allowed = {
        # NCHW is what the model_fn consumes and the kernel writes; 'NHWC' returns the permuted view
        ("preprocessing.common_preprocessing", "light_head_preprocess_for_test", "data_format"),
        ("preprocessing.common_preprocessing", "light_head_preprocess_for_eval", "data_format"),
        # numpy float64 in place of tf.float64
        ("utility.metrics", "precision_recall", "dtype"),
    }
    checked = 0
    for modname, sigs in gold.items():
        mod = importlib.import_module("xdet_b200." + modname)
        for qual, ref in sigs.items():
            obj = mod
            for part in qual.split("."):
                obj = getattr(obj, part, None)
                if obj is None:
                    break
            if obj is None or not callable(obj):
                continue                       # not part of the hot path's surface (other detectors' helpers, ...)
            params = list(inspect.signature(obj).parameters.items())
            names = [n for n, _ in params]
            for pos, (name, dflt) in enumerate(ref):
                assert name in names, (modname, qual, name)
                p = dict(params)[name]
                if dflt is None:               # required in the reference: same position here
                    assert names.index(name) == pos, (modname, qual, name)
                    continue
                assert p.default is not inspect.Parameter.empty, (modname, qual, name)
                if (modname, qual, name) in allowed:
                    continue
                if "value" in dflt:
                    mine = list(p.default) if isinstance(p.default, (tuple, list)) else p.default
                    assert mine == dflt["value"], (modname, qual, name, p.default, dflt)
            checked += 1
    assert checked >= 25, checked


def test_separation_of_product_oracle_and_reference():
    """Static guards on the three rules the parity claims rest on:
    (1) nothing under x-detector_b200/ or tools/ imports or executes oracle/ (test infrastructure only);
    (2) nothing that runs on the GPU box -- the product, bench.py, __graft_entry__, the -m gpu tests -- reads
        /root/reference (only the golden-minting scripts and the CPU-only 'reference mounted?' probes mention it);
    (3) the stand-in TensorFlow (oracle/tf_shim) is imported only by tests/golden/make_*.py: no test and no oracle
        module does ``import tensorflow``."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    product = glob.glob(os.path.join(root, "x-detector_b200", "**", "*.py"), recursive=True)
    tools = glob.glob(os.path.join(root, "tools", "*.py"))
    imp = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b|from\s+\.+\s*oracle\b)", re.M)
    for f in product + tools:
        src = open(f).read()
        assert not imp.search(src), f
        assert "tf_shim" not in src and "/root/reference" not in src, f
    gpu_side = [os.path.join(root, "bench.py")] + glob.glob(os.path.join(root, "tests", "*_gpu.py"))
    for f in gpu_side:
        src = open(f).read()
        assert "/root/reference" not in src and "tf_shim" not in src, f
    entry = open(os.path.join(root, "__graft_entry__.py")).read()
    assert entry.count("/root/reference") == 1 and 'isdir("/root/reference")' in entry   # build(): compile oracle/_ref if present
    for f in glob.glob(os.path.join(root, "tests", "test_*.py")) + glob.glob(os.path.join(root, "oracle", "*.py")):
        src = open(f).read()
        assert not re.search(r"^\s*import tensorflow|^\s*from tensorflow", src, re.M), f
