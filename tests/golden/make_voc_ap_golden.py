#!/usr/bin/env python
"""Mint golden vectors for the VOC AP arithmetic from the REFERENCE'S OWN code: the pure-numpy ``voc_ap`` of
/root/reference/voc_eval.py:98-130 (the rest of that module imports TensorFlow through dataset_common, so only this
function is extracted -- by its source text, with ast -- and executed unmodified).  Run in the build container only
(the GPU box has no /root/reference); the .npz it writes is committed.
    python tests/golden/make_voc_ap_golden.py"""
import ast
import os

import numpy as np

REF = "/root/reference/voc_eval.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "voc_ap_golden.npz")


def reference_voc_ap():
    src = open(REF).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "voc_ap"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["voc_ap"]


def main():
    voc_ap = reference_voc_ap()
    rng = np.random.default_rng(2018)
    cases = {}
    for i in range(24):
        n = int(rng.integers(1, 400))
        tp = rng.random(n) < rng.uniform(0.05, 0.9)
        if i == 0:
            tp[:] = True
        if i == 1:
            tp[:] = False
        npos = int(tp.sum() + rng.integers(0, 30)) or 1
        ctp, cfp = np.cumsum(tp.astype(np.float64)), np.cumsum((~tp).astype(np.float64))
        rec = ctp / float(npos)
        prec = ctp / np.maximum(ctp + cfp, np.finfo(np.float64).eps)
        cases["tp_%d" % i] = tp
        cases["npos_%d" % i] = np.int64(npos)
        cases["ap07_%d" % i] = np.float64(voc_ap(rec, prec, True))
        cases["ap12_%d" % i] = np.float64(voc_ap(rec, prec, False))
    np.savez_compressed(OUT, n_cases=np.int64(24), **cases)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
