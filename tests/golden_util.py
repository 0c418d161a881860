"""helpers shared by the oracle tests and the GPU parity tests (numpy only)"""
import json
import os

import numpy as np

import cases as C

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    d = dict(np.load(os.path.join(GOLD, name + ".npz")))
    d["info"] = json.loads(str(d["info"]))
    return d


def ref_testdata():
    return dict(np.load(os.path.join(GOLD, "ref_testdata.npz")))


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def case_coef(case, spec, k, gold, NC, NQ, GD):
    """coefficient for integrator #k of the case: None | float | ndarray | callable (cartesian)"""
    coef = spec.get("coef")
    if coef is None or isinstance(coef, (int, float)):
        return coef
    if coef in C.COEF_FUNCS:
        return C.COEF_FUNCS[coef]
    return gold[f"coef_{k}"]
