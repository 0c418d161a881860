"""Simplex quadrature rules (host tables).

Mirrors `mesh.quadrature_formula(q).get_quadrature_points_and_weights()` of the reference
(fealpy/quadrature/triangle.py:16-329, tetrahedron.py:7-243, stroud_quadrature.py:5-40).
The truncated-digit tables are data dumped from the reference by tools/gen_tables.py into
data/quadrature.npz, so the element matrices see the same bits.
"""
import os

import numpy as np

_TABLES = None


def _tables():
    global _TABLES
    if _TABLES is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "quadrature.npz")
        _TABLES = dict(np.load(path))
    return _TABLES


class Quadrature:
    def __init__(self, bcs, ws):
        self.bcs, self.ws = bcs, ws

    def get_quadrature_points_and_weights(self):
        return self.bcs, self.ws

    def number_of_quadrature_points(self):
        return self.ws.shape[0]


def simplex_quadrature(TD: int, q: int) -> Quadrature:
    name = {2: "tri", 3: "tet"}.get(TD)
    if name is None:
        raise ValueError(f"unsupported simplex dimension {TD}")
    t = _tables()
    key = f"{name}_q{q}_bcs"
    if key not in t:
        raise NotImplementedError(f"no {name} quadrature table for q={q}")
    return Quadrature(t[key], t[f"{name}_q{q}_ws"])
