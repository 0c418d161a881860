"""Dump the reference's simplex quadrature tables to .npz (run in the build container only).

The reference hard-codes truncated-digit tables (fealpy/quadrature/triangle.py:16-329,
fealpy/quadrature/tetrahedron.py:7-243, stroud_quadrature.py:5-40 beyond the tables).
Parity to the last bit needs the same digits, so they are obtained by *calling* the
reference and storing the arrays as data:
    fealpy_b200/data/quadrature.npz   (product copy)
    oracle/quadrature.npz             (oracle copy; the product may not import oracle/)
Keys: tri_q{q}_bcs, tri_q{q}_ws, tet_q{q}_bcs, tet_q{q}_ws.
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import ref_import
ref_import.install()

from fealpy.quadrature import TriangleQuadrature, TetrahedronQuadrature  # noqa: E402
from fealpy.mesh import TriangleMesh, TetrahedronMesh  # noqa: E402

out = {}
tm = TriangleMesh.from_box([0, 1, 0, 1], 1, 1)
for q in range(1, 13):
    bcs, ws = tm.quadrature_formula(q).get_quadrature_points_and_weights()
    out[f"tri_q{q}_bcs"] = np.ascontiguousarray(bcs, dtype=np.float64)
    out[f"tri_q{q}_ws"] = np.ascontiguousarray(ws, dtype=np.float64)
tt = TetrahedronMesh.from_box([0, 1, 0, 1, 0, 1], 1, 1, 1)
for q in range(1, 10):
    bcs, ws = tt.quadrature_formula(q).get_quadrature_points_and_weights()
    out[f"tet_q{q}_bcs"] = np.ascontiguousarray(bcs, dtype=np.float64)
    out[f"tet_q{q}_ws"] = np.ascontiguousarray(ws, dtype=np.float64)

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rel in ("fealpy_b200/data/quadrature.npz", "oracle/quadrature.npz"):
    path = os.path.join(root, rel)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
