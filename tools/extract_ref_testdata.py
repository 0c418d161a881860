"""Extract the reference's OWN golden vectors for the hot path into one .npz
(tests/golden/ref_testdata.npz).  Build-container only.

Sources (python-literal data modules of the reference test-suite):
  test/fem/scalar_diffusion_integrator_data.py, scalar_mass_integrator_data.py,
  test/fem/bilinear_form_data.py, test/mesh/tetrahedron_mesh_data.py,
  test/mesh/triangle_mesh_data.py (only the part before its syntax error at line 2793),
  test/functionspace/lagrange_fe_space_data.py, tensor_space_data.py,
  test/backend/backend_data.py, and the inline vectors of
  test/sparse/test_coo_tensor.py:29-48,377-392.
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402
ref_import.install()
T = os.path.join(ref_import.REF_ROOT, "test")


def load(rel, upto_line=None):
    path = os.path.join(T, rel)
    if upto_line is None:
        spec = importlib.util.spec_from_file_location("m_" + os.path.basename(rel)[:-3], path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return vars(m)
    src = "".join(open(path).readlines()[:upto_line])
    ns = {}
    exec(compile(src, path, "exec"), ns)
    return ns


out = {}
d = load("fem/scalar_diffusion_integrator_data.py")["triangle_mesh_one_box"][0]
out["diffusion_p2_q3_box1_Ke"] = d["assembly_cell_matrix"]
d = load("fem/scalar_mass_integrator_data.py")["triangle_mesh_one_box"][0]
out["mass_p2_q3_box1_Ke"] = d["assembly_cell_matrix"]
for k, md in enumerate(load("fem/bilinear_form_data.py")["mesh_data"]):
    out[f"bform_mesh{k}_node"] = md["node"]
    out[f"bform_mesh{k}_cell"] = md["cell"]

tet = load("mesh/tetrahedron_mesh_data.py")
fb = tet["from_box"][0]                       # from_box(3,2,1) with threshold x,y,z<0.5 removed
for k in ("node", "edge", "face", "cell", "face2cell"):
    out[f"tet_from_box_321_thr_{k}"] = fb[k]
out["tet_321_cm"] = tet["entity_measure"][0]["cm"]
out["tet_321_glambda"] = tet["grad_lambda"][0]["glambda"]
out["tet_111_gsf_x_p2_q3"] = tet["grad_shape_function"][0]["grad_shape_function"]
out["tet_111_gsf_u_p2_q3"] = tet["grad_shape_function"][1]["grad_shape_function"]
out["tet_321_cell2ipoint_p4"] = tet["cell_to_ipoint"][0]["cell2ipoint"]
out["tet_321_face2ipoint_p4"] = tet["face_to_ipoint"][0]["f2p"]
out["tet_321_ipoints_p4"] = tet["interpolation_points"][0]["ipoint"]

tri = load("mesh/triangle_mesh_data.py", upto_line=2792)
out["tri_22_cell"] = tri["from_box_data"][0]["cell"] if "from_box_data" in tri else np.zeros(0)
out["tri_glambda_m11_22"] = tri["grad_lambda_data"][0]["val"]
out["tri_22_gphi_p2_q3"] = tri["grad_shape_function_data"][0]["gphi"]
ipd = tri["interpolation_point_data"][0]
out["tri_22_ips_p4"] = ipd["ips"]
out["tri_22_cip_p4"] = ipd["cip"]

ls = load("functionspace/lagrange_fe_space_data.py")["triangle_mesh_one_box"][0]
out["lfs_box1_p2_cell_to_dof"] = ls["cell_to_dof"]
out["lfs_box1_p2_is_boundary_dof"] = ls["is_boundary_dof"]
out["lfs_box1_p2_bcs"] = ls["bcs"]
out["lfs_box1_p2_basis"] = ls["basis"]
out["lfs_box1_p2_grad_basis"] = ls["grad_basis"]
out["lfs_box1_p2_ipoints"] = ls["interpolation points"]

ts = load("functionspace/tensor_space_data.py")["triangle_mesh"][0]
out["tensor_tri_dims"] = ts["triangle_mesh"]
out["tensor_tcell2dof"] = ts["tcell2dof"]

bd = load("backend/backend_data.py")
for k, mdat in enumerate(bd["multi_index_data"]):
    out[f"multi_index_p{mdat['p']}_d{mdat['dim']}"] = mdat["result"]
t2 = bd["triangle_mesh2d_data"][0]
for k in ("node", "cell", "bcs", "simple_measure", "simple_shape_function", "simple_grad_shape_function",
          "triangle_grad_lambda_2d", "bc_to_points"):
    out["bk_tri2d_" + k] = np.asarray(t2[k])
out["bk_tri2d_p"] = np.array(t2["p"])
t3 = bd["tetrahedron_mesh_data"][0]
for k in ("node", "cell", "tetrahedron_grad_lambda_3d"):
    out["bk_tet_" + k] = np.asarray(t3[k])

# test/sparse/test_coo_tensor.py:29-48
out["coo_indices"] = np.array([[0, 0, 1, 2, 0, 1], [1, 2, 0, 0, 2, 0]])
out["coo_values"] = np.array([1, 2, 3, 4, 5, 6], dtype=np.float64)
out["coo_expected_indices"] = np.array([[0, 0, 1, 2], [1, 2, 0, 0]])
out["coo_expected_values"] = np.array([1, 7, 9, 4], dtype=np.float64)
# test/sparse/test_coo_tensor.py:377-392
out["tocsr_dense"] = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 0], [0, 3, 4]], dtype=np.float64)
out["tocsr_crow"] = np.array([0, 0, 1, 2, 2, 4])
out["tocsr_col"] = np.array([0, 1, 1, 2])
out["tocsr_values"] = np.array([1, 2, 3, 4], dtype=np.float64)

path = os.path.join(ROOT, "tests", "golden", "ref_testdata.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")
for k, v in out.items():
    print(f"  {k:36s} {np.asarray(v).shape} {np.asarray(v).dtype}")
