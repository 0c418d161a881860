"""The parity case ladder shared by tools/gen_golden.py (reference side, build container)
and the tests (oracle / CUDA side).  Pure data + numpy; no reference, oracle or product imports."""
import numpy as np


def kappa_cart(p):
    """smooth positive variable coefficient, evaluated at physical points (..., GD)"""
    x, y = p[..., 0], p[..., 1]
    return 1.0 + 0.5 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)


kappa_cart.coordtype = "cartesian"


def reaction_cart(p):
    return 2.0 + p[..., 0] * p[..., -1]


reaction_cart.coordtype = "cartesian"


def source_cart(p):
    return np.sin(np.pi * p[..., 0]) * np.cos(np.pi * p[..., 1]) + 1.0


source_cart.coordtype = "cartesian"

COEF_FUNCS = {"kappa_cart": kappa_cart, "reaction_cart": reaction_cart, "source_cart": source_cart}


def perturb(node, dims, seed, amp=0.15):
    """deterministic node jitter (keeps from_box topology, breaks the uniform geometry)"""
    rng = np.random.default_rng(seed)
    h = 1.0 / max(dims)
    return node + amp * h * rng.uniform(-1.0, 1.0, size=node.shape)


def coef_array(kind, NC, NQ, GD, seed):
    rng = np.random.default_rng(seed)
    if kind == "cell":
        return rng.uniform(0.5, 1.5, size=(NC,))
    if kind == "cellquad":
        return rng.uniform(0.5, 1.5, size=(NC, NQ))
    if kind == "matrix":
        B = rng.uniform(-0.3, 0.3, size=(NC, NQ, GD, GD))
        return np.eye(GD) + 0.5 * (B + np.swapaxes(B, -1, -2))
    raise ValueError(kind)


# integrator spec: (kind, dict(q=?, coef=None|float|'cell'|'cellquad'|'matrix'|<func name>, method=None|'fast'))
# groups: list of lists; every inner list is ONE add_integrator(...) call
CASES = [
    dict(name="tri_p1_n8_diff_q3", mesh="tri", dims=(8, 8), p=1,
         groups=[[("diffusion", dict(q=3))]], cg=False, elem=True),
    dict(name="tri_p1_n8_diffmass_q3", mesh="tri", dims=(8, 8), p=1,
         groups=[[("diffusion", dict(q=3))], [("mass", dict(q=3))]], cg=True, elem=True),
    dict(name="tri_p2_5x4_diffmass", mesh="tri", dims=(5, 4), p=2,
         groups=[[("diffusion", dict())], [("mass", dict())]], cg=True, elem=True),
    dict(name="tri_p3_4x3_diffmass_jit", mesh="tri", dims=(4, 3), p=3, jitter=7,
         groups=[[("diffusion", dict())], [("mass", dict())]], cg=True, elem=True),
    dict(name="tri_p3_6x6_varcoef_q6", mesh="tri", dims=(6, 6), p=3,
         groups=[[("diffusion", dict(q=6, coef="kappa_cart"))], [("mass", dict(q=6, coef="reaction_cart"))]],
         cg=True, elem=True),
    dict(name="tri_p2_4x4_coefs_jit", mesh="tri", dims=(4, 4), p=2, jitter=3,
         groups=[[("diffusion", dict(coef="cellquad"))], [("mass", dict(coef="cell"))],
                 [("diffusion", dict(coef=2.5))]], cg=False, elem=True),
    dict(name="tri_p2_4x4_matrixcoef", mesh="tri", dims=(4, 4), p=2,
         groups=[[("diffusion", dict(coef="matrix"))]], cg=False, elem=True),
    dict(name="tri_p2_4x4_fast_grouped", mesh="tri", dims=(4, 4), p=2, jitter=11,
         groups=[[("diffusion", dict(method="fast")), ("mass", dict())]], cg=True, elem=True),
    dict(name="tet_p1_3x2x2_diffmass", mesh="tet", dims=(3, 2, 2), p=1,
         groups=[[("diffusion", dict())], [("mass", dict())]], cg=True, elem=True),
    dict(name="tet_p1_4_diff", mesh="tet", dims=(4, 4, 4), p=1,
         groups=[[("diffusion", dict())]], cg=False, elem=False),
    dict(name="tet_p2_3x2x1_diffmass", mesh="tet", dims=(3, 2, 1), p=2,
         groups=[[("diffusion", dict())], [("mass", dict())]], cg=True, elem=True),
    dict(name="tet_p2_4_diffmass_jit", mesh="tet", dims=(4, 4, 4), p=2, jitter=5,
         groups=[[("diffusion", dict())], [("mass", dict())]], cg=True, elem=False),
    dict(name="tet_p2_8_diffmass", mesh="tet", dims=(8, 8, 8), p=2,
         groups=[[("diffusion", dict())], [("mass", dict())]], cg=True, elem=False, values_only_checksum=True),
    dict(name="tet_p2_3_varcoef", mesh="tet", dims=(3, 3, 3), p=2, jitter=9,
         groups=[[("diffusion", dict(coef="cellquad", q=4))], [("mass", dict(coef="reaction_cart", q=4))]],
         cg=True, elem=True),
    dict(name="tet_p3_2_diffmass", mesh="tet", dims=(2, 2, 2), p=3,
         groups=[[("diffusion", dict())], [("mass", dict())]], cg=True, elem=True),
    dict(name="tet_p1_3_elasticity", mesh="tet", dims=(3, 3, 3), p=1, jitter=2,
         groups=[[("elasticity", dict(q=4, E=1.0, nu=0.3, hypo="3D"))]], tensor=dict(dof_priority=False),
         cg=False, elem=True),
    dict(name="tet_p1_3_elasticity_prio", mesh="tet", dims=(3, 2, 2), p=1,
         groups=[[("elasticity", dict(q=4, E=2.0, nu=0.25, hypo="3D"))]], tensor=dict(dof_priority=True),
         cg=False, elem=True),
    dict(name="tet_p2_2_elasticity", mesh="tet", dims=(2, 2, 2), p=2, jitter=4,
         groups=[[("elasticity", dict(E=1.0, nu=0.3, hypo="3D"))]], tensor=dict(dof_priority=False),
         cg=False, elem=True),
    dict(name="tri_p2_4x3_elasticity", mesh="tri", dims=(4, 3), p=2, jitter=6,
         groups=[[("elasticity", dict(E=1.0, nu=0.3, hypo="plane_strain"))]], tensor=dict(dof_priority=False),
         cg=False, elem=True),
    dict(name="tri_p1_5x5_elasticity_prio", mesh="tri", dims=(5, 5), p=1,
         groups=[[("elasticity", dict(E=1.0, nu=0.3, hypo="plane_stress"))]], tensor=dict(dof_priority=True),
         cg=False, elem=True),
]

def thr_partial(p):
    """partial-boundary predicate (x < 0.38 or y > 0.62): evaluated on face barycentres (method None / 'centroid') and on
    interpolation points (method 'interp') it selects different dof sets along the border of the region"""
    return (p[..., 0] < 0.38) | (p[..., 1] > 0.62)


def gd_vector(p):
    """vector-valued Dirichlet data (npoints, GD) for the elasticity case"""
    return np.stack([0.1 * p[..., 1], -0.2 * p[..., 0] * p[..., -1], 0.05 + 0.0 * p[..., 0]], axis=-1)[..., :p.shape[-1]]


gd_vector.coordtype = "cartesian"


THRESHOLDS = {"thr_partial": thr_partial}

# Poisson problems with Dirichlet BC + source (rows f1/f2): -div(grad u) = f, u = g on the boundary
# (threshold / method: the Dirichlet part of the boundary, functionspace/dofs.py:23-55; the rest is natural)
BC_CASES = [
    dict(name="bc_tri_p1_8", mesh="tri", dims=(8, 8), p=1),
    dict(name="bc_tri_p2_6_jit", mesh="tri", dims=(6, 6), p=2, jitter=13),
    dict(name="bc_tet_p2_3", mesh="tet", dims=(3, 3, 3), p=2),
    dict(name="bc_tri_p2_7_thr_centroid", mesh="tri", dims=(7, 7), p=2, threshold="thr_partial", method=None, reaction=True),
    dict(name="bc_tet_p2_4_thr_interp", mesh="tet", dims=(4, 4, 4), p=2, threshold="thr_partial", method="interp", reaction=True),
    dict(name="bc_tri_p3_5_thr_centroid", mesh="tri", dims=(5, 5), p=3, jitter=17, threshold="thr_partial", method="centroid",
         reaction=True),
]

# linear elasticity with a clamped / displaced part of the boundary (TensorFunctionSpace boundary API) + Jacobi CG
TENSOR_BC_CASES = [
    dict(name="bc_tet_p1_4_elasticity", mesh="tet", dims=(4, 3, 3), p=1, dof_priority=False, threshold="thr_partial",
         E=1.0, nu=0.3, hypo="3D", q=4),
    dict(name="bc_tri_p2_5_elasticity_prio", mesh="tri", dims=(5, 5), p=2, dof_priority=True, threshold="thr_partial",
         E=2.0, nu=0.25, hypo="plane_strain", q=None, jitter=21),
]


def box_of(case):
    return [0, 1, 0, 1] if case["mesh"] == "tri" else [0, 1, 0, 1, 0, 1]


def by_name(name):
    for c in CASES + BC_CASES + TENSOR_BC_CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)
