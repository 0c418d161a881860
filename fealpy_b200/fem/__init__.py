from .integrators import Integrator, ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator
from .bilinear_form import BilinearForm, GroupIntegrator
from .linear_form import LinearForm, ScalarSourceIntegrator, VectorSourceIntegrator, DirichletBC, DirichletBCOperator

__all__ = ["Integrator", "ScalarDiffusionIntegrator", "ScalarMassIntegrator", "LinearElasticityIntegrator",
           "BilinearForm", "GroupIntegrator", "LinearForm", "ScalarSourceIntegrator", "VectorSourceIntegrator", "DirichletBC", "DirichletBCOperator"]
