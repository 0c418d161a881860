from .integrators import Integrator, ScalarDiffusionIntegrator, ScalarMassIntegrator, LinearElasticityIntegrator
from .bilinear_form import BilinearForm, GroupIntegrator

__all__ = ["Integrator", "ScalarDiffusionIntegrator", "ScalarMassIntegrator", "LinearElasticityIntegrator",
           "BilinearForm", "GroupIntegrator"]
