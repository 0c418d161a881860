"""Isotropic linear-elastic material (fealpy/material/elastic_material.py:184-319):
Lame parameters from (E, nu) or given directly, and the elastic matrix D in Voigt form."""
from __future__ import annotations

import torch


class LinearElasticMaterial:
    def __init__(self, name: str, elastic_modulus=None, poisson_ratio=None, lame_lambda=None, shear_modulus=None,
                 density=None, hypo: str = "3D", device=None):
        self.name = name
        self.hypo = hypo
        self.device = device
        if elastic_modulus is not None and poisson_ratio is not None and lame_lambda is None and shear_modulus is None:
            E, nu = elastic_modulus, poisson_ratio
            lam = nu * E / ((1 + nu) * (1 - 2 * nu))
            mu = E / (2 * (1 + nu))
        elif lame_lambda is not None and shear_modulus is not None and elastic_modulus is None and poisson_ratio is None:
            lam, mu = lame_lambda, shear_modulus
            E = mu * (3 * lam + 2 * mu) / (lam + mu)
            nu = lam / (2 * (lam + mu))
        elif None not in (elastic_modulus, poisson_ratio, lame_lambda, shear_modulus):
            E, nu, lam, mu = elastic_modulus, poisson_ratio, lame_lambda, shear_modulus
            cE = mu * (3 * lam + 2 * mu) / (lam + mu)
            cnu = lam / (2 * (lam + mu))
            if abs(cE - E) > 1e-5 or abs(cnu - nu) > 1e-5:
                raise ValueError("The input elastic modulus and Poisson's ratio are inconsistent with "
                                 "the values calculated from the provided Lame's lambda and shear modulus.")
        else:
            raise ValueError("You must provide either (elastic_modulus, poisson_ratio) or "
                             "(lame_lambda, shear_modulus), or all four.")
        self.E, self.nu, self.lam, self.mu, self.rho = E, nu, lam, mu, density
        if hypo == "3D":
            D = [[2 * mu + lam, lam, lam, 0, 0, 0], [lam, 2 * mu + lam, lam, 0, 0, 0], [lam, lam, 2 * mu + lam, 0, 0, 0],
                 [0, 0, 0, mu, 0, 0], [0, 0, 0, 0, mu, 0], [0, 0, 0, 0, 0, mu]]
            self._D = torch.tensor(D, dtype=torch.float64)
        elif hypo == "plane_stress":
            self._D = E / (1 - nu ** 2) * torch.tensor([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]], dtype=torch.float64)
        elif hypo == "plane_strain":
            self._D = torch.tensor([[2 * mu + lam, lam, 0], [lam, 2 * mu + lam, 0], [0, 0, mu]], dtype=torch.float64)
        else:
            raise NotImplementedError("Only 3D, plane_stress, and plane_strain are supported.")

    elastic_modulus = property(lambda s: s.E)
    poisson_ratio = property(lambda s: s.nu)
    lame_lambda = property(lambda s: s.lam)
    shear_modulus = property(lambda s: s.mu)
    density = property(lambda s: s.rho)
    hypothesis = property(lambda s: s.hypo)

    @property
    def D(self):
        return self._D if self.device is None else self._D.to(self.device)

    def elastic_matrix(self, bcs=None):
        """(1, 1, 3, 3) in 2-D, (1, 1, 6, 6) in 3-D"""
        return self.D[None, None, ...]
