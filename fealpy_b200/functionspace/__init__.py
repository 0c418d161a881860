from .lagrange import LagrangeFESpace, TensorFunctionSpace

__all__ = ["LagrangeFESpace", "TensorFunctionSpace"]
