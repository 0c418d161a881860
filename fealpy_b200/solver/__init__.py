from .cg import cg

__all__ = ["cg"]
