from .fealpy_plugin import (install, assemble_with_b200, cg_with_b200, adapt_space, adapt_integrator, on_cuda,
                            B200CGSolver, B200JacobiPreconditioner)

__all__ = ["install", "assemble_with_b200", "cg_with_b200", "adapt_space", "adapt_integrator", "on_cuda",
           "B200CGSolver", "B200JacobiPreconditioner"]
