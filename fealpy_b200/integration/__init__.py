from .fealpy_plugin import install, assemble_with_b200, cg_with_b200, adapt_space

__all__ = ["install", "assemble_with_b200", "cg_with_b200", "adapt_space"]
