from .tensors import COOTensor, CSRTensor, coo_matrix, csr_matrix

__all__ = ["COOTensor", "CSRTensor", "coo_matrix", "csr_matrix"]
