from .box_partition import BoxSlab, box_edges_before, box_number_of_edges, box_number_of_nodes
from .dist_cg import dist_cg, halo_exchange, CudaCgOps
from .slab_problem import SlabProblem

__all__ = ["BoxSlab", "box_edges_before", "box_number_of_edges", "box_number_of_nodes", "dist_cg", "halo_exchange",
           "CudaCgOps", "SlabProblem"]
