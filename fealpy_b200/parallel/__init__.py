from .box_partition import BoxSlab, box_edges_before, box_number_of_edges, box_number_of_nodes
from .dist_cg import dist_cg, halo_exchange, CudaCgOps, DistCG, make_dist_solver
from .slab_problem import SlabProblem
from .verify import verify_slab, verify_partition
from .mesh_partition import MeshPartition, PartitionedProblem, PackedExchange, morton_codes

__all__ = ["BoxSlab", "box_edges_before", "box_number_of_edges", "box_number_of_nodes", "dist_cg", "halo_exchange",
           "CudaCgOps", "DistCG", "make_dist_solver", "SlabProblem", "verify_slab", "verify_partition", "MeshPartition", "PartitionedProblem", "PackedExchange", "morton_codes"]
