from .simplex import SimplexMesh, TriangleMesh, TetrahedronMesh

__all__ = ["SimplexMesh", "TriangleMesh", "TetrahedronMesh"]
