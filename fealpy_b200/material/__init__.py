from .elastic import LinearElasticMaterial

__all__ = ["LinearElasticMaterial"]
