"""Coordinate-type tags for coefficient callables (fealpy/decorator/coordinates.py:10-22)."""


def _tag(kind):
    def deco(func):
        func.coordtype = kind
        return func
    return deco


cartesian = _tag("cartesian")
barycentric = _tag("barycentric")
