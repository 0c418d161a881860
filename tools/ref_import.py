"""Import helper for the *reference* FEALPy tree (only usable in the build container).

The reference lives read-only under /root/reference and imports GUI packages
(matplotlib, vtk, ...) that are absent here.  This installs a meta-path finder that
fabricates empty stand-ins for them, so that `fealpy.mesh`, `fealpy.fem`, ... import.
Used only by the golden-vector generators in tools/; never by the product, the
tests, smoke() or bench.py (the reference does not exist on the GPU box).
"""
import os
import sys
import types
import importlib.abc
import importlib.machinery

REF_ROOT = os.environ.get("FEALPY_REFERENCE", "/root/reference")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        if n.startswith('__'):
            raise AttributeError(n)
        return _Dummy()


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []

        def _ga(n):
            if n.startswith('__'):
                raise AttributeError(n)
            return type(n, (_Dummy,), {})
        m.__getattr__ = _ga
        return m

    def exec_module(self, m):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    ROOTS = ('matplotlib', 'mpl_toolkits', 'vtk', 'gmsh', 'pyevtk')

    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, _Loader(), is_package=True)
        return None


def install():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
