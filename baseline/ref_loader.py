"""Loader for the UNMODIFIED reference (FEALPy 3.4.0) used as the baseline arm.

The reference is installed once, offline, into the git-ignored `baseline/_ref/`
(`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`,
see `__graft_entry__.build()`); that directory is NOT gpurun-ignored, so it travels to the GPU box.
`fealpy.mesh` imports GUI packages that this image does not have (matplotlib, vtk, gmsh, pyevtk):
a meta-path finder fabricates empty stand-ins so that the numerical path imports and runs unchanged.

Test / benchmark infrastructure only: nothing under `fealpy_b200/` imports this module.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []

        def _ga(n):
            if n.startswith("__"):          # keep inspect / torch introspection working
                raise AttributeError(n)
            return type(n, (_Dummy,), {})
        m.__getattr__ = _ga
        return m

    def exec_module(self, m):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    ROOTS = ("matplotlib", "mpl_toolkits", "vtk", "gmsh", "pyevtk")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, _Loader(), is_package=True)
        return None


def reference_root(root=None):
    """directory that holds the `fealpy` package: explicit argument, $FEALPY_REFERENCE, else baseline/_ref"""
    root = root or os.environ.get("FEALPY_REFERENCE") or REF_DIR
    return root if os.path.isdir(os.path.join(root, "fealpy")) else None


def available(root=None):
    return reference_root(root) is not None


def install(root=None):
    """make `import fealpy` resolve to the reference; returns the root used"""
    r = reference_root(root)
    if r is None:
        raise RuntimeError("reference not installed: baseline/_ref/fealpy is missing (run __graft_entry__.build() in the "
                           "build container, where /root/reference exists)")
    sys.dont_write_bytecode = True
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())
    if r not in sys.path:
        sys.path.insert(0, r)
    return r


def pip_install(src="/root/reference", dest=REF_DIR):
    """the one offline install: wheel of the pure-Python reference into baseline/_ref (no dependencies resolved:
    numpy/scipy/sympy/torch come from the image).  /root/reference is read-only, so the build runs on a /tmp copy."""
    import shutil
    import subprocess
    import tempfile
    if os.path.isdir(os.path.join(dest, "fealpy")):
        return dest
    if not os.path.isdir(src):
        raise RuntimeError(f"{src} not found")
    tmp = tempfile.mkdtemp(prefix="fealpy_ref_")
    try:
        cp = os.path.join(tmp, "src")
        shutil.copytree(src, cp, symlinks=True, ignore=shutil.ignore_patterns(".git", "docs", "notebook", "example", "kb"))
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
                        "--find-links", "/opt/wheelhouse", "--target", dest, cp], check=True,
                       env=dict(os.environ, HOME=tmp))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return dest
