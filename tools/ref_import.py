"""Import helper for the *reference* FEALPy tree in the build container (golden-vector generators in tools/).

Delegates to baseline/ref_loader.py (stub finder for the absent GUI packages); the default root here is the
read-only source tree /root/reference, which does not exist on the GPU box -- tests and bench.py use the
installed copy under baseline/_ref through baseline.ref_loader directly.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

REF_ROOT = os.environ.get("FEALPY_REFERENCE", "/root/reference")


def install():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    return ref_loader.install(REF_ROOT)
