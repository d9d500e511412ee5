"""Imports the unmodified reference package from ``baseline/_ref`` (installed by ``baseline/install_ref.py``; the directory
is git-ignored but travels to the GPU box).

The reference needs ``polars`` at import time (``torchctr/__init__.py:1`` -> ``transformer.py:7``), which is not in this
image; only ``FeatureTransformer`` uses it, and that class is outside the hot path, so a placeholder module
stands in for it.  Everything else -- ``torchctr.models.DNN``, ``torchctr.trainer.Trainer``, ``torchctr.dataset``,
``torchctr.nn``, ``torchctr.utils.hash_bucket`` -- is the reference's own code, unchanged.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "torchctr", "trainer.py"))


def load_reference():
    """-> the ``torchctr`` module of the reference.  Raises ImportError when ``baseline/_ref`` is missing."""
    if "torchctr" in sys.modules and getattr(sys.modules["torchctr"], "__file__", "").startswith(REF_DIR):
        return sys.modules["torchctr"]
    if not available():
        raise ImportError("baseline/_ref is missing: run `python baseline/install_ref.py` where /root/reference exists")
    try:
        import datasets  # noqa: F401  (must be imported before the polars placeholder exists: it probes for polars itself)
    except ImportError:
        pass
    if "polars" not in sys.modules:
        m = types.ModuleType("polars")
        m.__spec__ = importlib.machinery.ModuleSpec("polars", None)
        m.DataFrame = m.Expr = type("X", (), {})
        sys.modules["polars"] = m
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torchctr
    import torchctr.dataset    # noqa: F401  (the package's __init__ does not pull its sub-packages in)
    import torchctr.models     # noqa: F401
    import torchctr.nn         # noqa: F401
    import torchctr.trainer    # noqa: F401
    return torchctr
