"""TEST INFRASTRUCTURE ONLY -- makes the *unmodified* reference importable in this container.

Puts ``oracle/refstubs`` (stand-ins for the two absent third-party wheels, see
``oracle/refstubs/torch_scatter/__init__.py``) and ``/root/reference`` on ``sys.path``.
Used by ``tests/golden/generate_golden.py`` and by the CPU tests that pin the oracle against the
live reference.  ``/root/reference`` does not exist on the GPU box: callers must check
``reference_available()`` and skip.  Nothing under ``ptgnn_b200/`` imports this module.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("PTGNN_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstubs")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ptgnn", "neuralmodels", "gnn"))


def import_reference():
    """Returns the reference's ``ptgnn`` package (raises if /root/reference is absent)."""
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ptgnn  # noqa: F401
    import ptgnn.neuralmodels.gnn  # noqa: F401
    import ptgnn.neuralmodels.gnn.messagepassing  # noqa: F401

    return ptgnn
