"""Import the UNMODIFIED reference modules from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container: /root/reference does not exist on the GPU box, so
nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is used by
``tests/golden/make_golden.py`` (fixture generation) and by the optional live differential
tests, which skip themselves when the reference is absent.

``lib/losses.py:4-5`` imports ``pytorch_metric_learning`` (not installed, never used by the
code): three empty stub modules are registered so the import succeeds.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("WEALY_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "tensor_ops.py"))


def load():
    """-> (tensor_ops module, losses module) of the reference, imported unchanged."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in ("pytorch_metric_learning", "pytorch_metric_learning.losses",
                 "pytorch_metric_learning.miners"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    tops = importlib.import_module("lib.tensor_ops")
    losses = importlib.import_module("lib.losses")
    return tops, losses
