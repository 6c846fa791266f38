"""Import the reference's own functions from oracle/_ref (TEST INFRASTRUCTURE; see oracle/fetch_ref.py).

`load()` returns None when oracle/_ref is absent (the callers then fall back to the oracle port and say so).
The reference's `DistanseAwareCalibration.predict` calls `.cuda()` on its tensors
(trainers/calibration/distanse_aware_calibration.py:52-53); the CPU legs run it with `torch.Tensor.cuda`
temporarily replaced by the identity - `cpu_only()` is that context manager.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_cached = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "tools", "metrics.py"))


@contextlib.contextmanager
def cpu_only():
    import torch
    orig_cuda, orig_to = torch.Tensor.cuda, torch.Tensor.to
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else orig_to(self, *a, **k)
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.Tensor.to = orig_cuda, orig_to


def load():
    """-> namespace(metrics, DistanseAwareCalibration, proximity) of the unmodified reference modules, or None."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    for name in ("tools", "tools.metrics", "trainers", "trainers.calibration",
                 "trainers.calibration.distanse_aware_calibration", "trainers.calibration.proximity"):
        if name in sys.modules and REF_DIR not in (getattr(sys.modules[name], "__file__", None) or REF_DIR):
            raise RuntimeError(f"module {name} is already imported from elsewhere; cannot load oracle/_ref")
    sys.path.insert(0, REF_DIR)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            metrics = importlib.import_module("tools.metrics")
            dac = importlib.import_module("trainers.calibration.distanse_aware_calibration")
            prox = importlib.import_module("trainers.calibration.proximity")
    finally:
        sys.path.remove(REF_DIR)
    _cached = SimpleNamespace(metrics=metrics, DistanseAwareCalibration=dac.DistanseAwareCalibration, proximity=prox)
    return _cached
