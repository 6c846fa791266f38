import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 GPU (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The C-ABI library is a build artefact (git-ignored): make sure it exists and matches the sources
    before any test touches it (no-op when the stamp is current; nvcc cross-compiles without a GPU)."""
    from clip_calibration_b200 import build as _build
    _build.build()
    _build.build_c_demo()
    return _build.LIB


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


_case_cache = {}


@pytest.fixture(scope="session")
def synth_case(golden):
    """Regenerate a golden case's inputs from its seed and check they are the inputs the golden
    outputs were produced from (checksums stored next to the outputs)."""
    from clip_calibration_b200 import synth

    def make(name):
        if name not in _case_cache:
            g = golden(name)
            case = synth.make_config(name, seed=int(g["seed"]), n_override=int(g["N"]))
            assert abs(case.img.astype(np.float64).sum() - float(g["img_checksum"])) < 1e-6, "generator drifted"
            assert abs(case.txt_tuned.astype(np.float64).sum() - float(g["txt_checksum"])) < 1e-6
            _case_cache[name] = case
        return _case_cache[name]
    return make


@pytest.fixture(scope="session")
def cuda_lib():
    """Skip-free guard for -m gpu tests: the CUDA library must load and the device must be sm_100."""
    import torch
    from clip_calibration_b200 import _lib
    assert torch.cuda.is_available(), "gpu-marked test running without a GPU"
    lib = _lib.load()
    rc = lib.ccal_check_device()
    assert rc == 0, _lib.last_error()
    return lib
