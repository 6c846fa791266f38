# Re-export stub for <reference>/tools/metrics.py: same import path, CUDA implementation.
from clip_calibration_b200.tools.metrics import *  # noqa: F401,F403
from clip_calibration_b200.tools.metrics import ECE, MCE, AdaptiveECE, PIECE, compute_acc_bin  # noqa: F401
