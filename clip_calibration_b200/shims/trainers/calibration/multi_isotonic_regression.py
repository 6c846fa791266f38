# Re-export stub for <reference>/trainers/calibration/multi_isotonic_regression.py.
from clip_calibration_b200.trainers.calibration.multi_isotonic_regression import MultiIsotonicRegression  # noqa: F401
