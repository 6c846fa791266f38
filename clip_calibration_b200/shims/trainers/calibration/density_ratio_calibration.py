# Re-export stub for <reference>/trainers/calibration/density_ratio_calibration.py (no statsmodels needed).
from clip_calibration_b200.trainers.calibration.density_ratio_calibration import DensityRatioCalibration  # noqa: F401
