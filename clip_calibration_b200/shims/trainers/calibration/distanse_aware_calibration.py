# Re-export stub for <reference>/trainers/calibration/distanse_aware_calibration.py.
from clip_calibration_b200.trainers.calibration.distanse_aware_calibration import DistanseAwareCalibration  # noqa: F401
