# Re-export stub for <reference>/trainers/calibration/multi_proximity_isotonic.py.
from clip_calibration_b200.trainers.calibration.multi_proximity_isotonic import BinMeanShift  # noqa: F401
