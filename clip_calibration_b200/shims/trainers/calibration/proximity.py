# Re-export stub for <reference>/trainers/calibration/proximity.py.
from clip_calibration_b200.trainers.calibration.proximity import get_knn_dists, get_val_image_knn_dists  # noqa: F401
