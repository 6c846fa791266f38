# Re-export stub for `netcal.binning` (see netcal/__init__.py next to this file).
from clip_calibration_b200.trainers.calibration.netcal_binning import HistogramBinning, IsotonicRegression  # noqa: F401
