# Re-export stub for <reference>/tools/plot.py.
from clip_calibration_b200.tools.plot import plot_reliability_diagram, reliability_diagram_data  # noqa: F401
