# Stand-in for the `netcal` package inside a reference checkout (installed only by
# `python -m clip_calibration_b200.install_shims <tree> --with-netcal`).  It provides the two
# names the reference imports - `from netcal.binning import HistogramBinning, IsotonicRegression`
# (trainers/calibration/vl_calibrator.py:20-21) - as the GPU restatement of netcal's one-vs-all scheme.  A real netcal
# installation is shadowed while the checkout's directory comes first on sys.path; leave this shim out to keep it.
