"""Apply INTEGRATION.md section 1 to a checkout of the reference: overwrite the hot-path modules with the re-export
stubs under clip_calibration_b200/shims/ (same relative paths), keeping a `.orig` copy of every file replaced.

    python -m clip_calibration_b200.install_shims /path/to/CLIP_Calibration [--with-netcal] [--revert]

After this the reference's own call sites (`VLCalibration.build_dac_calibrator`, `VLClassification.evaluate`,
`VLBaseLearner.test`) reach the CUDA path through their usual imports - `tools.metrics`,
`trainers.calibration.distanse_aware_calibration`, ... - with this repository on PYTHONPATH.
`--with-netcal` additionally drops a `netcal/` stand-in package into the checkout, so that the reference's own
`vl_calibrator.py` (`from netcal.binning import HistogramBinning, IsotonicRegression`, :20-21) gets the GPU calibrators
of trainers/calibration/netcal_binning.py without netcal being installed (it shadows a real netcal - hence opt-in).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
SHIMS_NETCAL = os.path.join(HERE, "shims_netcal")


def shim_files(root: str = SHIMS):
    out = []
    for dirpath, _, files in os.walk(root):
        for f in sorted(files):
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(dirpath, f), root))
    return sorted(out)


def _roots(with_netcal: bool):
    return [SHIMS, SHIMS_NETCAL] if with_netcal else [SHIMS]


def install(reference_dir: str, with_netcal: bool = False) -> list:
    done = []
    for root in _roots(with_netcal):
        for rel in shim_files(root):
            dst = os.path.join(reference_dir, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if os.path.exists(dst) and not os.path.exists(dst + ".orig"):
                shutil.copyfile(dst, dst + ".orig")
            shutil.copyfile(os.path.join(root, rel), dst)
            done.append(rel)
    return done


def revert(reference_dir: str) -> list:
    """Put every `.orig` back; files the netcal stand-in created (no `.orig`: the reference has no such files) are removed."""
    done = []
    for rel in shim_files(SHIMS):
        dst = os.path.join(reference_dir, rel)
        if os.path.exists(dst + ".orig"):
            shutil.move(dst + ".orig", dst)
            done.append(rel)
    for rel in shim_files(SHIMS_NETCAL):
        dst = os.path.join(reference_dir, rel)
        if os.path.exists(dst + ".orig"):
            shutil.move(dst + ".orig", dst)
            done.append(rel)
        elif os.path.exists(dst):
            with open(dst) as fh:
                ours = "clip_calibration_b200" in fh.read()
            if ours:
                os.remove(dst)
                done.append(rel)
    pkg = os.path.join(reference_dir, "netcal")
    if os.path.isdir(pkg) and not [f for f in os.listdir(pkg) if f != "__pycache__"]:
        shutil.rmtree(pkg)
    return done


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    files = revert(sys.argv[1]) if "--revert" in sys.argv[2:] else install(sys.argv[1], "--with-netcal" in sys.argv[2:])
    print("\n".join(files))
