"""Apply INTEGRATION.md section 1 to a checkout of the reference: overwrite the hot-path modules with the re-export
stubs under clip_calibration_b200/shims/ (same relative paths), keeping a `.orig` copy of every file replaced.

    python -m clip_calibration_b200.install_shims /path/to/CLIP_Calibration [--revert]

After this the reference's own call sites (`VLCalibration.build_dac_calibrator`, `VLClassification.evaluate`,
`VLBaseLearner.test`) reach the CUDA path through their usual imports - `tools.metrics`,
`trainers.calibration.distanse_aware_calibration`, ... - with this repository on PYTHONPATH.
"""
from __future__ import annotations

import os
import shutil
import sys

SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def shim_files():
    out = []
    for dirpath, _, files in os.walk(SHIMS):
        for f in sorted(files):
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(dirpath, f), SHIMS))
    return sorted(out)


def install(reference_dir: str) -> list:
    done = []
    for rel in shim_files():
        dst = os.path.join(reference_dir, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(dst) and not os.path.exists(dst + ".orig"):
            shutil.copyfile(dst, dst + ".orig")
        shutil.copyfile(os.path.join(SHIMS, rel), dst)
        done.append(rel)
    return done


def revert(reference_dir: str) -> list:
    done = []
    for rel in shim_files():
        dst = os.path.join(reference_dir, rel)
        if os.path.exists(dst + ".orig"):
            shutil.move(dst + ".orig", dst)
            done.append(rel)
    return done


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    files = revert(sys.argv[1]) if "--revert" in sys.argv[2:] else install(sys.argv[1])
    print("\n".join(files))
