"""Bring the reference's own CPU implementation of the hot path next to the oracle (TEST INFRASTRUCTURE).

    python oracle/fetch_ref.py [--reference /root/reference]

The reference (ml-stat-Sustech/CLIP_Calibration) is a Python repo; the three modules below import with
numpy / scipy / scikit-learn / pandas / torch alone.  This script places unmodified
copies under oracle/_ref/ (git-ignored, NOT gpurun-ignored: like a built .so it travels to the GPU box, where
/root/reference does not exist), so that `bench.py --impl reference` and the `cpu_baseline` leg time the real
reference functions and tests/test_oracle_ref.py can check the oracle port against them.  Nothing in the product
path (clip_calibration_b200/) may import oracle/_ref; only tests/, bench.py's reference legs and
__graft_entry__.smoke() do.  Reference sources are never committed to this repository.

  tools/metrics.py                                        ECE / MCE / AdaptiveECE / PIECE          (:90-236)
  trainers/calibration/distanse_aware_calibration.py      DistanseAwareCalibration.fit / .predict  (:13-58)
  trainers/calibration/proximity.py                       get_knn_dists / get_val_image_knn_dists  (:19-70)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["tools/metrics.py", "trainers/calibration/distanse_aware_calibration.py", "trainers/calibration/proximity.py"]


def fetch(reference: str = "/root/reference", quiet: bool = False) -> bool:
    """Copy the files; returns False (and leaves oracle/_ref untouched) when the reference tree is absent."""
    if not all(os.path.exists(os.path.join(reference, f)) for f in FILES):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": reference, "sha256": manifest}, fh, indent=1)
    if not quiet:
        print(f"oracle/_ref: {len(FILES)} reference modules copied from {reference}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("CCAL_REFERENCE", "/root/reference"))
    args = ap.parse_args()
    sys.exit(0 if fetch(args.reference) else 1)
