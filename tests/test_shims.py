"""INTEGRATION.md section 1 applied for real: the re-export stubs are written over (a copy of) the reference's
module files and the CUDA path is reached through the REFERENCE import paths - `tools.metrics`,
`trainers.calibration.distanse_aware_calibration`, ... - in a fresh interpreter."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from clip_calibration_b200 import install_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("CCAL_REFERENCE", "/root/reference")


def _tree(tmp_path):
    """A copy of the reference's hot-path packages when the reference exists (build container), else an empty tree
    (GPU box): install() creates the stub files either way."""
    dst = tmp_path / "CLIP_Calibration"
    dst.mkdir()
    if os.path.isdir(REFERENCE):
        shutil.copytree(os.path.join(REFERENCE, "tools"), dst / "tools")
        os.makedirs(dst / "trainers")
        shutil.copytree(os.path.join(REFERENCE, "trainers", "calibration"), dst / "trainers" / "calibration")
    return str(dst)


def _run(tree, code):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([tree, ROOT]))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=tree)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


def test_install_and_revert_shims(tmp_path):
    tree = _tree(tmp_path)
    had_reference = os.path.exists(os.path.join(tree, "tools", "metrics.py"))
    files = install_shims.install(tree)
    assert "tools/metrics.py" in files and "trainers/calibration/distanse_aware_calibration.py" in files
    assert os.path.exists(os.path.join(tree, "tools", "metrics.py.orig")) == had_reference
    out = _run(tree, "import tools.metrics as m, trainers.calibration.distanse_aware_calibration as d, "
                     "trainers.calibration.proximity as p\n"
                     "import clip_calibration_b200.tools.metrics as mm\n"
                     "from clip_calibration_b200.trainers.calibration.distanse_aware_calibration import DistanseAwareCalibration as D\n"
                     "assert m.ECE is mm.ECE and m.MCE is mm.MCE and m.AdaptiveECE is mm.AdaptiveECE and m.PIECE is mm.PIECE\n"
                     "assert m.compute_acc_bin is mm.compute_acc_bin and d.DistanseAwareCalibration is D\n"
                     "assert callable(p.get_knn_dists) and callable(p.get_val_image_knn_dists)\n"
                     "print(m.__file__)")
    assert out.strip().startswith(tree)
    assert not os.path.exists(os.path.join(tree, "netcal"))          # the netcal stand-in is opt-in
    files = install_shims.install(tree, with_netcal=True)
    assert "netcal/binning.py" in files
    out = _run(tree, "from netcal.binning import HistogramBinning, IsotonicRegression\n"
                     "import clip_calibration_b200.trainers.calibration.netcal_binning as nb\n"
                     "assert HistogramBinning is nb.HistogramBinning and IsotonicRegression is nb.IsotonicRegression\n"
                     "print(HistogramBinning(bins=10).bins)")
    assert out.strip() == "10"
    install_shims.revert(tree)
    assert not os.path.exists(os.path.join(tree, "netcal"))
    if had_reference:
        with open(os.path.join(tree, "tools", "metrics.py")) as fh:
            assert "KBinsDiscretizer" in fh.read()           # the reference's own file is back


@pytest.mark.gpu
def test_reference_import_paths_reach_the_cuda_path(tmp_path, golden, cuda_lib):
    tree = _tree(tmp_path)
    install_shims.install(tree)
    gpath = os.path.join(ROOT, "tests", "golden", "eurosat.npz")
    code = f"""
import numpy as np
from tools.metrics import ECE, MCE, AdaptiveECE
from trainers.calibration.distanse_aware_calibration import DistanseAwareCalibration
from clip_calibration_b200 import synth, native
g = np.load({gpath!r})
case = synth.make_config("eurosat", seed=int(g["seed"]), n_override=int(g["N"]))
n0 = native.launch_count()
dac = DistanseAwareCalibration()
dac.fit(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, 5)
np.testing.assert_allclose(dac.class_confidence, g["cc_k5"], rtol=2e-6)
logits = (np.float32(100.0) * case.img[:2000]) @ case.txt_tuned.T
scaled = dac.predict(logits.astype(np.float64))
assert scaled.dtype == np.float32 and scaled.shape == logits.shape
conf, pred = g["dac_conf"], g["dac_pred"]
assert abs(ECE(conf, pred, case.labels, 10) - float(g["dac_ece10"])) < 1e-7
assert abs(MCE(conf, pred, case.labels, 10) - float(g["dac_mce10"])) < 1e-7
assert abs(AdaptiveECE(conf, pred, case.labels, 10) - float(g["dac_ace10"])) < 1e-7
assert native.launch_count() > n0, "no CUDA kernel was launched through the reference import paths"
print("ok")
"""
    assert _run(tree, code).strip().endswith("ok")
