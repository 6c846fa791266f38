"""Drop-in for the DAC + softmax branch of the reference's trainers/calibration/vl_calibrator.py
(`VLCalibration.__init__ / fit / predict / build_dac_calibrator`, reference :30-109, :155-180).

In scope: DAC (dac_flag) followed by softmax; the `scaling_based` base calibrator with `procal_flag` =
DensityRatioCalibration (reference :116-119, :95-96; CUDA KDE, density_ratio_calibration.py); the `bin_based`
base calibrators 'multi_isotonic_regression', 'histogram_binning' and 'isotonic_regression', plain or
proximity-binned through BinMeanShift (reference :121-148, :97-102; GPU isotonic fit, multi_isotonic_regression.py /
multi_proximity_isotonic.py; the latter two are netcal's one-vs-all calibrators restated in netcal_binning.py -
netcal is an unpinned pip dependency that is not installed, so parity with netcal itself is unpinned).

`predict(logits, proximity)` keeps the reference contract and returns probabilities [N, C].
`predict_confidence` / `predict_from_features` are the additive routes that return only
(pred, conf) - all the evaluator reads (evaluators/vl_evaluator.py:68, :83).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import native
from .density_ratio_calibration import DensityRatioCalibration
from .distanse_aware_calibration import DistanseAwareCalibration
from .multi_isotonic_regression import MultiIsotonicRegression
from .multi_proximity_isotonic import BinMeanShift
from .netcal_binning import HistogramBinning, IsotonicRegression

# name -> (class, constructor kwargs): reference :125-148
_BIN_CALIBRATORS = {
    "histogram_binning": (HistogramBinning, {"bins": 10}),
    "isotonic_regression": (IsotonicRegression, {}),
    "multi_isotonic_regression": (MultiIsotonicRegression, {}),
}


class VLCalibration():

    def __init__(self, cfg, base_calibration_mode=None, base_bin_calibrator_name=None, dac_flag=False,
                 procal_flag=False, val_dict=None, text_feature_dict=None):
        if base_calibration_mode not in (None, "scaling_based", "bin_based"):
            raise ValueError(f"unknown base_calibration_mode {base_calibration_mode!r}")
        self.cfg = cfg
        self.base_calibration_mode = base_calibration_mode
        self.base_bin_calibrator_name = base_bin_calibrator_name
        self.dac_flag = dac_flag
        self.procal_flag = procal_flag
        self.text_feature_dict = text_feature_dict
        self.k_dac = cfg.CALIBRATION.DAC.K if cfg is not None else 5
        self.val_dict = val_dict
        self.val_image_proximity = None          # only the proximity-informed branches need it
        if val_dict is not None and "val_image_knn_dists" in val_dict:
            # distances to proximity (reference :68)
            self.val_image_proximity = np.exp(-np.mean(val_dict["val_image_knn_dists"], axis=-1))
        self.dac_calibrator = None
        self.base_calibrator = None

    def fit(self):
        self.dac_calibrator = None
        self.base_calibrator = None
        if self.dac_flag:
            self.dac_calibrator = self.build_dac_calibrator(self.text_feature_dict, self.k_dac)
        if self.base_calibration_mode is not None:
            self.base_calibrator = self.build_base_calibrator(self.base_bin_calibrator_name, self.val_image_proximity)

    def build_base_calibrator(self, base_bin_calibrator_name, val_image_proximity):
        """reference :112-150, `scaling_based` branch: the density-ratio calibrator is fitted on the validation
        (calibration) set - softmax of the un-scaled validation logits (:59-60), their argmax, labels, proximity."""
        if self.base_calibration_mode == "bin_based":
            if base_bin_calibrator_name not in _BIN_CALIBRATORS:
                return None                         # the reference's if/elif chain builds nothing for other names
            method, kwargs = _BIN_CALIBRATORS[base_bin_calibrator_name]
            val_probs, _, _ = self._val_probs()
            labels = self._val_labels()
            if self.procal_flag:
                self._need_proximity(val_image_proximity, "bin_based calibration with procal_flag")
                base_calibrator = BinMeanShift(base_bin_calibrator_name, method, bin_strategy="quantile",
                                               normalize_conf=False, proximity_bin=5, **kwargs)
                base_calibrator.fit_transform(val_probs, val_image_proximity, labels)
            elif base_bin_calibrator_name == "multi_isotonic_regression":
                base_calibrator = method()
                base_calibrator.fit_transform(val_probs, labels)
            else:
                base_calibrator = method(**kwargs)
                base_calibrator.fit(val_probs, labels)
            return base_calibrator
        if not (self.base_calibration_mode == "scaling_based" and self.procal_flag):
            return None                             # the reference builds nothing in this case either
        self._need_proximity(val_image_proximity, "scaling_based (density-ratio) calibration")
        val_probs, val_preds, _ = self._val_probs()
        labels = self._val_labels().cpu().numpy()
        base_calibrator = DensityRatioCalibration()
        base_calibrator.fit(val_probs, val_preds.cpu().numpy(), labels, val_image_proximity)
        return base_calibrator

    @staticmethod
    def _need_proximity(val_image_proximity, what: str) -> None:
        if val_image_proximity is None:
            raise ValueError(f"{what} needs val_dict['val_image_knn_dists'] (the validation images' kNN distances, "
                             "reference vl_calibrator.py:68)")

    def _val_probs(self):
        """softmax of the un-scaled validation logits (reference :59-60) as a float32 CUDA matrix, with argmax / max."""
        val_logits = self.val_dict["val_logits"]
        x = val_logits.detach() if isinstance(val_logits, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(val_logits))
        val_probs = x.to(device="cuda", dtype=torch.float32, copy=True).contiguous()
        pred, conf = native.dac_softmax_logits_(val_probs, None)
        return val_probs, pred, conf

    def _val_labels(self) -> torch.Tensor:
        labels = self.val_dict["val_labels"]
        t = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(labels))
        return t.to(torch.int64)

    def build_dac_calibrator(self, text_feature_dict, k_dac):
        dac_calibrator = DistanseAwareCalibration()
        dac_calibrator.fit(text_feature_dict["base_text_features_zs"], text_feature_dict["current_text_features_zs"],
                           text_feature_dict["base_text_features_tuned"],
                           text_feature_dict["current_text_features_tuned"], k=k_dac)
        return dac_calibrator

    # ------------------------------------------------------------------ reference contract
    def predict(self, logits, test_proximity):
        assert logits.shape[0] == test_proximity.shape[0], \
            f"Shape mismatch: logits shape {logits.shape[0]} != test_proximity shape {test_proximity.shape[0]}"
        as_numpy = not isinstance(logits, torch.Tensor)
        x = torch.from_numpy(np.ascontiguousarray(logits)) if as_numpy else logits.detach()
        # DAC scaling + softmax happen in place in one CUDA kernel (ccal_dac_softmax_logits); the
        # [N, C] probability matrix exists only because the reference contract returns it.
        # Without DAC the reference keeps float64 probabilities; here they are float32.
        work = x.to(device="cuda", dtype=torch.float32, copy=True).contiguous()
        cc = self.dac_calibrator._cc() if self.dac_calibrator is not None else None
        native.dac_softmax_logits_(work, cc)
        if self.base_calibrator is not None:          # float64 [N, C] like the reference
            if self.base_calibration_mode == "scaling_based":
                prox = test_proximity if isinstance(test_proximity, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(test_proximity))
                work, _, _ = self.base_calibrator.predict_device(work, prox.cuda())
            elif self.procal_flag:
                work = self.base_calibrator.transform(work, test_proximity)
            else:
                work = self.base_calibrator.transform(work)
        return work.cpu().numpy() if as_numpy else work

    # ------------------------------------------------------------------ additive
    def predict_confidence(self, logits):
        cc = self.dac_calibrator._cc() if self.dac_calibrator is not None else None
        as_numpy = not isinstance(logits, torch.Tensor)
        x = torch.from_numpy(np.ascontiguousarray(logits)) if as_numpy else logits.detach()
        pred, conf = native.logits_confidence(x.to(device="cuda", dtype=torch.float32).contiguous(), cc)
        return (pred.cpu().numpy().astype(np.int64), conf.cpu().numpy()) if as_numpy else (pred, conf)

    def predict_from_features(self, image_features, text_features=None, logit_scale=100.0):
        dac = self.dac_calibrator if self.dac_calibrator is not None else DistanseAwareCalibration()
        if text_features is None:
            text_features = self.text_feature_dict["current_text_features_tuned"]
        return dac.predict_from_features(image_features, text_features, logit_scale)
