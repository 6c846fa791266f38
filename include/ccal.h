/*
 * ccal.h — C ABI of the B200-native scoring + calibration + calibration-metrics path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.  The
 * reference (ml-stat-Sustech/CLIP_Calibration) is pure Python, so the binding a maintainer
 * adds is a ctypes stub (shown in INTEGRATION.md); each entry point names the reference
 * code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless its name ends in _host;
 *   - row-major, contiguous; feature matrices are [rows, D] with D contiguous;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it, the caller
 *     synchronises (the one exception, ccal_isotonic_fit_binary, returns a size to the host and says so);
 *   - scratch memory: a call that needs any takes it from the device's stream-ordered pool
 *     (cudaMallocAsync / cudaFreeAsync on `stream`) and returns it before the call ends in stream order; the
 *     library keeps no device state between calls (TMA descriptors are built per call) and is re-entrant;
 *   - return value 0 = OK, otherwise a CCAL_ERR_* code; ccal_last_error() gives the
 *     thread-local message.  There is no CPU fallback: a device that is not sm_100 fails.
 *   - bin tables are [(n_thr+1)][3] unsigned 64-bit {count, n_correct, sum(round(conf*2^40))},
 *     bin index of a confidence x = number of thresholds <= x (compared in double).
 *     Kernels ACCUMULATE into the table (the caller zeroes it), so shards / chunks / ranks
 *     add up exactly and in any order.
 */
#ifndef CCAL_H_
#define CCAL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCAL_VERSION 100

enum {
  CCAL_OK = 0,
  CCAL_ERR_BAD_ARG = 1,      /* null pointer, negative size, unsupported shape / alignment */
  CCAL_ERR_UNSUPPORTED = 2,  /* device is not sm_100 / feature width not supported          */
  CCAL_ERR_CUDA = 3          /* a CUDA runtime / driver call failed                         */
};

/* operand dtypes of the feature matrices */
enum { CCAL_F32 = 0, CCAL_F16 = 1, CCAL_BF16 = 2 };

#define CCAL_FX_SHIFT 40        /* fixed-point scale of the per-bin confidence sums        */
#define CCAL_MAX_THRESHOLDS 63  /* n_thr upper bound (64 bins)                              */
#define CCAL_MAX_K 16           /* nearest-neighbour count upper bound                      */

typedef void* ccal_stream_t;

#if defined(__GNUC__)
#define CCAL_API __attribute__((visibility("default")))
#else
#define CCAL_API
#endif

CCAL_API int ccal_version(void);
CCAL_API const char* ccal_last_error(void);

/* Number of CUDA kernels this library has launched in this process (monotonic). */
CCAL_API long long ccal_launch_count(void);

/* Development aid: with CCAL_TRACE_MARKS=1 in the environment the DAC-fit launch chain records named events on its
 * streams; this writes "seq:stream:id[*PENDING*]" for the most recent ones into buf (NUL-terminated) and returns the
 * length - the marks the device has not reached yet show where a stream stopped.  Empty without the variable. */
CCAL_API int ccal_trace_marks_report(char* buf, int cap);

/* 0 if the current device can run this library (compute capability 10.x). */
CCAL_API int ccal_check_device(void);

/* ---- K2: fused scoring --------------------------------------------------------------
 * Replaces, without ever materialising logits:
 *   logits = logit_scale * image_features @ text_features.t()
 *                                   (trainers/classification/zsclip.py:97-102, coop.py:215-217,
 *                                    trainers/calibration/tempscaling.py:53-56)
 *   DistanseAwareCalibration.predict (trainers/calibration/distanse_aware_calibration.py:49-58)
 *   scipy softmax                    (trainers/calibration/vl_calibrator.py:91)
 *   argmax + confidence gather       (evaluators/vl_evaluator.py:68, :83)
 *   and, when `table` is given, the binning of tools/metrics.py:104-127.
 *
 * img [n,d], txt [c,d] in `dtype`, d a multiple of 64 and <= 1024, base pointers 16-byte aligned.
 * CCAL_BF16 / CCAL_F16: operands go to the tensor cores as they are (products exact, fp32 accumulate).
 * CCAL_F32: every operand is split on the fly into an fp16 pair (x*2^e = hi + lo) and each K step issues
 * hi.hi + hi.lo + lo.hi - fp32-grade logits (error < 2^-21 |a||b|) at 3x the tensor work; uses a transient
 * stream-ordered workspace and scores 262,144 rows per launch.  class_conf [c] float or NULL (= all ones, plain softmax).
 * pred_out [n] int32, conf_out [n] float, rowmax_out [n] float (logit_scale * max cosine) may
 * each be NULL.  labels [n] int64 + thresholds_host [n_thr] (HOST doubles) + table (device)
 * enable the fused binning; pass table = NULL to skip it.
 * Pass 1 = row max / argmax over all text tiles; pass 2 recomputes the tiles and sums
 * exp(cc[pred] * (logit - max)); confidence = 1 / sum.  Ties: lowest class index.
 */
CCAL_API int ccal_score_fused(const void* img, const void* txt, const float* class_conf, float logit_scale,
                     int64_t n, int c, int d, int dtype,
                     int32_t* pred_out, float* conf_out, float* rowmax_out,
                     const int64_t* labels, const double* thresholds_host, int n_thr,
                     unsigned long long* table, ccal_stream_t stream);

/* Large 16-bit shards (n >= 18,944 rows, n*c >= 1e9, d a multiple of 128 and <= 768; CCAL_SCORE_FP8=0/1 in the
 * environment forces it off / on) run ccal_score_fused as FP8-guess -> bf16-verify -> redo: pass 1 is done on e4m3
 * copies of the operands and only guesses the argmax; the bf16 pass sums exp at the guessed class's multiplier while
 * tracking the exact bf16 maximum / first argmax; rows whose exact argmax carries another multiplier are redone
 * (pass 2 only).  pred_out is exactly the two-pass result; conf_out agrees to a few ulp.
 * ccal_score_guess_stats: cumulative {rows scored through that pipeline, rows redone} on the current device
 * (synchronises the device); reset != 0 zeroes the counters afterwards. */
CCAL_API int ccal_score_guess_stats(unsigned long long* out2_host, int reset);

/* In-kernel timing of the scoring kernels since the last reset, per kind k = 0 FP8 guess pass, 1 bf16 verify pass,
 * 2 redo, 3 two-pass kernel, 4 temperature-scaling kernel: out_host[4k..4k+3] = {launches, summed launch spans in ns
 * (first CTA in -> last CTA out, %globaltimer), summed CTA busy ns, summed CTA SM cycles (clock64)} - busy cycles /
 * busy ns is the SM clock the kernel really ran at.  Synchronises the device.  out_host holds 20 values. */
CCAL_API int ccal_score_trace(unsigned long long* out_host, int reset);

/* Two-launch form of ccal_score_fused, for pipelines in which the features are on the device before the per-class
 * multipliers are (the DAC fit still running on another stream): ccal_score_pass1 needs only the features and writes,
 * per image, the maximum of the RAW dot products (not scaled) and the first argmax; ccal_score_pass2 takes both back
 * and finishes exactly like ccal_score_fused - confidence, optional scaled row maximum, optional bin table - with
 * bit-identical results.  fp16 / bf16 operands only; no column-split small-batch mode.
 */
CCAL_API int ccal_score_pass1(const void* img, const void* txt, int64_t n, int c, int d, int dtype,
                     float* rowdot_max_out, int32_t* pred_out, ccal_stream_t stream);
CCAL_API int ccal_score_pass2(const void* img, const void* txt, const float* class_conf, float logit_scale,
                     int64_t n, int c, int d, int dtype, const float* rowdot_max_in, const int32_t* pred_in,
                     float* conf_out, float* rowmax_out, const int64_t* labels, const double* thresholds_host,
                     int n_thr, unsigned long long* table, ccal_stream_t stream);

/* ---- K5: temperature-scaling objective ------------------------------------------------
 * loss = F.cross_entropy(exp(log_scale) * img @ txt.T, labels) and d loss / d log_scale, for
 * trainers/calibration/tempscaling.py:31-41 (ScaleLearner), :53-56, :155-160.
 * Same operand rules as ccal_score_fused.  row_ws [2*n] float scratch (per-row loss and
 * gradient terms, reduced in a fixed order => deterministic).  out2 [2] double = {loss, grad}.
 */
CCAL_API int ccal_ts_loss_grad(const void* img, const void* txt, const int64_t* labels, float log_scale,
                      int64_t n, int c, int d, int dtype, float* row_ws, double* out2,
                      ccal_stream_t stream);

/* Device-side training loop of the scalar (tempscaling.py:146-169 + dassl's SGD): ccal_ts_loss_grad_dev is
 * ccal_ts_loss_grad with the log-scale read from DEVICE memory (a double), and ccal_sgd_scalar_step applies one
 * momentum-SGD step to it on the device: state = {t, velocity, sum of batch losses, batches};
 * g = grad + weight_decay * t; v = momentum * v + g; t -= lr * v.  A whole 20-epoch fit is then a stream of launches
 * with no host synchronisation.  fp16 / bf16 operands only. */
CCAL_API int ccal_ts_loss_grad_dev(const void* img, const void* txt, const int64_t* labels, const double* log_scale_dev,
                          int64_t n, int c, int d, int dtype, float* row_ws, double* out2, ccal_stream_t stream);
CCAL_API int ccal_sgd_scalar_step(double* state, const double* loss_grad, double lr, double momentum,
                         double weight_decay, ccal_stream_t stream);

/* ---- K1: k nearest rows by Euclidean distance + the DAC map ---------------------------
 * ccal_knn_l2: for every query row q_i [nq,d] the kk = min(k, nr) smallest ||r_j - q_i||_2 over
 * ref rows [nr,d] (fp32), ascending, ties by lowest j.  dist_out [nq,k] (unused tail = +inf),
 * idx_out [nq,k] (unused tail = -1) may be NULL.  drop_first != 0 computes k+1 and drops the
 * nearest (trainers/calibration/proximity.py:49-70); otherwise proximity.py:19-46 and the
 * distance/sort lines of distanse_aware_calibration.py:28-30, :34-36.
 */
CCAL_API int ccal_knn_l2(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k,
                int drop_first, float* dist_out, int32_t* idx_out, ccal_stream_t stream);
/* Same contract, always the exhaustive fp32 scan (ccal_knn_l2 switches to a tcgen05 GEMM filter +
 * exact verification for large problems and uses this scan for the rows it cannot prove). */
CCAL_API int ccal_knn_l2_exhaustive(const float* ref, const float* query, int64_t nr, int64_t nq, int d, int k,
                int drop_first, float* dist_out, int32_t* idx_out, ccal_stream_t stream);

/* ccal_dac_fit = DistanseAwareCalibration.fit (distanse_aware_calibration.py:13-46):
 * class_conf_out[i] = 1 if nearest tuned distance < 0.05 else
 *                     exp(-sum(top-k tuned)/k) / exp(-sum(top-k zero-shot)/k).
 * Inputs fp32 [b,d] / [c,d].  The two [c,k] distance buffers are required (they double as
 * workspace); the two [c,k] index buffers may be NULL.
 */
CCAL_API int ccal_dac_fit(const float* base_zs, const float* cur_zs, const float* base_tuned,
                 const float* cur_tuned, int b, int c, int d, int k,
                 float* class_conf_out, int32_t* knn_idx_zs_out, int32_t* knn_idx_tuned_out,
                 float* knn_dist_zs_out, float* knn_dist_tuned_out, ccal_stream_t stream);

/* ccal_dac_fit_f16: the same fit in the reference's OWN arithmetic when its inputs are float16 numpy arrays (its default
 * precision: train.py:152; np.linalg.norm / np.sum / np.exp keep float16, distanse_aware_calibration.py:28-42): every
 * elementwise result rounded to half, squares summed in float32 in numpy's pairwise order, one rounding per row, half
 * sqrt / sum / exp / ratio, the "< 0.05" test in half.  Inputs are IEEE half [b,d] / [c,d]; outputs as ccal_dac_fit
 * (class_conf_out holds half values widened to float; distance / index buffers may be NULL).  b <= 16,384.
 * Opt-in exact-parity mode (DistanseAwareCalibration.fit(..., arithmetic="input")); ccal_dac_fit is the accurate one. */
CCAL_API int ccal_dac_fit_f16(const void* base_zs, const void* cur_zs, const void* base_tuned, const void* cur_tuned,
                     int b, int c, int d, int k, float* class_conf_out, int32_t* knn_idx_zs_out,
                     int32_t* knn_idx_tuned_out, float* knn_dist_zs_out, float* knn_dist_tuned_out,
                     ccal_stream_t stream);

/* ---- K4: materialised-logits drop-ins -------------------------------------------------
 * ccal_dac_predict_logits = DistanseAwareCalibration.predict (:49-58): in place,
 * logits[i,:] *= class_conf[argmax_j logits[i,j]].  pred_out may be NULL.
 * ccal_logits_confidence: pred/conf of softmax(class_conf[pred] * logits) without writing the
 * scaled logits or the probabilities (vl_calibrator.py:91 + vl_evaluator.py:68,:83);
 * class_conf may be NULL.
 */
CCAL_API int ccal_dac_predict_logits(float* logits, const float* class_conf, int64_t n, int c,
                            int32_t* pred_out, ccal_stream_t stream);
CCAL_API int ccal_logits_confidence(const float* logits, const float* class_conf, int64_t n, int c,
                           int32_t* pred_out, float* conf_out, ccal_stream_t stream);
/* ccal_dac_softmax_logits: VLCalibration.predict (vl_calibrator.py:83-109, DAC + softmax branch):
 * in place, logits[i,:] <- softmax(class_conf[pred_i] * logits[i,:]); class_conf may be NULL.
 * ccal_row_argmax: first argmax and row maximum of an [n,c] matrix (evaluators/vl_evaluator.py:68,
 * :83 on a probability matrix).  Output pointers may be NULL. */
CCAL_API int ccal_dac_softmax_logits(float* logits, const float* class_conf, int64_t n, int c,
                            int32_t* pred_out, float* conf_out, ccal_stream_t stream);
CCAL_API int ccal_row_argmax(const float* values, int64_t n, int c, int32_t* pred_out, float* max_out,
                    ccal_stream_t stream);

/* ---- K3: bin statistics for ECE / MCE / ACE / PIECE -----------------------------------
 * tools/metrics.py:104-127 (ECE), :195-206 (MCE), :228-234 (AdaptiveECE) all reduce to
 * per-bin {count, n_correct, sum conf}.  conf is float (conf_f64 = 0) or double (1).
 * pred may be int32 (pred_i64 = 0) or int64 (1); gt is int64.
 * Optional second key (PIECE, :152-168): key2 [n] float with thresholds2_host [n_thr2]; the
 * table is then [(n_thr2+1)][(n_thr+1)][3].  Pass key2 = NULL, n_thr2 = 0 for the 1-D table.
 */
CCAL_API int ccal_bin_stats(const void* conf, int conf_f64, const void* pred, int pred_i64,
                   const int64_t* gt, int64_t n, const double* thresholds_host, int n_thr,
                   const float* key2, const double* thresholds2_host, int n_thr2,
                   unsigned long long* table, ccal_stream_t stream);

/* ---- exact order statistics of float keys (quantile bin edges for ACE / PIECE) ---------
 * 16-bit radix histograms over the order-preserving bit pattern of non-negative floats.
 * level 0: hist[65536] += count of (bits >> 16).  level 1: for each of the n_prefix given
 * high halves, hist[p][65536] += count of (bits & 0xffff) among keys with that high half.
 * Counts are uint32 (n < 2^32 per call); tables accumulate (caller zeroes) so ranks can be
 * all-reduced.
 */
CCAL_API int ccal_radix_hist(const float* keys, int64_t n, int level, const uint32_t* prefixes_host,
                    int n_prefix, uint32_t* hist, ccal_stream_t stream);

/* ---- per-class confusion counts (macro-F1, evaluators/vl_evaluator.py:74-79) ---------------
 * counts[cls][3] += {tp: pred == gt == cls, fp: pred == cls != gt, fn: gt == cls != pred}; uint64, accumulating
 * (caller zeroes), so per-rank tables can be all-reduced.  Labels / predictions outside [0, c) are ignored.
 */
CCAL_API int ccal_class_counts(const void* pred, int pred_i64, const int64_t* gt, int64_t n, int c,
                      unsigned long long* counts, ccal_stream_t stream);

/* ---- density-ratio (proximity-informed) calibration --------------------------------------
 * Reference: trainers/calibration/density_ratio_calibration.py, DensityRatioCalibration.
 * ccal_kde2_pdf = `sm.nonparametric.KDEMultivariate(data=[conf, proximity], var_type='cc').pdf(points)`
 * (:66, :70, :104-105): pdf_out[q] = 1/(m 2 pi bw_x bw_y) * sum_i exp(-(x_i-qx)^2/(2 bw_x^2) - (y_i-qy)^2/(2 bw_y^2)),
 * float64 in and out (device pointers), valid down to the float64 underflow limit.
 * ccal_density_ratio_apply = the rest of .predict (:108-117): conf_cal = t / max(t + f*ratio, 1e-10); in every row
 * the first-argmax class gets conf_cal and the other classes are rescaled to sum to 1 - conf_cal.  probs is fp32
 * or fp64 (exactly one pointer non-NULL); probs_out is float64 [n,c]; conf_cal_out / pred_out may be NULL.
 */
CCAL_API int ccal_kde2_pdf(const double* data_x, const double* data_y, int64_t m, const double* query_x,
                  const double* query_y, int64_t n, double bw_x, double bw_y, double* pdf_out,
                  ccal_stream_t stream);
CCAL_API int ccal_density_ratio_apply(const float* probs_f32, const double* probs_f64, int64_t n, int c,
                             const double* pdf_true, const double* pdf_false, double false_true_ratio,
                             double* probs_out, double* conf_cal_out, int32_t* pred_out, ccal_stream_t stream);

/* ---- multi-class isotonic-regression calibrator -------------------------------------------
 * Reference: trainers/calibration/multi_isotonic_regression.py:14-35 (scikit-learn IsotonicRegression(out_of_bounds=
 * 'clip') on the flattened probabilities against the flattened one-hot labels) and BinMeanShift,
 * multi_proximity_isotonic.py:130-247.
 * ccal_exp_normalise_rows: out[i,:] = exp(v[i,:]) / sum_j exp(v[i,j]) in float64 without a max shift (:26, :33); v is
 *   fp32 or fp64 (exactly one pointer non-NULL); with labels [n] int64, onehot_out [n,c] uint8 gets (labels[i] == j).
 * ccal_isotonic_fit_binary: scikit-learn's fit for 0/1 targets y [n] uint8 at points x [n] float64: knots_x / knots_y
 *   (device, capacity n) receive X_thresholds_ / y_thresholds_, *n_knots_host their number.  SYNCHRONISES the stream.
 * ccal_isotonic_transform: out[i] = f(clip(t[i], knots_x[0], knots_x[n_knots-1])) + residual_scale * t[i], f = linear
 *   interpolation between the knots (constant for a single knot); the reference uses residual_scale = 1e-9.
 */
CCAL_API int ccal_exp_normalise_rows(const float* v_f32, const double* v_f64, int64_t n, int c, double* out,
                            const int64_t* labels, unsigned char* onehot_out, ccal_stream_t stream);
CCAL_API int ccal_isotonic_fit_binary(const double* x, const unsigned char* y, int64_t n, double* knots_x,
                             double* knots_y, int64_t* n_knots_host, ccal_stream_t stream);
CCAL_API int ccal_isotonic_transform(const double* knots_x, const double* knots_y, int64_t n_knots, const double* t,
                            int64_t n, double residual_scale, double* out, ccal_stream_t stream);

/* ---- one-vs-all bin-based calibrators (netcal.binning.HistogramBinning / IsotonicRegression) -----------------
 * Reference call sites: trainers/calibration/vl_calibrator.py:20-21, :125-131 (under BinMeanShift), :137-143 (plain).
 * netcal itself is a pip dependency (requirements.txt:6, unpinned, not vendored): the kernels follow netcal 1.3's
 * published multi-class scheme - class j gets a BINARY calibrator fitted on (X[:, j], y == j); transform applies
 * calibrator j to column j and divides each row by its sum (AbstractCalibration._create_one_vs_all_models /
 * _calibrate_multiclass).  Parity with netcal is unpinned; the scikit-learn isotonic fit inside is pinned.
 * ccal_ova_hist_fit: count[j, b] / hits[j, b] (uint32 [c, n_bins], zeroed here) = number of rows whose p[i, j] falls
 *   into bin b of `edges` (n_bins + 1 float64 values: edges[b] <= x < edges[b+1], last bin closed) / those with
 *   labels[i] == j.  p is fp32 or fp64 (exactly one pointer non-NULL), n_bins <= 64.
 * ccal_ova_apply: out[i, j] = bin_map[j, bin(p[i, j])] (bin_map float64 [c, n_bins] given) or the isotonic function of
 *   class j at p[i, j] (knot_off int32 [c + 1] given: knots of class j are knots_x/knots_y[knot_off[j] .. knot_off[j+1]),
 *   clip + linear interpolation as ccal_isotonic_transform; a class without knots gives 0); normalise != 0 divides each
 *   row by its sum.
 */
CCAL_API int ccal_ova_hist_fit(const float* p_f32, const double* p_f64, int64_t n, int c, const int64_t* labels,
                      const double* edges, int n_bins, uint32_t* count, uint32_t* hits, ccal_stream_t stream);
CCAL_API int ccal_ova_apply(const float* p_f32, const double* p_f64, int64_t n, int c, const double* edges, int n_bins,
                   const double* bin_map, const double* knots_x, const double* knots_y, const int32_t* knot_off,
                   int normalise, double* out, ccal_stream_t stream);

/* ---- device primitives of the bin-based calibrators (csrc/sort_scan.cuh), exported for tests ------------------
 * The isotonic fit (sklearn/isotonic.py: `order = np.lexsort((y, X))`, `_make_unique`, trim_duplicates) needs a stable
 * sort by x, prefix sums and a flagged compaction; these are the library's own kernels, not a vendor library's.
 * ccal_sort_pairs_f64_u8: stable ascending sort of (keys[i], vals[i]), n < 2^31, not in place; -0.0 < +0.0.
 *   SYNCHRONISES the stream (which of the 8 digit passes run is decided on the host).
 * ccal_prefix_sum_i32: out[i] = in[0] + ... + in[i] (inclusive != 0) or ... + in[i-1]; out may be in; sums must fit int32.
 */
CCAL_API int ccal_sort_pairs_f64_u8(const double* keys, const unsigned char* vals, int64_t n, double* keys_out,
                           unsigned char* vals_out, ccal_stream_t stream);
CCAL_API int ccal_prefix_sum_i32(const int* in, int* out, int64_t n, int inclusive, ccal_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CCAL_H_ */
