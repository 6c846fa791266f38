/*
 * c_abi_smoke.c - the C ABI of libccal.so used from plain C: no Python, no torch.
 *
 *   nvcc -O2 -o c_abi_smoke examples/c_abi_smoke.c -Iinclude -Lclip_calibration_b200 -lccal \
 *        -Xlinker -rpath -Xlinker '$ORIGIN'        (built by clip_calibration_b200/build.py)
 *
 * Scores a small synthetic problem with ccal_score_fused (fused binning on), re-bins the returned
 * (pred, conf) with ccal_bin_stats, and checks everything against a double-precision loop on the host:
 * logits = s * img . txt, pred = first argmax, conf = 1 / sum exp(cc[pred] * (l_j - l_max)).
 * Exit code 0 = all checks passed.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ccal.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define CHECK_CCAL(x) do { int r_ = (x); if (r_ != 0) { \
  fprintf(stderr, "ccal error %d: %s (%s:%d)\n", r_, ccal_last_error(), __FILE__, __LINE__); return 3; } } while (0)

static uint16_t f32_to_bf16(float f) {              /* round to nearest even */
  uint32_t u; memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double rnd(void) {                            /* xorshift64*, uniform in [0,1) */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (double)((rng_state * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}
static double gauss(void) { double u = rnd() + 1e-12, v = rnd(); return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }

int main(void) {
  const int n = 777, c = 301, d = 128, n_bins = 10;
  const float scale = 100.0f;
  CHECK_CCAL(ccal_check_device());
  printf("libccal version %d\n", ccal_version());

  uint16_t* img = (uint16_t*)malloc((size_t)n * d * 2);
  uint16_t* txt = (uint16_t*)malloc((size_t)c * d * 2);
  float* cc = (float*)malloc((size_t)c * 4);
  int64_t* labels = (int64_t*)malloc((size_t)n * 8);
  double* row = (double*)malloc((size_t)d * 8);
  for (int j = 0; j < c; ++j) {
    double nrm = 0; for (int k = 0; k < d; ++k) { row[k] = gauss() + 0.8; nrm += row[k] * row[k]; }
    for (int k = 0; k < d; ++k) txt[(size_t)j * d + k] = f32_to_bf16((float)(row[k] / sqrt(nrm)));
    cc[j] = (float)(0.95 + 0.05 * rnd());
  }
  for (int i = 0; i < n; ++i) {
    labels[i] = (int64_t)(rnd() * c);
    double nrm = 0;
    for (int k = 0; k < d; ++k) { row[k] = 0.6 * bf16_to_f32(txt[(size_t)labels[i] * d + k]) + gauss() / sqrt((double)d); nrm += row[k] * row[k]; }
    for (int k = 0; k < d; ++k) img[(size_t)i * d + k] = f32_to_bf16((float)(row[k] / sqrt(nrm)));
  }

  void *d_img, *d_txt; float *d_cc, *d_conf; int32_t* d_pred; int64_t* d_labels; unsigned long long *d_table, *d_table2;
  CHECK_CUDA(cudaMalloc(&d_img, (size_t)n * d * 2)); CHECK_CUDA(cudaMalloc(&d_txt, (size_t)c * d * 2));
  CHECK_CUDA(cudaMalloc((void**)&d_cc, c * 4)); CHECK_CUDA(cudaMalloc((void**)&d_conf, n * 4));
  CHECK_CUDA(cudaMalloc((void**)&d_pred, n * 4)); CHECK_CUDA(cudaMalloc((void**)&d_labels, n * 8));
  CHECK_CUDA(cudaMalloc((void**)&d_table, (n_bins + 1) * 3 * 8)); CHECK_CUDA(cudaMalloc((void**)&d_table2, (n_bins + 1) * 3 * 8));
  CHECK_CUDA(cudaMemcpy(d_img, img, (size_t)n * d * 2, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_txt, txt, (size_t)c * d * 2, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_cc, cc, c * 4, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_labels, labels, n * 8, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemset(d_table, 0, (n_bins + 1) * 3 * 8)); CHECK_CUDA(cudaMemset(d_table2, 0, (n_bins + 1) * 3 * 8));

  double thr[16];
  for (int b = 0; b < n_bins; ++b) thr[b] = (double)(b + 1) / n_bins;      /* ~ np.linspace(0,1,11)[1:] */
  cudaStream_t stream; CHECK_CUDA(cudaStreamCreate(&stream));
  CHECK_CCAL(ccal_score_fused(d_img, d_txt, d_cc, scale, n, c, d, CCAL_BF16, d_pred, d_conf, NULL, d_labels, thr, n_bins,
                              d_table, stream));
  CHECK_CCAL(ccal_bin_stats(d_conf, 0, d_pred, 0, d_labels, n, thr, n_bins, NULL, NULL, 0, d_table2, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));

  int32_t* pred = (int32_t*)malloc(n * 4); float* conf = (float*)malloc(n * 4);
  unsigned long long t1[33], t2[33];
  CHECK_CUDA(cudaMemcpy(pred, d_pred, n * 4, cudaMemcpyDeviceToHost)); CHECK_CUDA(cudaMemcpy(conf, d_conf, n * 4, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(t1, d_table, 33 * 8, cudaMemcpyDeviceToHost)); CHECK_CUDA(cudaMemcpy(t2, d_table2, 33 * 8, cudaMemcpyDeviceToHost));

  /* host reference in double */
  int bad_pred = 0, ties = 0; double max_rel = 0; unsigned long long count = 0, correct = 0;
  double* l = (double*)malloc((size_t)c * 8);
  for (int i = 0; i < n; ++i) {
    int arg = 0; double best = -1e300, second = -1e300;
    for (int j = 0; j < c; ++j) {
      double s = 0; for (int k = 0; k < d; ++k) s += (double)bf16_to_f32(img[(size_t)i * d + k]) * bf16_to_f32(txt[(size_t)j * d + k]);
      l[j] = s * scale;
      if (l[j] > best) { second = best; best = l[j]; arg = j; } else if (l[j] > second) second = l[j];
    }
    double sum = 0; for (int j = 0; j < c; ++j) sum += exp((double)cc[arg] * (l[j] - best));
    if (best - second < 4e-5) { ++ties; continue; }
    if (pred[i] != arg) ++bad_pred;
    double rel = fabs(conf[i] - 1.0 / sum) * sum; if (rel > max_rel) max_rel = rel;
  }
  for (int b = 0; b <= n_bins; ++b) { count += t1[3 * b]; correct += t1[3 * b + 1]; }
  unsigned long long agree = 0; for (int i = 0; i < n; ++i) agree += (pred[i] == labels[i]);
  int tables_equal = memcmp(t1, t2, 33 * 8) == 0;
  printf("n=%d c=%d d=%d: label mismatches %d (ties %d), max conf rel err %.3e, table count %llu correct %llu (%llu), "
         "fused table == bin_stats table: %s, kernels launched %lld\n", n, c, d, bad_pred, ties, max_rel, count, correct, agree,
         tables_equal ? "yes" : "NO", ccal_launch_count());
  int ok = bad_pred == 0 && max_rel < 1e-4 && count == (unsigned long long)n && correct == agree && tables_equal;
  /* error path: a feature width that is not a multiple of 64 must be refused with a message */
  int rc = ccal_score_fused(d_img, d_txt, d_cc, scale, n, c, 100, CCAL_BF16, d_pred, d_conf, NULL, NULL, NULL, 0, NULL, stream);
  printf("bad-shape call -> code %d: %s\n", rc, ccal_last_error());
  ok = ok && rc == CCAL_ERR_BAD_ARG;
  printf(ok ? "C ABI smoke: OK\n" : "C ABI smoke: FAILED\n");
  return ok ? 0 : 1;
}
