/* Pure-C client of libsola_maskpath.so: no Python, no torch — the drop-in boundary is a C ABI.
 * Build: nvcc (or gcc + -lcudart) tests/c_abi_smoke.c -Iinclude -Lsola_b200/lib -lsola_maskpath -o c_abi_smoke
 * Checks K1 (planes + the three stability counts), K3 (per-frame counts), the J&F accumulators, the fused J&F sweep (unit table)
 * and the fused K1+R1 entry point against plain C loops. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sola_maskpath.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(void) {
  const int n = 3, H = 70, W = 150, Wp = (W + 31) / 32;            /* W % 32 != 0 on purpose */
  const size_t px = (size_t)n * H * W;
  float* h = (float*)malloc(px * sizeof(float));
  unsigned s = 12345u;
  for (size_t i = 0; i < px; ++i) { s = s * 1664525u + 1013904223u; h[i] = ((int)(s >> 8) % 4001 - 2000) / 1000.0f; }   /* [-2, 2] */
  h[5] = 0.0f; h[6] = -0.0f; h[7] = 1.0f; h[8] = -1.0f; h[9] = NAN;
  float *d; uint32_t *packed; int *cnt;
  CK(cudaMalloc((void**)&d, px * sizeof(float)));
  CK(cudaMalloc((void**)&packed, (size_t)n * H * Wp * 4));
  CK(cudaMalloc((void**)&cnt, 3 * n * sizeof(int)));
  CK(cudaMemcpy(d, h, px * sizeof(float), cudaMemcpyHostToDevice));
  if (strcmp(sola_build_arch(), "sm_100a") != 0) { printf("unexpected arch %s\n", sola_build_arch()); return 1; }
  int rc = sola_binarize_pack_f32(d, n, H, W, 0.0, 1.0, packed, cnt, cnt + n, cnt + 2 * n, 0);
  if (rc) { printf("sola_binarize_pack_f32: %s\n", sola_last_error_string()); return 1; }
  uint32_t* hp = (uint32_t*)malloc((size_t)n * H * Wp * 4);
  int hc[9];
  CK(cudaMemcpy(hp, packed, (size_t)n * H * Wp * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost));
  for (int f = 0; f < n; ++f) {
    int hi = 0, mid = 0, lo = 0;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float v = h[((size_t)f * H + y) * W + x];
        hi += v > 1.0f; mid += v > 0.0f; lo += v > -1.0f;
        const int bit = (hp[((size_t)f * H + y) * Wp + (x >> 5)] >> (x & 31)) & 1;
        if (bit != (v > 0.0f)) { printf("plane mismatch f=%d y=%d x=%d\n", f, y, x); return 1; }
      }
    if (hc[f] != hi || hc[n + f] != mid || hc[2 * n + f] != lo) { printf("count mismatch frame %d\n", f); return 1; }
  }
  /* K3 on raw fp32 {0,1} planes: |A & B|, |A|, |B| per frame */
  float *a = (float*)malloc(px * sizeof(float)), *b = (float*)malloc(px * sizeof(float)), *da, *db;
  for (size_t i = 0; i < px; ++i) { a[i] = h[i] > 0.0f; b[i] = h[i] > 0.5f || h[i] < -1.5f; }
  CK(cudaMalloc((void**)&da, px * sizeof(float))); CK(cudaMalloc((void**)&db, px * sizeof(float)));
  CK(cudaMemcpy(da, a, px * sizeof(float), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b, px * sizeof(float), cudaMemcpyHostToDevice));
  rc = sola_frame_counts_f32(da, db, n, (long long)H * W, cnt, cnt + n, cnt + 2 * n, 0);
  if (rc) { printf("sola_frame_counts_f32: %s\n", sola_last_error_string()); return 1; }
  CK(cudaMemcpy(hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost));
  for (int f = 0; f < n; ++f) {
    int i_ = 0, na = 0, nb = 0;
    for (size_t k = (size_t)f * H * W; k < (size_t)(f + 1) * H * W; ++k) { i_ += a[k] != 0 && b[k] != 0; na += a[k] != 0; nb += b[k] != 0; }
    if (hc[f] != i_ || hc[n + f] != na || hc[2 * n + f] != nb) { printf("K3 mismatch frame %d\n", f); return 1; }
  }
  /* J&F accumulators: per-frame inter / union and the exact tp / fp / fn of the unit (evaluator.py:227-247) */
  {
    int* iu; long long* tot; int hiu[6]; long long htot[3], tp = 0, fp = 0, fn = 0;
    CK(cudaMalloc((void**)&iu, 2 * n * sizeof(int))); CK(cudaMalloc((void**)&tot, 3 * sizeof(long long)));
    rc = sola_jf_f32(da, db, n, (long long)H * W, iu, iu + n, tot, 0);
    if (rc) { printf("sola_jf_f32: %s\n", sola_last_error_string()); return 1; }
    CK(cudaMemcpy(hiu, iu, sizeof(hiu), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(htot, tot, sizeof(htot), cudaMemcpyDeviceToHost));
    for (int f = 0; f < n; ++f) {
      int i_ = 0, u_ = 0;
      for (size_t k = (size_t)f * H * W; k < (size_t)(f + 1) * H * W; ++k) {
        const int pa = a[k] != 0, pb = b[k] != 0;
        i_ += pa && pb; u_ += pa || pb; tp += pa && pb; fp += pa && !pb; fn += !pa && pb;
      }
      if (hiu[f] != i_ || hiu[n + f] != u_) { printf("sola_jf mismatch frame %d\n", f); return 1; }
    }
    if (htot[0] != tp || htot[1] != fp || htot[2] != fn) { printf("sola_jf totals mismatch\n"); return 1; }
  }
  /* fused K1 + R1: the resized planes must equal R1 applied to K1's planes */
  const int oh = 54, ow = 96, owp = (ow + 31) / 32;
  uint32_t *r_fused, *r_two, *packed2;
  CK(cudaMalloc((void**)&r_fused, (size_t)n * oh * owp * 4)); CK(cudaMalloc((void**)&r_two, (size_t)n * oh * owp * 4));
  CK(cudaMalloc((void**)&packed2, (size_t)n * H * Wp * 4));
  rc = sola_binarize_pack_resize_f32(d, n, H, W, oh, ow, 0.0, 1.0, packed2, r_fused, NULL, NULL, NULL, NULL, 0);
  if (rc) { printf("sola_binarize_pack_resize_f32: %s\n", sola_last_error_string()); return 1; }
  rc = sola_resize_bilinear_bin_packed(packed, n, H, W, oh, ow, r_two, NULL, 0);
  if (rc) { printf("sola_resize_bilinear_bin_packed: %s\n", sola_last_error_string()); return 1; }
  uint32_t *h1 = (uint32_t*)malloc((size_t)n * oh * owp * 4), *h2 = (uint32_t*)malloc((size_t)n * oh * owp * 4);
  CK(cudaMemcpy(h1, r_fused, (size_t)n * oh * owp * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2, r_two, (size_t)n * oh * owp * 4, cudaMemcpyDeviceToHost));
  if (memcmp(h1, h2, (size_t)n * oh * owp * 4) != 0) { printf("fused != K1 then R1\n"); return 1; }
  /* fused J&F sweep through the unit table: the struct layout as a C client sees it; region rows and |b(pred)| vs plain C loops */
  {
    uint32_t* packed_b; int* jf; sola_jf_unit unit, *unit_dev; sola_jf_plan plan;
    CK(cudaMalloc((void**)&packed_b, (size_t)n * H * Wp * 4));
    rc = sola_threshold_pack_f32(d, n, H, W, 0.5, packed_b, NULL, 0);
    if (rc) { printf("sola_threshold_pack_f32: %s\n", sola_last_error_string()); return 1; }
    memset(&unit, 0, sizeof(unit)); memset(&plan, 0, sizeof(plan));      /* plan.reserved = 0: automatic tile class */
    unit.pred = packed; unit.gt = packed_b; unit.T = n; unit.H = H; unit.W = W; unit.radius = 2;
    if (sizeof(sola_jf_unit) != 64) { printf("sola_jf_unit is %zu bytes\n", sizeof(sola_jf_unit)); return 1; }
    rc = sola_jf_sweep_plan(&unit, 1, &plan);
    if (rc || plan.total_frames != n || plan.n_items < n) { printf("sola_jf_sweep_plan: %s\n", sola_last_error_string()); return 1; }
    CK(cudaMalloc((void**)&unit_dev, sizeof(unit))); CK(cudaMemcpy(unit_dev, &unit, sizeof(unit), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&jf, 7 * n * sizeof(int)));
    rc = sola_jf_sweep(unit_dev, 1, &plan, jf, 0);
    if (rc) { printf("sola_jf_sweep: %s\n", sola_last_error_string()); return 1; }
    int hjf[21];
    CK(cudaMemcpy(hjf, jf, sizeof(hjf), cudaMemcpyDeviceToHost));
    for (int f = 0; f < n; ++f) {
      int i_ = 0, na = 0, nb = 0, nbf = 0;
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          const float* fr = h + (size_t)f * H * W;
          const int pa = fr[y * W + x] > 0.0f, pb = fr[y * W + x] > 0.5f;
          i_ += pa && pb; na += pa; nb += pb;
          /* seg2bmap of the prediction (DAVIS): xor with east / south / south-east, last row / column / corner rules */
          const int e = x + 1 < W ? fr[y * W + x + 1] > 0.0f : 0, so = y + 1 < H ? fr[(y + 1) * W + x] > 0.0f : 0;
          const int se = (x + 1 < W && y + 1 < H) ? fr[(y + 1) * W + x + 1] > 0.0f : 0;
          int bb = (pa ^ e) | (pa ^ so) | (pa ^ se);
          if (y == H - 1) bb = pa ^ e;
          if (x == W - 1) bb = pa ^ so;
          if (y == H - 1 && x == W - 1) bb = 0;
          nbf += bb;
        }
      if (hjf[f] != i_ || hjf[n + f] != na || hjf[2 * n + f] != nb || hjf[3 * n + f] != nbf) { printf("sola_jf_sweep mismatch frame %d\n", f); return 1; }
      if (hjf[5 * n + f] > hjf[3 * n + f] || hjf[6 * n + f] > hjf[4 * n + f]) { printf("match counts exceed boundary sizes\n"); return 1; }
    }
  }
  /* error path: status + message, no exception */
  if (sola_binarize_pack_f32(NULL, 1, 4, 4, 0.0, 1.0, NULL, NULL, NULL, NULL, 0) != SOLA_ERR_INVALID) { printf("expected SOLA_ERR_INVALID\n"); return 1; }
  printf("c_abi_smoke ok (version %d, %llu launches, last error: %s)\n", sola_version(), sola_launch_count(), sola_last_error_string());
  return 0;
}
