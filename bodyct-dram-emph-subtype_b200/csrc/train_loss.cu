// K11 — the training loss of ScanRegLightningModule.shared_step(TRAIN) with its gradient, and K12 — Adam.
//
// K11 replaces, in one reduction pass + one scalar pass (forward) and one element-wise pass (backward), what the
// reference spreads over ~60 ATen calls and their autograd nodes:
//   * the lobe-masked mean of each dense map (med3d.py:383-387: lungs -> nearest to the map's size, sum(map*m)/sum(m));
//   * `_interval_regression_loss` on both scores (models.py:495-506, beta/gamma models.py:412-413);
//   * `_segmentation_loss` (models.py:508-518): BinaryDice(1e-7) of the two masked maps (metrics.py:33-47) and
//     BinaryCrossEntropy (metrics.py:10-30) of clamp(cle + pse, 0, 1) against the LAA mask of scans that carry a
//     positive label (models.py:553-556), smoothness 0.85 inside the lungs;
//   * total = loss_cle + loss_pse + 2 * mul_loss + seg_loss (models.py:565).
// All masks are 0/1 bytes at scan resolution, read through ATen's legacy `nearest` index (F.interpolate, models.py:555-556).
// Sums are fp64, two-phase and fixed-order (deterministic).  K12 is torch.optim.Adam's update (models.py:685-698:
// lr only, betas (0.9, 0.999), eps 1e-8, no weight decay) over one flat fp32 buffer — one launch per step.
#include "common.h"

namespace dram {

constexpr int LS_THREADS = 256;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_SUMS = 8;  // S0, S1, M, I, T, A1, A0, (pad) per sample
constexpr int LS_MAX_B = 64;
constexpr int LS_HEAD = DRAM_LOSS_COEF_HEAD;
constexpr float BCE_EPS = 1e-6f, BCE_SMOOTH = 0.85f, DICE_SMOOTH = 1e-7f;

__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {  // ATen nearest_idx (legacy mode)
  if (out_size == in_size) return dst;
  const float scale = (float)in_size / (float)out_size;
  const int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

struct LossGeom {
  int b, d, h, w, d2, h2, w2;
};

// One warp per (d2, h2) row of a sample's half-resolution maps; blockIdx.y = sample.
__global__ void __launch_bounds__(LS_THREADS) loss_sums_kernel(const float *__restrict__ d0, const float *__restrict__ d1,
                                                              const uint8_t *__restrict__ lungs, const uint8_t *__restrict__ ems,
                                                              const int64_t *__restrict__ cle_labels,
                                                              const int64_t *__restrict__ pse_labels, LossGeom g,
                                                              double *__restrict__ partial) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool flag = cle_labels[b] > 0 || pse_labels[b] > 0;  // models.py:553
  const int rows = g.d2 * g.h2;
  const size_t v2 = (size_t)rows * g.w2;
  const float *m0 = d0 + (size_t)b * v2, *m1 = d1 + (size_t)b * v2;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int r = blockIdx.x * LS_WARPS + warp; r < rows; r += gridDim.x * LS_WARPS) {
    const int zd = r / g.h2, zh = r - zd * g.h2;
    const size_t src_row = (((size_t)b * g.d + nearest_src(zd, g.d, g.d2)) * g.h + nearest_src(zh, g.h, g.h2)) * g.w;
    const uint8_t *lrow = lungs + src_row, *erow = ems + src_row;
    float s0 = 0.f, s1 = 0.f, mm = 0.f, ii = 0.f, tt = 0.f, a1 = 0.f, a0 = 0.f;  // one row: <= a few hundred terms
    for (int x = lane; x < g.w2; x += 32) {
      const int sx = nearest_src(x, g.w, g.w2);
      const float m = lrow[sx] ? 1.f : 0.f;
      const bool t = flag && erow[sx];
      const float p0 = m0[(size_t)r * g.w2 + x], p1 = m1[(size_t)r * g.w2 + x];
      const float y0 = p0 * m, y1 = p1 * m;
      s0 += y0;
      s1 += y1;
      mm += m;
      ii += y0 * y1;
      const float both = fminf(fmaxf(p0 + p1, 0.f), 1.f);
      const float pt = t ? both : 1.f - both;
      const float lp = logf(fminf(fmaxf(pt, BCE_EPS), 1.f - BCE_EPS)) * (m != 0.f ? BCE_SMOOTH : 1.f);
      if (t) {
        tt += 1.f;
        a1 += lp;
      } else {
        a0 += lp;
      }
    }
    acc[0] += s0, acc[1] += s1, acc[2] += mm, acc[3] += ii, acc[4] += tt, acc[5] += a1, acc[6] += a0;
  }
  __shared__ double red[LS_WARPS][7];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double s = 0.0;
    for (int k = 0; k < LS_WARPS; ++k) s += red[k][threadIdx.x];
    partial[((size_t)b * gridDim.x + blockIdx.x) * LS_SUMS + threadIdx.x] = s;
  }
}

__device__ __forceinline__ float interval_term(float x, float lo, float hi, float wgt, float beta, float gamma, float *grad) {
  // models.py:495-506 on one sample; *grad = d term / d x
  const float dx = beta * powf(x, gamma), dlo = beta * powf(lo, gamma), dhi = beta * powf(hi, gamma);
  const float k = 0.5f * (dhi - dlo), mid = (dhi + dlo) / 2.0f;
  const float u = (dx - mid) * (dx - mid) - k * k;
  if (u > 0.f) {
    *grad = 10.0f * wgt * 2.0f * (dx - mid) * beta * gamma * powf(x, gamma - 1.0f);
    return 10.0f * u * wgt;
  }
  *grad = 0.f;
  return 0.f;
}

// One CTA: fixed-order reduction of the per-CTA partials, then the scalar part of the loss and the coefficients
// the backward pass needs.  coef layout: see DRAM_LOSS_COEF_* in dram_b200.h.
__global__ void __launch_bounds__(LS_THREADS) loss_finalize_kernel(const double *__restrict__ partial, int ctas_per_sample,
                                                                  int nb, double voxels_per_sample,
                                                                  const float *__restrict__ cle_bands,
                                                                  const float *__restrict__ pse_bands,
                                                                  const float *__restrict__ cle_weights,
                                                                  const float *__restrict__ pse_weights, float beta,
                                                                  float gamma, float *__restrict__ coef) {
  __shared__ double sums[LS_MAX_B][LS_SUMS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = warp; u < nb * 7; u += LS_WARPS) {
    const int b = u / 7, j = u - b * 7;
    double s = 0.0;
    for (int c = lane; c < ctas_per_sample; c += 32) s += partial[((size_t)b * ctas_per_sample + c) * LS_SUMS + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) sums[b][j] = s;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double inter = 0.0, ysum = 0.0, yhsum = 0.0, tsum = 0.0, a1 = 0.0, a0 = 0.0;
  float loss_cle = 0.f, loss_pse = 0.f;
  for (int b = 0; b < nb; ++b) {
    const float m = (float)sums[b][2];
    const float r0 = (float)sums[b][0] / m, r1 = (float)sums[b][1] / m;  // med3d.py:387
    float g0, g1;
    loss_cle += interval_term(r0, cle_bands[2 * b], cle_bands[2 * b + 1], cle_weights[b], beta, gamma, &g0);
    loss_pse += interval_term(r1, pse_bands[2 * b], pse_bands[2 * b + 1], pse_weights[b], beta, gamma, &g1);
    float *row = coef + LS_HEAD + 4 * b;
    row[0] = r0, row[1] = r1, row[2] = g0 / m, row[3] = g1 / m;
    ysum += sums[b][0], yhsum += sums[b][1], inter += sums[b][3];
    tsum += sums[b][4], a1 += sums[b][5], a0 += sums[b][6];
  }
  // BinaryDice: (2*I + s) / (sum y + sum y_hat + s), metrics.py:33-37
  const float num = 2.0f * (float)inter + DICE_SMOOTH, den = (float)ysum + (float)yhsum + DICE_SMOOTH;
  const float mul_loss = num / den;
  // BinaryCrossEntropy with a mask, metrics.py:10-30; alpha divides by the batch size (y.shape[0])
  float alpha = 1.0f - (float)tsum / (float)nb;
  alpha = fminf(fmaxf(alpha, 0.3f), 0.7f);
  const double total = voxels_per_sample * nb;
  const float wsum = (float)((double)alpha * tsum + (double)(1.0f - alpha) * (total - tsum));
  const float seg_loss = -(float)((double)alpha * a1 + (double)(1.0f - alpha) * a0) / wsum;
  coef[DRAM_LOSS_COEF_TOTAL] = loss_cle + loss_pse + 2.0f * mul_loss + seg_loss;  // models.py:565
  coef[DRAM_LOSS_COEF_CLE] = loss_cle;
  coef[DRAM_LOSS_COEF_PSE] = loss_pse;
  coef[DRAM_LOSS_COEF_MUL] = mul_loss;
  coef[DRAM_LOSS_COEF_SEG] = seg_loss;
  coef[5] = 2.0f * 2.0f / den;         // d(2*mul)/d y      = c1 * y_hat - c2
  coef[6] = 2.0f * num / (den * den);  //                     (and symmetrically for y_hat)
  coef[7] = alpha;
  coef[8] = 1.0f / wsum;
  for (int i = 9; i < LS_HEAD; ++i) coef[i] = 0.f;
}

__global__ void __launch_bounds__(LS_THREADS) loss_backward_kernel(const float *__restrict__ d0, const float *__restrict__ d1,
                                                                  const uint8_t *__restrict__ lungs,
                                                                  const uint8_t *__restrict__ ems,
                                                                  const int64_t *__restrict__ cle_labels,
                                                                  const int64_t *__restrict__ pse_labels,
                                                                  const float *__restrict__ coef,
                                                                  const float *__restrict__ grad_loss, LossGeom g,
                                                                  float *__restrict__ g0, float *__restrict__ g1) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool flag = cle_labels[b] > 0 || pse_labels[b] > 0;
  const float up = grad_loss ? *grad_loss : 1.0f;
  const float c1 = coef[5], c2 = coef[6], alpha = coef[7], inv_w = coef[8];
  const float ga0 = coef[LS_HEAD + 4 * b + 2], ga1 = coef[LS_HEAD + 4 * b + 3];
  const int rows = g.d2 * g.h2;
  const size_t base = (size_t)b * rows * g.w2;
  for (int r = blockIdx.x * LS_WARPS + warp; r < rows; r += gridDim.x * LS_WARPS) {
    const int zd = r / g.h2, zh = r - zd * g.h2;
    const size_t src_row = (((size_t)b * g.d + nearest_src(zd, g.d, g.d2)) * g.h + nearest_src(zh, g.h, g.h2)) * g.w;
    const uint8_t *lrow = lungs + src_row, *erow = ems + src_row;
    for (int x = lane; x < g.w2; x += 32) {
      const int sx = nearest_src(x, g.w, g.w2);
      const float m = lrow[sx] ? 1.f : 0.f;
      const bool t = flag && erow[sx];
      const size_t i = base + (size_t)r * g.w2 + x;
      const float p0 = d0[i], p1 = d1[i];
      // masked mean -> interval regression, and the dice term: both only inside the lungs
      float q0 = m * (ga0 + c1 * (p1 * m) - c2);
      float q1 = m * (ga1 + c1 * (p0 * m) - c2);
      // BCE through clamp(p0 + p1, 0, 1) and clamp(pt, eps, 1 - eps): ATen's clamp passes the gradient on [min, max]
      const float s = p0 + p1;
      if (s >= 0.f && s <= 1.f) {
        const float pt = t ? s : 1.f - s;
        if (pt >= BCE_EPS && pt <= 1.f - BCE_EPS) {
          const float wv = t ? alpha : 1.f - alpha;
          const float f = m != 0.f ? BCE_SMOOTH : 1.f;
          const float dl = -(wv * f * inv_w) / pt * (t ? 1.f : -1.f);
          q0 += dl;
          q1 += dl;
        }
      }
      g0[i] = up * q0;
      g1[i] = up * q1;
    }
  }
}

static int loss_ctas_per_sample(int b, int rows) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int per = (4 * sms + b - 1) / b;
  const int need = (rows + LS_WARPS - 1) / LS_WARPS;
  if (per > need) per = need;
  return per < 1 ? 1 : per;
}

// K12: torch.optim.Adam (single tensor form, amsgrad off, weight decay 0) over flat buffers.
__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ gr, float *__restrict__ m,
                                                   float *__restrict__ v, long long n, float w1, float beta2, float w2,
                                                   float step_size, float bc2_sqrt, float eps, float grad_scale) {
  const long long n4 = n >> 2, stride = (long long)gridDim.x * 256;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4 *>(p)[i], mv = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
    const float4 gv = reinterpret_cast<const float4 *>(gr)[i];
    float *pp = &pv.x, *mp = &mv.x, *vp = &vv.x;
    const float *gp = &gv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * grad_scale;
      mp[j] = mp[j] + w1 * (gj - mp[j]);         // exp_avg.lerp_(grad, 1 - beta1)
      vp[j] = vp[j] * beta2 + w2 * gj * gj;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
      const float denom = sqrtf(vp[j]) / bc2_sqrt + eps;
      pp[j] = pp[j] - step_size * (mp[j] / denom);  // param.addcdiv_(exp_avg, denom, value = -lr / bias_correction1)
    }
    reinterpret_cast<float4 *>(p)[i] = pv;
    reinterpret_cast<float4 *>(m)[i] = mv;
    reinterpret_cast<float4 *>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + blockIdx.x * 256LL + threadIdx.x; i < n; i += stride) {
    const float gj = gr[i] * grad_scale;
    const float mj = m[i] + w1 * (gj - m[i]);
    const float vj = v[i] * beta2 + w2 * gj * gj;
    m[i] = mj;
    v[i] = vj;
    p[i] = p[i] - step_size * (mj / (sqrtf(vj) / bc2_sqrt + eps));
  }
}

}  // namespace dram

using namespace dram;

extern "C" int64_t dram_train_loss_workspace_bytes(int32_t b) {
  if (b < 1) b = 1;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  // loss_ctas_per_sample() never exceeds ceil(4*sms / b)
  return (int64_t)b * ((4 * sms + b - 1) / b) * LS_SUMS * (int64_t)sizeof(double);
}

static int loss_check_geom(const char *who, int32_t b, int32_t d, int32_t h, int32_t w, int32_t d2, int32_t h2, int32_t w2) {
  DRAM_REQUIRE(b >= 1 && b <= LS_MAX_B, "%s: batch %d outside 1..%d", who, b, LS_MAX_B);
  DRAM_REQUIRE(d > 0 && h > 0 && w > 0 && d2 > 0 && h2 > 0 && w2 > 0, "%s: empty volume", who);
  DRAM_REQUIRE((int64_t)d2 * h2 < (1LL << 31), "%s: map too large", who);
  return DRAM_OK;
}

extern "C" int dram_train_loss_forward(const float *cle_map, const float *pse_map, const uint8_t *lungs, const uint8_t *ems,
                                       const int64_t *cle_labels, const int64_t *pse_labels, const float *cle_bands,
                                       const float *pse_bands, const float *cle_weights, const float *pse_weights,
                                       int32_t b, int32_t d, int32_t h, int32_t w, int32_t d2, int32_t h2, int32_t w2,
                                       float beta, float gamma, float *coef, void *workspace, void *stream) {
  DRAM_REQUIRE(cle_map && pse_map && lungs && ems && cle_labels && pse_labels && cle_bands && pse_bands && cle_weights &&
                   pse_weights && coef && workspace,
               "dram_train_loss_forward: null pointer");
  int rc = loss_check_geom("dram_train_loss_forward", b, d, h, w, d2, h2, w2);
  if (rc != DRAM_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int per = loss_ctas_per_sample(b, d2 * h2);
  const LossGeom g{b, d, h, w, d2, h2, w2};
  loss_sums_kernel<<<dim3(per, b), LS_THREADS, 0, st>>>(cle_map, pse_map, lungs, ems, cle_labels, pse_labels, g,
                                                        reinterpret_cast<double *>(workspace));
  DRAM_CHECK_LAUNCH("loss_sums_kernel launch");
  loss_finalize_kernel<<<1, LS_THREADS, 0, st>>>(reinterpret_cast<const double *>(workspace), per, b,
                                                 (double)d2 * h2 * w2, cle_bands, pse_bands, cle_weights, pse_weights, beta,
                                                 gamma, coef);
  DRAM_CHECK_LAUNCH("loss_finalize_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_train_loss_backward(const float *cle_map, const float *pse_map, const uint8_t *lungs, const uint8_t *ems,
                                        const int64_t *cle_labels, const int64_t *pse_labels, const float *coef,
                                        const float *grad_loss, int32_t b, int32_t d, int32_t h, int32_t w, int32_t d2,
                                        int32_t h2, int32_t w2, float *grad_cle, float *grad_pse, void *stream) {
  DRAM_REQUIRE(cle_map && pse_map && lungs && ems && cle_labels && pse_labels && coef && grad_cle && grad_pse,
               "dram_train_loss_backward: null pointer");
  int rc = loss_check_geom("dram_train_loss_backward", b, d, h, w, d2, h2, w2);
  if (rc != DRAM_OK) return rc;
  const int per = loss_ctas_per_sample(b, d2 * h2);
  const LossGeom g{b, d, h, w, d2, h2, w2};
  loss_backward_kernel<<<dim3(per, b), LS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      cle_map, pse_map, lungs, ems, cle_labels, pse_labels, coef, grad_loss, g, grad_cle, grad_pse);
  DRAM_CHECK_LAUNCH("loss_backward_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr,
                              double beta1, double beta2, double eps, int32_t step, float grad_scale, void *stream) {
  DRAM_REQUIRE(param && grad && exp_avg && exp_avg_sq, "dram_adam_step: null pointer");
  DRAM_REQUIRE(n > 0, "dram_adam_step: empty parameter buffer");
  DRAM_REQUIRE(step >= 1, "dram_adam_step: step counts from 1 (got %d)", step);
  DRAM_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "dram_adam_step: bad hyper-parameters");
  DRAM_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
               "dram_adam_step: buffers must be 16-byte aligned");
  // bias corrections in double, as torch's Python front end does (_single_tensor_adam)
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2);
  adam_kernel<<<stream_grid((n + 3) / 4, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, w1, (float)beta2, w2, step_size, bc2_sqrt, (float)eps, grad_scale);
  DRAM_CHECK_LAUNCH("adam_kernel launch");
  return DRAM_OK;
}
