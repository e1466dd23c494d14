// K10 — train-mode BatchNorm3d (+ ReLU, + residual add) on NDHWC 16-bit activations, forward and backward
// (training path).  Replaces, for the training step, every `bn(conv(x))` / `relu(...)` /
// `out += residual` of med3d.py:129-144, 164-184, 74-80, 371-373 as autograd sees them in `model.train()`:
//
//   forward   mean_c, var_c (biased) over the M = N*D*H*W rows;  y = relu((x - mean) * rstd * gamma + beta (+ res))
//             running_mean/var updated with momentum (unbiased variance), as nn.BatchNorm3d
//   backward  dz = dy * (y > 0);  dbeta = sum dz;  dgamma = sum dz * xhat;
//             dx = gamma * rstd * (dz - dbeta / M - xhat * dgamma / M);  dres = dz
//
// Memory-bound: every kernel streams [M][C] tensors with 16-byte accesses, a thread owning 8 channels of a row.
// The two reductions are two-phase and deterministic: per-CTA fp32 partials over a contiguous row range, written as
// fp64, then a fixed-order fp64 sum over the CTAs.  The per-channel sums are exposed between the phases so that the
// caller can all-reduce them over the process group (SyncBatchNorm, train.py:101).
#include "umma_common.cuh"

namespace dram {

static constexpr int BN_THREADS = 256;
static constexpr int BN_MAX_C = 2048;          // 256 threads x 8 channels

struct BnGeom {
  int tpr;    // threads per row = C / 8
  int rpi;    // rows per iteration = 256 / tpr
};
__host__ __device__ inline BnGeom bn_geom(int c) {
  BnGeom g;
  g.tpr = c / 8;
  g.rpi = BN_THREADS / g.tpr;
  return g;
}

__device__ __forceinline__ void unpack8(const uint4 &v, int is_f16, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 p = unpack2(w[q], is_f16);
    f[2 * q] = p.x;
    f[2 * q + 1] = p.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int is_f16) {
  return make_uint4(pack2(f[0], f[1], is_f16), pack2(f[2], f[3], is_f16), pack2(f[4], f[5], is_f16),
                    pack2(f[6], f[7], is_f16));
}

// Block reduction of per-thread (a[8], b[8]) over the threads that share a channel group; result to
// partial[cta][0][c], partial[cta][1][c] (fp64).
__device__ __forceinline__ void bn_block_reduce(const float (&a)[8], const float (&b)[8], int c, double *partial) {
  __shared__ float red[BN_THREADS][17];
  const BnGeom g = bn_geom(c);
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[t][j] = a[j];
    red[t][8 + j] = b[j];
  }
  __syncthreads();
  // thread u < 2 * C handles one (which, channel) pair: channel ch lives in threads (ch / 8) + k * tpr, slot ch % 8
  for (int u = t; u < 2 * c; u += BN_THREADS) {
    const int which = u / c, ch = u - which * c;
    const int cg = ch >> 3, j = ch & 7;
    double s = 0.0;
    for (int k = 0; k < g.rpi; ++k) s += (double)red[cg + k * g.tpr][which * 8 + j];
    partial[((size_t)blockIdx.x * 2 + which) * c + ch] = s;
  }
}

// phase 1 of the forward statistics: sum x, sum x^2 per channel over this CTA's rows
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const uint4 *__restrict__ x, long long m, int c, int is_f16,
                                                             double *__restrict__ partial) {
  const BnGeom g = bn_geom(c);
  const int cg = threadIdx.x % g.tpr, rl = threadIdx.x / g.tpr;
  const long long r0 = (m * blockIdx.x) / gridDim.x, r1 = (m * (blockIdx.x + 1)) / gridDim.x;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.0f;
  if (rl < g.rpi)
    for (long long r = r0 + rl; r < r1; r += g.rpi) {
      float f[8];
      unpack8(__ldg(x + r * g.tpr + cg), is_f16, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
  bn_block_reduce(s, q, c, partial);
}

// phase 2: sums[which][ch] = sum over CTAs in fp64 — one warp per output, lanes stride over the CTAs and a fixed
// shuffle tree combines them, so the order (and the result) does not depend on anything but the CTA count
__global__ void bn_partials_reduce_kernel(const double *__restrict__ partial, int ctas, int c, double *__restrict__ sums) {
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (u >= 2 * c) return;
  double s = 0.0;
  for (int b = lane; b < ctas; b += 32) s += partial[(size_t)b * 2 * c + u];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) sums[u] = s;
}

__global__ void bn_finalize_kernel(const double *__restrict__ sums, double count, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, float momentum, float *running_mean,
                                   float *running_var, float *__restrict__ scale, float *__restrict__ shift,
                                   float *__restrict__ mean, float *__restrict__ rstd, int c) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double mu = sums[ch] / count;
  double var = sums[c + ch] / count - mu * mu;
  if (var < 0.0) var = 0.0;
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  mean[ch] = (float)mu;
  rstd[ch] = rs;
  const float sc = gamma[ch] * rs;
  scale[ch] = sc;
  shift[ch] = beta[ch] - (float)mu * sc;
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[ch] = (1.0f - momentum) * running_mean[ch] + momentum * (float)mu;
    running_var[ch] = (1.0f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

// y = act(x * scale + shift (+ res)).  The grid stride is a multiple of the threads per row (256 % (c/8) == 0), so a
// thread keeps the same 8 channels for its whole loop and holds their coefficients in registers.
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const uint4 *__restrict__ x, const float *__restrict__ scale,
                                                             const float *__restrict__ shift, const uint4 *__restrict__ res,
                                                             int relu, uint4 *__restrict__ out, long long m, int c,
                                                             int is_f16) {
  const int tpr = c / 8;
  const long long total = m * tpr;
  const int cg = (int)(threadIdx.x % tpr);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + cg * 8 + j);
    sh[j] = __ldg(shift + cg * 8 + j);
  }
  for (long long i = blockIdx.x * (long long)BN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BN_THREADS) {
    float f[8];
    unpack8(__ldg(x + i), is_f16, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
    if (res != nullptr) {
      float r[8];
      unpack8(__ldg(res + i), is_f16, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.0f);
    }
    out[i] = pack8(f, is_f16);
  }
}

// phase 1 of the backward reduction: sum dz, sum dz * xhat, dz = dy masked by y > 0 (y = the forward output, or null)
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x,
                                                                  const uint4 *__restrict__ y, const float *__restrict__ mean,
                                                                  const float *__restrict__ rstd, long long m, int c,
                                                                  int is_f16, double *__restrict__ partial) {
  const BnGeom g = bn_geom(c);
  const int cg = threadIdx.x % g.tpr, rl = threadIdx.x / g.tpr;
  const long long r0 = (m * blockIdx.x) / gridDim.x, r1 = (m * (blockIdx.x + 1)) / gridDim.x;
  float a[8], b[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = b[j] = 0.0f;
    mu[j] = __ldg(mean + cg * 8 + j);
    rs[j] = __ldg(rstd + cg * 8 + j);
  }
  if (rl < g.rpi)
    for (long long r = r0 + rl; r < r1; r += g.rpi) {
      const long long i = r * g.tpr + cg;
      float d[8], xv[8];
      unpack8(__ldg(dy + i), is_f16, d);
      unpack8(__ldg(x + i), is_f16, xv);
      if (y != nullptr) {
        float yv[8];
        unpack8(__ldg(y + i), is_f16, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = yv[j] > 0.0f ? d[j] : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] += d[j];
        b[j] = fmaf(d[j], (xv[j] - mu[j]) * rs[j], b[j]);
      }
    }
  bn_block_reduce(a, b, c, partial);
}

// dx = gamma * rstd * (dz - sum_dz / M - xhat * sum_dz_xhat / M) = A * dz + B * x + C per channel;  dres = dz (optional).
// As in bn_apply_kernel a thread keeps its 8 channels, so A, B, C are computed once per thread.
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x,
                                                                 const uint4 *__restrict__ y, const float *__restrict__ mean,
                                                                 const float *__restrict__ rstd, const float *__restrict__ gamma,
                                                                 const double *__restrict__ sums, double count,
                                                                 uint4 *__restrict__ dx, uint4 *__restrict__ dres, long long m,
                                                                 int c, int is_f16) {
  const int tpr = c / 8;
  const long long total = m * tpr;
  const int cg = (int)(threadIdx.x % tpr);
  float A[8], B[8], Cc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cg * 8 + j;
    const double rs = (double)__ldg(rstd + ch), mu = (double)__ldg(mean + ch), g = (double)__ldg(gamma + ch);
    const double s1 = sums[ch] / count, s2 = sums[c + ch] / count;
    A[j] = (float)(g * rs);
    B[j] = (float)(-g * rs * rs * s2);
    Cc[j] = (float)(-g * rs * s1 + g * rs * rs * s2 * mu);
  }
  for (long long i = blockIdx.x * (long long)BN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * BN_THREADS) {
    float d[8], xv[8], o[8];
    unpack8(__ldg(dy + i), is_f16, d);
    unpack8(__ldg(x + i), is_f16, xv);
    if (y != nullptr) {
      float yv[8];
      unpack8(__ldg(y + i), is_f16, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = yv[j] > 0.0f ? d[j] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], d[j], fmaf(B[j], xv[j], Cc[j]));
    dx[i] = pack8(o, is_f16);
    if (dres != nullptr) dres[i] = pack8(d, is_f16);
  }
}

static int bn_reduce_ctas(long long m, int c) {
  const BnGeom g = bn_geom(c);
  long long want = (m + (long long)g.rpi * 8 - 1) / ((long long)g.rpi * 8);  // at least ~8 rows per thread
  long long cap = (long long)(sm_count() > 0 ? sm_count() : 148) * 4;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}
static bool bn_shape_ok(long long m, int c) { return m > 0 && c >= 8 && c <= BN_MAX_C && c % 8 == 0 && BN_THREADS % (c / 8) == 0; }

}  // namespace dram

using namespace dram;

extern "C" int64_t dram_bn_workspace_bytes(int32_t c) {
  if (c <= 0) return -1;
  return (int64_t)((sm_count() > 0 ? sm_count() : 148) * 8) * 2 * c * (int64_t)sizeof(double);
}

extern "C" int dram_bn_stats(const void *x, int64_t m, int32_t c, int32_t dtype, double *sums, void *workspace, void *stream) {
  DRAM_REQUIRE(x && sums && workspace, "dram_bn_stats: null pointer");
  DRAM_REQUIRE(bn_shape_ok(m, c), "dram_bn_stats: channels must be 8, 16, ..., 2048 with 256 %% (c/8) == 0 (got m %lld c %d)",
               (long long)m, c);
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_bn_stats: bad dtype");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ctas = bn_reduce_ctas(m, c);
  bn_stats_kernel<<<ctas, BN_THREADS, 0, st>>>(reinterpret_cast<const uint4 *>(x), m, c, dtype == DRAM_DTYPE_F16,
                                               reinterpret_cast<double *>(workspace));
  DRAM_CHECK_LAUNCH("bn_stats_kernel launch");
  bn_partials_reduce_kernel<<<ceil_div(2 * c * 32, 256), 256, 0, st>>>(reinterpret_cast<const double *>(workspace), ctas, c, sums);
  DRAM_CHECK_LAUNCH("bn_partials_reduce_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_bn_finalize(const double *sums, double count, const float *gamma, const float *beta, float eps,
                                float momentum, float *running_mean, float *running_var, float *scale, float *shift,
                                float *mean, float *rstd, int32_t c, void *stream) {
  DRAM_REQUIRE(sums && gamma && beta && scale && shift && mean && rstd, "dram_bn_finalize: null pointer");
  DRAM_REQUIRE(c > 0 && count >= 1.0, "dram_bn_finalize: bad size");
  DRAM_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "dram_bn_finalize: running_mean and running_var go together");
  bn_finalize_kernel<<<ceil_div(c, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      sums, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, rstd, c);
  DRAM_CHECK_LAUNCH("bn_finalize_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_bn_apply(const void *x, const float *scale, const float *shift, const void *res, int32_t relu, void *out,
                             int64_t m, int32_t c, int32_t dtype, void *stream) {
  DRAM_REQUIRE(x && scale && shift && out, "dram_bn_apply: null pointer");
  DRAM_REQUIRE(bn_shape_ok(m, c), "dram_bn_apply: unsupported shape (m %lld c %d)", (long long)m, c);
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_bn_apply: bad dtype");
  bn_apply_kernel<<<stream_grid(m * (c / 8), BN_THREADS), BN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4 *>(x), scale, shift, reinterpret_cast<const uint4 *>(res), relu,
      reinterpret_cast<uint4 *>(out), m, c, dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("bn_apply_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_bn_backward_reduce(const void *dy, const void *x, const void *y, const float *mean, const float *rstd,
                                       int64_t m, int32_t c, int32_t dtype, double *sums, void *workspace, void *stream) {
  DRAM_REQUIRE(dy && x && mean && rstd && sums && workspace, "dram_bn_backward_reduce: null pointer");
  DRAM_REQUIRE(bn_shape_ok(m, c), "dram_bn_backward_reduce: unsupported shape (m %lld c %d)", (long long)m, c);
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_bn_backward_reduce: bad dtype");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ctas = bn_reduce_ctas(m, c);
  bn_bwd_reduce_kernel<<<ctas, BN_THREADS, 0, st>>>(reinterpret_cast<const uint4 *>(dy), reinterpret_cast<const uint4 *>(x),
                                                    reinterpret_cast<const uint4 *>(y), mean, rstd, m, c,
                                                    dtype == DRAM_DTYPE_F16, reinterpret_cast<double *>(workspace));
  DRAM_CHECK_LAUNCH("bn_bwd_reduce_kernel launch");
  bn_partials_reduce_kernel<<<ceil_div(2 * c * 32, 256), 256, 0, st>>>(reinterpret_cast<const double *>(workspace), ctas, c, sums);
  DRAM_CHECK_LAUNCH("bn_partials_reduce_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_bn_backward_apply(const void *dy, const void *x, const void *y, const float *mean, const float *rstd,
                                      const float *gamma, const double *sums, double count, void *dx, void *dres, int64_t m,
                                      int32_t c, int32_t dtype, void *stream) {
  DRAM_REQUIRE(dy && x && mean && rstd && gamma && sums && dx, "dram_bn_backward_apply: null pointer");
  DRAM_REQUIRE(bn_shape_ok(m, c) && count >= 1.0, "dram_bn_backward_apply: unsupported shape (m %lld c %d)", (long long)m, c);
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_bn_backward_apply: bad dtype");
  bn_bwd_apply_kernel<<<stream_grid(m * (c / 8), BN_THREADS), BN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4 *>(dy), reinterpret_cast<const uint4 *>(x), reinterpret_cast<const uint4 *>(y), mean, rstd,
      gamma, sums, count, reinterpret_cast<uint4 *>(dx), reinterpret_cast<uint4 *>(dres), m, c, dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("bn_bwd_apply_kernel launch");
  return DRAM_OK;
}
