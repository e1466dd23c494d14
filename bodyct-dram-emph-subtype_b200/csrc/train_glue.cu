// Backward of the memory-bound forward kernels (training path).
//
// K4T — adjoint of nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True) (med3d.py:83, 86) on NDHWC
// 16-bit tensors: dx[s] = sum over the outputs o that read source voxel s of w(o, s) * dy[o], with ATen's index rule
// (lin_index_ac: the same fp32 arithmetic as K4, so forward and backward use identical weights).  A thread owns one
// (source voxel, 8-channel group) and GATHERS its <= 4 x 4 x 4 contributing outputs — no atomics, deterministic;
// neighbouring threads share the dy rows through L1/L2.  fp32 accumulation, one rounding to the storage type.
#include "umma_common.cuh"

namespace dram {

static constexpr int UPB_MAX = 6;  // contributing outputs per axis (scale 2: at most 4; margin for the clamped ends)

struct AxisTaps {
  int n;
  int o[UPB_MAX];
  float w[UPB_MAX];
};
// Outputs along one axis whose interpolation reads source index s, with their weights.
__device__ __forceinline__ AxisTaps axis_taps(int s, float scale, int in_size, int out_size) {
  AxisTaps t;
  t.n = 0;
  int lo = 2 * s - 3, hi = 2 * s + 3;  // scale is (in-1)/(2*in-1) ~ 1/2: o/2 - 1 < s < o/2 + 1
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
  for (int o = lo; o <= hi; ++o) {
    const LinIdx id = lin_index_ac(o, scale, in_size);
    float w = 0.0f;
    if (id.i0 == s) w += id.w0;
    if (id.i1 == s) w += id.w1;
    if (w != 0.0f && t.n < UPB_MAX) {
      t.o[t.n] = o;
      t.w[t.n] = w;
      ++t.n;
    }
  }
  return t;
}

__global__ void __launch_bounds__(256) upsample2x_backward_kernel(const uint4 *__restrict__ dy, uint4 *__restrict__ dx, int n,
                                                                 int d, int h, int w, int c, int is_f16) {
  const int cg_n = c / 8;
  const int D = 2 * d, H = 2 * h, W = 2 * w;
  const float sd = ac_scale(d, D), sh = ac_scale(h, H), sw = ac_scale(w, W);
  const long long total = (long long)n * d * h * w * cg_n;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = (int)(i % cg_n);
    long long v = i / cg_n;
    const int xw = (int)(v % w);
    v /= w;
    const int xh = (int)(v % h);
    v /= h;
    const int xd = (int)(v % d);
    const int sample = (int)(v / d);
    const AxisTaps td = axis_taps(xd, sd, d, D), th = axis_taps(xh, sh, h, H), tw = axis_taps(xw, sw, w, W);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    for (int a = 0; a < td.n; ++a)
      for (int b = 0; b < th.n; ++b) {
        const float wab = td.w[a] * th.w[b];
        const uint4 *row = dy + ((((long long)sample * D + td.o[a]) * H + th.o[b]) * W) * cg_n + cg;
        for (int e = 0; e < tw.n; ++e) {
          const float wt = wab * tw.w[e];
          const uint4 q = __ldg(row + (long long)tw.o[e] * cg_n);
          const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 p = unpack2(u[k], is_f16);
            acc[2 * k] = fmaf(wt, p.x, acc[2 * k]);
            acc[2 * k + 1] = fmaf(wt, p.y, acc[2 * k + 1]);
          }
        }
      }
    dx[i] = make_uint4(pack2(acc[0], acc[1], is_f16), pack2(acc[2], acc[3], is_f16), pack2(acc[4], acc[5], is_f16),
                       pack2(acc[6], acc[7], is_f16));
  }
}

// K3T — backward of MaxPool3d(kernel 3, stride 2, padding 1) (med3d.py:305, 374) on NDHWC 16-bit tensors.
// dx[i] = sum of dy[o] over the windows o that contain i and whose maximum sits at i.  ATen's max_pool3d_with_indices
// takes the FIRST maximum of a window in (d, h, w) scan order (strict >), and after a ReLU ties are the rule, so the
// same choice is made here, in two passes without atomics:
//   (1) arg-max pass: a thread owns one (window, 8-channel group), scans its <= 27 in-bounds inputs in order and
//       records the local position (kd*9 + kh*3 + kw) of the first maximum, one byte per channel;
//   (2) gather pass: a thread owns one (input voxel, 8-channel group) and adds dy of the <= 2 x 2 x 2 windows that
//       contain it where the recorded position is its own.
__global__ void __launch_bounds__(256) maxpool3d_argmax_kernel(const uint4 *__restrict__ x, uint2 *__restrict__ idx, int n, int d,
                                                              int h, int w, int c, int is_f16) {
  const int cg_n = c / 8;
  const int Do = (d - 1) / 2 + 1, Ho = (h - 1) / 2 + 1, Wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * Do * Ho * Wo * cg_n;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = (int)(i % cg_n);
    long long v = i / cg_n;
    const int ow = (int)(v % Wo);
    v /= Wo;
    const int oh = (int)(v % Ho);
    v /= Ho;
    const int od = (int)(v % Do);
    const int sample = (int)(v / Do);
    float best[8];
    uint32_t pos[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      pos[j] = 255u;
    }
    for (int kd = 0; kd < 3; ++kd) {
      const int zd = 2 * od - 1 + kd;
      if (zd < 0 || zd >= d) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int zh = 2 * oh - 1 + kh;
        if (zh < 0 || zh >= h) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int zw = 2 * ow - 1 + kw;
          if (zw < 0 || zw >= w) continue;
          const uint32_t local = (uint32_t)(kd * 9 + kh * 3 + kw);
          const uint4 q = __ldg(x + ((((long long)sample * d + zd) * h + zh) * w + zw) * cg_n + cg);
          const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 p = unpack2(u[k], is_f16);
            if (p.x > best[2 * k] || pos[2 * k] == 255u) {
              best[2 * k] = p.x;
              pos[2 * k] = local;
            }
            if (p.y > best[2 * k + 1] || pos[2 * k + 1] == 255u) {
              best[2 * k + 1] = p.y;
              pos[2 * k + 1] = local;
            }
          }
        }
      }
    }
    idx[i] = make_uint2(pos[0] | (pos[1] << 8) | (pos[2] << 16) | (pos[3] << 24),
                        pos[4] | (pos[5] << 8) | (pos[6] << 16) | (pos[7] << 24));
  }
}

__global__ void __launch_bounds__(256) maxpool3d_backward_kernel(const uint2 *__restrict__ idx, const uint4 *__restrict__ dy,
                                                                uint4 *__restrict__ dx, int n, int d, int h, int w, int c,
                                                                int is_f16) {
  const int cg_n = c / 8;
  const int Do = (d - 1) / 2 + 1, Ho = (h - 1) / 2 + 1, Wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * d * h * w * cg_n;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = (int)(i % cg_n);
    long long v = i / cg_n;
    const int xw = (int)(v % w);
    v /= w;
    const int xh = (int)(v % h);
    v /= h;
    const int xd = (int)(v % d);
    const int sample = (int)(v / d);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    // windows o with 2*o - 1 <= i <= 2*o + 1: o = i/2 and, for odd i, (i+1)/2
    for (int od = xd / 2; od <= (xd + 1) / 2; ++od) {
      if (od >= Do) continue;
      for (int oh = xh / 2; oh <= (xh + 1) / 2; ++oh) {
        if (oh >= Ho) continue;
        for (int ow = xw / 2; ow <= (xw + 1) / 2; ++ow) {
          if (ow >= Wo) continue;
          const uint32_t local = (uint32_t)((xd - (2 * od - 1)) * 9 + (xh - (2 * oh - 1)) * 3 + (xw - (2 * ow - 1)));
          const long long o = ((((long long)sample * Do + od) * Ho + oh) * Wo + ow) * cg_n + cg;
          const uint2 p = __ldg(idx + o);
          const uint4 g = __ldg(dy + o);
          const uint32_t u[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = unpack2(u[k], is_f16);
            const uint32_t word = k < 2 ? p.x : p.y;
            const uint32_t p0 = (word >> (16 * (k & 1))) & 255u, p1 = (word >> (16 * (k & 1) + 8)) & 255u;
            if (p0 == local) acc[2 * k] += f.x;
            if (p1 == local) acc[2 * k + 1] += f.y;
          }
        }
      }
    }
    dx[i] = make_uint4(pack2(acc[0], acc[1], is_f16), pack2(acc[2], acc[3], is_f16), pack2(acc[4], acc[5], is_f16),
                       pack2(acc[6], acc[7], is_f16));
  }
}

}  // namespace dram

using namespace dram;

extern "C" int dram_upsample2x_backward(const void *dy, void *dx, int32_t n, int32_t d, int32_t h, int32_t w, int32_t c,
                                        int32_t dtype, void *stream) {
  DRAM_REQUIRE(dy && dx, "dram_upsample2x_backward: null pointer");
  DRAM_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "dram_upsample2x_backward: bad shape (c %% 8 != 0?)");
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_upsample2x_backward: bad dtype");
  const long long total = (long long)n * d * h * w * (c / 8);
  upsample2x_backward_kernel<<<stream_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4 *>(dy), reinterpret_cast<uint4 *>(dx), n, d, h, w, c, dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("upsample2x_backward_kernel launch");
  return DRAM_OK;
}

extern "C" int64_t dram_maxpool3d_backward_workspace_bytes(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c) {
  if (n <= 0 || d <= 0 || h <= 0 || w <= 0 || c <= 0) return -1;
  return (int64_t)n * ((d - 1) / 2 + 1) * ((h - 1) / 2 + 1) * ((w - 1) / 2 + 1) * c;
}

extern "C" int dram_maxpool3d_backward(const void *x, const void *dy, void *dx, void *workspace, int32_t n, int32_t d,
                                       int32_t h, int32_t w, int32_t c, int32_t dtype, void *stream) {
  DRAM_REQUIRE(x && dy && dx && workspace, "dram_maxpool3d_backward: null pointer");
  DRAM_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "dram_maxpool3d_backward: bad shape (c %% 8 != 0?)");
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_maxpool3d_backward: bad dtype");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long outs = (long long)n * ((d - 1) / 2 + 1) * ((h - 1) / 2 + 1) * ((w - 1) / 2 + 1) * (c / 8);
  maxpool3d_argmax_kernel<<<stream_grid(outs, 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(x),
                                                                  reinterpret_cast<uint2 *>(workspace), n, d, h, w, c,
                                                                  dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("maxpool3d_argmax_kernel launch");
  const long long total = (long long)n * d * h * w * (c / 8);
  maxpool3d_backward_kernel<<<stream_grid(total, 256), 256, 0, st>>>(reinterpret_cast<const uint2 *>(workspace),
                                                                     reinterpret_cast<const uint4 *>(dy),
                                                                     reinterpret_cast<uint4 *>(dx), n, d, h, w, c,
                                                                     dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("maxpool3d_backward_kernel launch");
  return DRAM_OK;
}

// K5T — the two 1x1x1 regression heads with their sigmoid (med3d.py:329-332, 382: `fcs`, torch.sigmoid) in training:
// forward from the 32-channel NDHWC 16-bit feature map to two fp32 maps, and the backward of both
// (dx 16-bit, dW [2][32], db [2] as a deterministic two-phase reduction).  In inference the same heads are fused into
// the us3 epilogue (K5); in training us3 is followed by a train-mode BatchNorm, so they run on its output.
namespace dram {

static constexpr int HD_C = 32, HD_THREADS = 256;

__device__ __forceinline__ void load_row32(const uint4 *x, long long m, int is_f16, float (&f)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = __ldg(x + m * 4 + q);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 p = unpack2(u[k], is_f16);
      f[q * 8 + 2 * k] = p.x;
      f[q * 8 + 2 * k + 1] = p.y;
    }
  }
}

__global__ void __launch_bounds__(HD_THREADS) heads_forward_kernel(const uint4 *__restrict__ x, const float *__restrict__ w,
                                                                  const float *__restrict__ b, float *__restrict__ out0,
                                                                  float *__restrict__ out1, long long m_total, int is_f16) {
  __shared__ float sw[2 * HD_C + 2];
  if (threadIdx.x < 2 * HD_C) sw[threadIdx.x] = w[threadIdx.x];
  if (threadIdx.x < 2) sw[2 * HD_C + threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  for (long long m = blockIdx.x * (long long)HD_THREADS + threadIdx.x; m < m_total; m += (long long)gridDim.x * HD_THREADS) {
    float f[32];
    load_row32(x, m, is_f16, f);
    float s0 = sw[2 * HD_C], s1 = sw[2 * HD_C + 1];
#pragma unroll
    for (int c = 0; c < HD_C; ++c) {
      s0 = fmaf(sw[c], f[c], s0);
      s1 = fmaf(sw[HD_C + c], f[c], s1);
    }
    out0[m] = 1.0f / (1.0f + expf(-s0));
    out1[m] = 1.0f / (1.0f + expf(-s1));
  }
}

// dlogit_k = g_k * s_k * (1 - s_k);  dx = dlogit_0 * w_0 + dlogit_1 * w_1;  partial[cta][k*33 + c] = sum dlogit_k * x_c,
// partial[cta][k*33 + 32] = sum dlogit_k   (66 fp64 values per CTA = the [2][33] layout bn_partials_reduce_kernel sums)
__global__ void __launch_bounds__(HD_THREADS) heads_backward_kernel(const uint4 *__restrict__ x, const float *__restrict__ w,
                                                                   const float *__restrict__ s0, const float *__restrict__ s1,
                                                                   const float *__restrict__ g0, const float *__restrict__ g1,
                                                                   uint4 *__restrict__ dx, double *__restrict__ partial,
                                                                   long long m_total, int is_f16) {
  __shared__ float sw[2 * HD_C];
  __shared__ float red[HD_THREADS / 32][2 * (HD_C + 1)];
  if (threadIdx.x < 2 * HD_C) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const long long r0 = (m_total * blockIdx.x) / gridDim.x, r1 = (m_total * (blockIdx.x + 1)) / gridDim.x;
  float a0[HD_C + 1], a1[HD_C + 1];
#pragma unroll
  for (int c = 0; c <= HD_C; ++c) a0[c] = a1[c] = 0.0f;
  for (long long m = r0 + threadIdx.x; m < r1; m += HD_THREADS) {
    float f[32];
    load_row32(x, m, is_f16, f);
    const float p0 = __ldg(s0 + m), p1 = __ldg(s1 + m);
    const float d0 = __ldg(g0 + m) * p0 * (1.0f - p0), d1 = __ldg(g1 + m) * p1 * (1.0f - p1);
    float o[32];
#pragma unroll
    for (int c = 0; c < HD_C; ++c) {
      o[c] = fmaf(d0, sw[c], d1 * sw[HD_C + c]);
      a0[c] = fmaf(d0, f[c], a0[c]);
      a1[c] = fmaf(d1, f[c], a1[c]);
    }
    a0[HD_C] += d0;
    a1[HD_C] += d1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dx[m * 4 + q] = make_uint4(pack2(o[q * 8], o[q * 8 + 1], is_f16), pack2(o[q * 8 + 2], o[q * 8 + 3], is_f16),
                                 pack2(o[q * 8 + 4], o[q * 8 + 5], is_f16), pack2(o[q * 8 + 6], o[q * 8 + 7], is_f16));
  }
  // warp shuffle tree, then the eight warps of the CTA in fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c <= HD_C; ++c) {
    float v0 = a0[c], v1 = a1[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_down_sync(0xffffffffu, v0, o);
      v1 += __shfl_down_sync(0xffffffffu, v1, o);
    }
    if (lane == 0) {
      red[warp][c] = v0;
      red[warp][HD_C + 1 + c] = v1;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * (HD_C + 1)) {
    double s = 0.0;
    for (int k = 0; k < HD_THREADS / 32; ++k) s += (double)red[k][threadIdx.x];
    partial[(size_t)blockIdx.x * 2 * (HD_C + 1) + threadIdx.x] = s;
  }
}

__global__ void heads_partials_reduce_kernel(const double *__restrict__ partial, int ctas, float *__restrict__ dw, float *__restrict__ db) {
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (u >= 2 * (HD_C + 1)) return;
  double s = 0.0;
  for (int b = lane; b < ctas; b += 32) s += partial[(size_t)b * 2 * (HD_C + 1) + u];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const int k = u / (HD_C + 1), c = u - k * (HD_C + 1);
    if (c < HD_C) dw[k * HD_C + c] = (float)s;
    else db[k] = (float)s;
  }
}

static int heads_ctas(long long m) {
  long long want = (m + HD_THREADS * 4 - 1) / (HD_THREADS * 4);
  const long long cap = (long long)(sm_count() > 0 ? sm_count() : 148) * 4;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace dram

extern "C" int64_t dram_heads_workspace_bytes(void) {
  return (int64_t)((sm_count() > 0 ? sm_count() : 148) * 4) * 2 * (HD_C + 1) * (int64_t)sizeof(double);
}

extern "C" int dram_heads_sigmoid_forward(const void *x, const float *w, const float *b, float *out0, float *out1, int64_t m,
                                          int32_t dtype, void *stream) {
  DRAM_REQUIRE(x && w && b && out0 && out1, "dram_heads_sigmoid_forward: null pointer");
  DRAM_REQUIRE(m > 0, "dram_heads_sigmoid_forward: empty input");
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_heads_sigmoid_forward: bad dtype");
  heads_forward_kernel<<<stream_grid(m, HD_THREADS), HD_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4 *>(x), w, b, out0, out1, m, dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("heads_forward_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_heads_sigmoid_backward(const void *x, const float *w, const float *s0, const float *s1, const float *g0,
                                           const float *g1, void *dx, float *dw, float *db, void *workspace, int64_t m,
                                           int32_t dtype, void *stream) {
  DRAM_REQUIRE(x && w && s0 && s1 && g0 && g1 && dx && dw && db && workspace, "dram_heads_sigmoid_backward: null pointer");
  DRAM_REQUIRE(m > 0, "dram_heads_sigmoid_backward: empty input");
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_heads_sigmoid_backward: bad dtype");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ctas = heads_ctas(m);
  heads_backward_kernel<<<ctas, HD_THREADS, 0, st>>>(reinterpret_cast<const uint4 *>(x), w, s0, s1, g0, g1,
                                                     reinterpret_cast<uint4 *>(dx), reinterpret_cast<double *>(workspace), m,
                                                     dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("heads_backward_kernel launch");
  heads_partials_reduce_kernel<<<ceil_div(2 * (HD_C + 1) * 32, 256), 256, 0, st>>>(reinterpret_cast<const double *>(workspace),
                                                                                  ctas, dw, db);
  DRAM_CHECK_LAUNCH("heads_partials_reduce_kernel launch");
  return DRAM_OK;
}

// Weight re-packing for the training step: the fp32 master weights [Cout][Cin][taps] (PyTorch layout) change every
// optimiser step and the kernels read 16-bit packed copies —
//   forward / wgrad-free layout  out[co][t][ci]            = w[co][cin0 + ci][t]               (pack_conv_weight)
//   data-gradient layout         out[ci][t][co (padded)]   = w[co][cin0 + ci][taps - 1 - t]    (pack_dgrad_weight)
// One thread per output element, coalesced 16-bit writes; the strided fp32 reads of a layer (<= 28 MB) hit in L2.
namespace dram {

__device__ __forceinline__ uint16_t to16(float v, int is_f16) {
  if (is_f16) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const uint16_t *>(&h);
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t *>(&h);
}

__global__ void __launch_bounds__(256) pack_weight_kernel(const float *__restrict__ w, uint16_t *__restrict__ out, int cout,
                                                         int cin_total, int taps, int cin0, int cin_n, int transpose,
                                                         int cout_pad, int is_f16) {
  const long long total = transpose ? (long long)cin_n * taps * cout_pad : (long long)cout * taps * cin_n;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    float v = 0.0f;
    if (transpose) {
      const int co = (int)(i % cout_pad);
      const long long r = i / cout_pad;
      const int t = (int)(r % taps), ci = (int)(r / taps);
      if (co < cout) v = __ldg(w + ((size_t)co * cin_total + cin0 + ci) * taps + (taps - 1 - t));
    } else {
      const int ci = (int)(i % cin_n);
      const long long r = i / cin_n;
      const int t = (int)(r % taps), co = (int)(r / taps);
      v = __ldg(w + ((size_t)co * cin_total + cin0 + ci) * taps + t);
    }
    out[i] = to16(v, is_f16);
  }
}

}  // namespace dram

extern "C" int dram_pack_conv_weight(const float *w, void *out, int32_t cout, int32_t cin_total, int32_t taps, int32_t cin0,
                                     int32_t cin_n, int32_t transpose, int32_t cout_pad, int32_t dtype, void *stream) {
  DRAM_REQUIRE(w && out, "dram_pack_conv_weight: null pointer");
  DRAM_REQUIRE(cout > 0 && cin_total > 0 && taps > 0 && cin0 >= 0 && cin_n > 0 && cin0 + cin_n <= cin_total,
               "dram_pack_conv_weight: bad shape");
  DRAM_REQUIRE(!transpose || cout_pad >= cout, "dram_pack_conv_weight: cout_pad < cout");
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_pack_conv_weight: bad dtype");
  const long long total = transpose ? (long long)cin_n * taps * cout_pad : (long long)cout * taps * cin_n;
  pack_weight_kernel<<<stream_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<uint16_t *>(out), cout, cin_total, taps, cin0, cin_n, transpose, cout_pad, dtype == DRAM_DTYPE_F16);
  DRAM_CHECK_LAUNCH("pack_weight_kernel launch");
  return DRAM_OK;
}
