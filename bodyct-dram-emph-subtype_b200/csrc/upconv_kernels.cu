// K13 — the three interpolation passes of the commuted "3x3x3 convolution of a x2 up-sampled tensor" (SURVEY H4),
// used for the first decoder convolution us1.0 (med3d.py:83-89: Upsample -> cat -> Conv3d) whose up-sampled operand
// has 512 (ResNet-18/34) or 2048 (ResNet-50) channels but only 64 outputs.
//
// Trilinear up-sampling is linear and acts on space only, the convolution's channel mixing acts on channels only, so
//     conv(up(x))[o, co] = sum_tap sum_ci W[co, ci, tap] * up(x)[o + tap, ci]          (zero outside the volume)
//                        = sum_tap up(z_tap)[o + tap, co],     z_tap[v, co] = sum_ci W[co, ci, tap] * x[v, ci]
// i.e. 27 channel-mixing products at LOW resolution (one 1x1x1 K1 launch with 27 * 64 output channels: 8x fewer
// MACs than the convolution at high resolution) followed by a gather that is separable per axis:
//     r[sd, sh, ow | td, th]  = sum_tw  lerp_w(z[sd, sh, . | td, th, tw])(ow + tw - 1)
//     q[sd, oh, ow | td]      = sum_th  lerp_h(r[sd, ., ow | td, th])(oh + th - 1)
//     G[od, oh, ow]           = sum_td  lerp_d(q[., oh, ow | td])(od + td - 1)
// with lerp = ATen's align_corners=True source index / weights (lin_index_ac) and taps that leave the high-resolution
// volume contributing zero (the convolution's zero padding).  One kernel does one axis; the tensors are
//     in  [outer][L_lo][inner][G][3][64]   ->   out [outer][L_hi][inner][G][64]       (16-bit, channels innermost)
// (W pass: outer = N*Dl*Hl, inner = 1, G = 9; H pass: outer = N*Dl, inner = W, G = 3; D pass: outer = N, inner = H*W,
// G = 1).  A thread produces 8 channels of one output position from 6 16-byte loads; fp32 accumulation, one rounding
// to the storage type per pass.  Memory-bound: the up-sampled 512/2048-channel tensor (K4's output, 268 MB / 6.7 GB
// per volume) is never written.
#include "common.h"

namespace dram {

template <bool F16>
__device__ __forceinline__ void upconv_fma8(float (&acc)[8], const uint4 u, float w) {
  const uint32_t v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float lo, hi;
    if constexpr (F16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&v[q]));
      lo = f.x;
      hi = f.y;
    } else {
      lo = __uint_as_float(v[q] << 16);
      hi = __uint_as_float(v[q] & 0xffff0000u);
    }
    acc[2 * q] = fmaf(w, lo, acc[2 * q]);
    acc[2 * q + 1] = fmaf(w, hi, acc[2 * q + 1]);
  }
}

// Thread block = (8 channel groups, G tap groups, P output positions), 8 * G * P <= 256: the channel / tap-group part of
// the index is the thread's coordinate, only the position is decoded (32-bit divisions, once per thread) — the first
// version decoded a flat 64-bit index with six divisions and was bound by them (2 TB/s).
template <bool F16>
__global__ void __launch_bounds__(256)
upconv_axis_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, unsigned positions, int L_lo, int L_hi, unsigned inner,
                   int G, int in_g_stride /* uint4 units per input position */, float scale) {
  const int c8 = threadIdx.x, g = threadIdx.y;
  for (unsigned pos = blockIdx.x * blockDim.z + threadIdx.z; pos < positions; pos += gridDim.x * blockDim.z) {
    const unsigned isp = pos % inner;
    const unsigned v = pos / inner;
    const int o = (int)(v % (unsigned)L_hi);
    const unsigned ou = v / (unsigned)L_hi;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const uint4 *base = in + ((int64_t)ou * L_lo * inner + isp) * (int64_t)in_g_stride + g * 24 + c8;
    const int64_t lstride = (int64_t)inner * in_g_stride;  // uint4 units between consecutive low-resolution positions
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
      const int p = o + tap - 1;
      if (p < 0 || p >= L_hi) continue;  // the convolution's zero padding at high resolution
      const LinIdx li = lin_index_ac(p, scale, L_lo);
      const uint4 a = __ldg(base + (int64_t)li.i0 * lstride + tap * 8);
      const uint4 b = __ldg(base + (int64_t)li.i1 * lstride + tap * 8);
      upconv_fma8<F16>(acc, a, li.w0);
      upconv_fma8<F16>(acc, b, li.w1);
    }
    uint32_t w[4];
    if constexpr (F16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[q]) : "f"(acc[2 * q + 1]), "f"(acc[2 * q]));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
        w[q] = *reinterpret_cast<const uint32_t *>(&h);
      }
    }
    out[((int64_t)pos * G + g) * 8 + c8] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace dram

using namespace dram;

extern "C" int dram_upconv_axis(const void *in, void *out, int64_t outer, int32_t l_lo, int32_t l_hi, int64_t inner,
                                int32_t groups, int32_t in_channels, int32_t dtype, void *stream) {
  DRAM_REQUIRE(in && out, "dram_upconv_axis: null pointer");
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_upconv_axis: dtype must be bf16 (0) or fp16 (1)");
  DRAM_REQUIRE(outer > 0 && l_lo > 0 && l_hi > 0 && inner > 0 && groups > 0, "dram_upconv_axis: bad shape");
  DRAM_REQUIRE(in_channels >= groups * 192 && in_channels % 8 == 0,
               "dram_upconv_axis: an input position holds groups * 3 taps * 64 channels (+ optional padding), got %d for %d groups",
               in_channels, groups);
  DRAM_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "dram_upconv_axis: buffers must be 16-byte aligned");
  const int64_t positions = outer * l_hi * inner;
  DRAM_REQUIRE(positions < 0xffffffffLL && inner < 0xffffffffLL && groups <= 32,
               "dram_upconv_axis: too many output positions for 32-bit indexing");
  const int per_block = 256 / (8 * groups) > 0 ? 256 / (8 * groups) : 1;
  const dim3 block(8, groups, per_block);
  int64_t want = ceil_div64(positions, per_block);
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(want < cap ? want : cap);
  const float scale = ac_scale(l_lo, l_hi);
  if (dtype == DRAM_DTYPE_F16)
    upconv_axis_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4 *>(in), reinterpret_cast<uint4 *>(out), (unsigned)positions, l_lo, l_hi, (unsigned)inner, groups,
        in_channels / 8, scale);
  else
    upconv_axis_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4 *>(in), reinterpret_cast<uint4 *>(out), (unsigned)positions, l_lo, l_hi, (unsigned)inner, groups,
        in_channels / 8, scale);
  DRAM_CHECK_LAUNCH("upconv_axis_kernel");
  return DRAM_OK;
}
