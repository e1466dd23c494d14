// Device versions of the CPU steps either side of the network.
//
// f1  dataset.py:66-80  lung = lobe > 0; 2 x binary_dilation with the full 3x3x3 structure (== a 5x5x5 box
//                       maximum with a zero border); scan[outside] = -2048; bounding box of the lung grown by
//                       ceil(5 mm / spacing) (utils.py:53-63, the growth is applied by the caller); crop;
//                       ess = (scan < -910) & lung.  All integer/byte work: bit-exact.
// f2  processor.py:111-158 + utils.py:28-37  trilinear (align_corners=True) resample of a dRAM to the crop
//                       size, paste into the zero-filled full volume, clip to [0,1], x255 in fp64, truncate
//                       to uint8 — one pass over the output volume, only uint8 leaves the GPU.
#include <math.h>

#include "common.h"

namespace dram {

// ---------------------------------------------------------------------------------------
// bounding box of mask != 0: bbox = {zmin, zmax+1, ymin, ymax+1, xmin, xmax+1}; one warp per (z, y) row
// ---------------------------------------------------------------------------------------
__global__ void bbox_init_kernel(int *bbox, int D, int H, int W) {
  bbox[0] = D; bbox[1] = 0; bbox[2] = H; bbox[3] = 0; bbox[4] = W; bbox[5] = 0;
}
__global__ void __launch_bounds__(256)
mask_bbox_kernel(const uint8_t *__restrict__ mask, int *__restrict__ bbox, int D, int H, int W) {
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int rows = D * H;
  int zlo = D, zhi = 0, ylo = H, yhi = 0, xlo = W, xhi = 0;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
    const uint8_t *row = mask + (int64_t)r * W;
    int lo = W, hi = 0;
    for (int x = lane; x < W; x += 32)
      if (row[x]) {
        lo = min(lo, x);
        hi = max(hi, x + 1);
      }
    if (hi > 0) {
      const int z = r / H, y = r - z * H;
      xlo = min(xlo, lo); xhi = max(xhi, hi);
      zlo = min(zlo, z); zhi = max(zhi, z + 1);
      ylo = min(ylo, y); yhi = max(yhi, y + 1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    zlo = min(zlo, __shfl_xor_sync(0xffffffffu, zlo, o)); zhi = max(zhi, __shfl_xor_sync(0xffffffffu, zhi, o));
    ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
    xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
  }
  if (lane == 0 && zhi > 0) {
    atomicMin(&bbox[0], zlo); atomicMax(&bbox[1], zhi);
    atomicMin(&bbox[2], ylo); atomicMax(&bbox[3], yhi);
    atomicMin(&bbox[4], xlo); atomicMax(&bbox[5], xhi);
  }
}

// ---------------------------------------------------------------------------------------
// f1 pass A: in-plane 5x5 maximum of (lobe > 0), zero outside the volume, for the planes
// [z0 - 2, z0 + cd + 2) of the crop window (planes outside the volume are written as zero).
// tmp: uint8 [(cd + 4)][ch][cw].  One thread per voxel, x fastest.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lung_dilate_plane_kernel(const uint8_t *__restrict__ lobe, uint8_t *__restrict__ tmp, int D, int H, int W, int z0,
                         int y0, int x0, int cd, int ch, int cw) {
  const int64_t total = (int64_t)(cd + 4) * ch * cw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(t % cw);
    const int64_t r = t / cw;
    const int y = (int)(r % ch);
    const int zz = (int)(r / ch);
    const int gz = z0 - 2 + zz, gy = y0 + y, gx = x0 + x;
    uint8_t m = 0;
    if (gz >= 0 && gz < D) {
      const uint8_t *pl = lobe + (int64_t)gz * H * W;
#pragma unroll
      for (int dy = -2; dy <= 2; ++dy) {
        const int yy = gy + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
          const int xx = gx + dx;
          if (xx >= 0 && xx < W) m |= pl[(int64_t)yy * W + xx];
        }
      }
    }
    tmp[t] = m ? 1 : 0;
  }
}
// f1 pass B: 5-plane maximum of pass A + blanking + crop + masks.
__global__ void __launch_bounds__(256)
lung_crop_kernel(const short *__restrict__ scan, const uint8_t *__restrict__ lobe, const uint8_t *__restrict__ tmp,
                 short *__restrict__ image_c, uint8_t *__restrict__ lung_c, uint8_t *__restrict__ ess_c, int D, int H,
                 int W, int z0, int y0, int x0, int cd, int ch, int cw, short blank, short ess_below) {
  const int64_t total = (int64_t)cd * ch * cw;
  const int64_t cplane = (int64_t)ch * cw;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(t % cw);
    const int64_t r = t / cw;
    const int y = (int)(r % ch);
    const int z = (int)(r / ch);
    const int64_t src = ((int64_t)(z0 + z) * H + (y0 + y)) * W + (x0 + x);
    const int64_t inplane = (int64_t)y * cw + x;
    uint8_t dil = 0;
#pragma unroll
    for (int dz = 0; dz < 5; ++dz) dil |= tmp[(int64_t)(z + dz) * cplane + inplane];
    const short v = dil ? scan[src] : blank;
    const uint8_t lung = lobe[src] ? 1 : 0;
    image_c[t] = v;
    lung_c[t] = lung;
    ess_c[t] = (v < ess_below && lung) ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------
// f2: full-volume uint8 heat-map from one network-size dRAM.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
heatmap_u8_kernel(const float *__restrict__ map, uint8_t *__restrict__ out, int d, int h, int w, int OD, int OH,
                  int OW, int z0, int y0, int x0, int cd, int ch, int cw, float sd, float sh, float sw) {
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int rows = OD * OH;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
    const int gz = r / OH, gy = r - gz * OH;
    uint8_t *orow = out + (int64_t)r * OW;
    const int z = gz - z0, y = gy - y0;
    if (z < 0 || z >= cd || y < 0 || y >= ch) {
      for (int x = lane; x < OW; x += 32) orow[x] = 0;
      continue;
    }
    const LinIdx id = lin_index_ac(z, sd, d), ih = lin_index_ac(y, sh, h);
    const float *r00 = map + ((int64_t)id.i0 * h + ih.i0) * w, *r01 = map + ((int64_t)id.i0 * h + ih.i1) * w;
    const float *r10 = map + ((int64_t)id.i1 * h + ih.i0) * w, *r11 = map + ((int64_t)id.i1 * h + ih.i1) * w;
    for (int gx = lane; gx < OW; gx += 32) {
      const int x = gx - x0;
      uint8_t u = 0;
      if (x >= 0 && x < cw) {
        const LinIdx iw = lin_index_ac(x, sw, w);
        float v = id.w0 * (ih.w0 * (iw.w0 * __ldg(r00 + iw.i0) + iw.w1 * __ldg(r00 + iw.i1)) +
                           ih.w1 * (iw.w0 * __ldg(r01 + iw.i0) + iw.w1 * __ldg(r01 + iw.i1))) +
                  id.w1 * (ih.w0 * (iw.w0 * __ldg(r10 + iw.i0) + iw.w1 * __ldg(r10 + iw.i1)) +
                           ih.w1 * (iw.w0 * __ldg(r11 + iw.i0) + iw.w1 * __ldg(r11 + iw.i1)));
        v = fminf(fmaxf(v, 0.0f), 1.0f);
        u = (uint8_t)(int)((double)v * 255.0);  // numpy: float64 product, astype(uint8) truncates
      }
      orow[gx] = u;
    }
  }
}

}  // namespace dram

using namespace dram;

static const int kThreads = 256;

extern "C" int dram_mask_bbox(const uint8_t *mask, int32_t D, int32_t H, int32_t W, int32_t *bbox, void *stream) {
  DRAM_REQUIRE(mask && bbox, "dram_mask_bbox: null pointer");
  DRAM_REQUIRE(D > 0 && H > 0 && W > 0 && (int64_t)D * H < 0x7fffffffLL, "dram_mask_bbox: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  bbox_init_kernel<<<1, 1, 0, st>>>(bbox, D, H, W);
  DRAM_CHECK_LAUNCH("bbox_init_kernel");
  mask_bbox_kernel<<<stream_grid((int64_t)D * H * 32, kThreads, 8), kThreads, 0, st>>>(mask, bbox, D, H, W);
  DRAM_CHECK_LAUNCH("mask_bbox_kernel");
  return DRAM_OK;
}

extern "C" size_t dram_lung_crop_workspace_bytes(int32_t cd, int32_t ch, int32_t cw) {
  if (cd <= 0 || ch <= 0 || cw <= 0) return 0;
  return (size_t)(cd + 4) * (size_t)ch * (size_t)cw;
}

extern "C" int dram_lung_crop(const int16_t *scan, const uint8_t *lobe, int32_t D, int32_t H, int32_t W, int32_t z0,
                              int32_t y0, int32_t x0, int32_t cd, int32_t ch, int32_t cw, int16_t *image_c,
                              uint8_t *lung_c, uint8_t *ess_c, void *workspace, void *stream) {
  DRAM_REQUIRE(scan && lobe && image_c && lung_c && ess_c && workspace, "dram_lung_crop: null pointer");
  DRAM_REQUIRE(D > 0 && H > 0 && W > 0, "dram_lung_crop: empty volume");
  DRAM_REQUIRE(z0 >= 0 && y0 >= 0 && x0 >= 0 && cd > 0 && ch > 0 && cw > 0 && z0 + cd <= D && y0 + ch <= H &&
                   x0 + cw <= W,
               "dram_lung_crop: crop [%d:%d, %d:%d, %d:%d] outside the %dx%dx%d volume", z0, z0 + cd, y0, y0 + ch, x0,
               x0 + cw, D, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t *tmp = reinterpret_cast<uint8_t *>(workspace);
  const int64_t tA = (int64_t)(cd + 4) * ch * cw, tB = (int64_t)cd * ch * cw;
  lung_dilate_plane_kernel<<<stream_grid(tA, kThreads), kThreads, 0, st>>>(lobe, tmp, D, H, W, z0, y0, x0, cd, ch, cw);
  DRAM_CHECK_LAUNCH("lung_dilate_plane_kernel");
  lung_crop_kernel<<<stream_grid(tB, kThreads), kThreads, 0, st>>>(scan, lobe, tmp, image_c, lung_c, ess_c, D, H, W, z0,
                                                                  y0, x0, cd, ch, cw, (short)-2048, (short)-910);
  DRAM_CHECK_LAUNCH("lung_crop_kernel");
  return DRAM_OK;
}

extern "C" int dram_heatmap_u8(const float *map, int32_t d, int32_t h, int32_t w, uint8_t *out, int32_t OD, int32_t OH,
                               int32_t OW, int32_t z0, int32_t y0, int32_t x0, int32_t cd, int32_t ch, int32_t cw,
                               void *stream) {
  DRAM_REQUIRE(map && out, "dram_heatmap_u8: null pointer");
  DRAM_REQUIRE(d > 0 && h > 0 && w > 0 && OD > 0 && OH > 0 && OW > 0 && (int64_t)OD * OH < 0x7fffffffLL,
               "dram_heatmap_u8: bad shape");
  DRAM_REQUIRE(z0 >= 0 && y0 >= 0 && x0 >= 0 && cd > 0 && ch > 0 && cw > 0 && z0 + cd <= OD && y0 + ch <= OH &&
                   x0 + cw <= OW,
               "dram_heatmap_u8: crop outside the output volume");
  heatmap_u8_kernel<<<stream_grid((int64_t)OD * OH * 32, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      map, out, d, h, w, OD, OH, OW, z0, y0, x0, cd, ch, cw, ac_scale(d, cd), ac_scale(h, ch), ac_scale(w, cw));
  DRAM_CHECK_LAUNCH("heatmap_u8_kernel");
  return DRAM_OK;
}
