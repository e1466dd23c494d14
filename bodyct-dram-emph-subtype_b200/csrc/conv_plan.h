// The opaque plan behind dram_conv_plan: tensor maps + kernel parameters of either conv kernel.
#pragma once
#include "umma_common.cuh"

namespace dram {

// Per-tap TMA tile kernel (conv3d_umma.cu)
struct ConvKParams {
  int n, Do, Ho, Wo, Di, Hi, Wi;
  int tw, th, td;              // tile extents (powers of two, product 128)
  int tw_log2, th_log2;        // row -> (w,h,d) decode
  int tiles_w, tiles_h, tiles_d, tiles_per_sample, num_n_tiles, total_tiles;
  int kd, kh, kw, sd, sh, sw, dd, dh, dw, pd, ph, pw;
  int chunks1, chunks_total;   // 64-channel chunks of source 1 / of both sources
  int staged_res;              // staged epilogue: 1 = the residual tile comes through TMA as well
  EpiParams epi;
};


// Plane-ring kernel (conv3d_slab.cu): 3x3x3, stride 1, dilation 1, Cout in {32, 64}
struct SlabParams {
  int n, D, H, W;              // input == output spatial dims
  int cols_w, cols_h, groups_d, items_total;
  int chunks1, chunks_total;
  int desc_base_offset_mode;   // experiment knob for the UMMA descriptor base-offset field (0 = leave 0)
  // source 1 given at half resolution and up-sampled x2 (trilinear, align_corners=True) by the kernel
  int up2x;                    // 0 / 1
  int Dl, Hl, Wl;              // low-resolution dims (D/2, H/2, W/2)
  float up_sd, up_sh, up_sw;   // ATen source scales (Dl-1)/(D-1) ...
  EpiParams epi;
};

}  // namespace dram

struct dram_conv_plan {
  int kind;  // 0 = per-tap tiles, 1 = plane ring
  CUtensorMap map_a1, map_a2, map_w;
  CUtensorMap map_out, map_res;  // staged (TMA store) epilogue of the tile kernel
  int staged;
  dram::ConvKParams p;
  dram::SlabParams sp;
  int block_n;
  int stages;
  size_t smem_bytes;
  int64_t flops;
  int m_tiles, n_tiles;
};
