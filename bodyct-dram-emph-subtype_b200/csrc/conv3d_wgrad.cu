// K9 — conv3d weight gradient (training path; the reference gets it from autograd through every
// nn.Conv3d of med3d.py:93-100, 67/76, 152-157 when train.py runs Lightning's automatic optimisation):
//
//     dW[cout][cin][kd][kh][kw] = sum over output voxels v of  dy[v][cout] * x[v*stride + tap*dilation - pad][cin]
//
// GEMM view: M = (tap, cin) rows, N = cout, K = output voxels.  Both operands are NDHWC 16-bit tensors, so the
// channels — the GEMM M and N dimensions — are the contiguous axis: a TMA box of 64 voxels x 64 channels lands in
// shared memory as 64 SWIZZLE_128B rows, which is exactly the MN-major operand layout of tcgen05.mma
// (cute: ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in elements) with the voxels as K.  Nothing is transposed anywhere.
//
//   * A operand = two "units" (a unit = one tap x one 64-channel chunk of x) stacked along M through the
//     descriptor's leading-dimension offset: M = 128.  The x box of a unit is the output tile shifted by the tap
//     (stride = tensor-map element stride, zero padding = out-of-bounds fill), exactly the box K1's tile kernel loads.
//   * B operand = the dy tile, N = min(Cout, 256) channels in 64-channel boxes.
//   * a work item = (unit pair, cout block, K slice); it streams its voxel tiles through a TMA/mbarrier ring,
//     accumulates D[128 x N] in TMEM (fp32) and writes it once to a per-slice partial buffer.  Tiles whose x boxes lie
//     entirely in the padding are skipped.
//   * dram_wgrad_finish_kernel adds the slices (fixed order: deterministic) and scatters to the PyTorch layout.
//
// Roles (192 threads): warps 0-3 epilogue, warp 4 lane 0 TMA producer, warp 5 MMA issuer (owns TMEM).
//
// PLANES variant (3x3x3, stride 1, dilation 1, pad 1, Cout <= 64: layer1 and the whole decoder, where the streaming
// kernel above is bound by re-reading x once per tap through L2):
//   * a CTA owns one (kd, 64-channel chunk) "key" and a contiguous range of (column, plane) units, a column being
//     8 (W) x 16 (H) voxels; a unit pairs output plane t with input plane t + kd - 1, so one stage = one x plane with
//     halo (10 x 18 voxels, 23 KiB, one TMA box) + one dy plane (128 voxels, 16 KiB): each is fetched ONCE for the
//     nine (kh, kw) taps of the key;
//   * the A operand of tap (kh, kw) is the x plane seen through an MN-major descriptor whose start is shifted by
//     (kh*10 + kw) rows and whose K-group stride is the plane pitch of an h row (10 rows) — nothing moves per tap;
//     taps (kh=0, kw) and (kh=1, kw) are stacked along M through the leading-dimension offset (10 rows): M = 128;
//     tap (kh=2, kw) is an M = 64 MMA;
//   * all nine accumulators (6 x 64 TMEM columns) stay resident for the CTA's whole range and are written once;
//     the CTAs of the three kd keys walk the same columns at the same pace, so x and dy come from DRAM once.
#include "conv_plan.h"

namespace dram {

static constexpr int WG_TW = 8, WG_TH = 8, WG_TD = 1;       // voxel tile = 64 output voxels = GEMM-K per stage
static constexpr int WG_KVOX = WG_TW * WG_TH * WG_TD;
static constexpr int WG_UNIT_BYTES = WG_KVOX * 128;         // 8 KiB: 64 voxels x 64 channels
static constexpr int WG_A_BYTES = 2 * WG_UNIT_BYTES;        // M = 128
static constexpr int WG_THREADS = 192;
static constexpr int WG_PRODUCER_WARP = 4, WG_MMA_WARP = 5;

// G = unit pairs per work item: with G = 2 one dy tile feeds two M = 128 accumulators (four units), which cuts the
// bytes pulled through L2 per MMA by a third (BLOCK_N = 256) — the streaming kernel is bound by exactly that.
template <int BLOCK_N, int G>
struct WgCfg {
  static constexpr int A_BYTES = G * WG_A_BYTES;
  static constexpr int B_BYTES = (BLOCK_N / 64) * WG_UNIT_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_FIT = (225 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int ACC_BUFS = (2 * G * BLOCK_N <= 512) ? 2 : 1;  // accumulator sets in TMEM
  static constexpr int TMEM_COLS_RAW = ACC_BUFS * G * BLOCK_N;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 64 ? 64 : (TMEM_COLS_RAW <= 128 ? 128 : (TMEM_COLS_RAW <= 256 ? 256 : 512));
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256;
};

struct WgParams {
  int n, Do, Ho, Wo, Di, Hi, Wi;
  int tiles_w, tiles_h, tiles_d, tiles_per_sample, total_vtiles;
  int kd, kh, kw, sd, sh, sw, dd, dh, dw, pd, ph, pw;
  int chunks;       // Cin / 64 of this source
  int units;        // taps * chunks
  int a_blocks;     // ceil(units / 2): unit pairs = 128-row blocks of the partial buffer
  int s_blocks;     // ceil(a_blocks / G): groups of G pairs = work items per (cout block, slice)
  int n_blocks;     // cout_pad / BLOCK_N
  int kslices;
  int items_total;  // s_blocks * n_blocks * kslices
  int cout, cout_pad;
  int rows_total;   // a_blocks * 128
  int is_f16;
  float *partial;   // [kslices][rows_total][cout_pad]
};

struct WgUnit {
  int c0, ow, oh, od;  // channel offset and input-coordinate offset of the tap: i = o*stride + off
};
__device__ __forceinline__ WgUnit decode_wg_unit(const WgParams &p, int u) {
  WgUnit r;
  const int tap = u / p.chunks;
  r.c0 = (u - tap * p.chunks) * 64;
  const int zd = tap / (p.kh * p.kw), rem = tap - zd * p.kh * p.kw;
  const int zh = rem / p.kw, zw = rem - zh * p.kw;
  r.od = zd * p.dd - p.pd;
  r.oh = zh * p.dh - p.ph;
  r.ow = zw * p.dw - p.pw;
  return r;
}

template <int G>
struct WgItem {
  int slice, a_blk0, n0, vt_begin, vt_end;
  int n_units;           // valid units of this item (1 .. 2G), units a_blk0*2 .. a_blk0*2 + n_units - 1
  WgUnit unit[2 * G];
};
template <int BLOCK_N, int G>
__device__ __forceinline__ WgItem<G> decode_wg_item(const WgParams &p, int item) {
  WgItem<G> it;
  const int n_blk = item % p.n_blocks;
  const int r = item / p.n_blocks;
  const int s_blk = r % p.s_blocks;
  it.slice = r / p.s_blocks;
  it.a_blk0 = s_blk * G;
  it.n0 = n_blk * BLOCK_N;
  it.vt_begin = (int)(((long long)it.slice * p.total_vtiles) / p.kslices);
  it.vt_end = (int)(((long long)(it.slice + 1) * p.total_vtiles) / p.kslices);
  const int u0 = 2 * it.a_blk0;
  it.n_units = p.units - u0 < 2 * G ? p.units - u0 : 2 * G;
#pragma unroll
  for (int q = 0; q < 2 * G; ++q) it.unit[q] = decode_wg_unit(p, q < it.n_units ? u0 + q : u0);
  return it;
}

struct WgTile {
  int sample, d0, h0, w0;
};
__device__ __forceinline__ WgTile decode_wg_tile(const WgParams &p, int vt) {
  WgTile t;
  t.sample = vt / p.tiles_per_sample;
  int r = vt - t.sample * p.tiles_per_sample;
  const int iw = r % p.tiles_w;
  r /= p.tiles_w;
  const int ih = r % p.tiles_h;
  t.w0 = iw * WG_TW;
  t.h0 = ih * WG_TH;
  t.d0 = (r / p.tiles_h) * WG_TD;
  return t;
}
// The x box of a unit for this tile lies entirely in the zero padding.
__device__ __forceinline__ bool wg_unit_is_padding(const WgParams &p, const WgTile &t, const WgUnit &u) {
  return tap_is_padding(t.d0 * p.sd + u.od, WG_TD, p.sd, p.Di) || tap_is_padding(t.h0 * p.sh + u.oh, WG_TH, p.sh, p.Hi) ||
         tap_is_padding(t.w0 * p.sw + u.ow, WG_TW, p.sw, p.Wi);
}
// A tile is streamed unless every unit of the item reads only padding; the first tile of a slice is always
// streamed so that the accumulators are initialised (padding boxes arrive as zeros).
template <int G>
__device__ __forceinline__ bool wg_tile_is_skipped(const WgParams &p, const WgItem<G> &it, int vt, const WgTile &t) {
  if (vt == it.vt_begin) return false;
  bool all_pad = true;
#pragma unroll
  for (int q = 0; q < 2 * G; ++q)
    if (q < it.n_units) all_pad = all_pad && wg_unit_is_padding(p, t, it.unit[q]);
  return all_pad;
}

// MN-major SWIZZLE_128B descriptor: 64-element groups along M/N every `lbo` bytes, 8-row groups along K every 1 KiB.
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BLOCK_N, int G>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv3d_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                    const __grid_constant__ WgParams p) {
  using Cfg = WgCfg<BLOCK_N, G>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC_BUFS = Cfg::ACC_BUFS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto smem_a = [&](int s) { return smem_base + (uint32_t)s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + (uint32_t)s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full(a), 1);
      mbar_init(tmem_empty(a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (warp == WG_PRODUCER_WARP && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_dy);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp == WG_PRODUCER_WARP) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.items_total; item += gridDim.x) {
        const WgItem<G> it = decode_wg_item<BLOCK_N, G>(p, item);
        const uint32_t bytes = (uint32_t)(Cfg::B_BYTES + it.n_units * WG_UNIT_BYTES);
        for (int vt = it.vt_begin; vt < it.vt_end; ++vt) {
          const WgTile t = decode_wg_tile(p, vt);
          if (wg_tile_is_skipped<G>(p, it, vt, t)) continue;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), bytes);
#pragma unroll
          for (int q = 0; q < 2 * G; ++q)
            if (q < it.n_units) {
              const WgUnit &u = it.unit[q];
              tma_load_5d(smem_a(stage) + (uint32_t)(q * WG_UNIT_BYTES), &map_x, full_bar(stage), u.c0, t.w0 * p.sw + u.ow,
                          t.h0 * p.sh + u.oh, t.d0 * p.sd + u.od, t.sample);
            }
#pragma unroll
          for (int g = 0; g < BLOCK_N / 64; ++g)
            tma_load_5d(smem_b(stage) + (uint32_t)(g * WG_UNIT_BYTES), &map_dy, full_bar(stage), it.n0 + g * 64, t.w0,
                        t.h0, t.d0, t.sample);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == WG_MMA_WARP) {
    // A and B MN-major (bits 15 and 16 of the instruction descriptor)
    const uint32_t idesc = make_idesc_16bit(128, BLOCK_N, p.is_f16) | (1u << 15) | (1u << 16);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < p.items_total; item += gridDim.x) {
      const WgItem<G> it = decode_wg_item<BLOCK_N, G>(p, item);
      const int pairs = (it.n_units + 1) / 2;
      mbar_wait(tmem_empty(acc), acc_phase ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * G * BLOCK_N);
      uint32_t accumulate = 0;
      for (int vt = it.vt_begin; vt < it.vt_end; ++vt) {
        const WgTile t = decode_wg_tile(p, vt);
        if (wg_tile_is_skipped<G>(p, it, vt, t)) continue;
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < WG_KVOX / 16; ++k) {  // 16 voxels = 16 rows of 128 bytes per K step
            const uint64_t db = make_sw128_mn_desc(smem_b(stage) + (uint32_t)(k * 16 * 128), WG_UNIT_BYTES);
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (g < pairs) {
                const uint64_t da =
                    make_sw128_mn_desc(smem_a(stage) + (uint32_t)(g * WG_A_BYTES + k * 16 * 128), WG_UNIT_BYTES);
                umma_bf16(tmem_d + (uint32_t)(g * BLOCK_N), da, db, idesc, (k > 0) ? 1u : accumulate);
              }
          }
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one_sync()) umma_commit(tmem_full(acc));
      __syncwarp();
      if (++acc == ACC_BUFS) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------- epilogue warps 0..3: accumulators -> partial[slice][row][cout] ----------
    const int row = warp * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.items_total; item += gridDim.x) {
      const WgItem<G> it = decode_wg_item<BLOCK_N, G>(p, item);
      mbar_wait(tmem_full(acc), acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        // rows of a missing unit (odd tail of the unit list) hold garbage and are not written
        const bool valid = 2 * g + (row >> 6) < it.n_units;
        float *dst = p.partial + ((size_t)it.slice * p.rows_total + (size_t)(it.a_blk0 + g) * 128 + row) * p.cout_pad + it.n0;
        const uint32_t taddr = tmem_base + (uint32_t)((acc * G + g) * BLOCK_N) + ((uint32_t)(warp * 32) << 16);
        if (2 * g >= it.n_units) break;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)c0, v);
          tmem_wait_ld();
          if (valid) {
            float4 *d4 = reinterpret_cast<float4 *>(dst + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                  __uint_as_float(v[4 * j + 3]));
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(tmem_empty(acc));
      if (++acc == ACC_BUFS) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ----------------------------------------------------------------------------------------
// PLANES variant
// ----------------------------------------------------------------------------------------
static constexpr int WP_W = 8, WP_H = 16;                        // column footprint
static constexpr int WP_PW = WP_W + 2, WP_PH = WP_H + 2;         // x plane with halo
static constexpr int WP_X_BYTES = WP_PW * WP_PH * 128;           // 23040
static constexpr int WP_X_PITCH = 23 * 1024;
static constexpr int WP_DY_BYTES = WP_W * WP_H * 128;            // 16384
static constexpr int WP_STAGE_BYTES = WP_X_PITCH + WP_DY_BYTES;  // 39936
static constexpr int WP_STAGES = 5;
static constexpr int WP_TMEM_COLS = 512;                         // 6 accumulators of 64 columns
static constexpr int WP_SMEM_BYTES = 1024 + WP_STAGES * WP_STAGE_BYTES + 256;
static constexpr int WP_PARTIAL_FLOATS = 9 * 64 * 64;            // per CTA: [kh*3+kw][cin][cout]

struct WpParams {
  int n, D, H, W;
  int cols_w, cols_h, columns;  // columns = n * cols_h * cols_w
  int chunks, keys;             // keys = 3 * chunks, key = kd * chunks + chunk
  int ctas_per_key;             // grid = keys * ctas_per_key, CTA b: key b % keys, rank b / keys
  long long upk;                // units per key = columns * D
  int is_f16;
  float *partial;               // [grid][9][64][64]
};

struct WpUnit {
  int sample, h0, w0, t;
};
__device__ __forceinline__ WpUnit decode_wp_unit(const WpParams &p, long long u) {
  WpUnit r;
  const int col = (int)(u / p.D);
  r.t = (int)(u - (long long)col * p.D);
  const int per_sample = p.cols_w * p.cols_h;
  r.sample = col / per_sample;
  const int c = col - r.sample * per_sample;
  const int ih = c / p.cols_w;
  r.h0 = ih * WP_H;
  r.w0 = (c - ih * p.cols_w) * WP_W;
  return r;
}

// MN-major SWIZZLE_128B descriptor with explicit leading (64-element M/N groups) and stride (8-row K groups) offsets.
__device__ __forceinline__ uint64_t make_sw128_mn_desc2(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv3d_wgrad_planes_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                           const __grid_constant__ WpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + WP_STAGES * WP_STAGE_BYTES;
  auto smem_x = [&](int s) { return smem_base + (uint32_t)s * WP_STAGE_BYTES; };
  auto smem_dy = [&](int s) { return smem_base + (uint32_t)s * WP_STAGE_BYTES + WP_X_PITCH; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WP_STAGES + s); };
  const uint32_t tmem_full = bar_base + 8u * (2 * WP_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * WP_STAGES + 1);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WP_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, WP_TMEM_COLS);
  if (warp == WG_PRODUCER_WARP && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_dy);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const int key = blockIdx.x % p.keys, rank = blockIdx.x / p.keys;
  const int kd = key / p.chunks, chunk = key - kd * p.chunks;
  const long long u_lo = (p.upk * rank) / p.ctas_per_key, u_hi = (p.upk * (rank + 1)) / p.ctas_per_key;
  // Output plane t pairs with input plane t + kd - 1; a unit whose input plane is outside the volume contributes
  // nothing and is skipped — except the first unit of the range, which initialises the accumulators (its x box is
  // out of bounds and arrives as zeros).
  auto skipped = [&](long long u, int t) {
    const int z = t + kd - 1;
    return (z < 0 || z >= p.D) && u != u_lo;
  };

  if (warp == WG_PRODUCER_WARP) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long u = u_lo; u < u_hi; ++u) {
        const WpUnit un = decode_wp_unit(p, u);
        if (skipped(u, un.t)) continue;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), WP_X_BYTES + WP_DY_BYTES);
        tma_load_5d(smem_x(stage), &map_x, full_bar(stage), chunk * 64, un.w0 - 1, un.h0 - 1, un.t + kd - 1, un.sample);
        tma_load_5d(smem_dy(stage), &map_dy, full_bar(stage), 0, un.w0, un.h0, un.t, un.sample);
        if (++stage == WP_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == WG_MMA_WARP) {
    const uint32_t mn = (1u << 15) | (1u << 16);
    const uint32_t idesc_pair = make_idesc_16bit(128, 64, p.is_f16) | mn;
    const uint32_t idesc_single = make_idesc_16bit(64, 64, p.is_f16) | mn;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (long long u = u_lo; u < u_hi; ++u) {
      const WpUnit un = decode_wp_unit(p, u);
      if (skipped(u, un.t)) continue;
      mbar_wait(full_bar(stage), phase);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint32_t xs = smem_x(stage), ds = smem_dy(stage);
#pragma unroll
        for (int j = 0; j < WP_H / 2; ++j) {  // K step = 16 voxels = h rows 2j, 2j+1 of the column
          const uint64_t db = make_sw128_mn_desc2(ds + (uint32_t)(j * 16 * 128), 1024u, 1024u);
          const uint32_t acc = (j > 0) ? 1u : accumulate;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t row_pair = (uint32_t)((2 * j) * WP_PW + kw), row_single = (uint32_t)((2 * j + 2) * WP_PW + kw);
            const uint64_t da_pair = make_sw128_mn_desc2(xs + row_pair * 128u, WP_PW * 128u, WP_PW * 128u);
            const uint64_t da_single = make_sw128_mn_desc2(xs + row_single * 128u, WP_PW * 128u, WP_PW * 128u);
            umma_bf16(tmem_base + (uint32_t)((2 * kw) * 64), da_pair, db, idesc_pair, acc);
            umma_bf16(tmem_base + (uint32_t)((2 * kw + 1) * 64), da_single, db, idesc_single, acc);
          }
        }
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == WP_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (elect_one_sync()) umma_commit(tmem_full);
    __syncwarp();
  } else {
    // ------------------------------- epilogue warps 0..3: nine accumulators -> partial[cta][tap][cin][cout] -------
    float *base = p.partial + (size_t)blockIdx.x * WP_PARTIAL_FLOATS;
    mbar_wait(tmem_full, 0u);
    tcgen05_fence_after();
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
    for (int a = 0; a < 6; ++a) {
      const int kw = a >> 1;
      const bool single = (a & 1) != 0;
      // M = 128: TMEM lane = row (rows 0-63: kh = 0, rows 64-127: kh = 1).  M = 64: row r sits in lane (r % 16) +
      // 32 * (r / 16), i.e. the first 16 lanes of each warp's quarter.
      const int row = warp * 32 + lane;
      const int kh = single ? 2 : (row >> 6);
      const int ci = single ? warp * 16 + lane : (row & 63);
      const bool valid = !single || lane < 16;
      float *dst = base + ((size_t)(kh * 3 + kw) * 64 + ci) * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + (uint32_t)(a * 64 + c0), v);
        tmem_wait_ld();
        if (valid) {
          float4 *d4 = reinterpret_cast<float4 *>(dst + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            d4[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                __uint_as_float(v[4 * q + 3]));
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, WP_TMEM_COLS);
  }
}

// dW[co][cin_offset + chunk*64 + ci][kd][kh][kw] (+)= sum over the key's CTAs of partial[cta][kh*3+kw][ci][co]
__global__ void wgrad_planes_finish_kernel(const float *__restrict__ partial, float *__restrict__ dw, int keys, int chunks,
                                           int ctas_per_key, int cout, int cin_total, int cin_offset, int accumulate) {
  const long long total = (long long)keys * 9 * 64 * cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    long long r = i / cout;
    const int ci = (int)(r & 63);
    r >>= 6;
    const int tap9 = (int)(r % 9);
    const int key = (int)(r / 9);
    const int kd = key / chunks, chunk = key - kd * chunks;
    float s = 0.0f;
    const float *src = partial + (size_t)key * WP_PARTIAL_FLOATS + ((size_t)tap9 * 64 + ci) * 64 + co;
    const size_t pitch = (size_t)keys * WP_PARTIAL_FLOATS;
    int k = 0;
    for (; k + 4 <= ctas_per_key; k += 4) {  // fixed order, four loads in flight
      const float v0 = __ldg(src + (size_t)k * pitch), v1 = __ldg(src + (size_t)(k + 1) * pitch);
      const float v2 = __ldg(src + (size_t)(k + 2) * pitch), v3 = __ldg(src + (size_t)(k + 3) * pitch);
      s += v0;
      s += v1;
      s += v2;
      s += v3;
    }
    for (; k < ctas_per_key; ++k) s += __ldg(src + (size_t)k * pitch);
    float *o = dw + ((size_t)co * cin_total + cin_offset + chunk * 64 + ci) * 27 + kd * 9 + tap9;
    *o = accumulate ? *o + s : s;
  }
}

// dW[co][cin_offset + ci][tap] (+)= sum over slices of partial[s][(tap*chunks + ci/64)*64 + ci%64][co]
// A CTA owns (chunk, 8 input channels, 8 output channels): it sums the slices of the taps x 8 x 8 block reading
// 32-byte runs along cout, parks the block in shared memory and writes, per output channel, the [8 ci][taps] run that
// is contiguous in PyTorch's layout — both sides of the transpose move whole sectors.
static constexpr int WF_CI = 8, WF_CO = 8;
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float *__restrict__ partial, float *__restrict__ dw, int kslices,
                                                           int rows_total, int cout_pad, int cout, int chunks, int taps,
                                                           int cin_total, int cin_offset, int accumulate) {
  extern __shared__ float tile[];  // [taps * WF_CI][WF_CO + 1]
  const int co0 = blockIdx.x * WF_CO;
  const int cib = blockIdx.y;                       // 8-channel block of this source
  const int chunk = cib / (64 / WF_CI), c0 = (cib % (64 / WF_CI)) * WF_CI;
  const int n_rows = taps * WF_CI;
  for (int e = threadIdx.x; e < n_rows * WF_CO; e += blockDim.x) {
    const int j = e % WF_CO, r = e / WF_CO;         // r = tap * 32 + ci_local
    const int tap = r / WF_CI, cl = r - tap * WF_CI;
    const size_t row = (size_t)(tap * chunks + chunk) * 64 + c0 + cl;
    float s = 0.0f;
    if (co0 + j < cout) {
      // the slices are summed in a fixed order; four loads are in flight at a time (the pass is latency-bound)
      const float *src = partial + row * cout_pad + co0 + j;
      const size_t pitch = (size_t)rows_total * cout_pad;
      int k = 0;
      for (; k + 4 <= kslices; k += 4) {
        const float v0 = __ldg(src + (size_t)k * pitch), v1 = __ldg(src + (size_t)(k + 1) * pitch);
        const float v2 = __ldg(src + (size_t)(k + 2) * pitch), v3 = __ldg(src + (size_t)(k + 3) * pitch);
        s += v0;
        s += v1;
        s += v2;
        s += v3;
      }
      for (; k < kslices; ++k) s += __ldg(src + (size_t)k * pitch);
    }
    tile[r * (WF_CO + 1) + j] = s;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n_rows * WF_CO; e += blockDim.x) {
    const int j = e / n_rows, q = e - j * n_rows;   // q = ci_local * taps + tap: contiguous in dw for a fixed co
    if (co0 + j >= cout) continue;
    const int cl = q / taps, tap = q - cl * taps;
    const float v = tile[(tap * WF_CI + cl) * (WF_CO + 1) + j];
    float *o = dw + ((size_t)(co0 + j) * cin_total + cin_offset + chunk * 64 + c0) * taps + q;
    *o = accumulate ? *o + v : v;
  }
}

}  // namespace dram

using namespace dram;

struct dram_wgrad_plan {
  CUtensorMap map_x, map_dy;
  int group;   // unit pairs per work item of the streaming variant (1 or 2)
  int planes;  // 1 = PLANES variant (wp), 0 = streaming variant (p)
  WgParams p;
  WpParams wp;
  int block_n;
  float *dw;
  int cin_total, cin_offset, taps;
  int64_t flops;
};

static int wg_conv_out(int in, int k, int s, int d, int p) { return (in + 2 * p - d * (k - 1) - 1) / s + 1; }

// Fills the geometry part of the parameters (everything except pointers); returns 0 or an error.
static int wg_geometry(const dram_conv_desc *d, WgParams *p, int *block_n, int *group) {
  DRAM_REQUIRE(d, "conv3d_wgrad: null descriptor");
  DRAM_REQUIRE(d->n > 0 && d->di > 0 && d->hi > 0 && d->wi > 0, "conv3d_wgrad: bad input dims");
  DRAM_REQUIRE(d->c1 > 0 && d->c1 % 64 == 0 && d->c2 == 0,
               "conv3d_wgrad: c1 must be a multiple of 64 and c2 == 0 (a concatenated input is two calls with "
               "cin_offset), got c1 %d c2 %d", d->c1, d->c2);
  DRAM_REQUIRE(d->cout > 0 && d->cout % 32 == 0 && (d->cout <= 256 || d->cout % 256 == 0),
               "conv3d_wgrad: cout must be 32, 64, 128, 256 or a multiple of 256, got %d", d->cout);
  DRAM_REQUIRE(d->kd > 0 && d->kh > 0 && d->kw > 0 && d->sd > 0 && d->sh > 0 && d->sw > 0 && d->dd > 0 && d->dh > 0 &&
                   d->dw > 0 && d->pd >= 0 && d->ph >= 0 && d->pw >= 0,
               "conv3d_wgrad: bad filter geometry");
  DRAM_REQUIRE(d->dtype == DRAM_DTYPE_BF16 || d->dtype == DRAM_DTYPE_F16, "conv3d_wgrad: bad dtype");
  DRAM_REQUIRE(d->kd * d->kh * d->kw <= 42, "conv3d_wgrad: at most 42 taps (the finish pass keeps taps x 32 x 8 values in "
               "shared memory); the 7^3 stem goes through its (kh,kw)-unfolded 7x1x1 form");
  memset(p, 0, sizeof(*p));
  p->n = d->n; p->Di = d->di; p->Hi = d->hi; p->Wi = d->wi;
  p->Do = wg_conv_out(d->di, d->kd, d->sd, d->dd, d->pd);
  p->Ho = wg_conv_out(d->hi, d->kh, d->sh, d->dh, d->ph);
  p->Wo = wg_conv_out(d->wi, d->kw, d->sw, d->dw, d->pw);
  DRAM_REQUIRE(p->Do > 0 && p->Ho > 0 && p->Wo > 0, "conv3d_wgrad: empty output");
  p->kd = d->kd; p->kh = d->kh; p->kw = d->kw; p->sd = d->sd; p->sh = d->sh; p->sw = d->sw;
  p->dd = d->dd; p->dh = d->dh; p->dw = d->dw; p->pd = d->pd; p->ph = d->ph; p->pw = d->pw;
  p->tiles_w = ceil_div(p->Wo, WG_TW); p->tiles_h = ceil_div(p->Ho, WG_TH); p->tiles_d = ceil_div(p->Do, WG_TD);
  p->tiles_per_sample = p->tiles_w * p->tiles_h * p->tiles_d;
  const int64_t vt = (int64_t)p->tiles_per_sample * d->n;
  DRAM_REQUIRE(vt <= 0x7fffffffLL, "conv3d_wgrad: too many voxel tiles");
  p->total_vtiles = (int)vt;
  p->chunks = d->c1 / 64;
  p->units = d->kd * d->kh * d->kw * p->chunks;
  p->a_blocks = (p->units + 1) / 2;
  p->cout = d->cout;
  p->cout_pad = (d->cout + 63) / 64 * 64;
  *block_n = p->cout_pad < 256 ? p->cout_pad : 256;
  p->n_blocks = p->cout_pad / *block_n;
  p->rows_total = p->a_blocks * 128;
  p->is_f16 = d->dtype == DRAM_DTYPE_F16;
  // two unit pairs per item whenever there are that many (DRAM_B200_WGRAD_GROUP=1 keeps one, for A/B measurements)
  const char *knob = getenv("DRAM_B200_WGRAD_GROUP");
  *group = (p->a_blocks >= 2 && !(knob && atoi(knob) == 1)) ? 2 : 1;
  p->s_blocks = ceil_div(p->a_blocks, *group);
  // K slices: fill the SMs in whole waves; at least 4 voxel tiles per slice
  const int items = p->s_blocks * p->n_blocks, sms = sm_count() > 0 ? sm_count() : 148;
  int ks = 1;
  for (int w = 1; w <= 4; ++w) {
    const int cand = sms * w / items;
    if (cand >= 1 && (double)cand * items >= 0.9 * sms * w) {
      ks = cand;
      break;
    }
    if (cand >= 1) ks = cand;
  }
  const int max_ks = p->total_vtiles / 4 > 0 ? p->total_vtiles / 4 : 1;
  if (ks > max_ks) ks = max_ks;
  p->kslices = ks;
  p->items_total = items * ks;
  return DRAM_OK;
}

// PLANES variant: geometry it supports and its launch shape (keys <= #SMs so that every CTA owns one key).
static bool wp_supported(const dram_conv_desc *d) {
  if (d->algo == DRAM_CONV_ALGO_TILES) return false;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  return d->kd == 3 && d->kh == 3 && d->kw == 3 && d->sd == 1 && d->sh == 1 && d->sw == 1 && d->dd == 1 && d->dh == 1 &&
         d->dw == 1 && d->pd == 1 && d->ph == 1 && d->pw == 1 && d->cout <= 64 && 3 * (d->c1 / 64) <= sms;
}
static void wp_geometry(const dram_conv_desc *d, WpParams *w) {
  memset(w, 0, sizeof(*w));
  w->n = d->n; w->D = d->di; w->H = d->hi; w->W = d->wi;
  w->cols_w = ceil_div(d->wi, WP_W); w->cols_h = ceil_div(d->hi, WP_H);
  w->columns = d->n * w->cols_w * w->cols_h;
  w->chunks = d->c1 / 64;
  w->keys = 3 * w->chunks;
  w->upk = (long long)w->columns * d->di;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long per_key = sms / w->keys;
  if (per_key > w->upk) per_key = w->upk;
  w->ctas_per_key = (int)per_key;
  w->is_f16 = d->dtype == DRAM_DTYPE_F16;
}

// One place for the (BLOCK_N, G) instantiations: ctas == 0 sets the shared-memory attribute, otherwise launches.
template <int BN, int G>
static int wg_set_or_launch(const dram_wgrad_plan *pl, int ctas, cudaStream_t st) {
  if (ctas == 0)
    return check_cuda(cudaFuncSetAttribute(conv3d_wgrad_kernel<BN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           WgCfg<BN, G>::SMEM_BYTES),
                      "cudaFuncSetAttribute(conv3d_wgrad_kernel)");
  conv3d_wgrad_kernel<BN, G><<<ctas, WG_THREADS, WgCfg<BN, G>::SMEM_BYTES, st>>>(pl->map_x, pl->map_dy, pl->p);
  return DRAM_OK;
}
static int wg_dispatch(const dram_wgrad_plan *pl, int ctas, cudaStream_t st) {
  const int g = pl->group;
  switch (pl->block_n) {
    case 64: return g == 2 ? wg_set_or_launch<64, 2>(pl, ctas, st) : wg_set_or_launch<64, 1>(pl, ctas, st);
    case 128: return g == 2 ? wg_set_or_launch<128, 2>(pl, ctas, st) : wg_set_or_launch<128, 1>(pl, ctas, st);
    default: return g == 2 ? wg_set_or_launch<256, 2>(pl, ctas, st) : wg_set_or_launch<256, 1>(pl, ctas, st);
  }
}

extern "C" int64_t dram_conv3d_wgrad_workspace_bytes(const dram_conv_desc *d) {
  WgParams p;
  int bn, grp;
  if (wg_geometry(d, &p, &bn, &grp) != DRAM_OK) return -1;
  if (wp_supported(d)) {
    WpParams w;
    wp_geometry(d, &w);
    return (int64_t)w.keys * w.ctas_per_key * WP_PARTIAL_FLOATS * 4;
  }
  return (int64_t)p.kslices * p.rows_total * p.cout_pad * 4;
}

extern "C" int dram_conv3d_wgrad_plan_create(const dram_conv_desc *d, const void *x, const void *dy, float *dw,
                                             int32_t cin_total, int32_t cin_offset, void *workspace,
                                             int64_t workspace_bytes, dram_wgrad_plan **plan) {
  DRAM_REQUIRE(plan, "dram_conv3d_wgrad_plan_create: null plan pointer");
  *plan = nullptr;
  DRAM_REQUIRE(x && dy && dw && workspace, "dram_conv3d_wgrad_plan_create: x, dy, dw and workspace are required");
  WgParams p;
  int bn, grp;
  int rc = wg_geometry(d, &p, &bn, &grp);
  if (rc != DRAM_OK) return rc;
  DRAM_REQUIRE(cin_offset >= 0 && cin_offset + d->c1 <= cin_total,
               "dram_conv3d_wgrad_plan_create: channel range [%d, %d) outside cin_total %d", cin_offset,
               cin_offset + d->c1, cin_total);
  const int64_t need = dram_conv3d_wgrad_workspace_bytes(d);
  DRAM_REQUIRE(workspace_bytes >= need, "dram_conv3d_wgrad_plan_create: workspace of %lld bytes, %lld needed",
               (long long)workspace_bytes, (long long)need);
  dram_wgrad_plan *pl = new dram_wgrad_plan();
  memset(pl, 0, sizeof(*pl));
  pl->p = p;
  pl->p.partial = reinterpret_cast<float *>(workspace);
  pl->block_n = bn;
  pl->group = grp;
  pl->dw = dw;
  pl->cin_total = cin_total;
  pl->cin_offset = cin_offset;
  pl->taps = d->kd * d->kh * d->kw;
  pl->flops = 2LL * d->n * p.Do * p.Ho * p.Wo * (int64_t)d->cout * d->c1 * pl->taps;
  pl->planes = wp_supported(d) ? 1 : 0;
  if (pl->planes) {
    wp_geometry(d, &pl->wp);
    pl->wp.partial = reinterpret_cast<float *>(workspace);
    rc = encode_act_map(&pl->map_x, x, d->n, d->di, d->hi, d->wi, d->c1, 64, WP_PW, WP_PH, 1, 1, 1, 1, p.is_f16);
    if (rc == DRAM_OK)
      rc = encode_act_map(&pl->map_dy, dy, d->n, p.Do, p.Ho, p.Wo, d->cout, 64, WP_W, WP_H, 1, 1, 1, 1, p.is_f16);
    if (rc == DRAM_OK)
      rc = check_cuda(cudaFuncSetAttribute(conv3d_wgrad_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           WP_SMEM_BYTES),
                      "cudaFuncSetAttribute(conv3d_wgrad_planes_kernel)");
    if (rc != DRAM_OK) {
      delete pl;
      return rc;
    }
    *plan = pl;
    return DRAM_OK;
  }
  rc = encode_act_map(&pl->map_x, x, d->n, d->di, d->hi, d->wi, d->c1, 64, WG_TW, WG_TH, WG_TD, d->sw, d->sh, d->sd,
                      p.is_f16);
  if (rc == DRAM_OK)
    rc = encode_act_map(&pl->map_dy, dy, d->n, p.Do, p.Ho, p.Wo, d->cout, 64, WG_TW, WG_TH, WG_TD, 1, 1, 1, p.is_f16);
  if (rc == DRAM_OK) rc = wg_dispatch(pl, 0, nullptr);
  if (rc != DRAM_OK) {
    delete pl;
    return rc;
  }
  *plan = pl;
  return DRAM_OK;
}

extern "C" int dram_conv3d_wgrad_plan_destroy(dram_wgrad_plan *plan) {
  delete plan;
  return DRAM_OK;
}

extern "C" int dram_conv3d_wgrad_plan_info(const dram_wgrad_plan *plan, int64_t *flops, int32_t *items, int32_t *kslices,
                                           int32_t *block_n) {
  DRAM_REQUIRE(plan, "dram_conv3d_wgrad_plan_info: null plan");
  if (flops) *flops = plan->flops;
  if (items) *items = plan->planes ? plan->wp.keys * plan->wp.ctas_per_key : plan->p.items_total;
  if (kslices) *kslices = plan->planes ? plan->wp.ctas_per_key : plan->p.kslices;
  if (block_n) *block_n = plan->planes ? -64 : plan->block_n;  // negative: PLANES variant
  return DRAM_OK;
}

extern "C" int dram_conv3d_wgrad_run(const dram_wgrad_plan *plan, int32_t accumulate, int32_t max_ctas, void *stream) {
  DRAM_REQUIRE(plan, "dram_conv3d_wgrad_run: null plan");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (plan->planes) {  // the grid is part of the plan (one key per CTA): max_ctas does not apply
    const WpParams &w = plan->wp;
    conv3d_wgrad_planes_kernel<<<w.keys * w.ctas_per_key, WG_THREADS, WP_SMEM_BYTES, st>>>(plan->map_x, plan->map_dy, w);
    DRAM_CHECK_LAUNCH("conv3d_wgrad_planes_kernel launch");
    const long long total = (long long)w.keys * 9 * 64 * plan->p.cout;
    wgrad_planes_finish_kernel<<<stream_grid(total, 256), 256, 0, st>>>(w.partial, plan->dw, w.keys, w.chunks,
                                                                       w.ctas_per_key, plan->p.cout, plan->cin_total,
                                                                       plan->cin_offset, accumulate ? 1 : 0);
    DRAM_CHECK_LAUNCH("wgrad_planes_finish_kernel launch");
    return DRAM_OK;
  }
  int ctas = sm_count();
  if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
  if (plan->p.items_total < ctas) ctas = plan->p.items_total;
  int lrc = wg_dispatch(plan, ctas, st);
  if (lrc != DRAM_OK) return lrc;
  DRAM_CHECK_LAUNCH("conv3d_wgrad_kernel launch");
  const WgParams &p = plan->p;
  const dim3 fgrid(ceil_div(p.cout, WF_CO), p.chunks * (64 / WF_CI));
  const size_t fsmem = (size_t)plan->taps * WF_CI * (WF_CO + 1) * sizeof(float);
  wgrad_finish_kernel<<<fgrid, 256, fsmem, st>>>(p.partial, plan->dw, p.kslices, p.rows_total, p.cout_pad, p.cout, p.chunks,
                                                 plan->taps, plan->cin_total, plan->cin_offset, accumulate ? 1 : 0);
  DRAM_CHECK_LAUNCH("wgrad_finish_kernel launch");
  return DRAM_OK;
}
