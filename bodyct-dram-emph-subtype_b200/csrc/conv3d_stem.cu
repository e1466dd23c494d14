// K2, fused stem — conv1 = Conv3d(1, 64, k7, s2, p3) + BN + ReLU (med3d.py:296-304, 371-373) straight from
// the fp32 image: no unfolded copy in HBM (K2a + K1 wrote and re-read 537 MB per 256^3 volume).
//
// GEMM view per input plane z: K = 64 pseudo-channels (kh*8 + j), j <-> input column 2*ow - 4 + j, so the
// filter tap kw = j - 1 and the j = 0 / kh = 7 weights are zero.  Because (kh, j) only index the INPUT plane,
// the A operand of plane z is the same for every (output plane, kd) pair that touches it:
//   * the CTA owns a column of 8 (W) x 16 (H) output voxels and marches along D;
//   * producer warps unfold kw only: slot[r][ow] = 8 halves x[z][2*h0 - 3 + r][2*(w0 + ow) - 4 .. + 3]
//     (38 rows x 8 x 16 B = 4.75 KiB per input plane, 20-slot ring);
//   * kh is NOT unfolded: output row oh reads slot rows 2*oh + kh, so the A tile of a (kd, kh pair) MMA is a
//     no-swizzle K-major UMMA view of the slot: 8-row group stride (SBO) = two slot rows, K core-matrix
//     stride (LBO) = one slot row, start = slot + 2*pair rows.  Nothing is copied per tap;
//   * all 7 x 64 x 64 packed weights (56 KiB) stay resident in shared memory, grouped by the parity of kd and
//     stored kd-descending inside a group;
//   * four consecutive output planes accumulate in TMEM (4 x 64 columns, double buffered); each input plane
//     is unfolded once per column and used by the 3-4 output planes that touch it — with ONE MMA per kh pair:
//     plane j feeds output planes t through kd = j - 2t, their accumulators are adjacent TMEM columns and the
//     matching weight blocks are adjacent rows, so N = 64 x (number of output planes): 64 instead of 112 MMAs
//     per item and 43 % less A-operand traffic, which is what bounds an N = 64 MMA.
//
// Roles (416 threads): warps 0-3 epilogue, warp 4 MMA issuer (owns TMEM), warps 5-12 producers (one input
// plane each, round robin).
#include "conv_plan.h"

namespace dram {

static constexpr int ST_W = 8, ST_H = 16, ST_GROUP = 4;
static constexpr int ST_ROWS = 2 * ST_H + 6;                 // 37 rows read by kh < 7, +1 read by the zero tap
static constexpr int ST_PLANE_BYTES = ST_ROWS * ST_W * 16;   // 4864
static constexpr int ST_RING = 20;
static constexpr int ST_ITEM_PLANES = 2 * ST_GROUP + 5;      // 13 input planes feed 4 output planes
static constexpr int ST_KEEP = 5;                            // planes shared with the next group of the column
static constexpr int ST_W_EVEN_BYTES = 8 * 4 * 64 * 16;      // [kh][kd = 6,4,2,0][cout][8 halves] = 32768
static constexpr int ST_W_ODD_BYTES = 8 * 3 * 64 * 16;       // [kh][kd = 5,3,1][cout][8 halves] = 24576
static constexpr int ST_WEIGHT_BYTES = ST_W_EVEN_BYTES + ST_W_ODD_BYTES;  // 57344
#ifndef DRAM_STEM_PROD_WARPS
#define DRAM_STEM_PROD_WARPS 11
#endif
static constexpr int ST_MMA_WARP = 4, ST_PROD_WARP0 = 5, ST_PROD_WARPS = DRAM_STEM_PROD_WARPS;  // 5 + 11 = 16 warps: 512 threads x 128 registers = the whole register file
static constexpr int ST_THREADS = (ST_PROD_WARP0 + ST_PROD_WARPS) * 32;  // 512
static constexpr int ST_TMEM_COLS = 2 * ST_GROUP * 64;       // 512
static constexpr int ST_OUT_TILE_BYTES = 128 * 128;           // staged epilogue: 128 voxels x 64 channels, SWIZZLE_128B
static constexpr int ST_SMEM_BYTES = 1024 + ST_RING * ST_PLANE_BYTES + ST_WEIGHT_BYTES + 2 * ST_OUT_TILE_BYTES + 512 + 512;
static constexpr int ST_SLICE_H = 4;                          // rows of the 8 x 16 tile one epilogue warp owns (32 voxels)

struct StemParams {
  const float *x;       // fp32 [n][D][H][W] (kernel variant HU = false)
  const short *hu;      // int16 HU [n][D][H][W] (HU = true): window + standardise happen in the producers
  const float *lut;     // HU: fp32 [n][lut_size], lut[v][i] = standardised value of HU lut_lo + i for volume v
  int lut_lo, lut_size; // HU: window lower bound (integral) and number of table entries (hi - lo + 1)
  const uint4 *weight;  // 16-bit, even kd [8][4][64][8] then odd kd [8][3][64][8] (ops.pack_stem_weight_fused)
  int n, D, H, W;       // input dims
  int cols_w, cols_h, groups_d, items_total;
  EpiParams epi;
};

struct StemItem {
  int sample, w0, h0, q0, g;
};
__device__ __forceinline__ StemItem decode_stem_item(const StemParams &p, int item) {
  StemItem it;
  const int col = item / p.groups_d;
  it.g = item - col * p.groups_d;
  it.q0 = it.g * ST_GROUP;
  const int per_sample = p.cols_w * p.cols_h;
  it.sample = col / per_sample;
  const int r = col - it.sample * per_sample;
  const int ih = r / p.cols_w;
  it.w0 = (r - ih * p.cols_w) * ST_W;
  it.h0 = ih * ST_H;
  return it;
}

// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; lbo = byte distance of the second K core
// matrix, sbo = byte distance between 8-row groups (cute::UMMA canonical layout ((8,m),(T,2)):((1T,SBO),(1,LBO))).
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint32_t pack_pair(float a, float b, int is_f16) {
  if (is_f16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&h);
}

// HU variant: the window clamps every voxel to one of (hi - lo + 1) integer values (851 for the reference's
// [-1150, -300]), so K8's per-voxel arithmetic ((clamp(hu) - lo) / (hi - lo) - mean) / sd — two IEEE divisions —
// is tabulated once per volume by dram_window_lut and the producers only clamp (packed 16-bit min/max) and
// gather from a 3.4 KB table that lives in L1: exactly K8's values, none of its arithmetic in this kernel.
__device__ __forceinline__ uint32_t stem_lut_index2(uint32_t pair, uint32_t lo2, uint32_t hi2) {
  return __vsub2(__vmins2(__vmaxs2(pair, lo2), hi2), lo2);  // two signed 16-bit lanes: clamp(hu, lo, hi) - lo
}

template <bool HU>
__global__ void __launch_bounds__(ST_THREADS, 1)
conv3d_stem_kernel(const __grid_constant__ CUtensorMap map_out, const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base + ST_RING * ST_PLANE_BYTES;
  const uint32_t out_base = w_base + ST_WEIGHT_BYTES;   // 1 KiB aligned: 95 KiB of planes + 56 KiB of weights
  const uint32_t bar_base = out_base + 2 * ST_OUT_TILE_BYTES;
  auto plane_addr = [&](int s) { return smem_base + (uint32_t)s * ST_PLANE_BYTES; };
  auto plane_full = [&](int s) { return bar_base + 8u * s; };
  auto plane_empty = [&](int s) { return bar_base + 8u * (ST_RING + s); };
  auto tmem_full = [&](int a) { return bar_base + 8u * (2 * ST_RING + a); };
  auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * ST_RING + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * ST_RING + 4);
  const uint32_t sb_base = bar_base + 512u;  // fp32 scale[64] then shift[64] of the epilogue

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_RING; ++s) {
      mbar_init(plane_full(s), 32);
      mbar_init(plane_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full(a), 1);
      mbar_init(tmem_empty(a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == ST_MMA_WARP) tmem_alloc(tmem_slot, ST_TMEM_COLS);
  if (threadIdx.x == 0) prefetch_tensormap(&map_out);
  // resident weights: plain 16-byte copies, made visible to the tensor-core (async) proxy below
  for (int i = threadIdx.x; i < ST_WEIGHT_BYTES / 16; i += ST_THREADS) {
    const uint4 v = __ldg(p.weight + i);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(w_base + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
  }
  if (threadIdx.x < 64) {  // the epilogue's per-channel scale / shift: read from shared memory (broadcast), not per group from global
    const float sc = p.epi.scale != nullptr ? __ldg(p.epi.scale + threadIdx.x) : 1.0f;
    const float sh = __ldg(p.epi.bias + threadIdx.x);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb_base + 4u * threadIdx.x), "f"(sc) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb_base + 256u + 4u * threadIdx.x), "f"(sh) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const int item_begin = (int)(((long long)blockIdx.x * p.items_total) / gridDim.x);
  const int item_end = (int)(((long long)(blockIdx.x + 1) * p.items_total) / gridDim.x);

  if (warp >= ST_PROD_WARP0) {
    // ------------------------------- producers: unfold kw of one input plane per step -------------
    // Each producer warp owns every ST_PROD_WARPS-th plane of the stream, so that many planes' global
    // loads are in flight at once; a lane builds its chunks in two batches of five (20 8-byte loads
    // outstanding per lane).
    const int pw = warp - ST_PROD_WARP0;
    const bool vec2 = (p.W & 1) == 0 && ((reinterpret_cast<uintptr_t>(HU ? (const void *)p.hu : (const void *)p.x) & (HU ? 3 : 7)) == 0);
    const int is_f16 = p.epi.is_f16;
    // Round-2 finding (ncu: 60 M warp instructions per 256^3 volume at IPC 1.2, nothing else saturated): this kernel
    // is bound by instruction issue, and the producers' per-chunk index arithmetic was the largest part.  A lane's ten
    // chunks of a plane are (slot row r0 + 4k, column owl) for k = 0..9 with r0 = lane / 8 and owl = lane % 8 fixed,
    // so everything but the row step 4k*W is computed once per item: the lane's offset inside the plane, which of its
    // four 2-element pairs lie inside the volume (W even: a pair never straddles the edge) and which rows do.
    const int owl = lane & 7, r0 = lane >> 3;
    unsigned seq_end = 0;
    for (int item = item_begin; item < item_end; ++item) {
      const StemItem it = decode_stem_item(p, item);
      const bool reuse = item > item_begin && it.g > 0;
      const unsigned seq_base = seq_end - (reuse ? (unsigned)ST_KEEP : 0u);
      seq_end = seq_base + ST_ITEM_PLANES;
      const int ih_base = 2 * it.h0 - 3, iw_base = 2 * it.w0 - 4;
      const int ih_l = ih_base + r0, iw_l = iw_base + 2 * owl;   // this lane's first row / first column
      unsigned wmask = 0, kmask = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (iw_l + 2 * q >= 0 && iw_l + 2 * q + 1 < p.W) wmask |= 1u << q;
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        const int ih = ih_l + 4 * k;
        if (r0 + 4 * k < ST_ROWS && ih >= 0 && ih < p.H) kmask |= 1u << k;
      }
      const long long lane_off = (long long)ih_l * p.W + iw_l;   // may be negative at the low borders (never dereferenced there)
      const long long row_step = 4LL * p.W;
      for (int j = reuse ? ST_KEEP : 0; j < ST_ITEM_PLANES; ++j) {
        const unsigned seq = seq_base + j;
        if ((int)(seq % ST_PROD_WARPS) != pw) continue;
        const int slot = seq % ST_RING;
        const int z = 2 * it.q0 - 3 + j;
        mbar_wait(plane_empty(slot), ((seq / ST_RING) & 1u) ^ 1u);
        const bool zok = z >= 0 && z < p.D;
        const size_t plane_off = ((size_t)it.sample * p.D + (zok ? z : 0)) * p.H * (size_t)p.W;
        const uint32_t dst0 = plane_addr(slot) + (uint32_t)lane * 16u;
        if (vec2) {
          const unsigned km = zok ? kmask : 0u;
          if constexpr (HU) {
            const short *base = p.hu + plane_off;
            const float *lut = p.lut + (size_t)it.sample * p.lut_size;
            const uint32_t lo2 = ((uint32_t)(uint16_t)(short)p.lut_lo) * 0x10001u;
            const uint32_t hi2 = ((uint32_t)(uint16_t)(short)(p.lut_lo + p.lut_size - 1)) * 0x10001u;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t u[5][4];
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const int k = half * 5 + c;
                const uint32_t *ptr = reinterpret_cast<const uint32_t *>(base + (lane_off + row_step * k));
#pragma unroll
                for (int q = 0; q < 4; ++q) u[c][q] = ((km >> k) & (wmask >> q) & 1u) ? __ldg(ptr + q) : 0u;
              }
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const int k = half * 5 + c;
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  float a = 0.0f, b = 0.0f;
                  if ((km >> k) & (wmask >> q) & 1u) {
                    const uint32_t ix = stem_lut_index2(u[c][q], lo2, hi2);
                    a = __ldg(lut + (ix & 0xffffu));
                    b = __ldg(lut + (ix >> 16));
                  }
                  w[q] = pack_pair(a, b, is_f16);
                }
                if (r0 + 4 * k < ST_ROWS)
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + (uint32_t)k * 512u), "r"(w[0]), "r"(w[1]),
                               "r"(w[2]), "r"(w[3])
                               : "memory");
              }
            }
          } else {
            const float *base = p.x + plane_off;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float2 v[5][4];
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const int k = half * 5 + c;
                const float2 *ptr = reinterpret_cast<const float2 *>(base + (lane_off + row_step * k));
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  v[c][q] = ((km >> k) & (wmask >> q) & 1u) ? __ldg(ptr + q) : make_float2(0.0f, 0.0f);
              }
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const int k = half * 5 + c;
                if (r0 + 4 * k < ST_ROWS)
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + (uint32_t)k * 512u),
                               "r"(pack_pair(v[c][0].x, v[c][0].y, is_f16)), "r"(pack_pair(v[c][1].x, v[c][1].y, is_f16)),
                               "r"(pack_pair(v[c][2].x, v[c][2].y, is_f16)), "r"(pack_pair(v[c][3].x, v[c][3].y, is_f16))
                               : "memory");
              }
            }
          }
        } else {
          // generic path (odd W or an unaligned volume): element-wise bounds checks
          const float *xz = HU ? nullptr : p.x + plane_off;
          const short *hz = HU ? p.hu + plane_off : nullptr;
          const float *lut = HU ? p.lut + (size_t)it.sample * p.lut_size : nullptr;
          const uint32_t lo2 = ((uint32_t)(uint16_t)(short)p.lut_lo) * 0x10001u;
          const uint32_t hi2 = ((uint32_t)(uint16_t)(short)(p.lut_lo + p.lut_size - 1)) * 0x10001u;
#pragma unroll 1
          for (int k = 0; k < 10; ++k) {
            const int r = r0 + 4 * k;
            if (r >= ST_ROWS) break;
            const int ih = ih_base + r;
            float f[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.0f;
            if (zok && ih >= 0 && ih < p.H) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int iw = iw_l + q;
                if (iw >= 0 && iw < p.W) {
                  if constexpr (HU)
                    f[q] = __ldg(lut + (stem_lut_index2((uint32_t)(uint16_t)__ldg(hz + (size_t)ih * p.W + iw), lo2, hi2) & 0xffffu));
                  else
                    f[q] = __ldg(xz + (size_t)ih * p.W + iw);
                }
              }
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + (uint32_t)k * 512u),
                         "r"(pack_pair(f[0], f[1], is_f16)), "r"(pack_pair(f[2], f[3], is_f16)),
                         "r"(pack_pair(f[4], f[5], is_f16)), "r"(pack_pair(f[6], f[7], is_f16))
                         : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(plane_full(slot));
      }
    }
  } else if (warp == ST_MMA_WARP) {
    // One elected lane issues; everything that does not change per MMA is hoisted (see conv3d_slab.cu): the weights are
    // resident, so every B descriptor is a compile-time offset from one descriptor built here, and a plane's A
    // descriptor is built once — two 64-bit uniform adds per tcgen05.mma.
    uint32_t idesc[4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) idesc[nb] = make_idesc_16bit(128, (nb + 1) * 64, p.epi.is_f16);
    uint32_t wb = __shfl_sync(0xffffffffu, w_base, 0);
    const uint64_t db_even = make_nosw_desc(wb, 4u * 1024u, 128u);                    // kd even: K (kh) stride 4 KiB
    const uint64_t db_odd = make_nosw_desc(wb + ST_W_EVEN_BYTES, 3u * 1024u, 128u);   // kd odd: 3 KiB
    unsigned seq_end = 0;
    int buf = 0;
    uint32_t buf_phase = 0;
    for (int item = item_begin; item < item_end; ++item) {
      const StemItem it = decode_stem_item(p, item);
      const bool reuse = item > item_begin && it.g > 0;
      const bool next_reuse = (item + 1 < item_end) && (it.g + 1 < p.groups_d);
      const unsigned seq_base = seq_end - (reuse ? (unsigned)ST_KEEP : 0u);
      seq_end = seq_base + ST_ITEM_PLANES;
      mbar_wait(tmem_empty(buf), buf_phase ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d0 = tmem_base + (uint32_t)(buf * ST_GROUP * 64);
#pragma unroll
      for (int j = 0; j < ST_ITEM_PLANES; ++j) {
        const unsigned seq = seq_base + j;
        const int slot = seq % ST_RING;
        mbar_wait(plane_full(slot), (seq / ST_RING) & 1u);
        tcgen05_fence_after();
        uint32_t a_base = plane_addr(slot);
        a_base = __shfl_sync(0xffffffffu, a_base, 0);
        const uint64_t da_plane = make_nosw_desc(a_base, 128u, 256u);
        // Input plane j feeds output planes t = t_lo .. t_hi through kd = j - 2t (same parity as j, descending as
        // t ascends).  Their accumulators are adjacent TMEM columns and the weights of one parity are stored
        // kd-descending, so ONE MMA with N = 64 * (t_hi - t_lo + 1) covers them all: the A tile (the bound
        // at N = 64, see conv3d_slab.cu) is read once for up to four output planes.
        const int t_lo = j > 6 ? (j - 5) / 2 : 0, t_hi = j / 2 < ST_GROUP - 1 ? j / 2 : ST_GROUP - 1;
        const int nblk = t_hi - t_lo + 1, kd_hi = j - 2 * t_lo;
        const bool odd = (j & 1) != 0;
        const uint64_t db_region = odd ? db_odd : db_even;
        const uint32_t kh_stride16 = (odd ? 3u * 1024u : 4u * 1024u) >> 4;
        const uint32_t blk0 = (uint32_t)(((odd ? 5 : 6) - kd_hi) / 2);   // block index of kd_hi inside the region
        // the accumulator of output plane t is first touched by kd = 0, i.e. by the LAST block of an even plane
        const bool first_touch = !odd && (j / 2 <= ST_GROUP - 1);
        const uint32_t d0 = tmem_d0 + (uint32_t)(t_lo * 64);
        if (elect_one_sync()) {
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {  // kh pairs (0,1) (2,3) (4,5) (6,7): K = 16 per MMA
            const uint64_t da = da_plane + (uint64_t)(((uint32_t)(2 * pr) * 128u) >> 4);
            const uint64_t db = db_region + (uint64_t)((uint32_t)(2 * pr) * kh_stride16 + ((blk0 * 1024u) >> 4));
            // only the very first K step of a fresh accumulator block has to overwrite: that one MMA is split off,
            // the other three kh pairs run the full kd stack (the split cost 12 % of the item's MMA clocks)
            if (first_touch && pr == 0) {
              if (nblk > 1) umma_bf16(d0, da, db, idesc[nblk - 2], 1u);
              umma_bf16(d0 + (uint32_t)((nblk - 1) * 64), da, db + (uint64_t)(((uint32_t)(nblk - 1) * 1024u) >> 4), idesc[0], 0u);
            } else {
              umma_bf16(d0, da, db, idesc[nblk - 1], 1u);
            }
          }
          if (j < ST_ITEM_PLANES - ST_KEEP || !next_reuse) umma_commit(plane_empty(slot));
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(tmem_full(buf));
      __syncwarp();
      if (++buf == 2) {
        buf = 0;
        buf_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------- epilogue warps 0..3 -------------------------------
    // Staged through shared memory: a lane owns one voxel (row) of the 8 x 16 tile and writes its 64 channels into a
    // SWIZZLE_128B tile; full 128-byte lines leave through TMA stores (the tensor map clips voxels outside the volume).
    // The direct form — four 16-byte stores per lane and 32-channel group, each to a different 128-byte line — kept
    // the L1/LSU pipe 70 % busy with half-filled sectors (ncu, round 2) in a kernel that writes 268 MB per 256^3 volume.
    // Round 2, second capture (profiles/stem_r2v.md): the producers and the MMA warp WAIT (for free slots / a free
    // accumulator); the four epilogue warps are the critical path — 2.7 k clocks per output plane against 1.4 k of MMAs.
    // So: each warp stores its own 32-voxel slice (4 rows of the tile, its own bulk groups: no CTA-level barrier),
    // both 32-channel groups are read from tensor memory before the first is used, and scale / shift come from
    // shared memory instead of sixteen global loads per group.
    const int row = warp * 32 + lane;
    int buf = 0, ob = 0;
    uint32_t buf_phase = 0;
    const bool relu_in_cvt = p.epi.relu && p.epi.sat_count == nullptr;
    const uint32_t row_off = (uint32_t)row * 128u;
    for (int item = item_begin; item < item_end; ++item) {
      const StemItem it = decode_stem_item(p, item);
      const bool slice_ok = it.h0 + ST_SLICE_H * warp < p.epi.Ho;
      mbar_wait(tmem_full(buf), buf_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int t = 0; t < ST_GROUP; ++t) {
        const int od = it.q0 + t;
        const uint32_t taddr = tmem_base + (uint32_t)((buf * ST_GROUP + t) * 64) + ((uint32_t)(warp * 32) << 16);
        const uint32_t tile = out_base + (uint32_t)ob * ST_OUT_TILE_BYTES;
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(taddr, v[0]);
        tmem_ld_32x32b_x32(taddr + 32u, v[1]);
        // the store that last read this slice (two planes ago) must be done with shared memory
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        tmem_wait_ld();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 sc, sh;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w)
                         : "r"(sb_base + (uint32_t)(half * 128 + j * 16)));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(sh.x), "=f"(sh.y), "=f"(sh.z), "=f"(sh.w)
                         : "r"(sb_base + 256u + (uint32_t)(half * 128 + j * 16)));
            y[4 * j + 0] = fmaf(__uint_as_float(v[half][4 * j + 0]), sc.x, sh.x);
            y[4 * j + 1] = fmaf(__uint_as_float(v[half][4 * j + 1]), sc.y, sh.y);
            y[4 * j + 2] = fmaf(__uint_as_float(v[half][4 * j + 2]), sc.z, sh.z);
            y[4 * j + 3] = fmaf(__uint_as_float(v[half][4 * j + 3]), sc.w, sh.w);
          }
          if (p.epi.relu && !relu_in_cvt) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.0f);
          }
          if (p.epi.sat_count != nullptr) note_saturation(p.epi, y);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              w[q] = relu_in_cvt ? pack2_relu(y[8 * j + 2 * q], y[8 * j + 2 * q + 1], p.epi.is_f16)
                                 : pack2(y[8 * j + 2 * q], y[8 * j + 2 * q + 1], p.epi.is_f16);
            const uint32_t chunk = (uint32_t)(((half * 4 + j) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + row_off + chunk), "r"(w[0]), "r"(w[1]),
                         "r"(w[2]), "r"(w[3])
                         : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (od < p.epi.Do && slice_ok)
            tma_store_5d(&map_out, tile + (uint32_t)warp * (32u * 128u), 0, it.w0, it.h0 + ST_SLICE_H * warp, od, it.sample);
          tma_store_commit();  // also when nothing was stored: wait_group.read<1> above counts one group per plane
        }
        ob ^= 1;
      }
      tcgen05_fence_before();
      mbar_arrive(tmem_empty(buf));
      if (++buf == 2) {
        buf = 0;
        buf_phase ^= 1u;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == ST_MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, ST_TMEM_COLS);
  }
}

}  // namespace dram

using namespace dram;

extern "C" size_t dram_stem_weight_bytes(void) { return ST_WEIGHT_BYTES; }

static int stem_launch(StemParams &p, const void *weight, const float *bias, const float *scale, void *out, int32_t n,
                       int32_t d, int32_t h, int32_t w, int32_t relu, int32_t dtype, int32_t max_ctas, void *stream,
                       bool from_hu) {
  DRAM_REQUIRE(weight && bias && out, "dram_stem_conv7: null pointer");
  DRAM_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, "dram_stem_conv7: empty volume");
  DRAM_REQUIRE(dtype == DRAM_DTYPE_BF16 || dtype == DRAM_DTYPE_F16, "dram_stem_conv7: dtype must be bf16 (0) or fp16 (1)");
  DRAM_REQUIRE((reinterpret_cast<uintptr_t>(weight) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "dram_stem_conv7: weight and out must be 16-byte aligned");
  p.weight = reinterpret_cast<const uint4 *>(weight);
  p.n = n; p.D = d; p.H = h; p.W = w;
  const int Do = (d - 1) / 2 + 1, Ho = (h - 1) / 2 + 1, Wo = (w - 1) / 2 + 1;
  p.cols_w = ceil_div(Wo, ST_W);
  p.cols_h = ceil_div(Ho, ST_H);
  p.groups_d = ceil_div(Do, ST_GROUP);
  const int64_t items = (int64_t)n * p.cols_w * p.cols_h * p.groups_d;
  DRAM_REQUIRE(items <= 0x7fffffffLL, "dram_stem_conv7: too many work items");
  p.items_total = (int)items;
  p.epi.Do = Do; p.epi.Ho = Ho; p.epi.Wo = Wo;
  p.epi.cout = 64;
  p.epi.relu = relu;
  p.epi.is_f16 = dtype == DRAM_DTYPE_F16;
  p.epi.bias = bias;
  p.epi.scale = scale;
  p.epi.out = reinterpret_cast<uint16_t *>(out);
  p.epi.res_stride = 1;
  p.epi.store_out = 1;
  p.epi = with_sat_counter(p.epi);
  int rc = from_hu ? check_cuda(cudaFuncSetAttribute(conv3d_stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ST_SMEM_BYTES), "cudaFuncSetAttribute(conv3d_stem_kernel)")
                   : check_cuda(cudaFuncSetAttribute(conv3d_stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ST_SMEM_BYTES), "cudaFuncSetAttribute(conv3d_stem_kernel)");
  if (rc != DRAM_OK) return rc;
  CUtensorMap map_out;  // one epilogue warp's slice of an output plane: 64 channels x 8 (W) x 4 (H) voxels, SWIZZLE_128B rows
  rc = encode_act_map(&map_out, out, n, Do, Ho, Wo, 64, 64, ST_W, ST_SLICE_H, 1, 1, 1, 1, p.epi.is_f16);
  if (rc != DRAM_OK) return rc;
  int ctas = sm_count();
  if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
  if (p.items_total < ctas) ctas = p.items_total;
  if (from_hu)
    conv3d_stem_kernel<true><<<ctas, ST_THREADS, ST_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(map_out, p);
  else
    conv3d_stem_kernel<false><<<ctas, ST_THREADS, ST_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(map_out, p);
  DRAM_CHECK_LAUNCH("conv3d_stem_kernel launch");
  return DRAM_OK;
}

extern "C" int dram_stem_conv7(const float *x, const void *weight, const float *bias, const float *scale,
                               void *out, int32_t n, int32_t d, int32_t h, int32_t w, int32_t relu,
                               int32_t dtype, int32_t max_ctas, void *stream) {
  DRAM_REQUIRE(x, "dram_stem_conv7: null pointer");
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  return stem_launch(p, weight, bias, scale, out, n, d, h, w, relu, dtype, max_ctas, stream, false);
}

extern "C" int dram_stem_conv7_hu(const int16_t *hu, const float *lut, int32_t lut_lo, int32_t lut_size,
                                  const void *weight, const float *bias, const float *scale, void *out, int32_t n,
                                  int32_t d, int32_t h, int32_t w, int32_t relu, int32_t dtype, int32_t max_ctas,
                                  void *stream) {
  DRAM_REQUIRE(hu && lut, "dram_stem_conv7_hu: null pointer");
  DRAM_REQUIRE(lut_size >= 2 && lut_size <= DRAM_WINDOW_LUT_MAX && lut_lo >= -32768 && lut_lo + lut_size - 1 <= 32767,
               "dram_stem_conv7_hu: bad table range [%d, %d]", lut_lo, lut_lo + lut_size - 1);
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.hu = reinterpret_cast<const short *>(hu);
  p.lut = lut;
  p.lut_lo = lut_lo;
  p.lut_size = lut_size;
  return stem_launch(p, weight, bias, scale, out, n, d, h, w, relu, dtype, max_ctas, stream, true);
}
