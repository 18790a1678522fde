// tcgen05 implicit-GEMM convolution (stride 1, TF 'SAME' zero padding, 3x3 or 1x1), NHWC fp16 in,
// fp32 accumulate in TMEM, fused bias / ReLU / residual / channel-offset (zero-copy concat) /
// pixel-shuffle (conv2d_transpose 2x2 s2) epilogue.
//
// Replaces tf.layers.conv2d / conv2d_transpose call sites of the reference:
//   UNet.py:29-31,56-58  Tiramisu.py:35-37,50-52,62-64,77-79  Architecture.py:238-243
//   MultiScalePrediction.py:64-66,73-75,88-90
//
// Mapping onto the hardware
//   GEMM view     D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * Wt[tap][cout][cin]
//   UMMA          M = 128 consecutive pixels of ONE image row, N = round16(cout) (<= 256), K = 16
//   tile          R image rows x 128 pixels  ->  R accumulators of N fp32 columns in TMEM,
//                 two accumulator sets so the epilogue of tile i overlaps the MMAs of tile i+1
//   A operand     one TMA box per 64-channel chunk: (R+2) rows x 130 pixels x 64 ch ("halo slab",
//                 OOB -> 0 gives SAME padding).  The 9 taps are 9 *views* of the same slab: tap (r,s)
//                 for accumulator j starts at slab row (j+r), pixel s, i.e. a 128-byte-granular offset
//                 into a 128B-swizzled buffer -> every activation byte is fetched from L2 once per
//                 chunk instead of nine times.
//   B operand     weights pre-packed [tap][N][Cin64] fp16, one TMA box (64 ch x N) per (chunk, tap)
//   warps         0: A producer   1: B producer   2: MMA issuer   3: TMEM allocator
//                 4-7: epilogue (TMEM lane quarter = warp % 4)
//   grid          persistent, min(#tiles, #SMs) CTAs, static round-robin over tiles
#pragma once
#include "dd_ptx.cuh"

namespace dd {

constexpr int kConvTileW = 128;   // UMMA M: pixels per accumulator
constexpr int kConvCH = 64;       // channels per K chunk (one 128-byte swizzle row)
constexpr int kConvThreads = 256;

struct ConvTcParams {
  int N, H, W;            // input spatial dims (== output dims unless ups == 2)
  int Cin;                // multiple of 16
  int n_umma;             // UMMA N (multiple of 16, <= 256) = ngroups * group_c
  int acc_stride;         // TMEM columns between accumulators (n_umma rounded up to 32)
  int taps;               // 9 or 1
  uint32_t tap_mask;      // bit (r*3+s) set: tap is computed (conv2d_transpose 3x3 phases use tap subsets)
  int R;                  // rows per tile
  int strips, bands, num_tiles;
  int n_chunks;           // ceil(Cin / 64)
  int shift_mode;         // 0: halo slab, base_offset 0   1: halo slab, base_offset=(addr>>7)&7
                          // 2: three column-shifted slabs (all descriptors 1024B aligned)
  int a_stages, b_stages;
  uint32_t a_stage_bytes, b_stage_bytes, a_tx_bytes, b_tx_bytes;
  int a_box_w;            // pixels per slab row (130, or 128 for shift_mode 2 / 1x1)
  uint32_t tmem_cols;     // power of two >= 2 * R * acc_stride
  // epilogue
  int ngroups;            // 1, 2 or 4 column groups; group g -> sub-pixel (sp0 + g)
  int group_c;            // columns per group (UMMA columns), multiple of 16
  int cout_store;         // channels stored per group (multiple of 8, <= group_c)
  int ups, sp0;           // ups == 2: pixel shuffle, sub-pixel sp -> (ay, ax) = (sp >> 1, sp & 1)
  int OH, OW;
  const float* bias;      // [group_c] (shared by all groups) or nullptr
  int relu;               // ReLU on the primary output
  int out_f32;            // primary output dtype: 0 fp16, 1 fp32
  void* out;
  int out_cstride, out_coff;
  __half* out_relu;       // optional second output = relu(primary), fp16 (Tiramisu / compose nets)
  int out_relu_cstride, out_relu_coff;
  const __half* residual; // optional: added before activation (x + conv(...)), fp16, same dims as out
  int res_cstride, res_coff;
};

__device__ __forceinline__ void conv_tile_coords(const ConvTcParams& p, int tile, int& n, int& y0, int& x0) {
  const int strip = tile % p.strips;
  const int t2 = tile / p.strips;
  const int band = t2 % p.bands;
  n = t2 / p.bands;
  y0 = band * p.R;
  x0 = strip * kConvTileW;
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = a_smem + static_cast<size_t>(p.a_stages) * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + static_cast<size_t>(p.b_stages) * p.b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.a_stages;
  uint64_t* b_full = a_empty + p.a_stages;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* t_full = b_empty + p.b_stages;   // [2]
  uint64_t* t_empty = t_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_sgroups = (p.taps == 9 && p.shift_mode == 2) ? 3 : 1;
  const int s_per_group = (p.taps == 9) ? (3 / n_sgroups) : 1;
  const int n_r = (p.taps == 9) ? 3 : 1;
  const int halo = (p.taps == 9) ? 1 : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ A producer (activation slabs)
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n, y0, x0; conv_tile_coords(p, tile, n, y0, x0);
        for (int c = 0; c < p.n_chunks; ++c) {
          for (int g = 0; g < n_sgroups; ++g) {
            mbar_wait(&a_empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&a_full[stage], p.a_tx_bytes);
            const int xs = x0 - halo + ((n_sgroups == 3) ? g : 0);
            tma_load_4d(a_smem + static_cast<size_t>(stage) * p.a_stage_bytes, &tmA, &a_full[stage],
                        c * kConvCH, xs, y0 - halo, n);
            if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ B producer (weight tiles)
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int c = 0; c < p.n_chunks; ++c) {
          for (int g = 0; g < n_sgroups; ++g) {
            for (int si = 0; si < s_per_group; ++si) {
              const int s = (n_sgroups == 3) ? g : si;
              for (int r = 0; r < n_r; ++r) {
                const int tap = (p.taps == 9) ? (r * 3 + s) : 0;
                if (!((p.tap_mask >> tap) & 1u)) continue;
                mbar_wait(&b_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&b_full[stage], p.b_tx_bytes);
                tma_load_3d(b_smem + static_cast<size_t>(stage) * p.b_stage_bytes, &tmB, &b_full[stage],
                            c * kConvCH, 0, tap);
                if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(kConvTileW, p.n_umma);
      const uint32_t a_base = smem_u32(a_smem);
      const uint32_t b_base = smem_u32(b_smem);
      const uint64_t desc_tmpl = make_desc_sw128(0, 0);
      const uint32_t a_row_pitch = static_cast<uint32_t>(p.a_box_w) * 128u;
      int as = 0; uint32_t aphase = 0;
      int bs = 0; uint32_t bphase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int accset = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&t_empty[accset], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(accset * p.R * p.acc_stride);
        bool first = true;
        for (int c = 0; c < p.n_chunks; ++c) {
          int ksteps = (p.Cin - c * kConvCH) / 16;
          if (ksteps > 4) ksteps = 4;
          for (int g = 0; g < n_sgroups; ++g) {
            mbar_wait(&a_full[as], aphase);
            tc_fence_after();
            const uint32_t a_stage = a_base + static_cast<uint32_t>(as) * p.a_stage_bytes;
            for (int si = 0; si < s_per_group; ++si) {
              const int s_off = (n_sgroups == 3) ? 0 : si;
              for (int r = 0; r < n_r; ++r) {
                const int s_tap = (n_sgroups == 3) ? g : si;
                if (!((p.tap_mask >> ((p.taps == 9) ? (r * 3 + s_tap) : 0)) & 1u)) continue;
                mbar_wait(&b_full[bs], bphase);
                tc_fence_after();
                const uint32_t b_stage = b_base + static_cast<uint32_t>(bs) * p.b_stage_bytes;
                // descriptors differ only in the 14-bit (address >> 4) field: plain 64-bit adds
                const uint64_t bdesc0 = desc_tmpl + (b_stage >> 4);
                const uint32_t a_tap = a_stage + static_cast<uint32_t>((r * p.a_box_w + s_off) * 128);
                const uint32_t acc0 = first ? 0u : 1u;
                for (int j = 0; j < p.R; ++j) {
                  const uint64_t adesc0 = desc_tmpl + ((a_tap + static_cast<uint32_t>(j) * a_row_pitch) >> 4);
                  const uint32_t d_tmem = tmem_acc + static_cast<uint32_t>(j * p.acc_stride);
                  umma_f16(d_tmem, adesc0, bdesc0, idesc, acc0);
                  if (ksteps > 1) umma_f16(d_tmem, adesc0 + 2, bdesc0 + 2, idesc, 1u);
                  if (ksteps > 2) umma_f16(d_tmem, adesc0 + 4, bdesc0 + 4, idesc, 1u);
                  if (ksteps > 3) umma_f16(d_tmem, adesc0 + 6, bdesc0 + 6, idesc, 1u);
                }
                first = false;
                umma_commit(&b_empty[bs]);
                if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
              }
            }
            umma_commit(&a_empty[as]);
            if (++as == p.a_stages) { as = 0; aphase ^= 1; }
          }
        }
        umma_commit(&t_full[accset]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int accset = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int n, y0, x0; conv_tile_coords(p, tile, n, y0, x0);
      mbar_wait(&t_full[accset], acc_phase);
      tc_fence_after();
      const int x = x0 + q * 32 + lane;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(accset * p.R * p.acc_stride);
      for (int j = 0; j < p.R; ++j) {
        const int y = y0 + j;
        const bool in_img = (y < p.H) && (x < p.W);
        for (int g = 0; g < p.ngroups; ++g) {
          const int sp = p.sp0 + g;
          const int oy = (p.ups == 2) ? (2 * y + (sp >> 1)) : y;
          const int ox = (p.ups == 2) ? (2 * x + (sp & 1)) : x;
          const size_t opix = (static_cast<size_t>(n) * p.OH + oy) * p.OW + ox;
          for (int cb = 0; cb < p.group_c; cb += 16) {
            uint32_t v[16];
            __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
            tmem_ld_32x16(t_lane + static_cast<uint32_t>(j * p.acc_stride + g * p.group_c + cb), v);
            tmem_ld_wait();
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int ch = cb + h8 * 8;
              if (!in_img || ch >= p.cout_store) continue;
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[i] = __uint_as_float(v[h8 * 8 + i]);
                if (p.bias) f[i] += __ldg(p.bias + ch + i);
              }
              if (p.residual) {
                const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + opix * p.res_cstride + p.res_coff + ch);
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 r2 = __half22float2(rh[i]);
                  f[2 * i] += r2.x; f[2 * i + 1] += r2.y;
                }
              }
              if (p.out_relu) {
                uint4 pk; __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                for (int i = 0; i < 4; ++i) ph[i] = __floats2half2_rn(fmaxf(f[2 * i], 0.f), fmaxf(f[2 * i + 1], 0.f));
                *reinterpret_cast<uint4*>(p.out_relu + opix * p.out_relu_cstride + p.out_relu_coff + ch) = pk;
              }
              if (p.relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
              }
              if (p.out_f32) {
                float* o = reinterpret_cast<float*>(p.out) + opix * p.out_cstride + p.out_coff + ch;
                *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
              } else if (p.out) {
                uint4 pk; __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                for (int i = 0; i < 4; ++i) ph[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + opix * p.out_cstride + p.out_coff + ch) = pk;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[accset]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace dd
