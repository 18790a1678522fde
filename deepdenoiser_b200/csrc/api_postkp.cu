// C-ABI: fused output head of one scale - AdjustNumberOfChannels (Architecture.py:230-244: conv1x1 C->O + ReLU,
// conv1x1 O->O) + tf.split per feature (Architecture.py:581-587) + KernelPrediction.kernel_prediction
// (KernelPrediction.py:11-63: softmax over K*K logits, symmetric-padded KxK weighted sum of the noisy source) in ONE
// kernel: the O = T*K*K logits of a pixel never reach HBM (unfused: C*2 read + O*2 write, O*2 read + O*4 write, O*4 read
// per pixel; fused: C*2 + 12 + 12 bytes).
//
//   tile        8 rows x 16 pixels; warp w owns row w: M = 16 pixels of an mma.sync m16n8k16 (fp16 in, fp32 accumulate)
//   GEMM 1      [16 x C] x [C x O16]: A by ldmatrix from the staged activation tile (row stride C*2+16 B: conflict free),
//               B by ldmatrix from W1^T [O16][C+8] in shared memory
//   GEMM 2      [16 x O16] x [O16 x O16]: A = relu(acc1 + b1) repacked from the accumulator fragments (the m16n8 C layout
//               of two adjacent n-tiles IS the m16k16 A layout), B from W2^T
//   filter      a quad of threads holds the 32 logits of a pixel (8 each): softmax max / sum and the weighted RGB sums are
//               reduced with two shuffles; sources come from a symmetric-padded halo tile in shared memory
//   grid        persistent: CTAs stride over the (image, tile) list, the weights are loaded once per CTA
// This is memory bound (AI ~ 30 FLOP/B): mma.sync keeps the arithmetic far below the HBM time, tcgen05 would add nothing.
#include <math.h>
#include <string.h>

#include "dd_internal.h"
#include "dd_ptx.cuh"

namespace dd {

constexpr int kPkTileH = 8, kPkTileW = 16, kPkThreads = 256;

struct PostKpParams {
  const __half* x; int xcs, xoff, C, Cpad;       // activations [B,h,w,C] fp16
  const __half* w1t; const __half* w2t;          // W1^T [O16][Cpad], W2^T [O16][O16] fp16 (zero padded)
  const float* b1; const float* b2;              // [O16] fp32 (zero padded)
  View src, out;                                  // fp32 rgb banks
  int B, h, w, ipt;
  int tiles_x, tiles_y; int total_tiles;
  int cp_shift;                                   // log2 of the power of two >= Cpad / 8 (activation-tile load mapping)
};

__device__ __forceinline__ void pk_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void pk_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk_pack_relu(float a, float b) {
  const __half2 h = __floats2half2_rn(fmaxf(a, 0.f), fmaxf(b, 0.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ int pk_sym(int i, int n) {
  while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - i - 1);
  return i;
}

__device__ __forceinline__ void pk_cp_async_16(uint32_t smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void pk_cp_async_4(uint32_t smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void pk_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pk_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float pk_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int K, int T>
__global__ void __launch_bounds__(kPkThreads) post_kp_fused_kernel(const PostKpParams p) {
  // every feature owns KP = K*K rounded up to 16 logit columns (NTF n-tiles); padded columns carry a bias of -inf
  constexpr int K2 = K * K, KP = (K2 + 15) / 16 * 16, O16 = T * KP, NT = O16 / 8, NTF = KP / 8, PAD = (K - 1) / 2;
  constexpr int TWs = kPkTileW + 2 * PAD, THs = kPkTileH + 2 * PAD;
  constexpr int SRC_ELEMS = T * THs * TWs;
  extern __shared__ __align__(16) uint8_t pk_smem[];
  const int xrow = p.Cpad * 2 + 16;                       // bytes per staged pixel (odd multiple of 16: conflict-free ldmatrix)
  const int w1row = p.Cpad * 2 + 16, w2row = O16 * 2 + 16;
  const int x_bytes = 128 * xrow;
  uint8_t* xs = pk_smem;                                  // [2][128][xrow]
  uint8_t* w1s = xs + 2 * x_bytes;                        // [O16][w1row]
  uint8_t* w2s = w1s + O16 * w1row;                       // [O16][w2row]
  float* b1s = reinterpret_cast<float*>(w2s + O16 * w2row);
  float* b2s = b1s + O16;
  float4* s_src = reinterpret_cast<float4*>(b2s + O16);   // [2][T][THs][TWs]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int cchunks = p.Cpad / 8;

  // weights: once per CTA
  for (int i = tid; i < O16 * cchunks; i += kPkThreads) {
    const int n = i / cchunks, c = i % cchunks;
    *reinterpret_cast<uint4*>(w1s + n * w1row + c * 16) = *reinterpret_cast<const uint4*>(p.w1t + static_cast<size_t>(n) * p.Cpad + c * 8);
  }
  for (int i = tid; i < O16 * (O16 / 8); i += kPkThreads) {
    const int n = i / (O16 / 8), c = i % (O16 / 8);
    *reinterpret_cast<uint4*>(w2s + n * w2row + c * 16) = *reinterpret_cast<const uint4*>(p.w2t + static_cast<size_t>(n) * O16 + c * 8);
  }
  for (int i = tid; i < O16; i += kPkThreads) { b1s[i] = p.b1[i]; b2s[i] = p.b2[i] * 1.4426950408889634f; }   // logits in log2 units

  const uint32_t xs_u = smem_u32(xs), w1_u = smem_u32(w1s), w2_u = smem_u32(w2s), src_u = smem_u32(s_src);
  // ldmatrix lane addressing.  A (x4): matrices (rows 0-7,k 0-7), (rows 8-15,k 0-7), (rows 0-7,k 8-15), (rows 8-15,k 8-15)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kof = (lane >> 4) * 16;
  // B (x4): matrices (n-tile j,k 0-7), (n-tile j,k 8-15), (n-tile j+1,k 0-7), (n-tile j+1,k 8-15); rows = n
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_kof = ((lane >> 3) & 1) * 16;
  const int valid_chunks = p.C / 8;
  // slot i = 2j + e of a feature (column 8j + 2*t4 + e of its KP) holds tap k = 4i + t4 (the host packs W2 / b2 that way):
  // the four threads of a quad read CONSECUTIVE taps of the source tile; offsets are fixed for the whole kernel
  int tap_off[NTF * 2];
#pragma unroll
  for (int i = 0; i < NTF * 2; ++i) {
    const int k = 4 * i + t4;
    tap_off[i] = (k < K2) ? (k / K) * TWs + (k % K) : 0;
  }
  // activation-tile load mapping: chunk = tid % CP, pixel = tid / CP (+ 256/CP per step), CP = power of two >= cchunks
  const int cp_shift = p.cp_shift, cp_mask = (1 << cp_shift) - 1, px_step = kPkThreads >> cp_shift;
  const int ld_c = tid & cp_mask, ld_px0 = tid >> cp_shift;

  // Tile coordinates advance by gridDim.x tiles per step WITHOUT divisions: (db, dty, dtx) is gridDim.x decomposed once.
  struct Coord { int b, ty, tx; };
  const int tiles_per_image = p.tiles_x * p.tiles_y;
  Coord delta;
  delta.b = static_cast<int>(gridDim.x) / tiles_per_image;
  { const int rem = static_cast<int>(gridDim.x) - delta.b * tiles_per_image; delta.ty = rem / p.tiles_x; delta.tx = rem - delta.ty * p.tiles_x; }
  auto advance = [&](Coord& c) {
    c.tx += delta.tx; if (c.tx >= p.tiles_x) { c.tx -= p.tiles_x; ++c.ty; }
    c.ty += delta.ty; if (c.ty >= p.tiles_y) { c.ty -= p.tiles_y; ++c.b; }
    c.b += delta.b;
  };
  auto first_image = [&](int b) {                      // bank image of feature 0 of logits image b; feature f: + f * ipt
    if (p.ipt == 1) return b * T;
    const int tup = b / p.ipt;
    return tup * T * p.ipt + (b - tup * p.ipt);
  };
  // source halo tile: element i = tid + it * 256 -> (feature, ly, lx), fixed for the whole kernel
  constexpr int SRC_ITERS = (SRC_ELEMS + kPkThreads - 1) / kPkThreads;
  int src_f[SRC_ITERS], src_ly[SRC_ITERS], src_lx[SRC_ITERS];
#pragma unroll
  for (int it = 0; it < SRC_ITERS; ++it) {
    const int i = tid + it * kPkThreads;
    const int f = i / (THs * TWs), r = i - f * (THs * TWs);
    src_f[it] = (i < SRC_ELEMS) ? f : -1; src_ly[it] = r / TWs - PAD; src_lx[it] = r % TWs - PAD;
  }

  // interior tiles (no clipping, no symmetric padding) take a path without per-element index arithmetic: all per-thread offsets
  // are fixed for the whole kernel
  const int im_h = p.h, im_w = p.w, x_cs = p.xcs;
  const int w_xcs = im_w * x_cs;
  const bool rows_step = (px_step & 15) == 0;                     // a thread keeps its tile column from load to load
  const int thr_x_off = (ld_px0 >> 4) * w_xcs + (ld_px0 & 15) * x_cs + p.xoff + ld_c * 8;
  const int step_x_off = (px_step >> 4) * w_xcs;
  const int src_cs = p.src.cstride;
  int src_off[SRC_ITERS];
#pragma unroll
  for (int it = 0; it < SRC_ITERS; ++it) src_off[it] = (src_ly[it] * im_w + src_lx[it]) * src_cs + p.src.coff;
  const float* src_base = reinterpret_cast<const float*>(p.src.ptr);

  auto issue_loads = [&](const Coord& c, int buf) {
    const int y0 = c.ty * kPkTileH, x0 = c.tx * kPkTileW;
    const bool inside = (y0 + kPkTileH <= im_h) && (x0 + kPkTileW <= im_w);
    if (ld_c < cchunks && inside && rows_step && ld_c < valid_chunks) {
      uint32_t dst = xs_u + static_cast<uint32_t>(buf * x_bytes + ld_c * 16 + ld_px0 * xrow);
      const uint32_t dst_step = static_cast<uint32_t>(px_step * xrow);
      const __half* src = p.x + ((static_cast<size_t>(c.b) * im_h + y0) * im_w + x0) * x_cs + thr_x_off;
      for (int px = ld_px0; px < 128; px += px_step) {
        pk_cp_async_16(dst, src);
        dst += dst_step; src += step_x_off;
      }
    } else if (ld_c < cchunks) {
      uint32_t dst = xs_u + static_cast<uint32_t>(buf * x_bytes + ld_c * 16 + ld_px0 * xrow);
      const uint32_t dst_step = static_cast<uint32_t>(px_step * xrow);
      const __half* tile_base = p.x + ((static_cast<size_t>(c.b) * p.h + y0) * p.w + x0) * p.xcs + p.xoff + ld_c * 8;
      const bool chunk_ok = ld_c < valid_chunks;
      for (int px = ld_px0; px < 128; px += px_step) {
        const int r = px >> 4, xx = px & 15;
        if (chunk_ok && y0 + r < p.h && x0 + xx < p.w) {
          pk_cp_async_16(dst, tile_base + r * w_xcs + xx * p.xcs);
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
        }
        dst += dst_step;
      }
    }
    // source halo tiles (symmetric padding, KernelPrediction.py:30 via Conv2dUtilities.pad_equally)
    const int img0 = first_image(c.b);
    if (y0 >= PAD && x0 >= PAD && y0 + kPkTileH + PAD <= im_h && x0 + kPkTileW + PAD <= im_w) {
#pragma unroll
      for (int it = 0; it < SRC_ITERS; ++it) {
        if (src_f[it] >= 0) {
          const float* sp = src_base + ((static_cast<size_t>(img0 + src_f[it] * p.ipt) * im_h + y0) * im_w + x0) * src_cs + src_off[it];
          const uint32_t dst = src_u + static_cast<uint32_t>((buf * SRC_ELEMS + tid + it * kPkThreads) * 16);
          pk_cp_async_4(dst, sp); pk_cp_async_4(dst + 4, sp + 1); pk_cp_async_4(dst + 8, sp + 2);
        }
      }
      pk_cp_async_commit();
      return;
    }
#pragma unroll
    for (int it = 0; it < SRC_ITERS; ++it) {
      if (src_f[it] >= 0) {
        const size_t sp = p.src.pix(img0 + src_f[it] * p.ipt, pk_sym(y0 + src_ly[it], p.h), pk_sym(x0 + src_lx[it], p.w));
        const float* s = reinterpret_cast<const float*>(p.src.ptr) + sp * p.src.cstride + p.src.coff;
        const uint32_t dst = src_u + static_cast<uint32_t>((buf * SRC_ELEMS + tid + it * kPkThreads) * 16);
        pk_cp_async_4(dst, s); pk_cp_async_4(dst + 4, s + 1); pk_cp_async_4(dst + 8, s + 2);
      }
    }
    pk_cp_async_commit();
  };

  int tile = blockIdx.x;
  int buf = 0;
  Coord cur;
  cur.b = tile / tiles_per_image;
  { const int rem = tile - cur.b * tiles_per_image; cur.ty = rem / p.tiles_x; cur.tx = rem - cur.ty * p.tiles_x; }
  if (tile < p.total_tiles) issue_loads(cur, 0);
  for (; tile < p.total_tiles; tile += gridDim.x, buf ^= 1) {
    const int b = cur.b;
    const int y0 = cur.ty * kPkTileH, x0 = cur.tx * kPkTileW;
    const int img0 = first_image(b);
    advance(cur);                                          // cur = the NEXT tile of this CTA from here on
    pk_cp_async_wait_all();
    __syncthreads();                                       // this tile's data (and the weights) visible; the other buffer is free
    if (tile + static_cast<int>(gridDim.x) < p.total_tiles) issue_loads(cur, buf ^ 1);   // in flight during the math below

    // ---------------------------------------------------------------- GEMM 1: hidden = relu(x W1 + b1)
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float bl = b1s[j * 8 + 2 * t4], bh = b1s[j * 8 + 2 * t4 + 1];
      acc[j][0] = bl; acc[j][1] = bh; acc[j][2] = bl; acc[j][3] = bh;
    }
    const uint32_t a_base = xs_u + static_cast<uint32_t>(buf * x_bytes + (warp * 16 + a_row) * xrow + a_kof);
    for (int ks = 0; ks < p.Cpad / 16; ++ks) {
      uint32_t a[4];
      pk_ldmatrix_x4(a, a_base + ks * 32);
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        uint32_t bf[4];
        pk_ldmatrix_x4(bf, w1_u + static_cast<uint32_t>((j * 8 + b_row) * w1row + ks * 32 + b_kof));
        pk_mma(acc[j], a, bf[0], bf[1]);
        pk_mma(acc[j + 1], a, bf[2], bf[3]);
      }
    }
    // ---------------------------------------------------------------- GEMM 2: logits = hidden W2 + b2 (then * log2 e)
    float lg[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) lg[j][e] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      uint32_t a[4];
      a[0] = pk_pack_relu(acc[2 * kk][0], acc[2 * kk][1]);
      a[1] = pk_pack_relu(acc[2 * kk][2], acc[2 * kk][3]);
      a[2] = pk_pack_relu(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
      a[3] = pk_pack_relu(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        uint32_t bf[4];
        pk_ldmatrix_x4(bf, w2_u + static_cast<uint32_t>((j * 8 + b_row) * w2row + kk * 32 + b_kof));
        pk_mma(lg[j], a, bf[0], bf[1]);
        pk_mma(lg[j + 1], a, bf[2], bf[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float bl = b2s[j * 8 + 2 * t4], bh = b2s[j * 8 + 2 * t4 + 1];      // already scaled by log2 e
      lg[j][0] = fmaf(lg[j][0], 1.4426950408889634f, bl); lg[j][1] = fmaf(lg[j][1], 1.4426950408889634f, bh);
      lg[j][2] = fmaf(lg[j][2], 1.4426950408889634f, bl); lg[j][3] = fmaf(lg[j][3], 1.4426950408889634f, bh);
    }
    // ---------------------------------------------------------------- softmax over K*K + filter apply, per feature
    // this thread: pixels (row `warp`, column g) [regs 0,1] and (row `warp`, column g + 8) [regs 2,3]; branch free: the
    // padded slots hold -inf (their bias), so they contribute exp2(-inf) = 0
    const float4* src_cur = s_src + buf * SRC_ELEMS;
#pragma unroll
    for (int f = 0; f < T; ++f) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int lx = g + half * 8;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NTF * 2; ++i) mx = fmaxf(mx, lg[f * NTF + (i >> 1)][half * 2 + (i & 1)]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f, r = 0.f, gg = 0.f, bb = 0.f;
        const float4* tile_f = src_cur + f * (THs * TWs) + warp * TWs + lx;
#pragma unroll
        for (int i = 0; i < NTF * 2; ++i) {
          const float ev = pk_ex2(lg[f * NTF + (i >> 1)][half * 2 + (i & 1)] - mx);
          const float4 sv = tile_f[tap_off[i]];
          sum += ev;
          r = fmaf(ev, sv.x, r); gg = fmaf(ev, sv.y, gg); bb = fmaf(ev, sv.z, bb);
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          r += __shfl_xor_sync(0xffffffffu, r, o);
          gg += __shfl_xor_sync(0xffffffffu, gg, o);
          bb += __shfl_xor_sync(0xffffffffu, bb, o);
        }
        const int y = y0 + warp, x = x0 + lx;
        if (t4 == half && y < p.h && x < p.w) {
          const float inv = 1.f / sum;
          float* o = reinterpret_cast<float*>(p.out.ptr) + p.out.pix(img0 + f * p.ipt, y, x) * p.out.cstride + p.out.coff;
          o[0] = r * inv; o[1] = gg * inv; o[2] = bb * inv;
        }
      }
    }
  }
  pk_cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------------------
// One feature per tuple (SINGLE tuples, the benchmarked configuration): same GEMMs, but the softmax / filter stage runs with ONE
// THREAD PER PIXEL.  A warp owns two tile rows (M = 32 as two m16 blocks: the weight fragments are shared), stages its 32 x K*K
// logits through the x rows it has just consumed ([column][32 pixels + 1] floats: conflict free both ways) and every lane then
// walks the K*K taps of its own pixel with compile-time offsets - no quad shuffles, no padded tap slots, no per-tap address
// arithmetic.  Measured on the quad form (ncu source counters, profiles/r02_ncu_postkp_compose_full.csv): 41.6 warp-instructions
// per pixel, 18 of them in the softmax / filter stage and 14 in the tile loads.
constexpr int kPxThreads = 128, kPxLgPitch = 33;

template <int K>
__global__ void __launch_bounds__(kPxThreads) post_kp_pixel_kernel(const PostKpParams p) {
  constexpr int K2 = K * K, O16 = (K2 + 15) / 16 * 16, NT = O16 / 8, PAD = (K - 1) / 2;
  constexpr int TWs = kPkTileW + 2 * PAD, THs = kPkTileH + 2 * PAD, SRC_ELEMS = THs * TWs;
  extern __shared__ __align__(16) uint8_t pk_smem[];
  const int xrow = p.Cpad * 2 + 16;                       // bytes per staged pixel (odd multiple of 16: conflict-free ldmatrix)
  const int w1row = p.Cpad * 2 + 16, w2row = O16 * 2 + 16;
  const int x_bytes = 128 * xrow;
  uint8_t* xs = pk_smem;                                  // [2][128][xrow]
  uint8_t* w1s = xs + 2 * x_bytes;                        // [O16][w1row]
  uint8_t* w2s = w1s + O16 * w1row;                       // [O16][w2row]
  float* b1s = reinterpret_cast<float*>(w2s + O16 * w2row);
  float* b2s = b1s + O16;
  float4* s_src = reinterpret_cast<float4*>(b2s + O16);   // [2][THs][TWs]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int cchunks = p.Cpad / 8;
  for (int i = tid; i < O16 * cchunks; i += kPxThreads) {
    const int n = i / cchunks, c = i % cchunks;
    *reinterpret_cast<uint4*>(w1s + n * w1row + c * 16) = *reinterpret_cast<const uint4*>(p.w1t + static_cast<size_t>(n) * p.Cpad + c * 8);
  }
  for (int i = tid; i < O16 * (O16 / 8); i += kPxThreads) {
    const int n = i / (O16 / 8), c = i % (O16 / 8);
    *reinterpret_cast<uint4*>(w2s + n * w2row + c * 16) = *reinterpret_cast<const uint4*>(p.w2t + static_cast<size_t>(n) * O16 + c * 8);
  }
  for (int i = tid; i < O16; i += kPxThreads) { b1s[i] = p.b1[i]; b2s[i] = p.b2[i] * 1.4426950408889634f; }   // logits in log2 units

  const uint32_t xs_u = smem_u32(xs), w1_u = smem_u32(w1s), w2_u = smem_u32(w2s), src_u = smem_u32(s_src);
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kof = (lane >> 4) * 16;
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_kof = ((lane >> 3) & 1) * 16;
  const int valid_chunks = p.C / 8;
  const int cp_shift = p.cp_shift, cp_mask = (1 << cp_shift) - 1, px_step = kPxThreads >> cp_shift;
  const int ld_c = tid & cp_mask, ld_px0 = tid >> cp_shift;
  const int im_h = p.h, im_w = p.w, x_cs = p.xcs, w_xcs = im_w * x_cs, src_cs = p.src.cstride;
  const float* src_base = reinterpret_cast<const float*>(p.src.ptr) + p.src.coff;
  float* out_base = reinterpret_cast<float*>(p.out.ptr) + p.out.coff;
  const int out_cs = p.out.cstride;

  struct Coord { int b, ty, tx; };
  const int tiles_per_image = p.tiles_x * p.tiles_y;
  Coord delta;
  delta.b = static_cast<int>(gridDim.x) / tiles_per_image;
  { const int rem = static_cast<int>(gridDim.x) - delta.b * tiles_per_image; delta.ty = rem / p.tiles_x; delta.tx = rem - delta.ty * p.tiles_x; }
  auto advance = [&](Coord& c) {
    c.tx += delta.tx; if (c.tx >= p.tiles_x) { c.tx -= p.tiles_x; ++c.ty; }
    c.ty += delta.ty; if (c.ty >= p.tiles_y) { c.ty -= p.tiles_y; ++c.b; }
    c.b += delta.b;
  };
  constexpr int SRC_ITERS = (SRC_ELEMS + kPxThreads - 1) / kPxThreads;
  int src_ly[SRC_ITERS], src_lx[SRC_ITERS];
#pragma unroll
  for (int it = 0; it < SRC_ITERS; ++it) {
    const int i = tid + it * kPxThreads;
    src_ly[it] = (i < SRC_ELEMS) ? i / TWs - PAD : -1000; src_lx[it] = i % TWs - PAD;
  }

  auto issue_loads = [&](const Coord& c, int buf) {
    const int y0 = c.ty * kPkTileH, x0 = c.tx * kPkTileW;
    if (ld_c < cchunks) {
      const bool inside = (y0 + kPkTileH <= im_h) && (x0 + kPkTileW <= im_w) && ld_c < valid_chunks;
      uint32_t dst = xs_u + static_cast<uint32_t>(buf * x_bytes + ld_c * 16 + ld_px0 * xrow);
      const uint32_t dst_step = static_cast<uint32_t>(px_step * xrow);
      const __half* tile_base = p.x + ((static_cast<size_t>(c.b) * im_h + y0) * im_w + x0) * x_cs + p.xoff + ld_c * 8;
      for (int px = ld_px0; px < 128; px += px_step) {
        const int r = px >> 4, xx = px & 15;
        if (inside || (ld_c < valid_chunks && y0 + r < im_h && x0 + xx < im_w)) {
          pk_cp_async_16(dst, tile_base + r * w_xcs + xx * x_cs);
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
        }
        dst += dst_step;
      }
    }
    const bool src_inside = y0 >= PAD && x0 >= PAD && y0 + kPkTileH + PAD <= im_h && x0 + kPkTileW + PAD <= im_w;
    const float* img = src_base + static_cast<size_t>(c.b) * im_h * im_w * src_cs;
#pragma unroll
    for (int it = 0; it < SRC_ITERS; ++it) {
      if (src_ly[it] > -1000) {
        int yy = y0 + src_ly[it], xx = x0 + src_lx[it];
        if (!src_inside) { yy = pk_sym(yy, im_h); xx = pk_sym(xx, im_w); }
        const float* sp = img + (static_cast<size_t>(yy) * im_w + xx) * src_cs;
        const uint32_t dst = src_u + static_cast<uint32_t>((buf * SRC_ELEMS + tid + it * kPxThreads) * 16);
        pk_cp_async_4(dst, sp); pk_cp_async_4(dst + 4, sp + 1); pk_cp_async_4(dst + 8, sp + 2);
      }
    }
    pk_cp_async_commit();
  };

  int tile = blockIdx.x;
  int buf = 0;
  Coord cur;
  cur.b = tile / tiles_per_image;
  { const int rem = tile - cur.b * tiles_per_image; cur.ty = rem / p.tiles_x; cur.tx = rem - cur.ty * p.tiles_x; }
  if (tile < p.total_tiles) issue_loads(cur, 0);
  for (; tile < p.total_tiles; tile += gridDim.x, buf ^= 1) {
    const int b = cur.b;
    const int y0 = cur.ty * kPkTileH, x0 = cur.tx * kPkTileW;
    advance(cur);                                          // cur = the NEXT tile of this CTA from here on
    pk_cp_async_wait_all();
    __syncthreads();                                       // this tile's data (and the weights) visible; the other buffer is free
    if (tile + static_cast<int>(gridDim.x) < p.total_tiles) issue_loads(cur, buf ^ 1);   // in flight during the math below

    // ---------------------------------------------------------------- GEMM 1: hidden = relu(x W1 + b1), two row blocks per warp
    float acc[2][NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float bl = b1s[j * 8 + 2 * t4], bh = b1s[j * 8 + 2 * t4 + 1];
#pragma unroll
      for (int m = 0; m < 2; ++m) { acc[m][j][0] = bl; acc[m][j][1] = bh; acc[m][j][2] = bl; acc[m][j][3] = bh; }
    }
    const uint32_t a_base = xs_u + static_cast<uint32_t>(buf * x_bytes + (warp * 32 + a_row) * xrow + a_kof);
    const uint32_t a_blk = static_cast<uint32_t>(16 * xrow);
    for (int ks = 0; ks < p.Cpad / 16; ++ks) {
      uint32_t a0[4], a1[4];
      pk_ldmatrix_x4(a0, a_base + ks * 32);
      pk_ldmatrix_x4(a1, a_base + a_blk + ks * 32);
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        uint32_t bf[4];
        pk_ldmatrix_x4(bf, w1_u + static_cast<uint32_t>((j * 8 + b_row) * w1row + ks * 32 + b_kof));
        pk_mma(acc[0][j], a0, bf[0], bf[1]);
        pk_mma(acc[0][j + 1], a0, bf[2], bf[3]);
        pk_mma(acc[1][j], a1, bf[0], bf[1]);
        pk_mma(acc[1][j + 1], a1, bf[2], bf[3]);
      }
    }
    // ---------------------------------------------------------------- GEMM 2: logits = hidden W2 + b2 (then * log2 e)
    float lg[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) lg[m][j][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        a[m][0] = pk_pack_relu(acc[m][2 * kk][0], acc[m][2 * kk][1]);
        a[m][1] = pk_pack_relu(acc[m][2 * kk][2], acc[m][2 * kk][3]);
        a[m][2] = pk_pack_relu(acc[m][2 * kk + 1][0], acc[m][2 * kk + 1][1]);
        a[m][3] = pk_pack_relu(acc[m][2 * kk + 1][2], acc[m][2 * kk + 1][3]);
      }
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        uint32_t bf[4];
        pk_ldmatrix_x4(bf, w2_u + static_cast<uint32_t>((j * 8 + b_row) * w2row + kk * 32 + b_kof));
        pk_mma(lg[0][j], a[0], bf[0], bf[1]);
        pk_mma(lg[0][j + 1], a[0], bf[2], bf[3]);
        pk_mma(lg[1][j], a[1], bf[0], bf[1]);
        pk_mma(lg[1][j + 1], a[1], bf[2], bf[3]);
      }
    }
    // ---------------------------------------------------------------- logits -> [column][pixel] in the warp's own (consumed) x rows
    __syncwarp();
    float* st = reinterpret_cast<float*>(xs + buf * x_bytes + warp * 32 * xrow);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int c0 = j * 8 + 2 * t4;
      const float bl = b2s[c0], bh = b2s[c0 + 1];            // already scaled by log2 e
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int px = m * 16 + g;
        st[c0 * kPxLgPitch + px] = fmaf(lg[m][j][0], 1.4426950408889634f, bl);
        st[(c0 + 1) * kPxLgPitch + px] = fmaf(lg[m][j][1], 1.4426950408889634f, bh);
        st[c0 * kPxLgPitch + px + 8] = fmaf(lg[m][j][2], 1.4426950408889634f, bl);
        st[(c0 + 1) * kPxLgPitch + px + 8] = fmaf(lg[m][j][3], 1.4426950408889634f, bh);
      }
    }
    __syncwarp();
    // ---------------------------------------------------------------- softmax over K*K + filter apply: lane = pixel
    {
      const int r = warp * 2 + (lane >> 4), cx = lane & 15;
      // three interleaved accumulator sets: one thread per pixel means 25-deep dependent chains otherwise
      float l[K2];
      float m3[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int k = 0; k < K2; ++k) { l[k] = st[k * kPxLgPitch + lane]; m3[k % 3] = fmaxf(m3[k % 3], l[k]); }
      const float mx = fmaxf(fmaxf(m3[0], m3[1]), m3[2]);
      const float4* tile_f = s_src + buf * SRC_ELEMS + r * TWs + cx;
      float sum[3] = {0.f, 0.f, 0.f}, rr[3] = {0.f, 0.f, 0.f}, gg[3] = {0.f, 0.f, 0.f}, bb[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < K2; ++k) {
        const float ev = pk_ex2(l[k] - mx);
        const float4 sv = tile_f[(k / K) * TWs + (k % K)];
        sum[k % 3] += ev;
        rr[k % 3] = fmaf(ev, sv.x, rr[k % 3]); gg[k % 3] = fmaf(ev, sv.y, gg[k % 3]); bb[k % 3] = fmaf(ev, sv.z, bb[k % 3]);
      }
      const int y = y0 + r, x = x0 + cx;
      if (y < im_h && x < im_w) {
        const float inv = 1.f / (sum[0] + sum[1] + sum[2]);
        float* o = out_base + ((static_cast<size_t>(b) * im_h + y) * im_w + x) * out_cs;
        o[0] = (rr[0] + rr[1] + rr[2]) * inv; o[1] = (gg[0] + gg[1] + gg[2]) * inv; o[2] = (bb[0] + bb[1] + bb[2]) * inv;
      }
    }
  }
  pk_cp_async_wait_all();
}

// logits staged in the consumed x rows: 32 pixels x xrow bytes per warp must hold [O16][33] floats
static bool post_kp_pixel_form(int cin, int ksize, int features) {
  if (features != 1 || !(ksize == 3 || ksize == 5)) return false;
  const int o16 = round_up(ksize * ksize, 16), cpad = round_up(cin, 16);
  return 32 * (cpad * 2 + 16) >= o16 * kPxLgPitch * 4;
}

template <int K>
static int launch_post_kp_pixel(dd_ctx* ctx, const PostKpParams& p, cudaStream_t s) {
  constexpr int K2 = K * K, O16 = (K2 + 15) / 16 * 16, PAD = (K - 1) / 2;
  const size_t smem = 2 * 128ull * (p.Cpad * 2 + 16) + static_cast<size_t>(O16) * (p.Cpad * 2 + 16) +
                      static_cast<size_t>(O16) * (O16 * 2 + 16) + 2ull * O16 * sizeof(float) +
                      2 * static_cast<size_t>(kPkTileH + 2 * PAD) * (kPkTileW + 2 * PAD) * sizeof(float4);
  DD_CHECK_ARG(smem <= ctx->max_smem_optin, "post_kp: tile does not fit shared memory");
  DD_CUDA(cudaFuncSetAttribute(post_kp_pixel_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  DD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, post_kp_pixel_kernel<K>, kPxThreads, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = ctx->sm_count * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  post_kp_pixel_kernel<K><<<static_cast<unsigned>(grid), kPxThreads, smem, s>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

template <int K, int T>
static int launch_post_kp(dd_ctx* ctx, const PostKpParams& p, cudaStream_t s) {
  constexpr int K2 = K * K, O16 = T * ((K2 + 15) / 16 * 16), PAD = (K - 1) / 2;
  const size_t smem = 2 * 128ull * (p.Cpad * 2 + 16) + static_cast<size_t>(O16) * (p.Cpad * 2 + 16) +
                      static_cast<size_t>(O16) * (O16 * 2 + 16) + 2ull * O16 * sizeof(float) +
                      2 * static_cast<size_t>(T) * (kPkTileH + 2 * PAD) * (kPkTileW + 2 * PAD) * sizeof(float4);
  DD_CHECK_ARG(smem <= ctx->max_smem_optin, "post_kp: tile does not fit shared memory");
  DD_CUDA(cudaFuncSetAttribute(post_kp_fused_kernel<K, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // persistent grid = exactly the CTAs that are resident at once (4 per SM at C = 64, 3 at 96, 2 at 128): a 4th CTA per SM that
  // has to wait for a slot would run its whole tile share after the others have finished
  int per_sm = 0;
  DD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, post_kp_fused_kernel<K, T>, kPkThreads, smem));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = ctx->sm_count * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  post_kp_fused_kernel<K, T><<<static_cast<unsigned>(grid), kPkThreads, smem, s>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_post_kp_supported(int ksize, int features) {
  return (ksize == 3 || ksize == 5) && (features == 1 || features == 3);
}

static int post_kp_o16(int ksize, int features) { return features * round_up(ksize * ksize, 16); }

size_t dd_post_kp_weights_bytes(int cin, int ksize, int features) {
  const int o16 = post_kp_o16(ksize, features), cpad = round_up(cin, 16);
  return static_cast<size_t>(o16) * cpad * 2 + static_cast<size_t>(o16) * o16 * 2 + 2ull * o16 * 4;
}

/* Host-side packing: w1 TF [1,1,cin,O], b1 [O], w2 TF [1,1,O,O], b2 [O] (fp32, host) -> blob = W1^T [O16][cpad] fp16,
 * W2^T [O16][O16] fp16, b1 [O16] fp32, b2 [O16] fp32 (zero padded). */
int dd_post_kp_pack_weights(const float* w1, const float* b1, const float* w2, const float* b2, int cin, int ksize, int features,
                            void* blob_host) {
  if (!w1 || !b1 || !w2 || !b2 || !blob_host || cin <= 0 || !dd_post_kp_supported(ksize, features)) {
    set_error("post_kp_pack_weights: bad argument");
    return DD_ERR_INVALID;
  }
  const int k2 = ksize * ksize, kp = round_up(k2, 16), o = features * k2, o16 = features * kp, cpad = round_up(cin, 16);
  memset(blob_host, 0, dd_post_kp_weights_bytes(cin, ksize, features));
  __half* w1t = reinterpret_cast<__half*>(blob_host);
  __half* w2t = w1t + static_cast<size_t>(o16) * cpad;
  float* b1p = reinterpret_cast<float*>(w2t + static_cast<size_t>(o16) * o16);
  float* b2p = b1p + o16;
  // hidden unit c -> row (c / K2) * KP + c % K2;  logit c = (feature f, tap k) -> column f * KP + 8j + 2*t4 + e with
  // k = 4 * (2j + e) + t4 (the kernel's quad layout); padded logit columns get a bias of -inf (softmax weight 0)
  auto hid = [&](int c) { return (c / k2) * kp + c % k2; };
  const bool pixel_form = post_kp_pixel_form(cin, ksize, features);     // one thread per pixel: logit column = tap
  auto col = [&](int c) {
    const int f = c / k2, k = c % k2, slot = k / 4, t4 = k % 4;
    return pixel_form ? f * kp + k : f * kp + 8 * (slot / 2) + 2 * t4 + (slot % 2);
  };
  for (int c = 0; c < cin; ++c)
    for (int n = 0; n < o; ++n) w1t[static_cast<size_t>(hid(n)) * cpad + c] = __float2half_rn(w1[static_cast<size_t>(c) * o + n]);
  for (int k = 0; k < o; ++k)
    for (int n = 0; n < o; ++n) w2t[static_cast<size_t>(col(n)) * o16 + hid(k)] = __float2half_rn(w2[static_cast<size_t>(k) * o + n]);
  for (int n = 0; n < o16; ++n) b2p[n] = -INFINITY;
  for (int n = 0; n < o; ++n) { b1p[hid(n)] = b1[n]; b2p[col(n)] = b2[n]; }
  return DD_OK;
}

int dd_post_kp_fwd(dd_ctx* ctx, const dd_tensor* x, const void* blob_dev, const dd_tensor* src, int ksize, int features,
                   int images_per_tuple, const dd_tensor* out, void* stream) {
  DD_CHECK_ARG(ctx && blob_dev && tensor_ok(x) && tensor_ok(src) && tensor_ok(out), "bad argument");
  DD_CHECK_ARG(dd_post_kp_supported(ksize, features), "post_kp: kernel size 3 or 5 and 1 or 3 features per tuple only");
  DD_CHECK_ARG(x->dtype == DD_F16 && x->c % 8 == 0 && x->coff % 8 == 0 && x->cstride % 8 == 0, "post_kp: x must be aligned fp16");
  DD_CHECK_ARG(src->c == 3 && out->c == 3 && src->dtype == DD_F32 && out->dtype == DD_F32, "src/out must be fp32 rgb");
  DD_CHECK_ARG(src->n == x->n * features && out->n == src->n && src->h == x->h && src->w == x->w && out->h == x->h &&
                   out->w == x->w, "post_kp: src/out must be [x.n * features, h, w, 3]");
  DD_CHECK_ARG(images_per_tuple >= 1 && x->n % images_per_tuple == 0, "x.n must be a multiple of images_per_tuple");
  PostKpParams p;
  memset(&p, 0, sizeof(p));
  p.x = reinterpret_cast<const __half*>(x->ptr); p.xcs = x->cstride; p.xoff = x->coff; p.C = x->c; p.Cpad = round_up(x->c, 16);
  const int o16 = post_kp_o16(ksize, features);
  p.w1t = reinterpret_cast<const __half*>(blob_dev);
  p.w2t = p.w1t + static_cast<size_t>(o16) * p.Cpad;
  p.b1 = reinterpret_cast<const float*>(p.w2t + static_cast<size_t>(o16) * o16);
  p.b2 = p.b1 + o16;
  p.src = make_view(src); p.out = make_view(out);
  p.B = x->n; p.h = x->h; p.w = x->w; p.ipt = images_per_tuple;
  p.tiles_x = (x->w + kPkTileW - 1) / kPkTileW; p.tiles_y = (x->h + kPkTileH - 1) / kPkTileH;
  DD_CHECK_ARG(static_cast<long long>(p.tiles_x) * p.tiles_y * x->n < (1ll << 30), "post_kp: too many tiles");
  p.total_tiles = p.tiles_x * p.tiles_y * x->n;
  while ((1 << p.cp_shift) < p.Cpad / 8) ++p.cp_shift;
  DD_CHECK_ARG(p.cp_shift <= 7, "post_kp: at most 1024 input channels");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (post_kp_pixel_form(x->c, ksize, features))       // must mirror dd_post_kp_pack_weights (the column order of W2 differs)
    return ksize == 5 ? launch_post_kp_pixel<5>(ctx, p, s) : launch_post_kp_pixel<3>(ctx, p, s);
  if (ksize == 5 && features == 1) return launch_post_kp<5, 1>(ctx, p, s);
  if (ksize == 5 && features == 3) return launch_post_kp<5, 3>(ctx, p, s);
  if (ksize == 3 && features == 1) return launch_post_kp<3, 1>(ctx, p, s);
  return launch_post_kp<3, 3>(ctx, p, s);
}

}  // extern "C"
