// C-ABI: MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93) as ONE kernel.
//
//   s_up = up2(small); x0 = relu(conv1x1_{6->24}(concat[s_up, large]))
//   x1 = x0 + conv3x3(relu(conv3x3(relu(x0))));  x2 = x1 + conv3x3(relu(conv3x3(relu(x1))))
//   w = sigmoid(relu(conv1x1_{24->1}(x2)));      out = large - w * up2(down2(large)) + w * s_up   (+ inverse standardisation)
//
// The unfused path moved six 24-channel tensors per pixel through HBM (~600 B/px) around 36 B/px of real input/output and
// ran the 24->24 convolutions as per-row-overhead-bound launches of the large-layer tcgen05 kernel.  Here every
// intermediate lives in shared memory: a CTA owns a 32x16 output tile, computes the 40x24 halo region of x0 and shrinks
// by one pixel per 3x3 layer (t1 38x22, x1 36x20, t3 34x18, x2 32x16).  The 3x3 layers are implicit GEMMs on the warp-level
// tensor-core path (mma.sync m16n8k16, fp16 operands, fp32 accumulate): M = 16 consecutive pixels of the layer's output
// region, N = 24 output channels (3 n-tiles), K = 9 taps x 24 (one k16 + one k8 step per tap).  A fragments come from ldmatrix over the [pixel][24 ch] fp16 activation buffers (48-byte pixel
// pitch: conflict-free), B fragments from the XOR-swizzled [tap][n][32] weights resident in shared memory.  ReLU on a
// layer INPUT is applied to the A fragments in registers, so x1 is stored once (raw) and serves both as the input of
// the third convolution and as the residual of the fourth.  Pixels outside the image are forced to zero after every
// layer (that is what SAME zero padding of the next layer sees in the reference).
// Activations between layers are fp16 (like the unfused tensor-core path); head, tail and blend are fp32.
#include <string.h>

#include "dd_internal.h"
#include "dd_ptx.cuh"

namespace dd {

constexpr int kCfTileW = 32, kCfTileH = 16;
constexpr int kCfC = 24;                         // channels of the compose net
constexpr int kCfThreads = 512, kCfWarps = 16;
constexpr int kCfW0 = kCfTileW + 8, kCfH0 = kCfTileH + 8;   // x0 region 40 x 24
constexpr int kCfPixBytes = kCfC * 2;            // 48-byte pixel pitch
constexpr int kCfConvBytes = 9 * kCfC * 64;      // one 3x3 layer: [tap][n = 24][k = 32] fp16, 64-byte rows, XOR swizzled
constexpr int kCfBufA = kCfW0 * kCfH0 * kCfPixBytes + 64;                     // x0, later t3
constexpr int kCfBufB = (kCfW0 - 2) * (kCfH0 - 2) * kCfPixBytes + 64;         // t1
constexpr int kCfBufC = (kCfW0 - 4) * (kCfH0 - 4) * kCfPixBytes + 64;         // x1 (raw)
constexpr int kCfStage = kCfW0 * kCfH0 * 6 * 4;                               // [pixel][small rgb, large rgb] fp32
// float parameters following the four weight blocks in the packed blob
constexpr int kCfHeadW = 0, kCfHeadB = 144, kCfConvB = 168, kCfTailW = 264, kCfTailB = 288, kCfFloats = 292;
constexpr int kCfBlobBytes = 4 * kCfConvBytes + kCfFloats * 4;
constexpr int kCfSmem = 4 * kCfConvBytes + kCfFloats * 4 + kCfBufA + kCfBufB + kCfBufC + kCfStage + 128;

struct ComposeFusedParams {
  View small, large, out;
  const uint8_t* blob;
  int tiles_x, tiles_y, total_tiles;
  int has_inv;
  dd_invert_params inv;
  float sqrt_var;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_1688(float (&d)[4], const uint32_t (&a)[2], uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(b0));
}
__device__ __forceinline__ uint32_t relu_h2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  h = __hmax2(h, __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void cf_cp_async_4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ float cf_signed_expm1(float v) { return copysignf(expm1f(fabsf(v)), v) * (v != 0.f); }

struct CfTile {
  int n, ty0, tx0;
};
__device__ __forceinline__ CfTile cf_tile(int lin, const ComposeFusedParams& p) {
  CfTile t;
  const int r = lin / p.tiles_x;
  t.tx0 = (lin - r * p.tiles_x) * kCfTileW;
  t.n = r / p.tiles_y;
  t.ty0 = (r - t.n * p.tiles_y) * kCfTileH;
  return t;
}

// One 3x3 layer over a WI x HI input region -> (WI-2) x (HI-2) output region whose top-left pixel is image pixel
// (gy0, gx0).  MODE 0: out = relu(acc + b) -> fp16 buffer;  MODE 1: out = acc + b + resid -> fp16 buffer (raw);
// MODE 2: x2 = acc + b + resid stays in registers and goes straight into the tail + blend.
template <int WI, int HI, bool RELU_IN, int MODE, int RW, int RO>
__device__ __forceinline__ void cf_conv_layer(const uint8_t* in, uint8_t* out, const uint8_t* resid, const uint8_t* wsm,
                                              const float* bias, const float* fl, int gy0, int gx0, int n,
                                              const ComposeFusedParams& p) {
  constexpr int WO = WI - 2, HO = HI - 2, NPIX = WO * HO, NMT = (NPIX + 15) / 16, NGRP = (NMT + 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t in_u32 = smem_u32(in), w_u32 = smem_u32(wsm);
  // B ldmatrix.x4 lane address inside a tap block: row n = nt*8 + lane%8, 16-byte chunk lane/8, swizzled with (n>>1)&3
  uint32_t b_off[3];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) {
    const int nrow = nt * 8 + (lane & 7);
    b_off[nt] = static_cast<uint32_t>(nrow * 64 + (((lane >> 3) ^ ((nrow >> 1) & 3)) << 4));
  }
  const int ld_row = (lane & 7) + ((lane >> 3) & 1) * 8;      // pixel of the m-tile this lane addresses for ldmatrix
  const uint32_t ld_k = static_cast<uint32_t>(lane >> 4) * 16;  // k chunk (0 / 8) inside a k-step
  const int h = p.large.h, w = p.large.w;
  for (int grp = warp; grp < NGRP; grp += kCfWarps) {
    float acc[2][3][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    uint32_t a_addr[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      int pl = (grp * 2 + mt) * 16 + ld_row;
      if (pl > NPIX - 1) pl = NPIX - 1;
      const int y = pl / WO, x = pl - y * WO;
      a_addr[mt] = in_u32 + static_cast<uint32_t>((y * WI + x) * kCfPixBytes);
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
      uint32_t b[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) ldmatrix_x4(b[nt], w_u32 + tap * (kCfC * 64) + b_off[nt]);
      // K = 24 exactly: channels 0..15 as one k16 step, channels 16..23 as one k8 step (the legacy tensor path is the
      // bound of this kernel, so no multiply-by-zero padding)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t a_tap = a_addr[mt] + static_cast<uint32_t>((dy * WI + dx) * kCfPixBytes);
        uint32_t a[4], a8[2];
        ldmatrix_x4(a, a_tap + ld_k);
        ldmatrix_x2(a8, a_tap + 32);
        if (RELU_IN) {
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = relu_h2(a[i]);
          a8[0] = relu_h2(a8[0]); a8[1] = relu_h2(a8[1]);
        }
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          mma_16816(acc[mt][nt], a, b[nt][0], b[nt][1]);
          mma_1688(acc[mt][nt], a8, b[nt][2]);
        }
      }
    }
    // ---- epilogue: this lane holds channels nt*8 + 2*t4 + {0,1} of pixels (row g) and (row g + 8) of both m-tiles
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int pl = (grp * 2 + mt) * 16 + g + hf * 8;
        const bool live = pl < NPIX;                 // uniform over the 4 lanes of a row
        const int plc = live ? pl : 0;
        const int y = plc / WO, x = plc - y * WO;
        const int gy = gy0 + y, gx = gx0 + x;
        const bool inside = live && gy >= 0 && gy < h && gx >= 0 && gx < w;
        float v[6];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          v[2 * nt] = acc[mt][nt][hf * 2] + bias[nt * 8 + 2 * t4];
          v[2 * nt + 1] = acc[mt][nt][hf * 2 + 1] + bias[nt * 8 + 2 * t4 + 1];
        }
        if (MODE != 0) {
          const uint8_t* rp = resid + ((y + RO) * RW + (x + RO)) * kCfPixBytes;
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(rp + (nt * 8 + 2 * t4) * 2));
            v[2 * nt] += r2.x; v[2 * nt + 1] += r2.y;
          }
        }
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 6; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (MODE != 2) {
          if (live) {
            uint8_t* op = out + (y * WO + x) * kCfPixBytes;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
              *reinterpret_cast<__half2*>(op + (nt * 8 + 2 * t4) * 2) =
                  inside ? __floats2half2_rn(v[2 * nt], v[2 * nt + 1]) : __floats2half2_rn(0.f, 0.f);
          }
        } else {
          // tail: a = relu(w_tail . x2 + b); wgt = sigmoid(a); the 24 channels of a pixel are spread over 4 lanes
          float s = 0.f;
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
            s += v[2 * nt] * fl[kCfTailW + nt * 8 + 2 * t4] + v[2 * nt + 1] * fl[kCfTailW + nt * 8 + 2 * t4 + 1];
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          const float a = fmaxf(s + fl[kCfTailB], 0.f);
          const float wgt = 1.f / (1.f + __expf(-a));
          if (inside && t4 < 3) {
            const int c = t4;
            const int yb = gy & ~1, xb = gx & ~1;
            const float low = 0.25f * (p.large.load(p.large.pix(n, yb, xb), c) + p.large.load(p.large.pix(n, yb, xb + 1), c) +
                                       p.large.load(p.large.pix(n, yb + 1, xb), c) + p.large.load(p.large.pix(n, yb + 1, xb + 1), c));
            const size_t opix = p.large.pix(n, gy, gx);
            float o = p.large.load(opix, c) - wgt * low + wgt * p.small.load(p.small.pix(n, gy >> 1, gx >> 1), c);
            if (p.has_inv) {
              if (p.inv.variance != 1.f) o *= p.sqrt_var;
              if (p.inv.mean != 0.f) o += p.inv.mean;
              if (p.inv.use_log1p) o = cf_signed_expm1(o);
            }
            p.out.store(opix, c, o);
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kCfThreads, 1) compose_fused_kernel(const __grid_constant__ ComposeFusedParams p) {
  extern __shared__ __align__(128) uint8_t cf_smem[];
  uint8_t* wsm = cf_smem;                                            // 4 conv weight blocks
  float* fl = reinterpret_cast<float*>(cf_smem + 4 * kCfConvBytes);  // head / bias / tail floats
  uint8_t* bufA = reinterpret_cast<uint8_t*>(fl + kCfFloats) + 48;   // 16-byte aligned: (4*13824 + 292*4 + 48) % 16 == 0
  uint8_t* bufB = bufA + kCfBufA;
  uint8_t* bufC = bufB + kCfBufB;
  float* stage = reinterpret_cast<float*>(bufC + kCfBufC);

  const int tid = threadIdx.x;
  // weights + parameters (L2 resident after the first CTA), activation buffers zeroed once (finite garbage only)
  for (int i = tid; i < kCfBlobBytes / 16; i += kCfThreads)
    reinterpret_cast<uint4*>(cf_smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.blob) + i);
  for (int i = tid; i < (kCfBufA + kCfBufB + kCfBufC) / 16; i += kCfThreads)
    reinterpret_cast<uint4*>(bufA)[i] = make_uint4(0, 0, 0, 0);
  const int h = p.large.h, w = p.large.w;
  const float* sm_ptr = reinterpret_cast<const float*>(p.small.ptr);
  const float* lg_ptr = reinterpret_cast<const float*>(p.large.ptr);

  auto prefetch_inputs = [&](const CfTile& t) {
    // region pixel i -> image pixel (ty0 - 4 + i / 40, tx0 - 4 + i % 40); outside pixels are skipped (x0 = 0 there)
    for (int i = tid; i < kCfW0 * kCfH0; i += kCfThreads) {
      const int y = i / kCfW0, x = i - y * kCfW0;
      const int gy = t.ty0 - 4 + y, gx = t.tx0 - 4 + x;
      if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
        const float* s = sm_ptr + p.small.pix(t.n, gy >> 1, gx >> 1) * p.small.cstride + p.small.coff;
        const float* l = lg_ptr + p.large.pix(t.n, gy, gx) * p.large.cstride + p.large.coff;
        float* d = stage + i * 6;
        cf_cp_async_4(d, s); cf_cp_async_4(d + 1, s + 1); cf_cp_async_4(d + 2, s + 2);
        cf_cp_async_4(d + 3, l); cf_cp_async_4(d + 4, l + 1); cf_cp_async_4(d + 5, l + 2);
      }
    }
  };

  int lin = blockIdx.x;
  if (lin < p.total_tiles) prefetch_inputs(cf_tile(lin, p));
  for (; lin < p.total_tiles; lin += gridDim.x) {
    const CfTile t = cf_tile(lin, p);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                   // inputs staged; previous tile completely done
    // ---- head: x0 = relu(W0 . [small_up, large] + b0) over the 40 x 24 region, zero outside the image
    for (int i = tid; i < kCfW0 * kCfH0; i += kCfThreads) {
      const int y = i / kCfW0, x = i - y * kCfW0;
      const int gy = t.ty0 - 4 + y, gx = t.tx0 - 4 + x;
      uint4 o[3] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
        float in[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) in[k] = stage[i * 6 + k];
        __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int c4 = 0; c4 < 6; ++c4) {
          float4 a = *reinterpret_cast<const float4*>(fl + kCfHeadB + c4 * 4);
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const float4 wk = *reinterpret_cast<const float4*>(fl + kCfHeadW + k * kCfC + c4 * 4);
            a.x = fmaf(in[k], wk.x, a.x); a.y = fmaf(in[k], wk.y, a.y); a.z = fmaf(in[k], wk.z, a.z); a.w = fmaf(in[k], wk.w, a.w);
          }
          oh[c4 * 2] = __floats2half2_rn(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
          oh[c4 * 2 + 1] = __floats2half2_rn(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(bufA + i * kCfPixBytes);
      dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
    __syncthreads();
    // the staging buffer is free again: fetch the inputs of the next tile while the convolutions run
    if (lin + static_cast<int>(gridDim.x) < p.total_tiles) prefetch_inputs(cf_tile(lin + gridDim.x, p));
    // t1 = relu(conv(x0))                       38 x 22 at (-3, -3)     [x0 >= 0 already]
    cf_conv_layer<kCfW0, kCfH0, false, 0, 1, 0>(bufA, bufB, nullptr, wsm, fl + kCfConvB, fl, t.ty0 - 3, t.tx0 - 3, t.n, p);
    __syncthreads();
    // x1 = x0 + conv(t1)                        36 x 20 at (-2, -2)     residual x0 at (+2, +2), pitch 40
    cf_conv_layer<kCfW0 - 2, kCfH0 - 2, false, 1, kCfW0, 2>(bufB, bufC, bufA, wsm + kCfConvBytes, fl + kCfConvB + kCfC, fl,
                                                             t.ty0 - 2, t.tx0 - 2, t.n, p);
    __syncthreads();
    // t3 = relu(conv(relu(x1)))                 34 x 18 at (-1, -1)     into the x0 buffer
    cf_conv_layer<kCfW0 - 4, kCfH0 - 4, true, 0, 1, 0>(bufC, bufA, nullptr, wsm + 2 * kCfConvBytes, fl + kCfConvB + 2 * kCfC, fl,
                                                        t.ty0 - 1, t.tx0 - 1, t.n, p);
    __syncthreads();
    // x2 = x1 + conv(t3); tail; blend           32 x 16 at (0, 0)       residual x1 at (+2, +2), pitch 36
    cf_conv_layer<kCfW0 - 6, kCfH0 - 6, false, 2, kCfW0 - 4, 2>(bufA, nullptr, bufC, wsm + 3 * kCfConvBytes,
                                                                 fl + kCfConvB + 3 * kCfC, fl, t.ty0, t.tx0, t.n, p);
  }
}

inline bool cf_same_spatial(const dd_tensor* a, const dd_tensor* b) { return a->n == b->n && a->h == b->h && a->w == b->w; }

}  // namespace dd

using namespace dd;

extern "C" {

size_t dd_compose_weights_bytes(void) { return kCfBlobBytes; }

int dd_compose_scales_fwd(dd_ctx* ctx, const dd_tensor* small, const dd_tensor* large, const void* packed_dev,
                          const dd_invert_params* inv, const dd_tensor* out, void* stream) {
  DD_CHECK_ARG(ctx && packed_dev && tensor_ok(small) && tensor_ok(large) && tensor_ok(out), "bad argument");
  DD_CHECK_ARG(small->c == 3 && large->c == 3 && out->c == 3 && small->dtype == DD_F32 && large->dtype == DD_F32 &&
                   out->dtype == DD_F32, "compose_scales: small / large / out must be fp32 rgb");
  DD_CHECK_ARG(large->h == 2 * small->h && large->w == 2 * small->w && small->n == large->n && cf_same_spatial(large, out),
               "compose_scales: dims");
  DD_CHECK_ARG(reinterpret_cast<uintptr_t>(packed_dev) % 16 == 0, "compose_scales: packed weights must be 16-byte aligned");
  ComposeFusedParams p;
  memset(&p, 0, sizeof(p));
  p.small = make_view(small); p.large = make_view(large); p.out = make_view(out);
  p.blob = static_cast<const uint8_t*>(packed_dev);
  p.tiles_x = (large->w + kCfTileW - 1) / kCfTileW; p.tiles_y = (large->h + kCfTileH - 1) / kCfTileH;
  const long long total = static_cast<long long>(p.tiles_x) * p.tiles_y * large->n;
  DD_CHECK_ARG(total < (1ll << 31), "compose_scales: too many tiles");
  p.total_tiles = static_cast<int>(total);
  if (inv) { p.has_inv = 1; p.inv = *inv; p.sqrt_var = sqrtf(inv->variance); }
  static bool configured = false;
  if (!configured) {
    DD_CUDA(cudaFuncSetAttribute(compose_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCfSmem));
    configured = true;
  }
  int grid = ctx->sm_count;
  if (grid > p.total_tiles) grid = p.total_tiles;
  compose_fused_kernel<<<grid, kCfThreads, kCfSmem, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // extern "C"
