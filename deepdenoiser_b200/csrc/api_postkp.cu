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
  int tiles_x, tiles_y; long long total_tiles;
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

template <int K, int T>
__global__ void __launch_bounds__(kPkThreads) post_kp_fused_kernel(const PostKpParams p) {
  constexpr int K2 = K * K, O = T * K2, O16 = (O + 15) / 16 * 16, NT = O16 / 8, PAD = (K - 1) / 2;
  constexpr int TWs = kPkTileW + 2 * PAD, THs = kPkTileH + 2 * PAD;
  extern __shared__ __align__(16) uint8_t pk_smem[];
  const int xrow = p.Cpad * 2 + 16;                       // bytes per staged pixel
  const int w1row = p.Cpad * 2 + 16, w2row = O16 * 2 + 16;
  uint8_t* xs = pk_smem;                                  // [128][xrow]
  uint8_t* w1s = xs + 128 * xrow;                         // [O16][w1row]
  uint8_t* w2s = w1s + O16 * w1row;                       // [O16][w2row]
  float* b1s = reinterpret_cast<float*>(w2s + O16 * w2row);
  float* b2s = b1s + O16;
  float4* s_src = reinterpret_cast<float4*>(b2s + O16);   // [T][THs][TWs]

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
  for (int i = tid; i < O16; i += kPkThreads) { b1s[i] = p.b1[i]; b2s[i] = p.b2[i]; }

  const uint32_t xs_u = smem_u32(xs), w1_u = smem_u32(w1s), w2_u = smem_u32(w2s);
  // ldmatrix lane addressing.  A (x4): matrices (rows 0-7,k 0-7), (rows 8-15,k 0-7), (rows 0-7,k 8-15), (rows 8-15,k 8-15)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kof = (lane >> 4) * 16;
  // B (x4): matrices (n-tile j,k 0-7), (n-tile j,k 8-15), (n-tile j+1,k 0-7), (n-tile j+1,k 8-15); rows = n
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_kof = ((lane >> 3) & 1) * 16;
  const int valid_chunks = p.C / 8;

  for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int tx = static_cast<int>(tile % p.tiles_x);
    const int ty = static_cast<int>((tile / p.tiles_x) % p.tiles_y);
    const int b = static_cast<int>(tile / (static_cast<long long>(p.tiles_x) * p.tiles_y));
    const int y0 = ty * kPkTileH, x0 = tx * kPkTileW;
    __syncthreads();                                       // previous tile fully consumed (and weights visible)
    // activation tile: 128 pixels x Cpad channels, zero outside the image / beyond C
    for (int i = tid; i < 128 * cchunks; i += kPkThreads) {
      const int px = i / cchunks, c = i % cchunks;
      const int y = y0 + (px >> 4), x = x0 + (px & 15);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (y < p.h && x < p.w && c < valid_chunks)
        v = __ldg(reinterpret_cast<const uint4*>(p.x + ((static_cast<size_t>(b) * p.h + y) * p.w + x) * p.xcs + p.xoff + c * 8));
      *reinterpret_cast<uint4*>(xs + px * xrow + c * 16) = v;
    }
    // source halo tiles (symmetric padding, KernelPrediction.py:30 via Conv2dUtilities.pad_equally)
    for (int i = tid; i < T * THs * TWs; i += kPkThreads) {
      const int f = i / (THs * TWs), r = i % (THs * TWs);
      const int ly = r / TWs, lx = r % TWs;
      const int img = ((b / p.ipt) * T + f) * p.ipt + (b % p.ipt);
      const size_t sp = p.src.pix(img, pk_sym(y0 + ly - PAD, p.h), pk_sym(x0 + lx - PAD, p.w));
      const float* s = reinterpret_cast<const float*>(p.src.ptr) + sp * p.src.cstride + p.src.coff;
      s_src[i] = make_float4(s[0], s[1], s[2], 0.f);
    }
    __syncthreads();

    // ---------------------------------------------------------------- GEMM 1: hidden = relu(x W1 + b1)
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float bl = b1s[j * 8 + 2 * t4], bh = b1s[j * 8 + 2 * t4 + 1];
      acc[j][0] = bl; acc[j][1] = bh; acc[j][2] = bl; acc[j][3] = bh;
    }
    const uint32_t a_base = xs_u + static_cast<uint32_t>((warp * 16 + a_row) * xrow + a_kof);
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
    // ---------------------------------------------------------------- GEMM 2: logits = hidden W2 + b2
    float lg[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float bl = b2s[j * 8 + 2 * t4], bh = b2s[j * 8 + 2 * t4 + 1];
      lg[j][0] = bl; lg[j][1] = bh; lg[j][2] = bl; lg[j][3] = bh;
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
    // ---------------------------------------------------------------- softmax over K*K + filter apply, per feature
    // this thread: pixels (row `warp`, column g) [regs 0,1] and (row `warp`, column g + 8) [regs 2,3]; logit columns
    // j*8 + 2*t4 + {0,1} of every n-tile j
#pragma unroll
    for (int f = 0; f < T; ++f) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int lx = g + half * 8;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = j * 8 + 2 * t4 + e;
            if (col >= f * K2 && col < (f + 1) * K2) mx = fmaxf(mx, lg[j][half * 2 + e]);
          }
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f, r = 0.f, gg = 0.f, bb = 0.f;
        const float4* tile_f = s_src + f * (THs * TWs) + warp * TWs + lx;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = j * 8 + 2 * t4 + e;
            if (col >= f * K2 && col < (f + 1) * K2) {
              const int k = col - f * K2;
              const float ev = __expf(lg[j][half * 2 + e] - mx);
              const float4 sv = tile_f[(k / K) * TWs + (k % K)];
              sum += ev;
              r = fmaf(ev, sv.x, r); gg = fmaf(ev, sv.y, gg); bb = fmaf(ev, sv.z, bb);
            }
          }
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
          const int img = ((b / p.ipt) * T + f) * p.ipt + (b % p.ipt);
          const float inv = 1.f / sum;
          float* o = reinterpret_cast<float*>(p.out.ptr) + p.out.pix(img, y, x) * p.out.cstride + p.out.coff;
          o[0] = r * inv; o[1] = gg * inv; o[2] = bb * inv;
        }
      }
    }
  }
}

template <int K, int T>
static int launch_post_kp(dd_ctx* ctx, const PostKpParams& p, cudaStream_t s) {
  constexpr int K2 = K * K, O16 = (T * K2 + 15) / 16 * 16, PAD = (K - 1) / 2;
  const size_t smem = 128ull * (p.Cpad * 2 + 16) + static_cast<size_t>(O16) * (p.Cpad * 2 + 16) +
                      static_cast<size_t>(O16) * (O16 * 2 + 16) + 2ull * O16 * sizeof(float) +
                      static_cast<size_t>(T) * (kPkTileH + 2 * PAD) * (kPkTileW + 2 * PAD) * sizeof(float4);
  DD_CHECK_ARG(smem <= ctx->max_smem_optin, "post_kp: tile does not fit shared memory");
  DD_CUDA(cudaFuncSetAttribute(post_kp_fused_kernel<K, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  long long grid = static_cast<long long>(ctx->sm_count) * 4;
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

size_t dd_post_kp_weights_bytes(int cin, int ksize, int features) {
  const int o16 = round_up(features * ksize * ksize, 16), cpad = round_up(cin, 16);
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
  const int o = features * ksize * ksize, o16 = round_up(o, 16), cpad = round_up(cin, 16);
  memset(blob_host, 0, dd_post_kp_weights_bytes(cin, ksize, features));
  __half* w1t = reinterpret_cast<__half*>(blob_host);
  __half* w2t = w1t + static_cast<size_t>(o16) * cpad;
  float* b1p = reinterpret_cast<float*>(w2t + static_cast<size_t>(o16) * o16);
  float* b2p = b1p + o16;
  for (int c = 0; c < cin; ++c)
    for (int n = 0; n < o; ++n) w1t[static_cast<size_t>(n) * cpad + c] = __float2half_rn(w1[static_cast<size_t>(c) * o + n]);
  for (int k = 0; k < o; ++k)
    for (int n = 0; n < o; ++n) w2t[static_cast<size_t>(n) * o16 + k] = __float2half_rn(w2[static_cast<size_t>(k) * o + n]);
  for (int n = 0; n < o; ++n) { b1p[n] = b1[n]; b2p[n] = b2[n]; }
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
  const int o16 = round_up(features * ksize * ksize, 16);
  p.w1t = reinterpret_cast<const __half*>(blob_dev);
  p.w2t = p.w1t + static_cast<size_t>(o16) * p.Cpad;
  p.b1 = reinterpret_cast<const float*>(p.w2t + static_cast<size_t>(o16) * o16);
  p.b2 = p.b1 + o16;
  p.src = make_view(src); p.out = make_view(out);
  p.B = x->n; p.h = x->h; p.w = x->w; p.ipt = images_per_tuple;
  p.tiles_x = (x->w + kPkTileW - 1) / kPkTileW; p.tiles_y = (x->h + kPkTileH - 1) / kPkTileH;
  p.total_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * x->n;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ksize == 5 && features == 1) return launch_post_kp<5, 1>(ctx, p, s);
  if (ksize == 5 && features == 3) return launch_post_kp<5, 3>(ctx, p, s);
  if (ksize == 3 && features == 1) return launch_post_kp<3, 1>(ctx, p, s);
  return launch_post_kp<3, 3>(ctx, p, s);
}

}  // extern "C"
