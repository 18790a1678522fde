// C-ABI: kernel-prediction apply (KernelPrediction.py:11-63) — softmax over the K*K logits of a pixel, then the weighted
// sum of the symmetric-padded K x K neighbourhood of the noisy source, the same weights for r, g and b.
//
// HBM-bound: K*K*sizeof(logit) + 12 B source + 12 B output per pixel.  The logits are the only large stream, so they
// are moved by TMA (cp.async.bulk.tensor) as [8 rows x 32 px x 128 B] boxes into 128B-swizzled shared memory while the
// CUDA cores only ever touch shared memory:
//   * one thread per pixel; its 128-byte logit row is read with LDS.128 (the swizzle spreads the eight threads of a
//     quarter warp over all banks), exp'd once, and every tap is one LDS.128 of the float4 source tile (rgb + 1.0, so
//     the softmax denominator is the fourth accumulator lane) and two packed FFMA2;
//   * K*K > 32 (fp32) / 64 (fp16) logits stream through a 2-stage ring of 128-byte channel chunks with an online
//     softmax (running maximum, accumulators rescaled when it grows);
//   * tap offsets are a table in the kernel parameter block: the tap index is warp-uniform, so the lookup runs on the
//     uniform datapath and the tile address is register + uniform register.
#include <string.h>

#include "dd_internal.h"
#include "dd_ptx.cuh"

namespace dd {

__device__ __forceinline__ int kp_sym_index(int i, int n) {
  while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - i - 1);
  return i;
}

// ------------------------------------------------------------------------------------------------ generic fallback
// LANES threads cooperate on one pixel, scalar logit loads.  Used only when the logits view cannot be described by a
// tensor map (pointer / pixel stride not 16-byte aligned).
constexpr int kKpTileH = 8, kKpTileW = 32, kKpThreads = 256;
struct KpParams {
  View src, logits, out;
  int K, F, ipt;     // kernel size, features per logits tensor, images per tuple
};

template <int LANES>
__global__ void __launch_bounds__(kKpThreads) kernel_predict_kernel(const KpParams p) {
  extern __shared__ float4 s_src[];
  const int K = p.K, K2 = K * K, pad = (K - 1) / 2;
  const int TW = kKpTileW + 2 * pad, TH = kKpTileH + 2 * pad;
  // logits image b = tuple * ipt + n, feature f  ->  src/out image (tuple * F + f) * ipt + n
  const int b = blockIdx.z / p.F, f = blockIdx.z % p.F;
  const int img = ((b / p.ipt) * p.F + f) * p.ipt + (b % p.ipt);
  const int ty0 = blockIdx.y * kKpTileH, tx0 = blockIdx.x * kKpTileW;
  const int h = p.src.h, w = p.src.w;
  for (int i = threadIdx.x; i < TW * TH; i += kKpThreads) {
    const int ly = i / TW, lx = i % TW;
    const int yy = kp_sym_index(ty0 + ly - pad, h), xx = kp_sym_index(tx0 + lx - pad, w);
    const size_t sp = p.src.pix(img, yy, xx);
    s_src[i] = make_float4(p.src.load(sp, 0), p.src.load(sp, 1), p.src.load(sp, 2), 0.f);
  }
  __syncthreads();
  constexpr int PIX_PER_PASS = kKpThreads / LANES;
  const int sub = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  const int coff = f * K2;
  for (int pp = grp; pp < kKpTileH * kKpTileW; pp += PIX_PER_PASS) {
    const int ly = pp / kKpTileW, lx = pp % kKpTileW;
    const int y = ty0 + ly, x = tx0 + lx;
    const bool valid = (y < h) && (x < w);   // uniform across the LANES of a pixel
    const size_t lpix = valid ? p.logits.pix(b, y, x) : 0;
    float mx = -INFINITY;
    if (valid) for (int t = sub; t < K2; t += LANES) mx = fmaxf(mx, p.logits.load(lpix, coff + t));
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f, r = 0.f, g = 0.f, bl = 0.f;
    if (valid) {
      for (int t = sub; t < K2; t += LANES) {
        const float e = __expf(p.logits.load(lpix, coff + t) - mx);
        const int i = t / K, j = t - i * K;
        const float4 sv = s_src[(ly + i) * TW + lx + j];
        sum += e;
        r = fmaf(e, sv.x, r); g = fmaf(e, sv.y, g); bl = fmaf(e, sv.z, bl);
      }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      r += __shfl_xor_sync(0xffffffffu, r, o);
      g += __shfl_xor_sync(0xffffffffu, g, o);
      bl += __shfl_xor_sync(0xffffffffu, bl, o);
    }
    if (valid && sub == 0) {
      const float inv = 1.f / sum;
      const size_t op = p.out.pix(img, y, x);
      p.out.store(op, 0, r * inv); p.out.store(op, 1, g * inv); p.out.store(op, 2, bl * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------ TMA-streamed kernel
// Persistent and warp specialised: warp 8 is the producer (one elected lane keeps a ring of logit boxes in flight, tile
// after tile), warps 0-7 are the consumers (warp = tile row, lane = pixel).  The consumers prefetch the source tile of
// their NEXT tile with 4-byte cp.async while they work on the current one.
constexpr int kKpConsumers = 256;
constexpr int kKpTmaThreads = kKpConsumers + 32;
constexpr int kKpStageBytes = kKpTileH * kKpTileW * 128;   // one 128-byte channel chunk of the 256 pixels of a tile
constexpr int kKpMaxK = 31;
constexpr int kKpTapTable = kKpMaxK * kKpMaxK + 64 + 3;    // slack: the last (partial) group of 32 indexes past the taps
constexpr float kLog2e = 1.4426950408889634f;

struct KpTmaParams {
  View src, out;
  int K, K2, F, ipt, H;   // kernel size, taps, features per logits tensor, images per tuple, rows per logits image
  int coff;               // first logit channel of feature 0 inside the pixel row
  int nstages;            // logit ring depth
  int nsrc;               // source tile buffers (2: the next tile is prefetched during the current one)
  int tiles_x, tiles_y;
  int total_tiles;
  uint16_t tap[kKpTapTable];   // tap k -> byte offset 16 * ((k / K) * TWs + (k % K)) inside the source tile
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kKpConsumers) : "memory"); }

// 32 consecutive logits of this thread's pixel, starting at 16-byte chunk `j0` of its swizzled 128-byte row
template <bool HALF>
__device__ __forceinline__ void kp_load32(const uint8_t* row, int sw, int j0, float (&v)[32]) {
  if (HALF) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 raw = *reinterpret_cast<const uint4*>(row + (((j0 + j) ^ sw) << 4));
      const __half2* hh = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f2 = __half22float2(hh[e]);
        v[j * 8 + 2 * e] = f2.x; v[j * 8 + 2 * e + 1] = f2.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 raw = *reinterpret_cast<const float4*>(row + (((j0 + j) ^ sw) << 4));
      v[j * 4] = raw.x; v[j * 4 + 1] = raw.y; v[j * 4 + 2] = raw.z; v[j * 4 + 3] = raw.w;
    }
  }
}

struct KpTile {
  int tx0, ty0, b, f, img;
};
// Walks the tiles lin = blockIdx.x, + gridDim.x, ... without per-tile divisions: the stride is decomposed once into
// (dz, dy, dx) and added with carries.
struct KpTileWalker {
  int tx, ty, z, dx, dy, dz, left;
  __device__ KpTileWalker(const KpTmaParams& p) {
    const int lin = blockIdx.x;
    int r = lin / p.tiles_x;
    tx = lin - r * p.tiles_x; z = r / p.tiles_y; ty = r - z * p.tiles_y;
    const int g = gridDim.x;
    r = g / p.tiles_x;
    dx = g - r * p.tiles_x; dz = r / p.tiles_y; dy = r - dz * p.tiles_y;
    left = (lin < p.total_tiles) ? (p.total_tiles - lin + g - 1) / g : 0;   // tiles of this CTA
  }
  __device__ __forceinline__ bool valid() const { return left > 0; }
  __device__ __forceinline__ KpTile get(const KpTmaParams& p) const {
    KpTile t;
    t.tx0 = tx * kKpTileW; t.ty0 = ty * kKpTileH;
    if (p.F == 1) {
      t.b = z; t.f = 0; t.img = z;
    } else {
      t.b = z / p.F; t.f = z - t.b * p.F;
      // logits image b = tuple * ipt + n, feature f  ->  src/out image (tuple * F + f) * ipt + n
      t.img = ((t.b / p.ipt) * p.F + t.f) * p.ipt + (t.b % p.ipt);
    }
    return t;
  }
  __device__ __forceinline__ void advance(const KpTmaParams& p) {
    --left;
    tx += dx; if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    ty += dy; if (ty >= p.tiles_y) { ty -= p.tiles_y; ++z; }
    z += dz;
  }
};

// KS > 0: kernel size known at compile time (tap offsets are immediates, no predicates); KS == 0: table driven.
template <bool HALF, int KS>
__global__ void __launch_bounds__(kKpTmaThreads, 2)
kernel_predict_tma_kernel(const __grid_constant__ CUtensorMap lmap, const __grid_constant__ KpTmaParams p) {
  extern __shared__ uint8_t kp_smem_raw[];
  constexpr int EPC = HALF ? 64 : 32;       // logits per 128-byte chunk
  constexpr int ALIGN = HALF ? 8 : 4;       // logits per 16 bytes: TMA box origins must be 16-byte aligned
  constexpr int SUBS = EPC / 32;
  const int K = KS ? KS : p.K;
  const int K2 = KS ? KS * KS : p.K2;
  const int pad = (K - 1) / 2;
  const int TWs = kKpTileW + 2 * pad, THs = kKpTileH + 2 * pad;
  const int nstages = p.nstages;
  const uint32_t base_u32 = (smem_u32(kp_smem_raw) + 1023u) & ~1023u;
  uint8_t* base = kp_smem_raw + (base_u32 - smem_u32(kp_smem_raw));
  float4* s_src = reinterpret_cast<float4*>(base + nstages * kKpStageBytes);      // [nsrc][THs][TWs]
  float* s_out = reinterpret_cast<float*>(s_src + p.nsrc * TWs * THs);              // [8 warps][96]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_out + kKpConsumers * 3);          // [nstages]
  uint64_t* empty = full + nstages;                                                 // [nstages]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tma_prefetch_desc(&lmap);
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kKpTileH); }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kKpTileH) {
    // ---------------------------------------------------------------- producer: (tile, chunk) items through the ring
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (KpTileWalker walk(p); walk.valid(); walk.advance(p)) {
        const KpTile t = walk.get(p);
        const int first = p.coff + t.f * K2;
        const int c0 = first & ~(ALIGN - 1);
        const int nchunks = (first - c0 + K2 + EPC - 1) / EPC;
        const int row0 = t.b * p.H + t.ty0;
        for (int c = 0; c < nchunks; ++c) {
          while (!mbar_try_wait(&empty[stage], phase ^ 1)) __nanosleep(64);   // polls must not steal issue slots
          mbar_arrive_expect_tx(&full[stage], kKpStageBytes);
          tma_load_3d(base + stage * kKpStageBytes, &lmap, &full[stage], c0 + c * EPC, t.tx0, row0);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int h = p.src.h, w = p.src.w;
  const float* sp = reinterpret_cast<const float*>(p.src.ptr);
  const int n_src = TWs * THs;
  // the 1.0 lane of the source tiles (softmax denominator) is written once; cp.async only ever fills rgb
  for (int i = tid; i < p.nsrc * n_src; i += kKpConsumers) s_src[i].w = 1.f;
  // Static kernel sizes: the (row, column) of the <= 3 tile elements this thread fetches never change.
  constexpr int NE = KS ? ((kKpTileW + KS - 1) * (kKpTileH + KS - 1) + kKpConsumers - 1) / kKpConsumers : 1;
  int e_ly[NE], e_lx[NE];
#pragma unroll
  for (int j = 0; j < NE; ++j) {
    const int i = tid + j * kKpConsumers;
    e_ly[j] = i / TWs - pad; e_lx[j] = i - (i / TWs) * TWs - pad;
  }
  auto gather_src = [&](const KpTile& t, float4* dst) {
    const float* img_base = sp + static_cast<size_t>(t.img) * h * w * p.src.cstride + p.src.coff;
    if (KS) {
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        const int i = tid + j * kKpConsumers;
        if (i < n_src) {
          const int yy = kp_sym_index(t.ty0 + e_ly[j], h), xx = kp_sym_index(t.tx0 + e_lx[j], w);
          const float* q = img_base + (yy * w + xx) * p.src.cstride;
          float* d = reinterpret_cast<float*>(dst + i);
          cp_async_4(d, q); cp_async_4(d + 1, q + 1); cp_async_4(d + 2, q + 2);
        }
      }
    } else {
      for (int i = tid; i < n_src; i += kKpConsumers) {
        const int ly = i / TWs, lx = i - ly * TWs;
        const int yy = kp_sym_index(t.ty0 + ly - pad, h), xx = kp_sym_index(t.tx0 + lx - pad, w);
        const float* q = img_base + (yy * w + xx) * p.src.cstride;
        float* d = reinterpret_cast<float*>(dst + i);
        cp_async_4(d, q); cp_async_4(d + 1, q + 1); cp_async_4(d + 2, q + 2);
      }
    }
  };
  const int sw = lane & 7;
  float* my_out = s_out + warp * (kKpTileW * 3);
  float* outp = reinterpret_cast<float*>(p.out.ptr);
  int stage = 0; uint32_t phase = 0;
  int sbuf = 0;
  KpTileWalker walk(p);
  KpTile t_next = walk.get(p);
  if (walk.valid()) gather_src(t_next, s_src);
  while (walk.valid()) {
    const KpTile t = t_next;
    walk.advance(p);
    const bool has_next = walk.valid();
    if (has_next) t_next = walk.get(p);
    cp_async_wait_all();
    consumer_barrier();                  // source tile of this tile complete; everyone is done with the previous tile
    if (p.nsrc == 2 && has_next) gather_src(t_next, s_src + (sbuf ^ 1) * n_src);
    const uint8_t* tile_b = reinterpret_cast<const uint8_t*>(s_src + sbuf * n_src + warp * TWs + lane);
    float2 acc_rg = make_float2(0.f, 0.f), acc_bs = make_float2(0.f, 0.f);
    float m = -INFINITY;
    // 32 logits of the pixel: running maximum, rescale, exp, taps.  kb = tap index of element 0 (warp-uniform).
    auto group32 = [&](const uint8_t* row, int sub, int kb) {
      const int nvalid = K2 - kb;
      if (nvalid <= 0 || kb <= -32) return;
      float v[32];
      kp_load32<HALF>(row, sw, sub * 4, v);
      if (kb >= 0 && nvalid >= 32) {
        float cm = v[0];
#pragma unroll
        for (int e = 1; e < 32; ++e) cm = fmaxf(cm, v[e]);
        const float mn = fmaxf(m, cm);
        const float sc = ex2_approx((m - mn) * kLog2e);
        acc_rg.x *= sc; acc_rg.y *= sc; acc_bs.x *= sc; acc_bs.y *= sc;
        m = mn;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float ex = ex2_approx((v[e] - mn) * kLog2e);
          const int off = KS ? 16 * (((kb + e) / K) * TWs + (kb + e) % K) : p.tap[kb + e];
          const float4 sv = *reinterpret_cast<const float4*>(tile_b + off);
          acc_rg = __ffma2_rn(make_float2(ex, ex), make_float2(sv.x, sv.y), acc_rg);
          acc_bs = __ffma2_rn(make_float2(ex, ex), make_float2(sv.z, sv.w), acc_bs);
        }
      } else {
        float cm = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (static_cast<unsigned>(kb + e) < static_cast<unsigned>(K2)) cm = fmaxf(cm, v[e]);
        const float mn = fmaxf(m, cm);
        const float sc = ex2_approx((m - mn) * kLog2e);
        acc_rg.x *= sc; acc_rg.y *= sc; acc_bs.x *= sc; acc_bs.y *= sc;
        m = mn;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (static_cast<unsigned>(kb + e) < static_cast<unsigned>(K2)) {
            const float ex = ex2_approx((v[e] - mn) * kLog2e);
            const int off = KS ? 16 * (((kb + e) / K) * TWs + (kb + e) % K) : p.tap[kb + e];
            const float4 sv = *reinterpret_cast<const float4*>(tile_b + off);
            acc_rg = __ffma2_rn(make_float2(ex, ex), make_float2(sv.x, sv.y), acc_rg);
            acc_bs = __ffma2_rn(make_float2(ex, ex), make_float2(sv.z, sv.w), acc_bs);
          }
        }
      }
    };
    auto chunk_done = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);      // this warp is done with the stage
      if (++stage == nstages) { stage = 0; phase ^= 1; }
    };
    // the first logit of this feature sits `shift` elements into the first (16-byte aligned) box
    const int shift = (p.coff + t.f * K2) & (ALIGN - 1);
    if (KS != 0 && shift == 0) {
      // every index below is a compile-time constant: tap offsets become immediates, no predicates
      constexpr int NCH = KS ? (KS * KS + EPC - 1) / EPC : 1;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        mbar_wait(&full[stage], phase);
        const uint8_t* row = base + stage * kKpStageBytes + tid * 128;
#pragma unroll
        for (int sub = 0; sub < SUBS; ++sub) group32(row, sub, c * EPC + sub * 32);
        chunk_done();
      }
    } else {
      const int nchunks = (shift + K2 + EPC - 1) / EPC;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[stage], phase);
        const uint8_t* row = base + stage * kKpStageBytes + tid * 128;
#pragma unroll
        for (int sub = 0; sub < SUBS; ++sub) group32(row, sub, c * EPC + sub * 32 - shift);
        chunk_done();
      }
    }
    // the 32 pixels of the warp leave as one contiguous run of 96 floats
    const float inv = 1.f / acc_bs.y;
    const int y = t.ty0 + warp;
    if (p.out.cstride == 3) {
      my_out[lane * 3] = acc_rg.x * inv; my_out[lane * 3 + 1] = acc_rg.y * inv; my_out[lane * 3 + 2] = acc_bs.x * inv;
      __syncwarp();
      if (y < h) {
        float* orow = outp + p.out.pix(t.img, y, t.tx0) * 3;
        const int nfl = ((w - t.tx0 < kKpTileW) ? (w - t.tx0) : kKpTileW) * 3;
#pragma unroll
        for (int q = 0; q < 3; ++q)
          if (q * 32 + lane < nfl) orow[q * 32 + lane] = my_out[q * 32 + lane];
      }
      __syncwarp();
    } else if (y < h && t.tx0 + lane < w) {
      const size_t op = p.out.pix(t.img, y, t.tx0 + lane);
      p.out.store(op, 0, acc_rg.x * inv); p.out.store(op, 1, acc_rg.y * inv); p.out.store(op, 2, acc_bs.x * inv);
    }
    if (p.nsrc == 2) {
      sbuf ^= 1;
    } else if (has_next) {
      consumer_barrier();                // single source buffer: everyone must be done before it is refilled
      gather_src(t_next, s_src);
    }
  }
}

typedef CUresult (*KpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline bool same_spatial(const dd_tensor* a, const dd_tensor* b) { return a->n == b->n && a->h == b->h && a->w == b->w; }

template <bool HALF, int KS>
static int launch_kp_tma(dd_ctx* ctx, const CUtensorMap& map, const KpTmaParams& p, size_t smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    DD_CUDA(cudaFuncSetAttribute(kernel_predict_tma_kernel<HALF, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(ctx->max_smem_optin)));
    configured = true;
  }
  // persistent grid: as many CTAs as fit (shared memory bound), every CTA strides over the tiles
  int per_sm = static_cast<int>((228 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  long long grid = static_cast<long long>(ctx->sm_count) * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  kernel_predict_tma_kernel<HALF, KS><<<static_cast<unsigned>(grid), kKpTmaThreads, smem, s>>>(map, p);
  return DD_OK;
}

template <bool HALF>
static int dispatch_kp_tma(dd_ctx* ctx, int ksize, const CUtensorMap& map, const KpTmaParams& p, size_t smem, cudaStream_t s) {
  switch (ksize) {
    case 3: return launch_kp_tma<HALF, 3>(ctx, map, p, smem, s);
    case 5: return launch_kp_tma<HALF, 5>(ctx, map, p, smem, s);
    case 7: return launch_kp_tma<HALF, 7>(ctx, map, p, smem, s);
    default: return launch_kp_tma<HALF, 0>(ctx, map, p, smem, s);
  }
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_kernel_predict_fwd(dd_ctx* ctx, const dd_tensor* src, const dd_tensor* logits, int ksize, int features,
                          int images_per_tuple, const dd_tensor* out, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(src) && tensor_ok(logits) && tensor_ok(out), "bad argument");
  DD_CHECK_ARG(ksize >= 1 && (ksize & 1) && ksize <= kKpMaxK, "kernel size must be odd and <= 31");
  DD_CHECK_ARG(features >= 1 && logits->c == features * ksize * ksize, "logits must have features*K*K channels");
  DD_CHECK_ARG(src->c == 3 && out->c == 3 && src->dtype == DD_F32 && out->dtype == DD_F32, "src/out must be fp32 rgb");
  DD_CHECK_ARG(src->n == logits->n * features && out->n == src->n, "src/out batch must be logits.n * features");
  DD_CHECK_ARG(src->h == logits->h && src->w == logits->w && same_spatial(src, out), "spatial dims differ");
  DD_CHECK_ARG(images_per_tuple >= 1 && logits->n % images_per_tuple == 0, "logits.n must be a multiple of images_per_tuple");
  const int tiles_x = (src->w + kKpTileW - 1) / kKpTileW, tiles_y = (src->h + kKpTileH - 1) / kKpTileH;
  const int pad = (ksize - 1) / 2;
  const int TWs = kKpTileW + 2 * pad, THs = kKpTileH + 2 * pad;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int k2 = ksize * ksize;
  DD_CHECK_ARG(logits->dtype == DD_F32 || logits->dtype == DD_F16, "kernel_predict: logits must be fp32 or fp16");
  const bool half = logits->dtype == DD_F16;
  const size_t es = elem_size(logits->dtype);
  const bool tma_ok = ctx->encode_tiled && (reinterpret_cast<uintptr_t>(logits->ptr) % 16 == 0) &&
                      (static_cast<size_t>(logits->cstride) * es % 16 == 0) &&
                      static_cast<long long>(logits->n) * logits->h < (1ll << 31);
  if (tma_ok) {
    const int epc = static_cast<int>(128 / es);
    const int nchunks = (k2 + epc - 1) / epc;
    KpTmaParams p;
    memset(&p, 0, sizeof(p));
    p.src = make_view(src); p.out = make_view(out);
    p.K = ksize; p.K2 = k2; p.F = features; p.ipt = images_per_tuple; p.H = logits->h; p.coff = logits->coff;
    p.tiles_x = tiles_x; p.tiles_y = tiles_y;
    p.total_tiles = static_cast<long long>(tiles_x) * tiles_y * src->n;
    for (int k = 0; k < k2; ++k) p.tap[k] = static_cast<uint16_t>(((k / ksize) * TWs + (k % ksize)) * 16);
    // shared memory plan: two CTAs per SM.  Small kernels double-buffer the source tile; large ones keep one
    const size_t src_bytes = static_cast<size_t>(TWs) * THs * sizeof(float4);
    const size_t fixed = 1024 + kKpConsumers * 3 * sizeof(float) + 64;
    const size_t budget = (ctx->max_smem_optin < 113 * 1024) ? ctx->max_smem_optin : 113 * 1024;
    p.nstages = 2; p.nsrc = 2;
    if (fixed + 2 * kKpStageBytes + 2 * src_bytes > budget) p.nsrc = 1;
    (void)nchunks;
    CUtensorMap map;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(logits->cstride), static_cast<cuuint64_t>(logits->w),
                          static_cast<cuuint64_t>(logits->n) * logits->h};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(logits->cstride) * es,
                             static_cast<cuuint64_t>(logits->w) * logits->cstride * es};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(epc), kKpTileW, kKpTileH};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<KpEncodeTiledFn>(ctx->encode_tiled)(
        &map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, logits->ptr, dims, strides, box,
        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("kernel_predict: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
      return DD_ERR_CUDA;
    }
    const size_t smem = fixed + static_cast<size_t>(p.nstages) * kKpStageBytes + p.nsrc * src_bytes;
    DD_CHECK_ARG(smem <= ctx->max_smem_optin, "kernel_predict: tile does not fit shared memory");
    int rc = half ? dispatch_kp_tma<true>(ctx, ksize, map, p, smem, s) : dispatch_kp_tma<false>(ctx, ksize, map, p, smem, s);
    if (rc != DD_OK) return rc;
  } else {
    KpParams p;
    p.src = make_view(src); p.logits = make_view(logits); p.out = make_view(out);
    p.K = ksize; p.F = features; p.ipt = images_per_tuple;
    dim3 grid(tiles_x, tiles_y, src->n);
    DD_CHECK_ARG(src->n <= 65535 && tiles_y <= 65535, "kernel_predict grid too large");
    const size_t smem = static_cast<size_t>(TWs) * THs * sizeof(float4);
    if (k2 >= 64) kernel_predict_kernel<32><<<grid, kKpThreads, smem, s>>>(p);
    else kernel_predict_kernel<4><<<grid, kKpThreads, smem, s>>>(p);
  }
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // extern "C"
