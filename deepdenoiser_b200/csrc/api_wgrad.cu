// Tensor-core weight gradient of the 3x3 / 1x1 stride-1 'SAME' convolutions (fp16 activations and activation
// gradients, fp32 accumulate) + the device-side weight (re)packing and the space-to-depth helper of the mixed-precision
// training path.  Replaces what TF autodiff derives for tf.layers.conv2d (Conv2DBackpropFilter), reached from
// Training.py:700-702 (optimizer.minimize) for the layers of UNet.py:29-31,56-58 / Tiramisu.py:35-37,50-52,77-79 /
// Architecture.py:238-243 / MultiScalePrediction.py:64-66,73-75.
//
//   dW[r,s,ci,co] += scale * sum_{n,j,i} x[n, j+r-1, i+s-1, ci] * dz[n, j, i, co]          (TF layout [kh,kw,cin,cout])
//
// Mapping onto tcgen05 (the reduction dimension K of the GEMM is the PIXEL axis, so both operands are "MN-major":
// a pixel is one 128-byte line of 64 channels - exactly what a 128B-swizzled TMA box of an NHWC tensor looks like):
//   work item      (64-channel cin chunk ic, 64-channel cout chunk oc, contiguous range of 128-pixel row strips)
//   B operand      x row b of chunk ic, box 64ch x 130px (halo; OOB -> 0 = SAME padding).  The three horizontal taps are
//                  the same slot shifted by s pixels = s*128 B: with LBO = 128 B they are STACKED along N (N = 192).
//   A operand      two consecutive dz rows of chunk oc (adjacent ring slots, LBO = slot size): M = 128 = 2 x 64 cout.
//   instructions   per x row b and 16-pixel k-step:  P += [dz_b ; dz_b+1] x_b   (taps r = 1 | r = 0 in the two lane halves)
//                                                    Q += [dz_b-1 ; dz_b] x_b   (tap r = 2 | duplicate, ignored)
//                  i.e. 3 of the 4 lane halves are useful (a 64x64 tap product fills half of the M = 128 datapath and
//                  the 9 taps of a chunk pair need 9 * 64 columns > the 512 of TMEM when laid out densely).
//   accumulators   P: TMEM columns [0,192), Q: [256,448); kept for the whole CTA, flushed once with fp32 atomics.
//   dz ring        rz slots + one tail slot that mirrors slot 0, so the pair (slot rz-1, slot 0) is contiguous too.
//   warps          0: x producer  1: dz producer  2: MMA issuer  3: TMEM allocator  4-7: epilogue
#include <string.h>

#include "dd_ptx.cuh"
#include "dd_internal.h"

namespace dd {

constexpr int kWgrThreads = 256;
constexpr int kWgrTileW = 128;
constexpr int kWgrKSteps = kWgrTileW / 16;
constexpr uint32_t kWgrColP = 0, kWgrColQ = 256;

struct WgrMaps {
  CUtensorMap x;    // [C, W, H, N] fp16, box [64, 130 | 128, 1, 1]
  CUtensorMap dz;   // [C, W, H, N] fp16, box [64, 128, 1, 1]
};

struct WgrParams {
  int N, H, W, strips;
  long long total_rows;      // N * strips * H
  int n_ic, n_oc, parts, rows_per_part;
  int k3;                    // 1: 3x3, 0: 1x1
  int cin, cout;
  int rx, rz;
  uint32_t x_slot_bytes, x_tx, dz_slot_bytes, dz_tx;
  uint32_t x_off, dz_off, bar_off;
  float* dw;
  int layout;                // 0: [tap][cin][cout]   1: [tap][cout][cin]
  float scale;
  int bf16;                  // operands are bfloat16
};

struct WgrSegment { int n, x0, y0, y1; };
__device__ __forceinline__ uint64_t wg_desc_from(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

struct WgrWalker {
  long long lin, lin_end;
  int H, strips;
  __device__ WgrWalker(const WgrParams& p, int part) {
    lin = static_cast<long long>(part) * p.rows_per_part;
    lin_end = lin + p.rows_per_part;
    if (lin_end > p.total_rows) lin_end = p.total_rows;
    H = p.H; strips = p.strips;
  }
  __device__ bool next(WgrSegment& s) {
    if (lin >= lin_end) return false;
    const long long col = lin / H;
    s.y0 = static_cast<int>(lin - col * H);
    s.n = static_cast<int>(col / strips);
    s.x0 = static_cast<int>(col % strips) * kWgrTileW;
    const long long left = lin_end - lin;
    s.y1 = (s.y0 + left > H) ? H : static_cast<int>(s.y0 + left);
    lin += s.y1 - s.y0;
    return true;
  }
};

// MN-major operand, 128-byte swizzle: 64 channels (one 128 B line) per pixel, 8-pixel groups 1024 B apart (SBO),
// 64-channel chunks LBO apart.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgrThreads, 1)
wgrad_rows_kernel(const __grid_constant__ WgrMaps maps, const WgrParams p) {
  extern __shared__ __align__(1024) uint8_t wg_smem_raw[];
  const uint32_t smem_base = (smem_u32(wg_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = wg_smem_raw + (smem_base - smem_u32(wg_smem_raw));
  uint8_t* x_smem = smem + p.x_off;
  uint8_t* dz_smem = smem + p.dz_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* x_full = bars;                 // [rx]
  uint64_t* x_empty = x_full + p.rx;       // [rx]
  uint64_t* dz_full = x_empty + p.rx;      // [rz]
  uint64_t* dz_empty = dz_full + p.rz;     // [rz]
  uint64_t* acc_full = dz_empty + p.rz;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x / p.parts, part = blockIdx.x % p.parts;
  const int ic = pair / p.n_oc, oc = pair % p.n_oc;
  const bool has_work = static_cast<long long>(part) * p.rows_per_part < p.total_rows;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.dz);
    for (int i = 0; i < p.rx; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < p.rz; ++i) { mbar_init(&dz_full[i], 1); mbar_init(&dz_empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ x producer: one row of chunk ic per slot
    if (elect_one()) {
      WgrWalker walk(p, part);
      WgrSegment sg;
      int slot = 0; uint32_t phase = 0;
      while (walk.next(sg)) {
        for (int b = sg.y0; b < sg.y1; ++b) {
          mbar_wait(&x_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&x_full[slot], p.x_tx);
          tma_load_4d(x_smem + static_cast<size_t>(slot) * p.x_slot_bytes, &maps.x, &x_full[slot], ic * 64, sg.x0 - p.k3, b, sg.n);
          if (++slot == p.rx) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ dz producer: rows y0-1 .. y1 of chunk oc (3x3)
    if (elect_one()) {
      WgrWalker walk(p, part);
      WgrSegment sg;
      int slot = 0; uint32_t phase = 0;
      while (walk.next(sg)) {
        const int a0 = p.k3 ? sg.y0 - 1 : sg.y0, a1 = p.k3 ? sg.y1 : sg.y1 - 1;
        for (int a = a0; a <= a1; ++a) {
          mbar_wait(&dz_empty[slot], phase ^ 1);
          const bool dup = (slot == 0);
          mbar_arrive_expect_tx(&dz_full[slot], dup ? 2u * p.dz_tx : p.dz_tx);
          tma_load_4d(dz_smem + static_cast<size_t>(slot) * p.dz_slot_bytes, &maps.dz, &dz_full[slot], oc * 64, sg.x0, a, sg.n);
          if (dup)
            tma_load_4d(dz_smem + static_cast<size_t>(p.rz) * p.dz_slot_bytes, &maps.dz, &dz_full[slot], oc * 64, sg.x0, a, sg.n);
          if (++slot == p.rz) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one() && has_work) {
      const uint32_t n_cols = p.k3 ? 192u : 64u;
      const uint32_t idesc = make_idesc_f16(128, static_cast<int>(n_cols)) | (1u << 15) | (1u << 16) | (p.bf16 ? kIdescBf16 : 0u);   // A, B MN-major
      const uint32_t x_base = smem_u32(x_smem), dz_base = smem_u32(dz_smem);
      const uint64_t a_tmpl = make_desc_mn_sw128(0, p.dz_slot_bytes, 1024);
      const uint64_t b_tmpl = make_desc_mn_sw128(0, 128, 1024);
      const uint32_t a_lo = static_cast<uint32_t>(a_tmpl), a_hi = static_cast<uint32_t>(a_tmpl >> 32);
      const uint32_t b_lo = static_cast<uint32_t>(b_tmpl), b_hi = static_cast<uint32_t>(b_tmpl >> 32);
      WgrWalker walk(p, part);
      WgrSegment sg;
      int xs = 0; uint32_t xph = 0;
      // dz rows are numbered q = 0,1,2,... in load order: slot q % rz, use q / rz
      int q0 = 0;                  // q of the first dz row of the segment
      int wq = 0, wslot = 0; uint32_t wph = 0;      // next dz row to wait for
      int rel_slot = 0;            // slot of the next dz row to release (rows are released in load order)
      uint32_t acc = 0;
      while (walk.next(sg)) {
        const int rows = sg.y1 - sg.y0;
        for (int i = 0; i < rows; ++i) {
          const int need = p.k3 ? q0 + i + 2 : q0 + i;
          while (wq <= need) {
            mbar_wait(&dz_full[wslot], wph);
            ++wq; if (++wslot == p.rz) { wslot = 0; wph ^= 1; }
          }
          mbar_wait(&x_full[xs], xph);
          tc_fence_after();
          const uint32_t x_addr = x_base + static_cast<uint32_t>(xs) * p.x_slot_bytes;
          const int qa = p.k3 ? q0 + i + 1 : q0 + i;                    // dz row a == b
          const uint32_t sa = static_cast<uint32_t>(qa % p.rz), sb = static_cast<uint32_t>((qa + p.rz - 1) % p.rz);
          const uint32_t pa = dz_base + sa * p.dz_slot_bytes;           // [dz_b ; dz_b+1]
          const uint32_t qa_addr = dz_base + sb * p.dz_slot_bytes;      // [dz_b-1 ; dz_b]
#pragma unroll
          for (int k = 0; k < kWgrKSteps; ++k) {
            const uint64_t bd = wg_desc_from(b_lo + ((x_addr + k * 2048u) >> 4), b_hi);
            umma_f16(tmem_base + kWgrColP, wg_desc_from(a_lo + ((pa + k * 2048u) >> 4), a_hi), bd, idesc, acc);
            if (p.k3) umma_f16(tmem_base + kWgrColQ, wg_desc_from(a_lo + ((qa_addr + k * 2048u) >> 4), a_hi), bd, idesc, acc);
            acc = 1u;
          }
          umma_commit(&x_empty[xs]);
          if (++xs == p.rx) { xs = 0; xph ^= 1; }
          umma_commit(&dz_empty[rel_slot]);                             // row b-1 (3x3) / row b (1x1) is done
          if (++rel_slot == p.rz) rel_slot = 0;
        }
        if (p.k3) {                                                     // the last two rows of the segment
          umma_commit(&dz_empty[rel_slot]); if (++rel_slot == p.rz) rel_slot = 0;
          umma_commit(&dz_empty[rel_slot]); if (++rel_slot == p.rz) rel_slot = 0;
        }
        q0 += p.k3 ? rows + 2 : rows;
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4 && has_work) {
    // ------------------------------------------------------------------ epilogue: TMEM -> scaled fp32 atomics
    const int wq = warp - 4;
    const int half = wq >> 1;
    const int co = oc * 64 + (wq & 1) * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const int n_s = p.k3 ? 3 : 1;
    for (int blk = 0; blk < (p.k3 ? 2 : 1); ++blk) {
      int r;
      if (!p.k3) r = (half == 0) ? 0 : -1;
      else if (blk == 0) r = (half == 0) ? 1 : 0;
      else r = (half == 0) ? 2 : -1;
      if (r < 0) continue;                                              // warp-uniform
      for (int s = 0; s < n_s; ++s) {
        const int tap = p.k3 ? r * 3 + s : 0;
        for (int cb = 0; cb < 64; cb += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_lane + (blk ? kWgrColQ : kWgrColP) + static_cast<uint32_t>(s * 64 + cb), v);
          tmem_ld_wait();
          if (co < p.cout) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int ci = ic * 64 + cb + i;
              if (ci < p.cin) {
                const size_t idx = p.layout ? (static_cast<size_t>(tap) * p.cout + co) * p.cin + ci
                                            : (static_cast<size_t>(tap) * p.cin + ci) * p.cout + co;
                atomicAdd(p.dw + idx, p.scale * __uint_as_float(v[i]));
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*WgEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_nhwc_f16(dd_ctx* ctx, CUtensorMap* map, const dd_tensor* t, uint32_t box_w) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(t->c), static_cast<cuuint64_t>(t->w), static_cast<cuuint64_t>(t->h),
                        static_cast<cuuint64_t>(t->n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(t->cstride) * 2, static_cast<cuuint64_t>(t->w) * t->cstride * 2,
                           static_cast<cuuint64_t>(t->h) * t->w * t->cstride * 2};
  cuuint32_t box[4] = {64, box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* base = reinterpret_cast<__half*>(t->ptr) + t->coff;
  CUresult r = reinterpret_cast<WgEncodeTiledFn>(ctx->encode_tiled)(
      map, t->dtype == DD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("wgrad: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return DD_ERR_CUDA;
  }
  return DD_OK;
}

// x, dz: fp16 NHWC views of the same spatial size; dw fp32, accumulated.
int launch_wgrad_rows(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* dz, int ksize, int layout, float* dw, float scale,
                      cudaStream_t stream) {
  DD_CHECK_ARG(ctx->encode_tiled, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  DD_CHECK_ARG(is_half_type(x->dtype) && dz->dtype == x->dtype, "tensor-core wgrad needs two fp16 or two bf16 operands");
  DD_CHECK_ARG(x->coff % 8 == 0 && x->cstride % 8 == 0 && dz->coff % 8 == 0 && dz->cstride % 8 == 0,
               "wgrad operand views must be 16-byte aligned");
  DD_CHECK_ARG(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
  WgrParams p;
  memset(&p, 0, sizeof(p));
  p.N = x->n; p.H = x->h; p.W = x->w;
  p.strips = (p.W + kWgrTileW - 1) / kWgrTileW;
  p.total_rows = static_cast<long long>(p.N) * p.strips * p.H;
  p.k3 = (ksize == 3);
  p.cin = x->c; p.cout = dz->c;
  p.n_ic = (p.cin + 63) / 64; p.n_oc = (p.cout + 63) / 64;
  const int pairs = p.n_ic * p.n_oc;
  long long parts = ctx->sm_count / pairs;
  if (parts < 1) parts = 1;
  const long long min_rows = 8;
  if (parts > (p.total_rows + min_rows - 1) / min_rows) parts = (p.total_rows + min_rows - 1) / min_rows;
  p.rows_per_part = static_cast<int>((p.total_rows + parts - 1) / parts);
  p.parts = static_cast<int>((p.total_rows + p.rows_per_part - 1) / p.rows_per_part);
  p.rx = 4; p.rz = 6;
  const uint32_t box_w = p.k3 ? kWgrTileW + 2 : kWgrTileW;
  p.x_tx = box_w * 128u;
  p.x_slot_bytes = static_cast<uint32_t>(round_up(static_cast<int>(p.x_tx) + 256, 1024));   // + room for the s-shifted view of the last k-step
  p.dz_tx = kWgrTileW * 128u;
  p.dz_slot_bytes = p.dz_tx;
  p.x_off = 0;
  p.dz_off = static_cast<uint32_t>(p.rx) * p.x_slot_bytes;
  p.bar_off = p.dz_off + static_cast<uint32_t>(p.rz + 1) * p.dz_slot_bytes;
  const size_t smem = 1024 + p.bar_off + 512;
  DD_CHECK_ARG(smem <= ctx->max_smem_optin, "wgrad: shared memory plan does not fit");
  p.dw = dw; p.layout = layout; p.scale = scale; p.bf16 = (x->dtype == DD_BF16);
  WgrMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = encode_nhwc_f16(ctx, &maps.x, x, box_w);
  if (rc) return rc;
  rc = encode_nhwc_f16(ctx, &maps.dz, dz, kWgrTileW);
  if (rc) return rc;
  DD_CUDA(cudaFuncSetAttribute(wgrad_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  wgrad_rows_kernel<<<pairs * p.parts, kWgrThreads, smem, stream>>>(maps, p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

// ------------------------------------------------------------------------------------------------ device-side packing
// fp32 master weights -> the fp16 [chunk][s][r][cpad][64] layout of conv_rows_kernel (api_conv.cu::dd_conv2d_pack_weights
// is the host-side twin).  mode 0: forward of TF [k,k,cin,cout];  mode 1: input-gradient convolution of the same layer
// (taps flipped, channels swapped: a conv with cin' = cout, cout' = cin);  mode 2: transposed 2x2, TF [2,2,cout,cin] ->
// [chunk][sub-pixel][cpad][64].  The destination's padding must have been zeroed once.
struct PackParams { const float* w; uint16_t* dst; int ksize, cin, cout, cpad, mode, bf16; };
__device__ __forceinline__ uint16_t pack_cvt(float v, int bf16) {
  if (bf16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
}
__global__ void __launch_bounds__(256) pack_f16_kernel(const PackParams p) {
  const int k = p.ksize, k2 = k * k;
  const size_t total = static_cast<size_t>(k2) * p.cin * p.cout;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const float v = p.w[idx];
  if (p.mode == 2) {           // idx = (sp * cout + o) * cin + c
    const int c = static_cast<int>(idx % p.cin), o = static_cast<int>((idx / p.cin) % p.cout), sp = static_cast<int>(idx / (static_cast<size_t>(p.cin) * p.cout));
    p.dst[((static_cast<size_t>(c / 64) * k2 + sp) * p.cpad + o) * 64 + (c % 64)] = pack_cvt(v, p.bf16);
    return;
  }
  // idx = ((r * k + s) * cin + c) * cout + o
  const int o = static_cast<int>(idx % p.cout), c = static_cast<int>((idx / p.cout) % p.cin);
  const int tap = static_cast<int>(idx / (static_cast<size_t>(p.cin) * p.cout));
  const int r = tap / k, s = tap % k;
  if (p.mode == 0) {
    p.dst[(((static_cast<size_t>(c / 64) * k + s) * k + r) * p.cpad + o) * 64 + (c % 64)] = pack_cvt(v, p.bf16);
  } else {
    const int rr = k - 1 - r, ss = k - 1 - s;
    p.dst[(((static_cast<size_t>(o / 64) * k + ss) * k + rr) * p.cpad + c) * 64 + (o % 64)] = pack_cvt(v, p.bf16);
  }
}

// ------------------------------------------------------------------------------------------------ space to depth (+ ReLU mask)
// out[n,i,j, sp*C + c] = dy[n, 2i+ay, 2j+ax, c] * [y[n, 2i+ay, 2j+ax, c] > 0],  sp = 2*ay + ax: turns the stride-2 2x2
// transposed convolution's backward into 1x1 GEMMs on the coarse grid (UNet.py:56-58).
struct S2dParams { View dy, y, out; int has_y; };
__global__ void __launch_bounds__(256) s2d_mask_kernel(const S2dParams p) {
  const int C = p.dy.c;
  const size_t total = static_cast<size_t>(p.out.n) * p.out.h * p.out.w * 4 * C;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = static_cast<int>(idx % (4 * C));
  const size_t opix = idx / (4 * C);
  const int sp = ch / C, c = ch % C;
  const int j = static_cast<int>(opix % p.out.w), i = static_cast<int>((opix / p.out.w) % p.out.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.out.w) * p.out.h));
  const size_t ipix = p.dy.pix(n, 2 * i + (sp >> 1), 2 * j + (sp & 1));
  float v = p.dy.load(ipix, c);
  if (p.has_y && !(p.y.load(ipix, c) > 0.f)) v = 0.f;
  p.out.store(opix, ch, v);
}


// ------------------------------------------------------------------------------------------------ ReLU backward + bias gradient
// dz = dy * [y > 0] (optional mask, optional store) and db[c] += scale * sum_pixels dz[.., c] in ONE pass over the gradient:
// tf.nn.relu's backward and the BiasAddGrad of the layer below it.  Fast path: fp16 tensors, 8 channels (16 B) per thread.
struct ReluBiasParams { View dy, y, dz; float* db; int has_y, has_dz; float scale; };

__global__ void __launch_bounds__(256) relu_bias_vec_kernel(const ReluBiasParams p) {
  __shared__ float sdb[1024];
  const int C = p.dy.c, G = C >> 3;
  for (int i = threadIdx.x; i < C; i += 256) sdb[i] = 0.f;
  __syncthreads();
  const size_t stride = static_cast<size_t>(gridDim.x) * 256;      // a multiple of G: a thread keeps its channel group
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  size_t idx = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  const int g = static_cast<int>(idx % G);
  // the grid stride is a multiple of G: the pixel index advances by a constant, no 64-bit division per 16-byte load (that
  // division kept this pass at a quarter of the HBM rate)
  const size_t npix = static_cast<size_t>(p.dy.n) * p.dy.h * p.dy.w, pix_step = stride / G;
  const uint16_t* dyp = reinterpret_cast<const uint16_t*>(p.dy.ptr) + p.dy.coff + g * 8;
  const uint16_t* yp = reinterpret_cast<const uint16_t*>(p.y.ptr) + p.y.coff + g * 8;
  uint16_t* dzp = reinterpret_cast<uint16_t*>(p.dz.ptr) + p.dz.coff + g * 8;
  const int bf = p.dy.bf16;
  auto body = [&](size_t pix) {
    float d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dyp + pix * p.dy.cstride)), bf, d);
    if (p.has_y) {
      float yv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(yp + pix * p.y.cstride)), p.y.bf16, yv);
#pragma unroll
      for (int i = 0; i < 8; ++i) if (!(yv[i] > 0.f)) d[i] = 0.f;
    }
    if (p.has_dz) {
      const uint4 packed = pack8(d, p.dz.bf16);
      *reinterpret_cast<uint4*>(dzp + pix * p.dz.cstride) = packed;
      if (p.has_y) unpack8(packed, p.dz.bf16, d);      // the bias gradient sums what was stored (rounded values)
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += d[i];
  };
  size_t pix = idx / G;
  // two independent pixels per iteration: more bytes in flight per thread
  for (; pix + pix_step < npix; pix += 2 * pix_step) { body(pix); body(pix + pix_step); }
  if (pix < npix) body(pix);
  if (p.db) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sdb[g * 8 + i], acc[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) atomicAdd(p.db + i, p.scale * sdb[i]);
  }
}

__global__ void __launch_bounds__(256) relu_bias_generic_kernel(const ReluBiasParams p) {
  __shared__ float sdb[1024];
  const int C = p.dy.c;
  for (int i = threadIdx.x; i < C; i += 256) sdb[i] = 0.f;
  __syncthreads();
  const size_t total = static_cast<size_t>(p.dy.n) * p.dy.h * p.dy.w * C;
  const size_t stride = static_cast<size_t>(gridDim.x) * 256;      // a multiple of C
  size_t idx = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  const int c = static_cast<int>(idx % C);
  float acc = 0.f;
  for (; idx < total; idx += stride) {
    const size_t pix = idx / C;
    float d = p.dy.load(pix, c);
    if (p.has_y && !(p.y.load(pix, c) > 0.f)) d = 0.f;
    if (p.has_dz) p.dz.store(pix, c, d);
    acc += d;
  }
  if (p.db) {
    atomicAdd(&sdb[c], acc);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) atomicAdd(p.db + i, p.scale * sdb[i]);
  }
}


// fast path: fp16 tensors, 8 channels (16 bytes) per thread
__global__ void __launch_bounds__(256) s2d_mask_vec_kernel(const S2dParams p) {
  const int C = p.dy.c, G = C >> 3;
  const size_t total = static_cast<size_t>(p.out.n) * p.out.h * p.out.w * 4 * G;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g4 = static_cast<int>(idx % (4 * G));
  const size_t opix = idx / (4 * G);
  const int sp = g4 / G, g = g4 % G;
  const int j = static_cast<int>(opix % p.out.w), i = static_cast<int>((opix / p.out.w) % p.out.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.out.w) * p.out.h));
  const size_t ipix = p.dy.pix(n, 2 * i + (sp >> 1), 2 * j + (sp & 1));
  uint4 d = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.dy.ptr) + ipix * p.dy.cstride + p.dy.coff + g * 8);
  if (p.has_y) {
    float dv[8], yv[8];
    unpack8(d, p.dy.bf16, dv);
    unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.y.ptr) + ipix * p.y.cstride + p.y.coff + g * 8), p.y.bf16, yv);
#pragma unroll
    for (int i = 0; i < 8; ++i) if (!(yv[i] > 0.f)) dv[i] = 0.f;
    d = pack8(dv, p.dy.bf16);
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out.ptr) + opix * p.out.cstride + p.out.coff + sp * C + g * 8) = d;
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_conv2d_wgrad_tc(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* dz, int ksize, int layout, float* dw_dev, float scale,
                       void* stream) {
  DD_CHECK_ARG(ctx && dw_dev && tensor_ok(x) && tensor_ok(dz), "bad argument");
  DD_CHECK_ARG(x->n == dz->n && x->h == dz->h && x->w == dz->w, "wgrad: spatial dims differ");
  DD_CHECK_ARG(layout == 0 || layout == 1, "layout must be 0 ([tap][cin][cout]) or 1 ([tap][cout][cin])");
  return launch_wgrad_rows(ctx, x, dz, ksize, layout, dw_dev, scale, static_cast<cudaStream_t>(stream));
}

int dd_conv2d_pack_weights_dev(dd_ctx* ctx, const float* w_dev, int ksize, int cin, int cout, int mode, void* packed_dev,
                               void* stream) {
  DD_CHECK_ARG(ctx && w_dev && packed_dev && cin > 0 && cout > 0, "bad argument");
  const int bf16 = (mode & DD_PACK_BF16) ? 1 : 0;
  mode &= ~DD_PACK_BF16;
  DD_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0 (forward), 1 (input gradient) or 2 (transposed 2x2)");
  DD_CHECK_ARG(mode == 2 ? ksize == 2 : (ksize == 1 || ksize == 3), "bad kernel size for this mode");
  PackParams p;
  p.w = w_dev; p.dst = reinterpret_cast<uint16_t*>(packed_dev); p.ksize = ksize; p.cin = cin; p.cout = cout; p.mode = mode; p.bf16 = bf16;
  p.cpad = round_up(mode == 1 ? cin : cout, 32);
  const size_t total = static_cast<size_t>(ksize) * ksize * cin * cout;
  pack_f16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_space_to_depth2_mask(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* out, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(dy) && tensor_ok(out) && (!y || tensor_ok(y)), "bad argument");
  DD_CHECK_ARG(dy->n == out->n && dy->h == 2 * out->h && dy->w == 2 * out->w && out->c == 4 * dy->c,
               "space_to_depth2: out must be [n, h/2, w/2, 4c]");
  DD_CHECK_ARG(!y || (y->n == dy->n && y->h == dy->h && y->w == dy->w && y->c == dy->c), "space_to_depth2: bad mask tensor");
  S2dParams p;
  p.dy = make_view(dy); p.out = make_view(out); p.has_y = y ? 1 : 0;
  p.y = y ? make_view(y) : p.dy;
  const size_t total = static_cast<size_t>(out->n) * out->h * out->w * out->c;
  auto aligned = [](const dd_tensor* t) { return is_half_type(t->dtype) && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0; };
  if (aligned(dy) && aligned(out) && dy->dtype == out->dtype && (!y || aligned(y)))
    s2d_mask_vec_kernel<<<static_cast<unsigned>((total / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    s2d_mask_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

static int relu_bwd_bias_launch(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* dz, float* db_dev, float scale,
                                cudaStream_t stream) {
  ReluBiasParams p;
  p.dy = make_view(dy); p.y = y ? make_view(y) : p.dy; p.dz = dz ? make_view(dz) : p.dy;
  p.db = db_dev; p.has_y = y ? 1 : 0; p.has_dz = dz ? 1 : 0; p.scale = scale;
  auto aligned = [](const dd_tensor* t) { return is_half_type(t->dtype) && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0; };
  const bool vec = aligned(dy) && (!y || aligned(y)) && (!dz || aligned(dz));
  const int per = vec ? dy->c / 8 : dy->c;                       // grid stride must be a multiple of this
  const size_t total = static_cast<size_t>(dy->n) * dy->h * dy->w * per;
  size_t blocks = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  blocks = (blocks + per - 1) / per * per;
  if (vec) relu_bias_vec_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  else relu_bias_generic_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_relu_bwd_bias(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* dz, float* db_dev, float scale,
                     void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(dy) && (!y || tensor_ok(y)) && (!dz || tensor_ok(dz)), "bad argument");
  DD_CHECK_ARG(dy->c <= 1024, "relu_bwd_bias: at most 1024 channels");
  DD_CHECK_ARG(!y || (y->n == dy->n && y->h == dy->h && y->w == dy->w && y->c == dy->c), "relu_bwd_bias: y differs from dy");
  DD_CHECK_ARG(!dz || (dz->n == dy->n && dz->h == dy->h && dz->w == dy->w && dz->c == dy->c), "relu_bwd_bias: dz differs from dy");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // channel counts like 25 or 441 (K*K logits) in 16-byte aligned 16-bit buffers: the first c & ~7 channels take the
  // vectorised kernel, the remaining <= 7 the generic one
  auto base_ok = [](const dd_tensor* t) { return is_half_type(t->dtype) && t->coff % 8 == 0 && t->cstride % 8 == 0; };
  const int head = dy->c & ~7;
  if (dy->c % 8 != 0 && head >= 8 && base_ok(dy) && (!y || base_ok(y)) && (!dz || base_ok(dz))) {
    dd_tensor a = *dy, b, c;
    if (y) b = *y;
    if (dz) c = *dz;
    a.c = head; if (y) b.c = head; if (dz) c.c = head;
    int rc = relu_bwd_bias_launch(ctx, &a, y ? &b : nullptr, dz ? &c : nullptr, db_dev, scale, s);
    if (rc) return rc;
    const int tail = dy->c - head;
    a.coff += head; a.c = tail;
    if (y) { b.coff += head; b.c = tail; }
    if (dz) { c.coff += head; c.c = tail; }
    return relu_bwd_bias_launch(ctx, &a, y ? &b : nullptr, dz ? &c : nullptr, db_dev ? db_dev + head : nullptr, scale, s);
  }
  return relu_bwd_bias_launch(ctx, dy, y, dz, db_dev, scale, s);
}

}  // extern "C"
