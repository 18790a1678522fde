// MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93) as ONE tcgen05 kernel: a row-streaming, layer-pipelined
// implicit GEMM.  Every intermediate of the weight network lives in shared memory / TMEM; HBM sees the two fp32 rgb inputs
// and the fp32 rgb output only.
//
//   s_up = up2(small); x0 = relu(conv1x1_{6->24}(concat[s_up, large]))                              head (CUDA cores, fp32)
//   x1 = x0 + conv3x3(relu(conv3x3(relu(x0))));  x2 = x1 + conv3x3(relu(conv3x3(relu(x1))))         tcgen05, fp32 accumulate
//   w = sigmoid(relu(conv1x1_{24->1}(x2)));      out = large - w * up2(down2(large)) + w * s_up    tail + blend (fp32)
//
// Mapping onto the hardware
//   unit of work   one image ROW of a 128-pixel column strip, streamed top to bottom through all four 3x3 layers: every
//                  layer free-runs on its own barriers, trailing its producer by the two rows a 3x3 window needs.
//                  A strip yields 122 finished pixels per row (each 3x3 layer eats one pixel on both sides: 130 -> 122).
//   UMMA           M = 128 pixels, K = 16, N = 96: like conv_rows.cuh the three vertical taps are STACKED along N, so the
//                  input row t of a layer updates the accumulators of its output rows t+1, t, t-1 in one instruction
//                  (two when the accumulator ring wraps).  24 channels = two K steps, the upper 8 channels of the second
//                  step are read from an all-zero region (its chunk stride points there).
//   operand layout K-major WITHOUT swizzle, 8-row core matrices contiguous (SBO = 128 B): element (pixel m, 16-byte channel
//                  chunk j) sits at  slot + j * plane + 16 * m,  so the horizontal tap s is the same slot at +16*s bytes
//                  and the epilogue's stores (one 16-byte chunk per thread = per pixel) are conflict free.
//   accumulators   TMEM: 128 columns per layer = ring of four 32-column blocks (24 used); row number g lives in block
//                  3 - g % 4.  A drained block is re-zeroed with tcgen05.st, so every UMMA accumulates (no first-touch split).
//   warps          0-3 / 4-7 / 8-11 / 12-15: epilogue of layer 1 / 2 / 3 / 4 (one TMEM lane = one pixel per thread):
//                  + bias, residual, ReLU, fp16 -> next layer's operand row in shared memory; layer 4 adds the tail + blend
//                  16-20: head (x0 rows from the fp32 inputs)      21-24: one UMMA-issuing thread per layer (21 also allocates TMEM)
//   grid           persistent: the N * strips * H output rows are split into gridDim contiguous ranges; a range adds 4 halo
//                  rows above and below (recomputed, not exchanged), clipped at the image border where SAME padding applies.
#pragma once
#include "dd_internal.h"
#include "dd_ptx.cuh"

namespace dd {

constexpr int kCrC = 24;                      // channels of the compose net
constexpr int kCrValid = 122;                 // finished pixels per strip row
constexpr int kCrRowPx = 130;                 // pixels of an operand row slot that are used
constexpr int kCrSlotPx = 136;
constexpr int kCrPlane = kCrSlotPx * 16;      // one 8-channel chunk of a row: 2176 B
constexpr int kCrSlot = 3 * kCrPlane;         // 6528 B
constexpr int kCrWChunk = 96 * 16;            // one 8-channel chunk of a weight tile: rows n = tap r * 32 + cout
constexpr int kCrWTile = 3 * kCrWChunk;       // 4608 B per (layer, horizontal tap s): input channels 0..23
// the centre column's tile (s = 1) has a 4th chunk: K index 24 / 25 of tap r = 1 holds the layer's bias (hi / lo halves) - the
// activations' 4th chunk is the constant (1, 1, 0, ...), so the tensor core adds the bias; the other tiles' 4th chunk is zeros
constexpr int kCrWLayer = 3 * kCrWTile + kCrWChunk;   // 15360 B: tiles at 0 (s=0), kCrWTile (s=1, 4 chunks), 2*kCrWTile+kCrWChunk (s=2)
constexpr int kCrOffIdent = 4 * kCrWLayer;            // identity [32 cout][24 cin]: the residual x + conv(..) as one more UMMA
constexpr int kCrIdentChunk = 32 * 16;
constexpr int kCrOffZeroB = kCrOffIdent + 3 * kCrIdentChunk;   // 96 rows of zeros: 4th chunk of every other weight tile
constexpr int kCrWBytes = kCrOffZeroB + kCrWChunk;    // 64512 B
constexpr int kCrX0Slots = 6, kCrMidSlots = 4, kCrResSlots = 6;
constexpr int kCrWarps = 25, kCrThreads = kCrWarps * 32;
constexpr int kCrWarpHead = 16, kCrWarpMma = 21;            // head: 5 warps = 130 pixels of a row slot (+ idle lanes); 4 issuing warps
// shared memory carve-up (bytes from the 1024-aligned base)
constexpr int kCrOffW = 0;
constexpr int kCrOffX0 = kCrOffW + kCrWBytes;
constexpr int kCrOffA1 = kCrOffX0 + kCrX0Slots * kCrSlot;
constexpr int kCrOffAx1 = kCrOffA1 + kCrMidSlots * kCrSlot;
constexpr int kCrOffA3 = kCrOffAx1 + kCrMidSlots * kCrSlot;
constexpr int kCrOffRx1 = kCrOffA3 + kCrMidSlots * kCrSlot;
constexpr int kCrOffZero = kCrOffRx1 + kCrResSlots * kCrSlot;        // the constant 4th activation chunk (1, 1, 0, ..) per pixel
                                                                     // (must be the highest operand address: LBO is unsigned)
constexpr int kCrOffBars = kCrOffZero + kCrPlane;
constexpr int kCrNumBars = 2 * kCrX0Slots + 6 * kCrMidSlots + 2 * kCrResSlots + 32;
constexpr int kCrSmem = 1024 + kCrOffBars + kCrNumBars * 8 + 16;
constexpr int kCrFloats = 292;                // head_w[6][24], head_b[24], conv_b[4][24], tail_w[24], tail_b, pad

struct ComposeRowsParams {
  const float* small; const float* large; float* out;
  int small_cs, small_co, large_cs, large_co, out_cs, out_co;      // channel stride / offset of the fp32 rgb views
  const uint8_t* wblob;
  int N, H, W, strips;
  long long total_rows;
  int rows_per_cta;
  int bf16;
  int has_inv;
  dd_invert_params inv;
  float sqrt_var;
  unsigned long long* trace;       // debug: [9 roles][64 rows][8] clock64 stamps of CTA 0 (NULL = off)
  float fl[kCrFloats];
};

__device__ __forceinline__ uint64_t cr_desc_from(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ void tmem_zero_32x32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      :
      : "r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct CrSeg {
  int n, xs, y0, y1, lo, hi;      // image, image x of slot pixel 0, finished rows [y0,y1), computed rows [lo,hi)
};
struct CrWalker {
  long long lin, lin_end;
  int H, strips;
  __device__ CrWalker(const ComposeRowsParams& p) {
    lin = static_cast<long long>(blockIdx.x) * p.rows_per_cta;
    lin_end = lin + p.rows_per_cta;
    if (lin_end > p.total_rows) lin_end = p.total_rows;
    H = p.H; strips = p.strips;
  }
  __device__ bool next(CrSeg& s) {
    if (lin >= lin_end) return false;
    const long long col = lin / H;
    s.y0 = static_cast<int>(lin - col * H);
    s.n = static_cast<int>(col / strips);
    s.xs = static_cast<int>(col % strips) * kCrValid - 4;
    const long long left = lin_end - lin;
    s.y1 = (s.y0 + left > H) ? H : static_cast<int>(s.y0 + left);
    s.lo = s.y0 - 4 < 0 ? 0 : s.y0 - 4;
    s.hi = s.y1 + 4 > H ? H : s.y1 + 4;
    lin += s.y1 - s.y0;
    return true;
  }
};

__device__ __forceinline__ float cr_signed_expm1(float v) { return copysignf(expm1f(fabsf(v)), v) * (v != 0.f); }

template <bool BF16>
__global__ void __launch_bounds__(kCrThreads, 1) compose_rows_kernel(const __grid_constant__ ComposeRowsParams p) {
  extern __shared__ __align__(1024) uint8_t cr_raw[];
  const uint32_t base_u32 = (smem_u32(cr_raw) + 1023u) & ~1023u;
  uint8_t* smem = cr_raw + (base_u32 - smem_u32(cr_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCrOffBars);
  uint64_t* x0_full = bars;                          // [6]   head -> L1 / WG2
  uint64_t* x0_empty = x0_full + kCrX0Slots;         // [6]   L1 commit + WG2 (count 5)
  uint64_t* mid_full = x0_empty + kCrX0Slots;        // [3][4] a1, ax1, a3: WG1/2/3 -> L2/3/4
  uint64_t* mid_empty = mid_full + 3 * kCrMidSlots;  // [3][4]
  uint64_t* rx1_full = mid_empty + 3 * kCrMidSlots;  // [6]   WG2 -> WG4
  uint64_t* rx1_empty = rx1_full + kCrResSlots;      // [6]
  uint64_t* acc_full = rx1_empty + kCrResSlots;      // [4 layers][4 blocks]
  uint64_t* acc_empty = acc_full + 16;               // [4][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kCrNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup: weights + zeroed operand rings, barriers, TMEM
  for (int i = threadIdx.x; i < kCrWBytes / 16; i += kCrThreads)
    reinterpret_cast<uint4*>(smem + kCrOffW)[i] = __ldg(reinterpret_cast<const uint4*>(p.wblob) + i);
  for (int i = threadIdx.x + kCrWBytes / 16; i < kCrOffZero / 16; i += kCrThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < kCrPlane / 16; i += kCrThreads)      // K index 24, 25 = 1.0 for every pixel (bias hi + lo)
    reinterpret_cast<uint4*>(smem + kCrOffZero)[i] = make_uint4(BF16 ? 0x3F803F80u : 0x3C003C00u, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kCrX0Slots; ++i) { mbar_init(&x0_full[i], 5); mbar_init(&x0_empty[i], 2); }   // layer 1 + layer 2 (residual)
    for (int i = 0; i < 3 * kCrMidSlots; ++i) { mbar_init(&mid_full[i], 4); mbar_init(&mid_empty[i], 1); }
    for (int i = 0; i < kCrResSlots; ++i) { mbar_init(&rx1_full[i], 4); mbar_init(&rx1_empty[i], 1); }  // layer 4 (residual)
    for (int i = 0; i < 16; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == kCrWarpMma) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();          // the generic-proxy fills above are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int W = p.W;

  if (warp >= kCrWarpHead && warp < kCrWarpMma) {
    // ------------------------------------------------------------------ head: x0 rows (one slot pixel per thread; the
    // six inputs of the NEXT row are in flight while this row is computed)
    const int q = threadIdx.x - kCrWarpHead * 32;
    uint8_t* ring = smem + kCrOffX0;
    CrWalker walk(p);
    CrSeg sg;
    uint32_t slot = 0, phase = 0, gh = 0;
    unsigned long long* tr = (p.trace && blockIdx.x == 0 && q == 0) ? p.trace + 4 * 64 * 8 : nullptr;
    while (walk.next(sg)) {
      const int x = sg.xs + q;
      const bool live = q < kCrRowPx && x >= 0 && x < W;
      const float* lbase = p.large + (static_cast<size_t>(sg.n) * p.H * W + x) * p.large_cs + p.large_co;
      const float* sbase = p.small + (static_cast<size_t>(sg.n) * (p.H >> 1) * (W >> 1) + (x >> 1)) * p.small_cs + p.small_co;
      float nxt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      auto fetch = [&](int t) {
        if (live) {
          const float* lp = lbase + static_cast<size_t>(t) * W * p.large_cs;
          const float* sp = sbase + static_cast<size_t>(t >> 1) * (W >> 1) * p.small_cs;
          nxt[0] = __ldg(sp); nxt[1] = __ldg(sp + 1); nxt[2] = __ldg(sp + 2);
          nxt[3] = __ldg(lp); nxt[4] = __ldg(lp + 1); nxt[5] = __ldg(lp + 2);
        }
      };
      fetch(sg.lo);
      for (int t = sg.lo; t < sg.hi; ++t, ++gh) {
        const bool trace = tr && gh < 64;
        if (trace) tr[gh * 8 + 0] = clock64();
        float in[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) in[k] = nxt[k];
        if (t + 1 < sg.hi) fetch(t + 1);
        uint4 o[3] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (live) {
          float a[kCrC];
#pragma unroll
          for (int c = 0; c < kCrC; ++c) {
            float v = p.fl[144 + c];
#pragma unroll
            for (int k = 0; k < 6; ++k) v = fmaf(in[k], p.fl[k * kCrC + c], v);
            a[c] = fmaxf(v, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float f8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f8[e] = a[j * 8 + e];
            o[j] = pack8(f8, BF16);
          }
        }
        if (trace) tr[gh * 8 + 1] = clock64();
        mbar_wait_sleep(&x0_empty[slot], phase ^ 1u);
        if (trace) tr[gh * 8 + 2] = clock64();
        if (q < kCrRowPx) {
          uint8_t* row = ring + slot * kCrSlot + q * 16;
#pragma unroll
          for (int j = 0; j < 3; ++j) *reinterpret_cast<uint4*>(row + j * kCrPlane) = o[j];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x0_full[slot]);
        if (trace) tr[gh * 8 + 3] = clock64();
        if (++slot == kCrX0Slots) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= kCrWarpMma) {
    // ------------------------------------------------------------------ UMMA issuers: warp kCrWarpMma + l drives layer l
    // One thread issues the 6-12 UMMAs of a row, so everything per instruction is two integer adds: descriptors are
    // (lo, hi) register pairs whose hi word (SBO = 128 B, version) never changes and whose lo word is a per-row base plus
    // compile-time offsets; the one or two accumulator "pieces" of a row are worked out once per row.  The four layers
    // are independent instruction streams (their own TMEM columns and operand rings) coupled by mbarriers only, so four
    // threads issue them: a single thread's scalar work per row step was the bound of the first version.
    if (elect_one()) {
      const int l = warp - kCrWarpMma;
      const uint32_t idesc0 = make_idesc_f16(128, 0) | (BF16 ? kIdescBf16 : 0u);
      constexpr uint32_t kHi = (128u >> 4) | (1u << 14);                      // SBO = 128 B, descriptor version 1, no swizzle
      const uint32_t w_q = (base_u32 + kCrOffW + l * kCrWLayer) >> 4;           // addresses in 16-byte units from here on
      const uint32_t zerob_q = (base_u32 + kCrOffW + kCrOffZeroB) >> 4, ident_q = (base_u32 + kCrOffW + kCrOffIdent) >> 4;
      const uint32_t zero_q = (base_u32 + kCrOffZero) >> 4;                   // the constant (1, 1, 0, ..) activation chunk
      // residual layers (2 and 4): x + conv(..) - the row of x (x0 / raw x1, already an operand row) times the identity
      const bool has_res = (l == 1 || l == 3);
      const uint32_t res_ring_q = (base_u32 + (l == 1 ? kCrOffX0 : kCrOffRx1)) >> 4;
      uint64_t* res_full0 = (l == 1) ? x0_full : rx1_full;
      uint64_t* res_empty0 = (l == 1) ? x0_empty : rx1_empty;
      uint32_t rslot = 0, rphase = 0;
      const uint32_t ring_q = (base_u32 + (l == 0 ? kCrOffX0 : (l == 1 ? kCrOffA1 : (l == 2 ? kCrOffAx1 : kCrOffA3)))) >> 4;
      const uint32_t slots = (l == 0) ? kCrX0Slots : kCrMidSlots;
      uint64_t* in_full0 = (l == 0) ? x0_full : mid_full + (l - 1) * kCrMidSlots;
      uint64_t* in_empty0 = (l == 0) ? x0_empty : mid_empty + (l - 1) * kCrMidSlots;
      uint64_t* accf = acc_full + l * 4;
      uint64_t* acce = acc_empty + l * 4;
      const uint32_t d_layer = tmem_base + static_cast<uint32_t>(l * 128);
      long long lin = static_cast<long long>(blockIdx.x) * p.rows_per_cta;
      long long lin_end = lin + p.rows_per_cta;
      if (lin_end > p.total_rows) lin_end = p.total_rows;
      uint32_t g = 0, untouched = 0, slot = 0, phase = 0;
      unsigned long long* tr = (p.trace && blockIdx.x == 0) ? p.trace + (5 + l) * 64 * 8 : nullptr;
      while (lin < lin_end) {
        // next segment (same arithmetic as CrWalker; only the row count matters here)
        const long long col = lin / p.H;
        const int y0 = static_cast<int>(lin - col * p.H);
        const long long left = lin_end - lin;
        const int y1 = (y0 + left > p.H) ? p.H : static_cast<int>(y0 + left);
        const int lo = y0 - 4 < 0 ? 0 : y0 - 4, hi = y1 + 4 > p.H ? p.H : y1 + 4;
        lin += y1 - y0;
        const uint32_t g_lo = g, g_hi = g + static_cast<uint32_t>(hi - lo);
        for (; g < g_hi; ++g) {
          const bool trace = tr && g < 64;
          if (trace) tr[g * 8 + 0] = clock64();
          mbar_wait_sleep(&in_full0[slot], phase);
          if (trace) tr[g * 8 + 1] = clock64();
          // output rows this input row feeds: q = g + 1 - r, clipped to the segment
          const uint32_t q_hi = (g + 1 < g_hi) ? g + 1 : g_hi - 1;
          const uint32_t q_lo = (g > g_lo) ? g - 1 : g_lo;
          while (untouched <= q_hi) {
            mbar_wait_sleep(&acce[3 - (untouched & 3u)], (untouched >> 2) & 1u);
            ++untouched;
          }
          tc_fence_after();
          if (trace) tr[g * 8 + 2] = clock64();
          // the taps r_lo..r_hi land in blocks (b0 + r) & 3: one instruction, or two when the ring wraps
          const uint32_t r_lo = g + 1 - q_hi, r_hi = g + 1 - q_lo;
          const uint32_t blk_a = (3u - ((g + 1) & 3u) + r_lo) & 3u;
          uint32_t cnt_a = r_hi - r_lo + 1;
          if (blk_a + cnt_a > 4u) cnt_a = 4u - blk_a;
          const uint32_t cnt_b = r_hi - r_lo + 1 - cnt_a;                      // second piece starts at block 0
          const uint32_t d_a = d_layer + blk_a * 32u, i_a = idesc0 + ((cnt_a * 4u) << 17), b_a = r_lo * 32u;
          const uint32_t i_b = idesc0 + ((cnt_b * 4u) << 17), b_b = (r_lo + cnt_a) * 32u;
          const uint32_t slot_q = ring_q + slot * (kCrSlot >> 4);
          const uint32_t a_k0 = slot_q | (static_cast<uint32_t>(kCrPlane >> 4) << 16);
          const uint32_t a_k1 = (slot_q + 2u * (kCrPlane >> 4)) | ((zero_q - slot_q - 2u * (kCrPlane >> 4)) << 16);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t a_lo = (ks == 0 ? a_k0 : a_k1) + static_cast<uint32_t>(s);
              // weight tile of column s (s = 2 sits behind the 4-chunk centre tile); second K step: chunk 2 + (bias chunk of the
              // centre tile | the zero rows for s = 0, 2)
              const uint32_t tile_q = w_q + static_cast<uint32_t>((s * kCrWTile + (s == 2 ? kCrWChunk : 0) + ks * 2 * kCrWChunk) >> 4);
              const uint32_t lbo = (ks == 0 || s == 1) ? static_cast<uint32_t>(kCrWChunk >> 4) : zerob_q - tile_q;
              const uint32_t b_lo = tile_q | (lbo << 16);
              umma_f16(d_a, cr_desc_from(a_lo, kHi), cr_desc_from(b_lo + b_a, kHi), i_a, 1u);
              if (cnt_b) umma_f16(d_layer, cr_desc_from(a_lo, kHi), cr_desc_from(b_lo + b_b, kHi), i_b, 1u);
            }
          }
          if (has_res) {
            // output row g (tap r = 1 of this input row) += identity . residual row g, centre column (s = 1)
            mbar_wait_sleep(&res_full0[rslot], rphase);
            tc_fence_after();
            const uint32_t rq = res_ring_q + rslot * (kCrSlot >> 4) + 1u;
            const uint32_t d_r = d_layer + ((3u - ((g + 1) & 3u) + 1u) & 3u) * 32u, i_r = idesc0 + (4u << 17);
            umma_f16(d_r, cr_desc_from(rq | (static_cast<uint32_t>(kCrPlane >> 4) << 16), kHi),
                     cr_desc_from(ident_q | (static_cast<uint32_t>(kCrIdentChunk >> 4) << 16), kHi), i_r, 1u);
            const uint32_t rq2 = rq + 2u * (kCrPlane >> 4), iq2 = ident_q + 2u * (kCrIdentChunk >> 4);
            umma_f16(d_r, cr_desc_from(rq2 | ((zero_q + 1u - rq2) << 16), kHi), cr_desc_from(iq2 | ((zerob_q - iq2) << 16), kHi), i_r, 1u);
            umma_commit(&res_empty0[rslot]);
            if (++rslot == kCrResSlots) { rslot = 0; rphase ^= 1u; }
          }
          umma_commit(&in_empty0[slot]);
          if (g > g_lo) umma_commit(&accf[3 - ((g - 1) & 3u)]);      // row g-1 has all three taps
          if (g + 1 == g_hi) umma_commit(&accf[3 - (g & 3u)]);      // last input row of the segment
          if (trace) tr[g * 8 + 3] = clock64();
          if (++slot == slots) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue of layer L = warp / 4
    const int L = warp >> 2, wq = warp & 3;
    const int m = wq * 32 + lane;                                  // TMEM lane == UMMA row == strip pixel (slot pixel m + 1)
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(L * 128);
    for (int b = 0; b < 4; ++b) tmem_zero_32x32(t_lane + b * 32);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int b = 0; b < 4; ++b) mbar_arrive(&acc_empty[L * 4 + b]);
    uint8_t* out_ring = smem + (L == 0 ? kCrOffA1 : (L == 1 ? kCrOffAx1 : kCrOffA3));
    uint64_t* out_full = mid_full + (L < 3 ? L : 0) * kCrMidSlots;
    uint64_t* out_empty = mid_empty + (L < 3 ? L : 0) * kCrMidSlots;
    CrWalker walk(p);
    CrSeg sg;
    uint32_t g = 0, oslot = 0, ophase = 0, rslot = 0, rphase = 0;
    unsigned long long* tr = (p.trace && blockIdx.x == 0 && wq == 0 && lane == 0) ? p.trace + L * 64 * 8 : nullptr;
    while (walk.next(sg)) {
      const int x = sg.xs + m + 1;
      const bool inside = x >= 0 && x < W;
      for (int t = sg.lo; t < sg.hi; ++t, ++g) {
        const uint32_t blk = 3u - (g & 3u), use = g >> 2;
        const bool trace = tr && g < 64;
        if (trace) tr[g * 8 + 0] = clock64();
        // layer 4: fetch the blend operands before waiting for the accumulator
        const bool emit = (L == 3) && inside && t >= sg.y0 && t < sg.y1 && m >= 3 && m < 3 + kCrValid;
        float lg[3] = {0.f, 0.f, 0.f}, low[3] = {0.f, 0.f, 0.f}, sm[3] = {0.f, 0.f, 0.f};
        if (emit) {
          const size_t img = static_cast<size_t>(sg.n) * p.H;
          const float* lp = p.large + ((img + t) * W + x) * p.large_cs + p.large_co;
          const float* l00 = p.large + ((img + (t & ~1)) * W + (x & ~1)) * p.large_cs + p.large_co;
          const float* l10 = l00 + static_cast<size_t>(W) * p.large_cs;
          const float* sp = p.small + ((static_cast<size_t>(sg.n) * (p.H >> 1) + (t >> 1)) * (W >> 1) + (x >> 1)) * p.small_cs + p.small_co;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            lg[c] = __ldg(lp + c);
            low[c] = 0.25f * (__ldg(l00 + c) + __ldg(l00 + p.large_cs + c) + __ldg(l10 + c) + __ldg(l10 + p.large_cs + c));
            sm[c] = __ldg(sp + c);
          }
        }
        mbar_wait_sleep(&acc_full[L * 4 + blk], use & 1u);
        tc_fence_after();
        if (trace) tr[g * 8 + 1] = clock64();
        // bias and (layers 2, 4) the residual were added by the tensor core: the accumulator IS the layer output
        uint32_t vr[32];
        tmem_ld_32x32(t_lane + blk * 32, vr);
        tmem_ld_wait();
        tmem_zero_32x32(t_lane + blk * 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[L * 4 + blk]);
        if (trace) tr[g * 8 + 2] = clock64();

        if (L < 3) {
          // operand row of the next layer: 16-bit, relu, zero outside the image (SAME padding of the next layer);
          // layer 2 also publishes the raw row: the residual operand of layer 4
          uint4 raw[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float f8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f8[e] = __uint_as_float(vr[j * 8 + e]);
            raw[j] = inside ? pack8(f8, BF16) : make_uint4(0, 0, 0, 0);
          }
          mbar_wait_sleep(&out_empty[oslot], ophase ^ 1u);
          if (trace) tr[g * 8 + 3] = clock64();
          uint8_t* orow = out_ring + oslot * kCrSlot + (m + 1) * 16;
          if (L == 1) {
            mbar_wait_sleep(&rx1_empty[rslot], rphase ^ 1u);
            uint8_t* rrow = smem + kCrOffRx1 + rslot * kCrSlot + (m + 1) * 16;
#pragma unroll
            for (int j = 0; j < 3; ++j) *reinterpret_cast<uint4*>(rrow + j * kCrPlane) = raw[j];
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            uint4 a = raw[j];
            if (BF16) {
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&a);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __hmax2(h[e], __float2bfloat162_rn(0.f));
            } else {
              __half2* h = reinterpret_cast<__half2*>(&a);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __hmax2(h[e], __float2half2_rn(0.f));
            }
            *reinterpret_cast<uint4*>(orow + j * kCrPlane) = a;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&out_full[oslot]);
            if (L == 1) mbar_arrive(&rx1_full[rslot]);
          }
          if (++oslot == kCrMidSlots) { oslot = 0; ophase ^= 1u; }
          if (L == 1 && ++rslot == kCrResSlots) { rslot = 0; rphase ^= 1u; }
          if (trace) tr[g * 8 + 4] = clock64();
        } else if (emit) {
          // tail + blend (MultiScalePrediction.py:45-52,73-77)
          float s = p.fl[288];
#pragma unroll
          for (int c = 0; c < kCrC; ++c) s = fmaf(__uint_as_float(vr[c]), p.fl[264 + c], s);
          const float a = fmaxf(s, 0.f);
          const float wgt = 1.f / (1.f + __expf(-a));
          float* op = p.out + ((static_cast<size_t>(sg.n) * p.H + t) * W + x) * p.out_cs + p.out_co;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float o = lg[c] - wgt * low[c] + wgt * sm[c];
            if (p.has_inv) {
              if (p.inv.variance != 1.f) o *= p.sqrt_var;
              if (p.inv.mean != 0.f) o += p.inv.mean;
              if (p.inv.use_log1p) o = cr_signed_expm1(o);
            }
            op[c] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kCrWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dd
