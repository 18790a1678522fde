// MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93) as ONE tcgen05 kernel: a row-streaming, layer-pipelined
// implicit GEMM.  Every intermediate of the weight network lives in shared memory / TMEM; HBM sees the two fp32 rgb inputs
// and the fp32 rgb output only.
//
//   s_up = up2(small); x0 = relu(conv1x1_{6->24}(concat[s_up, large]))                              head (CUDA cores, fp32)
//   x1 = x0 + conv3x3(relu(conv3x3(relu(x0))));  x2 = x1 + conv3x3(relu(conv3x3(relu(x1))))         tcgen05, fp32 accumulate
//   w = sigmoid(relu(conv1x1_{24->1}(x2)));      out = large - w * up2(down2(large)) + w * s_up    tail + blend (fp32)
//
// Mapping onto the hardware
//   unit of work   one image ROW of a 128-pixel column strip, streamed top to bottom through all four 3x3 layers:
//                  while layer 1 consumes x0 row t, layer 2 consumes its input row t-2, layer 3 row t-4, layer 4 row t-6.
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
//                  16-19: head (x0 rows from the fp32 inputs)      20: TMEM allocator + the single UMMA-issuing thread
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
constexpr int kCrWTile = 3 * kCrWChunk;       // 4608 B per (layer, horizontal tap s)
constexpr int kCrWBytes = 12 * kCrWTile;      // 55296 B
constexpr int kCrX0Slots = 6, kCrMidSlots = 4, kCrResSlots = 6;
constexpr int kCrSkew = 2;                    // rows by which each layer trails the previous one
constexpr int kCrWarps = 21, kCrThreads = kCrWarps * 32;
constexpr int kCrWarpHead = 16, kCrWarpMma = 20;
// shared memory carve-up (bytes from the 1024-aligned base)
constexpr int kCrOffW = 0;
constexpr int kCrOffX0 = kCrOffW + kCrWBytes + kCrWChunk;            // + finite pad behind the last weight tile
constexpr int kCrOffA1 = kCrOffX0 + kCrX0Slots * kCrSlot;
constexpr int kCrOffAx1 = kCrOffA1 + kCrMidSlots * kCrSlot;
constexpr int kCrOffA3 = kCrOffAx1 + kCrMidSlots * kCrSlot;
constexpr int kCrOffRx1 = kCrOffA3 + kCrMidSlots * kCrSlot;
constexpr int kCrOffZero = kCrOffRx1 + kCrResSlots * kCrSlot;        // the all-zero chunk (must be the highest operand address)
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
  int bf16, desc_swap;
  int has_inv;
  dd_invert_params inv;
  float sqrt_var;
  float fl[kCrFloats];
};

__device__ __forceinline__ uint64_t cr_desc(uint32_t saddr, uint32_t lbo_bytes, int swap) {
  // K-major, no swizzle: LBO = byte distance between the two 16-byte K chunks of a K=16 step, SBO = distance between 8-row groups
  const uint32_t lbo = lbo_bytes >> 4, sbo = 128u >> 4;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((swap ? sbo : lbo) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((swap ? lbo : sbo) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
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
  for (int i = threadIdx.x + kCrWBytes / 16; i < kCrOffBars / 16; i += kCrThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kCrX0Slots; ++i) { mbar_init(&x0_full[i], 4); mbar_init(&x0_empty[i], 5); }
    for (int i = 0; i < 3 * kCrMidSlots; ++i) { mbar_init(&mid_full[i], 4); mbar_init(&mid_empty[i], 1); }
    for (int i = 0; i < kCrResSlots; ++i) { mbar_init(&rx1_full[i], 4); mbar_init(&rx1_empty[i], 4); }
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
    // ------------------------------------------------------------------ head: x0 rows
    const int ht = threadIdx.x - kCrWarpHead * 32;
    uint8_t* ring = smem + kCrOffX0;
    CrWalker walk(p);
    CrSeg sg;
    uint32_t g = 0;
    while (walk.next(sg)) {
      for (int t = sg.lo; t < sg.hi; ++t, ++g) {
        const uint32_t slot = g % kCrX0Slots, use = g / kCrX0Slots;
        mbar_wait(&x0_empty[slot], (use & 1u) ^ 1u);
        uint8_t* row = ring + slot * kCrSlot;
        for (int q = ht; q < kCrRowPx; q += 128) {
          const int x = sg.xs + q;
          uint4 o[3] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
          if (x >= 0 && x < W) {
            const float* lp = p.large + ((static_cast<size_t>(sg.n) * p.H + t) * W + x) * p.large_cs + p.large_co;
            const float* sp = p.small + ((static_cast<size_t>(sg.n) * (p.H >> 1) + (t >> 1)) * (W >> 1) + (x >> 1)) * p.small_cs + p.small_co;
            float in[6];
            in[0] = __ldg(sp); in[1] = __ldg(sp + 1); in[2] = __ldg(sp + 2);
            in[3] = __ldg(lp); in[4] = __ldg(lp + 1); in[5] = __ldg(lp + 2);
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
              o[j] = pack8(f8, p.bf16);
            }
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) *reinterpret_cast<uint4*>(row + j * kCrPlane + q * 16) = o[j];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x0_full[slot]);
      }
    }
  } else if (warp == kCrWarpMma) {
    // ------------------------------------------------------------------ UMMA issuer: four layers, skewed by kCrSkew rows
    if (elect_one()) {
      const uint32_t fmt = p.bf16 ? kIdescBf16 : 0u;
      const uint32_t w_base = base_u32 + kCrOffW, zero_addr = base_u32 + kCrOffZero;
      const uint32_t ring_base[4] = {base_u32 + kCrOffX0, base_u32 + kCrOffA1, base_u32 + kCrOffAx1, base_u32 + kCrOffA3};
      // per-layer iterators over the same (segment, row) sequence
      long long lin[4]; uint32_t g[4], g_lo[4], g_hi[4], untouched[4];
      const long long lin0 = static_cast<long long>(blockIdx.x) * p.rows_per_cta;
      long long lin_end = lin0 + p.rows_per_cta;
      if (lin_end > p.total_rows) lin_end = p.total_rows;
      for (int l = 0; l < 4; ++l) { lin[l] = lin0; g[l] = 0; g_lo[l] = 0; g_hi[l] = 0; untouched[l] = 0; }
      int done = 0;
      for (uint32_t step = 0; done < 4; ++step) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          if (step < static_cast<uint32_t>(l * kCrSkew)) continue;
          if (g[l] == g_hi[l]) {
            // next segment of this layer (same arithmetic as CrWalker, only the row counts matter here)
            if (lin[l] >= lin_end) { if (g[l] != 0xffffffffu) { g[l] = g_hi[l] = 0xffffffffu; ++done; } continue; }
            const long long col = lin[l] / p.H;
            const int y0 = static_cast<int>(lin[l] - col * p.H);
            const long long left = lin_end - lin[l];
            const int y1 = (y0 + left > p.H) ? p.H : static_cast<int>(y0 + left);
            const int lo = y0 - 4 < 0 ? 0 : y0 - 4, hi = y1 + 4 > p.H ? p.H : y1 + 4;
            lin[l] += y1 - y0;
            g_lo[l] = g[l];
            g_hi[l] = g[l] + static_cast<uint32_t>(hi - lo);
          }
          const uint32_t gg = g[l];
          const uint32_t slots = (l == 0) ? kCrX0Slots : kCrMidSlots;
          const uint32_t slot = gg % slots, use = gg / slots;
          uint64_t* in_full = (l == 0) ? &x0_full[slot] : &mid_full[(l - 1) * kCrMidSlots + slot];
          uint64_t* in_empty = (l == 0) ? &x0_empty[slot] : &mid_empty[(l - 1) * kCrMidSlots + slot];
          mbar_wait(in_full, use & 1u);
          // output rows this input row feeds: q = gg + 1 - r, clipped to the segment
          const uint32_t q_hi = (gg + 1 < g_hi[l]) ? gg + 1 : g_hi[l] - 1;
          const uint32_t q_lo = (gg > g_lo[l]) ? gg - 1 : g_lo[l];
          while (untouched[l] <= q_hi) {
            const uint32_t nu = untouched[l]++;
            mbar_wait(&acc_empty[l * 4 + (3 - (nu & 3u))], (nu >> 2) & 1u);
          }
          tc_fence_after();
          const int r_lo = static_cast<int>(gg + 1 - q_hi), r_hi = static_cast<int>(gg + 1 - q_lo);
          const int b0 = 3 - static_cast<int>((gg + 1) & 3u);
          const uint32_t slot_addr = ring_base[l] + slot * kCrSlot;
          const uint32_t d_layer = tmem_base + static_cast<uint32_t>(l * 128);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint32_t w_tile = w_base + static_cast<uint32_t>((l * 3 + s) * kCrWTile);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t a_addr = slot_addr + static_cast<uint32_t>(ks * 2 * kCrPlane + s * 16);
              const uint32_t a_lbo = (ks == 0) ? static_cast<uint32_t>(kCrPlane) : zero_addr - (slot_addr + 2u * kCrPlane);
              const uint64_t ad = cr_desc(a_addr, a_lbo, p.desc_swap);
              int r = r_lo;
              while (r <= r_hi) {
                const int blk = (b0 + r) & 3;
                int cnt = r_hi - r + 1;
                if (blk + cnt > 4) cnt = 4 - blk;
                const uint64_t bd = cr_desc(w_tile + static_cast<uint32_t>(ks * 2 * kCrWChunk + r * 32 * 16), kCrWChunk, p.desc_swap);
                umma_f16(d_layer + static_cast<uint32_t>(blk * 32), ad, bd, make_idesc_f16(128, cnt * 32) | fmt, 1u);
                r += cnt;
              }
            }
          }
          umma_commit(in_empty);
          if (gg > g_lo[l]) umma_commit(&acc_full[l * 4 + (3 - ((gg - 1) & 3u))]);     // row gg-1 has all three taps
          if (gg + 1 == g_hi[l]) umma_commit(&acc_full[l * 4 + (3 - (gg & 3u))]);     // last input row of the segment
          g[l] = gg + 1;
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
    const float* bias = p.fl + 168 + L * kCrC;
    CrWalker walk(p);
    CrSeg sg;
    uint32_t g = 0;
    while (walk.next(sg)) {
      const int x = sg.xs + m + 1;
      const bool inside = x >= 0 && x < W;
      for (int t = sg.lo; t < sg.hi; ++t, ++g) {
        const uint32_t blk = 3u - (g & 3u), use = g >> 2;
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
        mbar_wait(&acc_full[L * 4 + blk], use & 1u);
        tc_fence_after();
        uint32_t vr[32];
        tmem_ld_32x32(t_lane + blk * 32, vr);
        tmem_ld_wait();
        tmem_zero_32x32(t_lane + blk * 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[L * 4 + blk]);
        float v[kCrC];
#pragma unroll
        for (int c = 0; c < kCrC; ++c) v[c] = __uint_as_float(vr[c]) + bias[c];

        if (L == 1 || L == 3) {
          // residual: x0 row g (layer 2) / x1 row g (layer 4), this thread's pixel
          const uint32_t rs = g % kCrResSlots, ru = g / kCrResSlots;     // kCrX0Slots == kCrResSlots
          uint64_t* rfull = (L == 1) ? &x0_full[rs] : &rx1_full[rs];
          uint64_t* rempty = (L == 1) ? &x0_empty[rs] : &rx1_empty[rs];
          mbar_wait(rfull, ru & 1u);
          const uint8_t* rrow = smem + (L == 1 ? kCrOffX0 : kCrOffRx1) + rs * kCrSlot + (m + 1) * 16;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rrow + j * kCrPlane);
            float f8[8];
            unpack8(rv, p.bf16, f8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j * 8 + e] += f8[e];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(rempty);
        }

        if (L < 3) {
          // operand row of the next layer: relu, zero outside the image (SAME padding of the next layer), 16-bit
          const uint32_t slot = g % kCrMidSlots, ou = g / kCrMidSlots;
          mbar_wait(&out_empty[slot], (ou & 1u) ^ 1u);
          uint8_t* orow = out_ring + slot * kCrSlot + (m + 1) * 16;
          uint8_t* rrow = nullptr;
          if (L == 1) {
            const uint32_t rs = g % kCrResSlots, ru = g / kCrResSlots;
            mbar_wait(&rx1_empty[rs], (ru & 1u) ^ 1u);
            rrow = smem + kCrOffRx1 + rs * kCrSlot + (m + 1) * 16;
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float f8[8], r8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float raw = inside ? v[j * 8 + e] : 0.f;
              r8[e] = raw;
              f8[e] = fmaxf(raw, 0.f);
            }
            *reinterpret_cast<uint4*>(orow + j * kCrPlane) = pack8(f8, p.bf16);
            if (L == 1) *reinterpret_cast<uint4*>(rrow + j * kCrPlane) = pack8(r8, p.bf16);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&out_full[slot]);
            if (L == 1) mbar_arrive(&rx1_full[g % kCrResSlots]);
          }
        } else if (emit) {
          // tail + blend (MultiScalePrediction.py:45-52,73-77)
          float s = p.fl[288];
#pragma unroll
          for (int c = 0; c < kCrC; ++c) s = fmaf(v[c], p.fl[264 + c], s);
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
