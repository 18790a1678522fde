// tcgen05 implicit-GEMM convolution, "row pipeline" formulation (stride 1, TF 'SAME' zero padding, 3x3 / 1x1, NHWC fp16
// in, fp32 accumulate in TMEM, fused bias / ReLU / ReLU-copy / residual / channel-window (zero-copy concat) /
// pixel-shuffle (conv2d_transpose) epilogue written with TMA stores).
//
// Replaces tf.layers.conv2d / conv2d_transpose call sites of the reference:
//   UNet.py:29-31,56-58  Tiramisu.py:35-37,50-52,62-64,77-79  Architecture.py:238-243
//   MultiScalePrediction.py:64-66,73-75,88-90
//
// Mapping onto the hardware
//   unit of work   one INPUT row t of a 128-pixel column strip.  It is fetched ONCE (TMA box 64ch x 130px, OOB -> 0
//                  gives SAME padding) and feeds all three vertical taps: out row j = t + 1 - r receives tap r.
//   UMMA           M = 128 pixels, K = 16, N = n_r * cpad: the weights of the vertical taps r that share this A view
//                  are STACKED along N, so one instruction updates the accumulators of up to three output rows.
//                  (A [128x16] is read from shared memory once per instruction; with N = 64 the operand traffic
//                  would be 192 B/clk - above the 128 B/clk of shared memory - stacked it is <= 107 B/clk.)
//   horizontal tap view of the same row slot shifted by s pixels = +s*128 B in the 128B-swizzled slot
//   accumulators   TMEM ring of `ring` blocks of cpad fp32 columns; output row number q (running) lives in block
//                  ring-1-(q % ring), so the blocks of rows j, j-1, j-2 are adjacent in increasing column order
//                  (= stacking order r = 0,1,2); a wrap of the ring splits the instruction in two.
//   weights        [chunk][shift s][tap r][cpad][64ch] fp16; resident in shared memory for the whole kernel when
//                  they fit, otherwise streamed per (chunk, s) and amortised over a group of G input rows.
//   epilogue       4 warps: tcgen05.ld -> +bias (+residual) -> ReLU -> fp16/fp32 -> 128B-swizzled staging row in
//                  shared memory -> one TMA store per 32 pixels x 64 channels (full-line writes; the tensor map
//                  clips channels / pixels outside the view, sub-pixel views implement the pixel shuffle).
//   warps          0-7: epilogue (two sets of four warps, one TMEM lane quarter each, output rows round-robin)
//                  8: A producer  9: B producer  10: TMEM allocator  11: MMA issuer
//   grid           persistent: the N*strips*H output rows are split into gridDim contiguous ranges (+-1 row).
#pragma once
#include "dd_ptx.cuh"

namespace dd {

constexpr int kRowsEpiSets = 2;          // epilogue warp sets (4 warps each), output rows round-robin over the sets
                                         // (3 sets measured slower: the issuing thread, not the epilogue, bounds small layers)
constexpr int kRowsEpiWarps = 4 * kRowsEpiSets;
constexpr int kRowsWarpA = kRowsEpiWarps, kRowsWarpB = kRowsEpiWarps + 1, kRowsWarpTmem = kRowsEpiWarps + 2,
              kRowsWarpMma = kRowsEpiWarps + 3;   // highest warp id: preferred by the scheduler
constexpr int kRowsThreads = 32 * (kRowsEpiWarps + 4);
constexpr int kRowsTileW = 128;
constexpr int kRowsMaxRing = 8;

struct ConvRowsMaps {
  CUtensorMap a;        // input  [C, W, H, N] fp16, box [64, a_box_w, 1, 1]
  CUtensorMap b;        // weights [64, cpad, n_r, tiles] fp16, box [64, cpad, n_r, 1]
  CUtensorMap out[4];   // primary output view per column group (sub-pixel), box [64 | 32, 32, 1, 1]
  CUtensorMap out_relu; // optional fp16 relu(primary) view
  CUtensorMap a_lo;     // split mode: the low halves of the input (x = hi + lo)
  CUtensorMap out_lo[4];// split mode: the low halves of the 16-bit output
};

struct ConvRowsParams {
  int N, H, W, strips;
  long long total_rows;   // N * strips * H output rows
  int rows_per_cta;
  int n_chunks, ksteps_last;
  int n_s;                // horizontal shifts iterated
  int s_list[3];
  int halo;               // 1: 3x3 (a_box_w = 130, x origin -1), 0: 1x1
  int rm_lo, rm_hi;       // vertical taps present (r = t - j + 1); 1x1 uses rm_lo = rm_hi = 1
  int cpad;               // UMMA N per stacked tap == accumulator block width in columns
  int ring;               // accumulator blocks in TMEM (blk_stride = cpad)
  int max_stack;          // taps stacked in one instruction: floor(256 / cpad)
  int tiles_per_chunk;    // weight tiles per 64-channel chunk in the packed tensor (3 for 3x3, 1 for 1x1)
  int b_r0, b_row0;       // B-map coordinates of tap rm_lo / of the first output-channel row of this launch
  int G;                  // input rows per weight pass
  int w_resident;
  int a_slots, b_stages;
  uint32_t a_slot_bytes, a_tx_bytes, b_tile_bytes, b_tx_bytes;
  uint32_t a_off, b_off, stage_off, bias_off, bar_off;   // shared memory carve-up (from the 1024B-aligned base)
  // epilogue
  int ngroups, group_c, cout_store, ups;
  int relu, out_f32, has_relu_copy;
  int bf16;               // 16-bit tensors are bfloat16 (operands, 16-bit outputs, residual / mask)
  // split-fp16 arithmetic (the high-accuracy tensor-core mode): x = x_hi + x_lo, W = W_hi + W_lo (each an fp16 pair with
  // ~22 significant bits); x.W ~= x_hi.W_hi + x_lo.W_hi + x_hi.W_lo is three chunk passes over the same accumulators:
  // A chunk c of the 3 * nb reads part c / nb in (hi, lo, hi) of the input, weight chunk (c / nb == 2 ? nb : 0) + c % nb.
  int split;              // 0: plain; 1: split input + split weights
  int nb;                 // 64-channel chunks of the input proper (n_chunks == nb, or 3 * nb when split)
  int split_out;          // the 16-bit output is written as hi (primary maps) + lo (out_lo maps)
  int bias_count;
  const float* bias;
  const __half* residual;
  int res_cstride, res_coff;
  int res_is_mask;        // the residual tensor is a ReLU mask source (DD_CONV_RESIDUAL_MASK) instead of an addend
  unsigned long long* trace;
  // ablation switches of tools/ablate_conv.py (results are WRONG with any of them set): 1 epilogue skips convert / stage /
  // store, 2 epilogue skips the TMEM load too, 4 no UMMAs are issued, 8 every A load fetches input row 0 of image 0 (L2 hits),
  // 16 only the first k-step of every (chunk, shift) is issued
  int dbg;
  float* colsum;          // optional: colsum[c] += sum over all pixels of the primary output (fp32, before the 16-bit rounding):
                          // BiasAddGrad fused into the input-gradient convolution of the tensor-core training path
  int epi_plain;          // one 16-bit output, optional ReLU, no residual / mask / relu copy / split output: the compact epilogue
};

// ---------------------------------------------------------------- extra PTX (bulk tensor store, x32 TMEM load)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint64_t desc_from(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

struct RowsSegment {
  int n, x0, y0, y1;
};

// Walks the contiguous output-row range of this CTA as (image, strip, row-range) segments.
struct RowsWalker {
  long long lin, lin_end;
  int H, strips;
  __device__ RowsWalker(const ConvRowsParams& p) {
    lin = static_cast<long long>(blockIdx.x) * p.rows_per_cta;
    lin_end = lin + p.rows_per_cta;
    if (lin_end > p.total_rows) lin_end = p.total_rows;
    H = p.H; strips = p.strips;
  }
  __device__ bool next(RowsSegment& s) {
    if (lin >= lin_end) return false;
    const long long col = lin / H;          // (n * strips + strip)
    s.y0 = static_cast<int>(lin - col * H);
    s.n = static_cast<int>(col / strips);
    s.x0 = static_cast<int>(col % strips) * kRowsTileW;
    long long left = lin_end - lin;
    s.y1 = (s.y0 + left > H) ? H : static_cast<int>(s.y0 + left);
    lin += s.y1 - s.y0;
    return true;
  }
};

// SPLIT: the float16x2 instantiation (three chunk passes, low-half output pass); the plain instantiation carries none of it
template <bool SPLIT>
__global__ void __launch_bounds__(kRowsThreads, 1)
conv_rows_kernel(const __grid_constant__ ConvRowsMaps maps, const ConvRowsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* a_smem = smem + p.a_off;
  uint8_t* b_smem = smem + p.b_off;
  uint8_t* st_smem = smem + p.stage_off;
  float* bias_smem = reinterpret_cast<float*>(smem + p.bias_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a_full = bars;                       // [a_slots]
  uint64_t* a_empty = a_full + p.a_slots;        // [a_slots]
  uint64_t* b_full = a_empty + p.a_slots;        // [b_stages] (resident: [1])
  uint64_t* b_empty = b_full + p.b_stages;       // [b_stages]
  uint64_t* acc_full = b_empty + p.b_stages;     // [ring]
  uint64_t* acc_empty = acc_full + kRowsMaxRing; // [ring]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kRowsMaxRing);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kRowsWarpA && lane == 0) {
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.out[0]);
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < kRowsMaxRing; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    tmem_slot[1] = 0u;                   // barrier watcher's progress counter
    fence_mbar_init();
  }
  if (warp == kRowsWarpTmem) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp < 8) {
    for (int i = threadIdx.x; i < 256; i += 256) bias_smem[i] = (p.bias && i < p.bias_count) ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int t_lo_off = p.rm_lo - 1;   // first input row of a segment = y0 + t_lo_off
  const int t_hi_off = p.rm_hi - 2;   // last input row           = y1 + t_hi_off
  const uint32_t ready_addr = smem_u32(tmem_slot + 1);

  if (warp == kRowsWarpA) {
    // ------------------------------------------------------------------ A producer: one input row x one 64ch chunk per item
    // (every parameter pinned in a register, see pin() in dd_ptx.cuh: this loop has to stay well ahead of the UMMA issuer)
    if (elect_one()) {
      const uint32_t scr = smem_u32(tmem_slot + 2 + warp);
      const int group_rows = pin(p.G, scr), n_chunks = pin(p.n_chunks, scr), nb = pin(p.nb, scr), a_slots = pin(p.a_slots, scr),
                halo = pin(p.halo, scr), fixed_row = pin(p.dbg & 8, scr);
      const uint32_t a_tx_bytes = pin(p.a_tx_bytes, scr), a_slot_bytes = pin(p.a_slot_bytes, scr);
      const uint32_t a_dst0 = pin(smem_u32(a_smem), scr), bar_full = pin(smem_u32(a_full), scr), bar_empty = pin(smem_u32(a_empty), scr);
      const int t_lo = pin(t_lo_off, scr), t_hi = pin(t_hi_off, scr);
      unsigned long long* const trace = (blockIdx.x == 0) ? pin(p.trace, scr) : nullptr;
      int item = 0;
      RowsWalker walk(p);
      RowsSegment sg;
      int slot = 0; uint32_t phase = 0;
      while (walk.next(sg)) {
        const int t_first = sg.y0 + t_lo, t_last = sg.y1 + t_hi;
        const int x0 = sg.x0 - halo, img = fixed_row ? 0 : sg.n;
        for (int tg = t_first; tg <= t_last; tg += group_rows) {
          const int gcur = (t_last + 1 - tg < group_rows) ? (t_last + 1 - tg) : group_rows;
          for (int c = 0; c < n_chunks; ++c) {
            const int part = SPLIT ? c / nb : 0;
            const CUtensorMap* map = (SPLIT && part == 1) ? &maps.a_lo : &maps.a;
            const int c0 = (c - part * nb) * 64;
            for (int g = 0; g < gcur; ++g) {
              mbar_wait_addr(bar_empty + 8u * slot, phase ^ 1);
              if (trace && item < 64) trace[896 + item] = clock64();
              mbar_arrive_expect_tx_addr(bar_full + 8u * slot, a_tx_bytes);
              tma_load_4d_addr(a_dst0 + static_cast<uint32_t>(slot) * a_slot_bytes, map, bar_full + 8u * slot, c0, x0,
                               fixed_row ? 0 : tg + g, img);
              if (trace && item < 64) trace[960 + item] = clock64();
              ++item;
              if (++slot == a_slots) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kRowsWarpB) {
    // ------------------------------------------------------------------ B producer
    if (elect_one()) {
      if (p.w_resident) {
        const int tiles = (SPLIT ? 2 * p.nb : p.n_chunks) * p.n_s;
        mbar_arrive_expect_tx(&b_full[0], p.b_tx_bytes * static_cast<uint32_t>(tiles));
        for (int i = 0; i < tiles; ++i)
          tma_load_4d(b_smem + static_cast<size_t>(i) * p.b_tile_bytes, &maps.b, &b_full[0], 0, p.b_row0, p.b_r0,
                      (i / p.n_s) * p.tiles_per_chunk + p.s_list[i % p.n_s]);
      } else {
        const uint32_t scr = smem_u32(tmem_slot + 2 + warp);
        const int group_rows = pin(p.G, scr), n_chunks = pin(p.n_chunks, scr), nb = pin(p.nb, scr), n_s = pin(p.n_s, scr),
                  b_stages = pin(p.b_stages, scr), b_row0 = pin(p.b_row0, scr), b_r0 = pin(p.b_r0, scr),
                  tiles_per_chunk = pin(p.tiles_per_chunk, scr), s0 = pin(p.s_list[0], scr), s1 = pin(p.s_list[1], scr),
                  s2 = pin(p.s_list[2], scr);
        const uint32_t b_tx_bytes = pin(p.b_tx_bytes, scr), b_tile_bytes = pin(p.b_tile_bytes, scr);
        const uint32_t b_dst0 = pin(smem_u32(b_smem), scr), bar_full = pin(smem_u32(b_full), scr), bar_empty = pin(smem_u32(b_empty), scr);
        const int t_lo = pin(t_lo_off, scr), t_hi = pin(t_hi_off, scr);
        RowsWalker walk(p);
        RowsSegment sg;
        int stage = 0; uint32_t phase = 0;
        while (walk.next(sg)) {
          const int t_first = sg.y0 + t_lo, t_last = sg.y1 + t_hi;
          for (int tg = t_first; tg <= t_last; tg += group_rows) {
            for (int c = 0; c < n_chunks; ++c) {
              const int bc = (SPLIT && c >= nb) ? c - nb : c;                          // weight chunk of A chunk c
              for (int si = 0; si < n_s; ++si) {
                mbar_wait_addr(bar_empty + 8u * stage, phase ^ 1);
                mbar_arrive_expect_tx_addr(bar_full + 8u * stage, b_tx_bytes);
                tma_load_4d_addr(b_dst0 + static_cast<uint32_t>(stage) * b_tile_bytes, &maps.b, bar_full + 8u * stage, 0, b_row0,
                                 b_r0, bc * tiles_per_chunk + (si == 0 ? s0 : (si == 1 ? s1 : s2)));
                if (++stage == b_stages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == kRowsWarpTmem) {
    // ------------------------------------------------------------------ barrier watcher
    // Polling an mbarrier costs the polling thread ~200-250 cycles even when the phase has long completed (measured, round 2,
    // tools/trace_conv.py: "wait a_full" 260 + "wait acc_empty" 210 cycles per row with no UMMA and no epilogue in flight), and
    // the tensor pipe queues only one or two UMMAs behind the one in flight - every cycle the issuing thread spends in a wait is
    // an idle cycle of the pipe.  So this otherwise idle thread does the polling: it walks the barrier waits of the issuing
    // thread in exactly the issuing thread's order and publishes how many have completed in a shared-memory counter; the
    // issuing thread compares a cached copy and re-reads the counter (one ~30-cycle shared-memory load) only when it runs dry.
    if (elect_one()) {
      const uint32_t scr = smem_u32(tmem_slot + 2 + warp);
      const int ring = pin(p.ring, scr), rm_lo = pin(p.rm_lo, scr), n_s = pin(p.n_s, scr), n_chunks = pin(p.n_chunks, scr),
                a_slots = pin(p.a_slots, scr), b_stages = pin(p.b_stages, scr), group_rows = pin(p.G, scr);
      const bool w_resident = pin(p.w_resident, scr) != 0;
      const int t_lo = pin(t_lo_off, scr), t_hi = pin(t_hi_off, scr);
      const uint32_t ready = pin(ready_addr, scr);
      const uint32_t bar_a_full = pin(smem_u32(a_full), scr), bar_b_full = pin(smem_u32(b_full), scr),
                     bar_acc_empty = pin(smem_u32(acc_empty), scr);
      unsigned long long* const trace = (blockIdx.x == 0) ? pin(p.trace, scr) : nullptr;
      int item = 0;
      uint32_t count = 0;
      auto publish = [&]() { st_release_shared(ready, ++count); };
      RowsWalker walk(p);
      RowsSegment sg;
      int a_slot = 0; uint32_t a_phase = 0;
      int b_stage = 0; uint32_t b_phase = 0;
      uint32_t q_seg = 0;
      if (w_resident) { mbar_wait_addr(bar_b_full, 0u); publish(); }
      while (walk.next(sg)) {
        const int t_first = sg.y0 + t_lo, t_last = sg.y1 + t_hi;
        const uint32_t q_top0 = q_seg + static_cast<uint32_t>(t_first + 1 - sg.y0);
        int blk_top = ring - 1 - static_cast<int>(q_top0 % static_cast<uint32_t>(ring));
        uint32_t use_top = q_top0 / static_cast<uint32_t>(ring);
        // block / wait parity of the output row an input row initialises (the first clauses of make_plan below)
        auto init_block = [&](int t, int bt, uint32_t ut, int& blk, uint32_t& par) {
          int r_lo = t + 2 - sg.y1; if (r_lo < rm_lo) r_lo = rm_lo;
          blk = bt + r_lo; uint32_t use = ut;
          if (blk >= ring) { blk -= ring; use -= 1u; }
          par = (use & 1u) ^ 1u;
          return r_lo == rm_lo;
        };
        if (group_rows == 1) {
          for (int t = t_first; t <= t_last; ++t) {
            int blk; uint32_t par;
            const bool init = init_block(t, blk_top, use_top, blk, par);
            for (int c = 0; c < n_chunks; ++c) {
              mbar_wait_addr(bar_a_full + 8u * a_slot, a_phase); publish();
              if (trace && item < 64) trace[768 + item] = clock64();
              if (c == 0 && init) { mbar_wait_addr(bar_acc_empty + 8u * blk, par); publish(); }
              if (trace && item < 64) trace[832 + item] = clock64();
              ++item;
              if (++a_slot == a_slots) { a_slot = 0; a_phase ^= 1; }
              if (!w_resident) {
                for (int si = 0; si < n_s; ++si) {
                  mbar_wait_addr(bar_b_full + 8u * b_stage, b_phase); publish();
                  if (++b_stage == b_stages) { b_stage = 0; b_phase ^= 1; }
                }
              }
            }
            if (--blk_top < 0) { blk_top = ring - 1; ++use_top; }
          }
        } else {
          for (int tg = t_first; tg <= t_last; tg += group_rows) {
            const int gcur = (t_last + 1 - tg < group_rows) ? (t_last + 1 - tg) : group_rows;
            const int a_slot0 = a_slot; const uint32_t a_phase0 = a_phase;
            for (int c = 0; c < n_chunks; ++c) {
              for (int si = 0; si < n_s; ++si) {
                mbar_wait_addr(bar_b_full + 8u * b_stage, b_phase); publish();
                if (++b_stage == b_stages) { b_stage = 0; b_phase ^= 1; }
                if (si == 0) {
                  int lin_slot = a_slot0 + c * gcur; uint32_t ph = a_phase0;
                  while (lin_slot >= a_slots) { lin_slot -= a_slots; ph ^= 1; }
                  int bt = blk_top; uint32_t ut = use_top;
                  for (int g = 0; g < gcur; ++g) {
                    mbar_wait_addr(bar_a_full + 8u * lin_slot, ph); publish();
                    if (c == 0) {
                      int blk; uint32_t par;
                      if (init_block(tg + g, bt, ut, blk, par)) { mbar_wait_addr(bar_acc_empty + 8u * blk, par); publish(); }
                      if (--bt < 0) { bt = ring - 1; ++ut; }
                    }
                    if (++lin_slot == a_slots) { lin_slot = 0; ph ^= 1; }
                  }
                }
              }
            }
            a_slot = a_slot0 + n_chunks * gcur; a_phase = a_phase0;
            while (a_slot >= a_slots) { a_slot -= a_slots; a_phase ^= 1; }
            for (int g = 0; g < gcur; ++g) { if (--blk_top < 0) { blk_top = ring - 1; ++use_top; } }
          }
        }
        q_seg += static_cast<uint32_t>(sg.y1 - sg.y0);
      }
    }
  } else if (warp == kRowsWarpMma) {
    // ------------------------------------------------------------------ MMA issuer (highest warp id: the scheduler
    // prefers it over the epilogue warp sharing its sub-partition)
    if (elect_one()) {
      // ONE thread issues every UMMA of the CTA, and its scalar instruction stream is the critical path of the kernel: a
      // lone thread retires a dependent instruction every ~4-10 cycles, the tensor pipe queues only one or two UMMAs, and an
      // N = 192 UMMA lasts 96 cycles.  Measured (round 2, tools/ablate_conv.py): with the table-driven generic issue loop of
      // round 1 a 64->64 row took ~3000 cycles whether 12, 3 or 0 of its UMMAs were issued (tensor work: 1152).  Hence:
      //  * the row's UMMA list is reduced to what the ring geometry allows - the taps of a row cover consecutive accumulator
      //    blocks that wrap around the ring at most once, and N <= 256 splits at most once, i.e. every (row, chunk, shift,
      //    k-step) is piece A [+ piece B] - derived with a dozen branch-free integer instructions per row and issued as
      //    predicated straight-line code: no plan tables, no shared-memory loads, no loops over pieces;
      //  * no mbarrier polling here (see the barrier watcher above): `await` is a register compare, rarely a shared load;
      //  * the plan of the next row is computed behind the first UMMAs of this row.
      // every parameter the loops below touch is pinned in a register (see pin() in dd_ptx.cuh)
      const uint32_t scr = smem_u32(tmem_slot + 2 + warp);
      const uint64_t desc_tmpl = make_desc_sw128(0, 0);
      const uint32_t d_lo = static_cast<uint32_t>(desc_tmpl), d_hi = static_cast<uint32_t>(desc_tmpl >> 32);
      const uint32_t a_desc0 = pin(d_lo + (smem_u32(a_smem) >> 4), scr), b_desc0 = pin(d_lo + (smem_u32(b_smem) >> 4), scr);
      const uint32_t a_slot_units = pin(p.a_slot_bytes >> 4, scr), b_tile_units = pin(p.b_tile_bytes >> 4, scr);
      const int ring = pin(p.ring, scr), rm_lo = pin(p.rm_lo, scr), rm_hi = pin(p.rm_hi, scr), n_s = pin(p.n_s, scr), n_chunks = pin(p.n_chunks, scr),
                a_slots = pin(p.a_slots, scr), max_stack = pin(p.max_stack, scr), b_stages = pin(p.b_stages, scr), nb = pin(p.nb, scr),
                ksteps_last = pin(p.ksteps_last, scr), dbg = pin(p.dbg, scr), group_rows = pin(p.G, scr);
      const bool w_resident = pin(p.w_resident, scr) != 0;
      const uint32_t cpad = pin(static_cast<uint32_t>(p.cpad), scr);
      const uint32_t idesc0 = pin(make_idesc_f16(kRowsTileW, 0) | (p.bf16 ? kIdescBf16 : 0u), scr);
      const uint32_t istep = pin((cpad >> 3) << 17, scr);  // one more stacked tap in the N field of the instruction descriptor
      const uint32_t bstep = pin(cpad * 8u, scr);          // one tap further in the weight tile (descriptor units of 16 B, 128 B rows)
      const uint32_t s_off0 = pin(static_cast<uint32_t>(p.s_list[0]) * 8u, scr), s_off1 = pin(static_cast<uint32_t>(p.s_list[1]) * 8u, scr),
                     s_off2 = pin(static_cast<uint32_t>(p.s_list[2]) * 8u, scr);              // shifts in descriptor units
      const uint32_t bar_a_empty = pin(smem_u32(a_empty), scr), bar_b_empty = pin(smem_u32(b_empty), scr),
                     bar_acc_full = pin(smem_u32(acc_full), scr), ready = pin(ready_addr, scr);
      const int t_lo = pin(t_lo_off, scr), t_hi = pin(t_hi_off, scr);
      unsigned long long* const trace = (blockIdx.x == 0) ? pin(p.trace, scr) : nullptr;
      uint4* plan = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 512);   // streamed path: [8 rows of a group][2]

      uint32_t needed = 0, seen = 0;       // barrier waits consumed / known complete (watcher's counter)
      auto await = [&]() {
        ++needed;
        while (static_cast<int32_t>(seen - needed) < 0) seen = ld_acquire_shared(ready);
      };

      struct RowPlan {
        uint32_t dA, iA, bA;   // piece A: TMEM address, instruction descriptor, offset of its first tap in the weight tile
        uint32_t dB, iB, bB;   // piece B (iB == 0: none): the taps behind the ring wrap / beyond N = 256
        uint32_t iR;           // first touch: piece A minus its first tap (0: piece A is a single tap)
        bool init;             // this input row is the first contribution to the output row served by its lowest tap
      };
      auto plan_taps = [&](int r_lo, int r_hi, int blk, RowPlan& pl) {      // taps r_lo..r_hi, tap r_lo on block blk
        const int n = r_hi - r_lo + 1;
        int cnt_a = ring - blk; if (cnt_a > n) cnt_a = n; if (cnt_a > max_stack) cnt_a = max_stack;
        const int cnt_b = n - cnt_a;
        int blk_b = blk + cnt_a; if (blk_b >= ring) blk_b -= ring;
        pl.dA = tmem_base + static_cast<uint32_t>(blk) * cpad;
        pl.iA = idesc0 + static_cast<uint32_t>(cnt_a) * istep;
        pl.bA = static_cast<uint32_t>(r_lo - rm_lo) * bstep;
        pl.dB = tmem_base + static_cast<uint32_t>(blk_b) * cpad;
        pl.iB = cnt_b ? idesc0 + static_cast<uint32_t>(cnt_b) * istep : 0u;
        pl.bB = pl.bA + static_cast<uint32_t>(cnt_a) * bstep;
        pl.iR = cnt_a > 1 ? idesc0 + static_cast<uint32_t>(cnt_a - 1) * istep : 0u;
        pl.init = (r_lo == rm_lo);
      };
      // Interior rows (all taps inside the segment) have one plan per accumulator block index: tabulated once, two 16-byte
      // shared-memory loads per row; only the first / last input rows of a segment are computed on the spot.
      uint4* table = plan + 16;                                                         // [ring][2]
      for (int bk = 0; bk < ring; ++bk) {
        RowPlan pl;
        plan_taps(rm_lo, rm_hi, bk, pl);
        table[2 * bk] = make_uint4(pl.dA, pl.iA, pl.bA, pl.dB);
        table[2 * bk + 1] = make_uint4(pl.iB, pl.bB, pl.iR, 1u);
      }
      auto make_plan = [&](int t, const RowsSegment& sg, int bt, RowPlan& pl) {
        int r_lo = t + 2 - sg.y1; if (r_lo < rm_lo) r_lo = rm_lo;     // taps landing on rows of the segment
        int r_hi = t + 1 - sg.y0; if (r_hi > rm_hi) r_hi = rm_hi;
        int blk = bt + r_lo;
        if (blk >= ring) blk -= ring;                                 // block of the row served by tap r_lo
        if (r_lo == rm_lo && r_hi == rm_hi) {
          const uint4 q0 = table[2 * blk], q1 = table[2 * blk + 1];
          pl.dA = q0.x; pl.iA = q0.y; pl.bA = q0.z; pl.dB = q0.w; pl.iB = q1.x; pl.bB = q1.y; pl.iR = q1.z; pl.init = true;
        } else {
          // the volatile moves keep the compiler from computing this rare path speculatively on every row
          asm volatile("mov.b32 %0, %0;\n\tmov.b32 %1, %1;\n\tmov.b32 %2, %2;" : "+r"(r_lo), "+r"(r_hi), "+r"(blk));
          plan_taps(r_lo, r_hi, blk, pl);
        }
      };
      // the UMMAs of one (input row, chunk, shift): k-steps x pieces, operands resident.  Real branches around the optional
      // pieces: a predicated-off tcgen05.mma still costs the thread its ~45-cycle issue slot (tools/umma_queue.cu: 45-56 cycles of
      // the issuing thread per UMMA on an idle pipe; the pipe queues 6-8 UMMAs before the issue is throttled to its own rate).
      auto issue_shift = [&](const RowPlan& pl, bool first, uint32_t a_lo, uint32_t b_lo, int ksteps) {
        if (dbg & 20) { if (dbg & 4) return; ksteps = 1; }
        const uint32_t b_a = b_lo + pl.bA, b_b = b_lo + pl.bB;
        if (first) {
          // first touch of an output row: its own block with accumulate = 0, then the rest of piece A
          umma_f16(pl.dA, desc_from(a_lo, d_hi), desc_from(b_a, d_hi), idesc0 + istep, 0u);
          if (pl.iR) umma_f16(pl.dA + cpad, desc_from(a_lo, d_hi), desc_from(b_a + bstep, d_hi), pl.iR, 1u);
        } else {
          umma_f16(pl.dA, desc_from(a_lo, d_hi), desc_from(b_a, d_hi), pl.iA, 1u);
        }
        if (pl.iB == 0u) {
          if (ksteps == 4) {
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_f16(pl.dA, desc_from(a_lo + 2u * k, d_hi), desc_from(b_a + 2u * k, d_hi), pl.iA, 1u);
          } else if (ksteps == 2) {
            umma_f16(pl.dA, desc_from(a_lo + 2u, d_hi), desc_from(b_a + 2u, d_hi), pl.iA, 1u);
          } else {
            for (int k = 1; k < ksteps; ++k) umma_f16(pl.dA, desc_from(a_lo + 2u * k, d_hi), desc_from(b_a + 2u * k, d_hi), pl.iA, 1u);
          }
        } else {
          umma_f16(pl.dB, desc_from(a_lo, d_hi), desc_from(b_b, d_hi), pl.iB, 1u);
          if (ksteps == 4) {
#pragma unroll
            for (int k = 1; k < 4; ++k) {
              umma_f16(pl.dA, desc_from(a_lo + 2u * k, d_hi), desc_from(b_a + 2u * k, d_hi), pl.iA, 1u);
              umma_f16(pl.dB, desc_from(a_lo + 2u * k, d_hi), desc_from(b_b + 2u * k, d_hi), pl.iB, 1u);
            }
          } else if (ksteps == 2) {
            umma_f16(pl.dA, desc_from(a_lo + 2u, d_hi), desc_from(b_a + 2u, d_hi), pl.iA, 1u);
            umma_f16(pl.dB, desc_from(a_lo + 2u, d_hi), desc_from(b_b + 2u, d_hi), pl.iB, 1u);
          } else {
            for (int k = 1; k < ksteps; ++k) {
              umma_f16(pl.dA, desc_from(a_lo + 2u * k, d_hi), desc_from(b_a + 2u * k, d_hi), pl.iA, 1u);
              umma_f16(pl.dB, desc_from(a_lo + 2u * k, d_hi), desc_from(b_b + 2u * k, d_hi), pl.iB, 1u);
            }
          }
        }
      };

      RowsWalker walk(p);
      RowsSegment sg;
      int a_slot = 0;
      int b_stage = 0;
      uint32_t q_seg = 0;                  // running index of the first output row of the segment
      int it = 0;
      if (w_resident) { await(); tc_fence_after(); }
      while (walk.next(sg)) {
        const int t_first = sg.y0 + t_lo, t_last = sg.y1 + t_hi;
        // accumulator block of the output row fed by tap r = 0 of input row t (running index q_seg + t + 1 - y0); tap r lands
        // r blocks further.  Maintained incrementally (one division per segment).
        const uint32_t q_top0 = q_seg + static_cast<uint32_t>(t_first + 1 - sg.y0);
        int blk_top = ring - 1 - static_cast<int>(q_top0 % static_cast<uint32_t>(ring));
        if (group_rows == 1) {
          // ---------------- one input row at a time (weights resident, or a single-row weight pass)
          RowPlan pl;
          make_plan(t_first, sg, blk_top, pl);
          for (int t = t_first; t <= t_last; ++t, ++it) {
            const bool tr = trace && it < 64;
            if (tr) trace[it * 8 + 0] = clock64();
            RowPlan nxt = pl;
            if (tr && (dbg & 256) && it == 8) {
              // in-situ latency probe of the issuing thread (tools/trace_conv.py probe): dependent integer chain, parameter
              // (constant bank) loads, shared-memory loads, back-to-back clock reads
              unsigned long long c0 = clock64();
              uint32_t x;
              asm volatile("mov.u32 %0, %1;" : "=r"(x) : "r"(it));
#pragma unroll
              for (int i = 0; i < 32; ++i) x = x * 3u + 1u;
              asm volatile("mov.u32 %0, %1;" : "=r"(x) : "r"(x));
              unsigned long long c1 = clock64();
              const uint32_t* pm = reinterpret_cast<const uint32_t*>(&maps);
#pragma unroll
              for (int i = 0; i < 8; ++i) x = pm[(x & 255u)] + i;
              asm volatile("mov.u32 %0, %1;" : "=r"(x) : "r"(x));
              unsigned long long c2 = clock64();
              const volatile uint32_t* ps = reinterpret_cast<const volatile uint32_t*>(plan);
#pragma unroll
              for (int i = 0; i < 8; ++i) x = ps[(x & 63u)] + i;
              unsigned long long c3 = clock64();
              unsigned long long c4 = clock64();
              unsigned long long c5 = clock64();
              trace[768] = c1 - c0; trace[769] = c2 - c1; trace[770] = c3 - c2; trace[771] = c4 - c3; trace[772] = c5 - c4;
              trace[773] = x;
            }
            for (int c = 0; c < n_chunks; ++c) {
              int cc = c;
              if (SPLIT) { while (cc >= nb) cc -= nb; }
              const int ksteps = (cc == nb - 1) ? ksteps_last : 4;
              const int bc = (SPLIT && c >= 2 * nb) ? c - nb : cc;                         // resident weight chunk of A chunk c
              await();                                                                    // A slot landed
              if (c == 0 && pl.init) await();                                             // previous user of the new row's block drained
              if (tr && c == 0) trace[it * 8 + 1] = clock64();
              tc_fence_after();
              if (tr && c == 0) trace[it * 8 + 2] = clock64();
              const uint32_t a_lo = a_desc0 + static_cast<uint32_t>(a_slot) * a_slot_units;
              uint32_t b_res = b_desc0 + static_cast<uint32_t>(bc * n_s) * b_tile_units;
#pragma unroll
              for (int si = 0; si < 3; ++si) {
                if (si < n_s) {
                  uint32_t b_lo;
                  if (w_resident) {
                    b_lo = b_res; b_res += b_tile_units;
                  } else {
                    await();
                    tc_fence_after();
                    b_lo = b_desc0 + static_cast<uint32_t>(b_stage) * b_tile_units;
                  }
                  const uint32_t s_off = (si == 0) ? s_off0 : ((si == 1) ? s_off1 : s_off2);
                  issue_shift(pl, si == 0 && c == 0 && pl.init, a_lo + s_off, b_lo, ksteps);
                  if (!w_resident) {
                    umma_commit_addr(bar_b_empty + 8u * b_stage);
                    if (++b_stage == b_stages) b_stage = 0;
                  }
                  if (si == 0 && c == 0 && t < t_last) {      // the next row's plan, behind the UMMAs queued so far
                    int nbt = blk_top - 1; if (nbt < 0) nbt = ring - 1;
                    make_plan(t + 1, sg, nbt, nxt);
                  }
                }
              }
              if (tr && c == n_chunks - 1) trace[512 + it] = clock64();
              umma_commit_addr(bar_a_empty + 8u * a_slot);
              if (tr && c == n_chunks - 1) trace[576 + it] = clock64();
              if (++a_slot == a_slots) a_slot = 0;
            }
            // output row completed by this input row: the one served by tap rm_hi, if it belongs to the segment
            {
              const int j = t + 1 - rm_hi;
              if (j >= sg.y0 && j < sg.y1) {
                int blk_done = blk_top + rm_hi; if (blk_done >= ring) blk_done -= ring;
                umma_commit_addr(bar_acc_full + 8u * blk_done);
              }
            }
            if (--blk_top < 0) blk_top = ring - 1;
            pl = nxt;
            if (tr) trace[it * 8 + 3] = clock64();
          }
        } else {
          // ---------------- groups of G input rows per weight pass (streamed weights)
          for (int tg = t_first; tg <= t_last; tg += group_rows, ++it) {
            const int gcur = (t_last + 1 - tg < group_rows) ? (t_last + 1 - tg) : group_rows;
            const bool tr = trace && it < 64;
            if (tr) trace[it * 8 + 0] = clock64();
            {
              int bt = blk_top;
              for (int g = 0; g < gcur; ++g) {
                RowPlan pl;
                make_plan(tg + g, sg, bt, pl);
                plan[2 * g] = make_uint4(pl.dA, pl.iA, pl.bA, pl.dB);
                plan[2 * g + 1] = make_uint4(pl.iB, pl.bB, pl.iR, pl.init ? 1u : 0u);
                if (--bt < 0) bt = ring - 1;
              }
            }
            const int a_slot0 = a_slot;
            if (tr) trace[it * 8 + 1] = clock64();                                      // plans built
            long long wait_b = 0, wait_a = 0;
            for (int c = 0; c < n_chunks; ++c) {
              int cc = c;
              if (SPLIT) { while (cc >= nb) cc -= nb; }
              const int ksteps = (cc == nb - 1) ? ksteps_last : 4;
              for (int si = 0; si < n_s; ++si) {
                const long long tb0 = tr ? clock64() : 0;
                await();                                                                  // weight stage landed
                tc_fence_after();
                if (tr) wait_b += clock64() - tb0;
                const uint32_t b_lo = b_desc0 + static_cast<uint32_t>(b_stage) * b_tile_units;
                const uint32_t s_off = (si == 0) ? s_off0 : ((si == 1) ? s_off1 : s_off2);
                int lin_slot = a_slot0 + c * gcur;      // slots of this chunk's rows: a_slot0 + c * gcur + g (mod a_slots)
                while (lin_slot >= a_slots) lin_slot -= a_slots;
                for (int g = 0; g < gcur; ++g) {
                  const uint4 q0 = plan[2 * g], q1 = plan[2 * g + 1];
                  RowPlan pl;
                  pl.dA = q0.x; pl.iA = q0.y; pl.bA = q0.z; pl.dB = q0.w; pl.iB = q1.x; pl.bB = q1.y; pl.iR = q1.z;
                  pl.init = q1.w != 0u;
                  const bool first = (c == 0 && si == 0 && pl.init);
                  if (si == 0) {
                    const long long ta0 = tr ? clock64() : 0;
                    await();                                                              // A slot landed
                    if (c == 0 && pl.init) await();                                       // the new row's block drained
                    tc_fence_after();
                    if (tr) wait_a += clock64() - ta0;
                    if (tr && c == 0 && g == 0) trace[it * 8 + 2] = clock64();
                  }
                  issue_shift(pl, first, a_desc0 + static_cast<uint32_t>(lin_slot) * a_slot_units + s_off, b_lo, ksteps);
                  if (si == n_s - 1) umma_commit_addr(bar_a_empty + 8u * lin_slot);
                  if (++lin_slot == a_slots) lin_slot = 0;
                }
                umma_commit_addr(bar_b_empty + 8u * b_stage);
                if (++b_stage == b_stages) b_stage = 0;
              }
            }
            // advance the A ring past this group
            a_slot = a_slot0 + n_chunks * gcur;
            while (a_slot >= a_slots) a_slot -= a_slots;
            // output rows completed by this group: tap rm_hi of each input row, if that row belongs to the segment
            for (int g = 0; g < gcur; ++g) {
              const int j = tg + g + 1 - rm_hi;
              if (j >= sg.y0 && j < sg.y1) {
                int blk_done = blk_top + rm_hi; if (blk_done >= ring) blk_done -= ring;
                umma_commit_addr(bar_acc_full + 8u * blk_done);
              }
              if (--blk_top < 0) blk_top = ring - 1;
            }
            if (tr) { trace[it * 8 + 3] = clock64(); trace[640 + it] = wait_b; trace[704 + it] = wait_a; }
          }
        }
        q_seg += static_cast<uint32_t>(sg.y1 - sg.y0);
      }
    }
  } else if (warp < kRowsEpiWarps) {
    // ------------------------------------------------------------------ epilogue: sets of 4 warps, rows round-robin
    const int wq = warp & 3;                       // TMEM lane quarter
    const int eset = warp >> 2;                    // rows with q % n_sets == eset
    // A set may only wait for use u of an accumulator barrier after use u-1 completed (parity waits cannot tell
    // phases two apart): its previous row q - n_sets must be at least as late as row q - ring, i.e. n_sets <= ring.
    const uint32_t scr = smem_u32(tmem_slot + 2 + warp);   // parameters of the row loop pinned in registers (see pin())
    const int ring_i = pin(p.ring, scr);
    const int n_sets = (ring_i < kRowsEpiSets) ? ring_i : kRowsEpiSets;
    const uint32_t cpad_e = pin(static_cast<uint32_t>(p.cpad), scr);
    const int epi_plain = pin(p.epi_plain, scr), e_ngroups = pin(p.ngroups, scr), e_group_c = pin(p.group_c, scr),
              e_cout = pin(p.cout_store, scr), e_dbg = pin(p.dbg, scr), e_bf16 = pin(p.bf16, scr);
    const uint32_t lo2 = pin(p.relu ? 0u : (p.bf16 ? 0xFF80FF80u : 0xFC00FC00u), scr);   // max(x, -inf) = x: ReLU as a constant
    unsigned long long* const e_trace = (blockIdx.x == 0 && wq == 0 && lane == 0) ? pin(p.trace, scr) : nullptr;
    const uint32_t bar_full_e = pin(smem_u32(acc_full), scr);
    uint8_t* stage = st_smem + static_cast<size_t>(warp) * 4096;   // one 4 KB staging row set per warp
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    RowsWalker walk(p);
    RowsSegment sg;
    int it = 0;
    // output row number q (running) lives in block ring-1-(q % ring), use q / ring, and belongs to set q % n_sets: maintained
    // incrementally (three integer divisions per row otherwise)
    int blk_next = ring_i - 1, set_next = 0; uint32_t par_next = 0;
    // fused column sums (BiasAddGrad): a private [128] accumulator per epilogue warp in shared memory (behind the barrier /
    // plan area, allocated by the host only when p.colsum is set): no registers are held across rows for it
    float* const cs_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 3072) + warp * 128;
    const bool has_colsum = !SPLIT && pin(p.colsum != nullptr ? 1 : 0, scr) != 0;
    if (has_colsum) {
#pragma unroll
      for (int i = 0; i < 4; ++i) cs_smem[i * 32 + lane] = 0.f;
      __syncwarp();
    }
    while (walk.next(sg)) {
      for (int j = sg.y0; j < sg.y1; ++j, ++it) {
        const int blk = blk_next; const uint32_t par = par_next; const bool mine = (set_next == eset);
        if (--blk_next < 0) { blk_next = ring_i - 1; par_next ^= 1u; }
        if (++set_next == n_sets) set_next = 0;
        if (!mine) continue;
        const bool tr = e_trace && it < 64;
        if (tr) e_trace[it * 8 + 4] = clock64();
        const int x_in = sg.x0 + wq * 32 + lane;              // input-grid pixel of this thread (TMEM lane)
        const int x_warp = sg.x0 + wq * 32;
        // residual of a narrow layer (<= 32 channels): fetched before the accumulator wait so that its HBM latency
        // overlaps the MMAs of this row, and kept for the relu-copy pass
        const bool res_once = (p.residual != nullptr) && p.cout_store <= 32;
        uint4 rv0[4];
        if (res_once && x_in < p.W) {
          const size_t opix = (static_cast<size_t>(sg.n) * p.H + j) * p.W + x_in;
          const __half* rp = p.residual + opix * p.res_cstride + p.res_coff;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rv0[i] = (i * 8 < p.cout_store) ? __ldg(reinterpret_cast<const uint4*>(rp + i * 8)) : make_uint4(0, 0, 0, 0);
        }
        mbar_wait_addr(bar_full_e + 8u * blk, par);
        tc_fence_after();
        if (tr) e_trace[it * 8 + 5] = clock64();
        const uint32_t t_blk = t_lane + static_cast<uint32_t>(blk) * cpad_e;
        uint8_t* row = stage + lane * 128;
        // the single staging set may be rewritten once the previous store of this warp has finished reading it
        auto begin_rows = [&]() {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        };
        auto end_rows = [&](const CUtensorMap* m, int c0) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(m, stage, c0, x_warp, j, sg.n);
            tma_store_commit();
          }
        };
        if (epi_plain) {
          // The common layer: one 16-bit output, optional ReLU, nothing else.  Kept small on purpose - an epilogue warp shares
          // the ~6 KB L0 instruction cache of its SM sub-partition with the UMMA-issuing thread, and the general path below
          // (residual / mask / relu copy / fp32 / split outputs) spreads one row over ~5 KB of code.
          for (int g = 0; g < e_ngroups; ++g) {
#pragma unroll 1
            for (int cb = 0; cb < e_cout; cb += 32) {
              uint32_t v[32];
              __syncwarp();
              const bool last = (g == e_ngroups - 1 && cb + 32 >= e_cout);
              if (!(e_dbg & 2)) { tmem_ld_32x32(t_blk + static_cast<uint32_t>(g * e_group_c + cb), v); tmem_ld_wait(); }
              if (last) {   // the accumulator block is free as soon as its last column chunk sits in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[blk]);
              }
              if (e_dbg & 3) continue;
              const int sub = (cb >> 5) & 1;
              if (sub == 0) begin_rows();
              const float4* b4 = reinterpret_cast<const float4*>(bias_smem + cb);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 pk; uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
                const float4 ba = b4[2 * i], bb = b4[2 * i + 1];
                const float f0 = __uint_as_float(v[8 * i]) + ba.x, f1 = __uint_as_float(v[8 * i + 1]) + ba.y,
                            f2 = __uint_as_float(v[8 * i + 2]) + ba.z, f3 = __uint_as_float(v[8 * i + 3]) + ba.w,
                            f4 = __uint_as_float(v[8 * i + 4]) + bb.x, f5 = __uint_as_float(v[8 * i + 5]) + bb.y,
                            f6 = __uint_as_float(v[8 * i + 6]) + bb.z, f7 = __uint_as_float(v[8 * i + 7]) + bb.w;
                if (e_bf16) {
                  const __nv_bfloat162 l = *reinterpret_cast<const __nv_bfloat162*>(&lo2);
                  __nv_bfloat162 h;
                  h = __hmax2(__floats2bfloat162_rn(f0, f1), l); pw[0] = *reinterpret_cast<uint32_t*>(&h);
                  h = __hmax2(__floats2bfloat162_rn(f2, f3), l); pw[1] = *reinterpret_cast<uint32_t*>(&h);
                  h = __hmax2(__floats2bfloat162_rn(f4, f5), l); pw[2] = *reinterpret_cast<uint32_t*>(&h);
                  h = __hmax2(__floats2bfloat162_rn(f6, f7), l); pw[3] = *reinterpret_cast<uint32_t*>(&h);
                } else {
                  const __half2 l = *reinterpret_cast<const __half2*>(&lo2);
                  __half2 h;
                  h = __hmax2(__floats2half2_rn(f0, f1), l); pw[0] = *reinterpret_cast<uint32_t*>(&h);
                  h = __hmax2(__floats2half2_rn(f2, f3), l); pw[1] = *reinterpret_cast<uint32_t*>(&h);
                  h = __hmax2(__floats2half2_rn(f4, f5), l); pw[2] = *reinterpret_cast<uint32_t*>(&h);
                  h = __hmax2(__floats2half2_rn(f6, f7), l); pw[3] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(row + (((sub * 4 + i) ^ (lane & 7)) << 4)) = pk;
              }
              if (sub == 1 || cb + 32 >= e_cout) end_rows(&maps.out[g], cb & ~63);
            }
          }
          __syncwarp();
          if (tr) e_trace[it * 8 + 6] = clock64();
          continue;
        }
        // pass 0: primary output; then (optional) the fp16 relu(primary) copy; then (split mode) the low halves of the primary:
        // every extra pass re-reads the accumulator from TMEM
        const int n_pass = 1 + (p.has_relu_copy ? 1 : 0) + ((SPLIT && p.split_out) ? 1 : 0);
        for (int g = 0; g < p.ngroups; ++g) {
          for (int pass = 0; pass < n_pass; ++pass) {
            const bool lo_pass = SPLIT && p.split_out && pass == n_pass - 1;
            const bool relu_pass = p.has_relu_copy && pass == 1;
            const bool f32_rows = p.out_f32 && pass == 0;
            const bool do_relu = p.relu || relu_pass;
            const CUtensorMap* omap = lo_pass ? &maps.out_lo[g] : (relu_pass ? &maps.out_relu : &maps.out[g]);
            for (int cb = 0; cb < p.cout_store; cb += 32) {
              uint32_t v[32];
              __syncwarp();
              if (p.dbg & 3) {
                if (!(p.dbg & 2)) { tmem_ld_32x32(t_blk + static_cast<uint32_t>(g * p.group_c + cb), v); tmem_ld_wait(); }
                if (g == p.ngroups - 1 && pass == n_pass - 1 && cb + 32 >= p.cout_store) {
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&acc_empty[blk]);
                }
                continue;
              }
              tmem_ld_32x32(t_blk + static_cast<uint32_t>(g * p.group_c + cb), v);
              uint4 rv[4];
              const bool has_res = (p.residual != nullptr) && (x_in < p.W);
              if (has_res && res_once) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rv[i] = rv0[i];
              } else if (has_res) {
                const size_t opix = (static_cast<size_t>(sg.n) * p.H + j) * p.W + x_in;
                const __half* rp = p.residual + opix * p.res_cstride + p.res_coff + cb;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  rv[i] = (cb + i * 8 < p.cout_store) ? __ldg(reinterpret_cast<const uint4*>(rp + i * 8))
                                                      : make_uint4(0, 0, 0, 0);
              }
              tmem_ld_wait();
              // the accumulator block is free as soon as its last column chunk sits in registers: the issuing thread can
              // start the next row on it while this row is still being converted and stored
              if (g == p.ngroups - 1 && pass == n_pass - 1 && cb + 32 >= p.cout_store) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[blk]);
              }
              float* f = reinterpret_cast<float*>(v);
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] += bias_smem[cb + i];
              if (has_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __half2* rh = reinterpret_cast<const __half2*>(&rv[i]);
                  const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rv[i]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 r2 = p.bf16 ? __bfloat1622float2(rb[e]) : __half22float2(rh[e]);
                    if (p.res_is_mask) {      // backward of a ReLU fused into the input-gradient conv: y = conv(x) * [mask > 0]
                      if (!(r2.x > 0.f)) f[i * 8 + 2 * e] = 0.f;
                      if (!(r2.y > 0.f)) f[i * 8 + 2 * e + 1] = 0.f;
                    } else {
                      f[i * 8 + 2 * e] += r2.x; f[i * 8 + 2 * e + 1] += r2.y;
                    }
                  }
                }
              }
              if (do_relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
              }
              if (f32_rows) {
                // 32 fp32 channels fill one 128-byte staging row
                begin_rows();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  *reinterpret_cast<float4*>(row + ((i ^ (lane & 7)) << 4)) =
                      make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                end_rows(omap, cb);
              } else {
                // 32 fp16 channels fill half a staging row; flush every 64 channels
                const int sub = (cb >> 5) & 1;
                if (sub == 0) begin_rows();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  uint4 pk; __half2* ph = reinterpret_cast<__half2*>(&pk);
                  __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&pk);
                  if (p.bf16) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) pb[e] = __floats2bfloat162_rn(f[i * 8 + 2 * e], f[i * 8 + 2 * e + 1]);
                  } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      ph[e] = __floats2half2_rn(f[i * 8 + 2 * e], f[i * 8 + 2 * e + 1]);
                      if (SPLIT && lo_pass) {      // what the fp16 rounding of the primary pass dropped
                        const float2 hi = __half22float2(ph[e]);
                        ph[e] = __floats2half2_rn(f[i * 8 + 2 * e] - hi.x, f[i * 8 + 2 * e + 1] - hi.y);
                      }
                    }
                  }
                  *reinterpret_cast<uint4*>(row + (((sub * 4 + i) ^ (lane & 7)) << 4)) = pk;
                }
                if (sub == 1 || cb + 32 >= p.cout_store) end_rows(omap, cb & ~63);
              }
              if (has_colsum && pass == 0) {
                // column sums of this 32-pixel x 32-channel block: a transpose-reduce over the warp (31 shuffles: in round
                // `half` a lane keeps the half of its values whose channel bit matches its lane bit and receives the partner's
                // partial sums for them), lane L ends with the sum of channel cb + L
                // (in place, after the block has been staged for the store: f is dead afterwards)
                const bool live = x_in < p.W;
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = live ? f[i] : 0.f;
#pragma unroll
                for (int half = 16; half >= 1; half >>= 1) {
                  const bool up = (lane & half) != 0;
#pragma unroll
                  for (int i = 0; i < half; ++i) {
                    const float send = up ? f[i] : f[i + half];
                    const float keep = up ? f[i + half] : f[i];
                    f[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
                  }
                }
                cs_smem[cb + lane] += f[0];
              }
            }
          }
        }
        __syncwarp();
        if (tr) e_trace[it * 8 + 6] = clock64();
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
    if (has_colsum) {
      const int cout = p.cout_store;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i * 32 + lane < cout) atomicAdd(p.colsum + i * 32 + lane, cs_smem[i * 32 + lane]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kRowsWarpTmem) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dd
