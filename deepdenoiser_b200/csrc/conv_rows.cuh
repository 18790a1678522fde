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

  if (warp == kRowsWarpA) {
    // ------------------------------------------------------------------ A producer: one input row x one 64ch chunk per item
    if (elect_one()) {
      RowsWalker walk(p);
      RowsSegment sg;
      int slot = 0; uint32_t phase = 0;
      while (walk.next(sg)) {
        const int t_first = sg.y0 + t_lo_off, t_last = sg.y1 + t_hi_off;
        for (int tg = t_first; tg <= t_last; tg += p.G) {
          const int gcur = (t_last + 1 - tg < p.G) ? (t_last + 1 - tg) : p.G;
          for (int c = 0; c < p.n_chunks; ++c) {
            for (int g = 0; g < gcur; ++g) {
              mbar_wait(&a_empty[slot], phase ^ 1);
              mbar_arrive_expect_tx(&a_full[slot], p.a_tx_bytes);
              const int part = SPLIT ? c / p.nb : 0;
              tma_load_4d(a_smem + static_cast<size_t>(slot) * p.a_slot_bytes, (SPLIT && part == 1) ? &maps.a_lo : &maps.a,
                          &a_full[slot], (c - part * p.nb) * 64, sg.x0 - p.halo, tg + g, sg.n);
              if (++slot == p.a_slots) { slot = 0; phase ^= 1; }
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
        RowsWalker walk(p);
        RowsSegment sg;
        int stage = 0; uint32_t phase = 0;
        while (walk.next(sg)) {
          const int t_first = sg.y0 + t_lo_off, t_last = sg.y1 + t_hi_off;
          for (int tg = t_first; tg <= t_last; tg += p.G) {
            for (int c = 0; c < p.n_chunks; ++c) {
              for (int si = 0; si < p.n_s; ++si) {
                mbar_wait(&b_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&b_full[stage], p.b_tx_bytes);
                const int bc = (SPLIT && c >= p.nb) ? c - p.nb : c;                        // weight chunk of A chunk c
                tma_load_4d(b_smem + static_cast<size_t>(stage) * p.b_tile_bytes, &maps.b, &b_full[stage], 0, p.b_row0,
                            p.b_r0, bc * p.tiles_per_chunk + p.s_list[si]);
                if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == kRowsWarpMma) {
    // ------------------------------------------------------------------ MMA issuer (highest warp id: the scheduler
    // prefers it over the epilogue warp sharing its sub-partition)
    if (elect_one()) {
      // Single issuing thread: scalar integer code on the critical path of the tensor pipe (one UMMA of N = 192
      // lasts 96 cycles), so the row loop avoids divisions, parameter loads and recomputation: the UMMA "pieces"
      // a row breaks into (ring wrap / N <= 256 / first touch of an output row) are tabulated once per block index.
      //   piece = {TMEM column, instruction descriptor, B offset >> 4, accumulate};  row plan = hdr + 3 normal + 4 first-touch
      const uint64_t desc_tmpl = make_desc_sw128(0, 0);
      const uint32_t d_lo = static_cast<uint32_t>(desc_tmpl), d_hi = static_cast<uint32_t>(desc_tmpl >> 32);
      const uint32_t a_base = smem_u32(a_smem), b_base = smem_u32(b_smem);
      uint4* plan = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [8 rows of a group][8]
      uint4* table = plan + 64;                                                         // [ring blocks][8], full tap range
      const int ring = p.ring, rm_lo = p.rm_lo, rm_hi = p.rm_hi, n_s = p.n_s, n_chunks = p.n_chunks, a_slots = p.a_slots;
      const uint32_t a_slot_bytes = p.a_slot_bytes, b_tile_bytes = p.b_tile_bytes;
      const uint32_t s_off0 = static_cast<uint32_t>(p.s_list[0]) * 8u, s_off1 = static_cast<uint32_t>(p.s_list[1]) * 8u,
                     s_off2 = static_cast<uint32_t>(p.s_list[2]) * 8u;                   // shifts in descriptor units

      const uint32_t fmt = p.bf16 ? kIdescBf16 : 0u;
      auto build_plan = [&](uint4* row_plan, int blk, uint32_t use, int r_lo, int r_hi) {
        const bool ft = (r_lo == rm_lo);   // this input row initialises the accumulator of the row served by tap r_lo
        int n_norm = 0, n_first = 0;
        int r = r_lo, bk = blk;
        while (r <= r_hi) {
          int cnt = r_hi - r + 1;
          if (bk + cnt > ring) cnt = ring - bk;
          if (cnt > p.max_stack) cnt = p.max_stack;
          row_plan[1 + n_norm++] = make_uint4(static_cast<uint32_t>(bk * p.cpad), make_idesc_f16(kRowsTileW, cnt * p.cpad) | fmt,
                                              static_cast<uint32_t>((r - rm_lo) * p.cpad) * 8u, 1u);
          r += cnt; bk += cnt; if (bk >= ring) bk -= ring;
        }
        if (ft) {
          row_plan[4 + n_first++] = make_uint4(static_cast<uint32_t>(blk * p.cpad), make_idesc_f16(kRowsTileW, p.cpad) | fmt,
                                               static_cast<uint32_t>((r_lo - rm_lo) * p.cpad) * 8u, 0u);
          r = r_lo + 1; bk = blk + 1; if (bk >= ring) bk -= ring;
          while (r <= r_hi) {
            int cnt = r_hi - r + 1;
            if (bk + cnt > ring) cnt = ring - bk;
            if (cnt > p.max_stack) cnt = p.max_stack;
            row_plan[4 + n_first++] = make_uint4(static_cast<uint32_t>(bk * p.cpad), make_idesc_f16(kRowsTileW, cnt * p.cpad) | fmt,
                                                 static_cast<uint32_t>((r - rm_lo) * p.cpad) * 8u, 1u);
            r += cnt; bk += cnt; if (bk >= ring) bk -= ring;
          }
        }
        row_plan[0] = make_uint4(static_cast<uint32_t>(n_norm), static_cast<uint32_t>(n_first), static_cast<uint32_t>(blk),
                                 (use & 1u) ^ 1u);
      };
      // issues the UMMAs of one (input row, chunk, shift): k-steps x pieces, operands already resident
      auto issue = [&](const uint4& hdr, const uint4 (&pn)[3], const uint4 (&pf)[4], bool first_pass, uint32_t a_lo,
                       uint32_t b_lo0, int ksteps) {
        int k0 = 0;
        if (first_pass) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < static_cast<int>(hdr.y))
              umma_f16(tmem_base + pf[i].x, desc_from(a_lo, d_hi), desc_from(b_lo0 + pf[i].z, d_hi), pf[i].y, pf[i].w);
          k0 = 1;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i < static_cast<int>(hdr.x)) {
            const uint32_t b_lo = b_lo0 + pn[i].z;
            const uint32_t d_col = tmem_base + pn[i].x;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k >= k0 && k < ksteps)
                umma_f16(d_col, desc_from(a_lo + 2u * k, d_hi), desc_from(b_lo + 2u * k, d_hi), pn[i].y, 1u);
          }
        }
      };

      for (int bk = 0; bk < ring; ++bk) build_plan(table + bk * 8, bk, 0u, rm_lo, rm_hi);
      RowsWalker walk(p);
      RowsSegment sg;
      int a_slot = 0; uint32_t a_phase = 0;
      int b_stage = 0; uint32_t b_phase = 0;
      uint32_t q_seg = 0;                  // running index of the first output row of the segment
      int it = 0;
      if (p.w_resident) { mbar_wait(&b_full[0], 0); tc_fence_after(); }
      while (walk.next(sg)) {
        const int t_first = sg.y0 + t_lo_off, t_last = sg.y1 + t_hi_off;
        // accumulator block / use count of the output row fed by tap r = 0 of input row t (index q_seg + t + 1 - y0);
        // tap r lands r blocks further.  Maintained incrementally (one division per segment).
        const uint32_t q_top0 = q_seg + static_cast<uint32_t>(t_first + 1 - sg.y0);
        int blk_top = ring - 1 - static_cast<int>(q_top0 % static_cast<uint32_t>(ring));
        uint32_t use_top = q_top0 / static_cast<uint32_t>(ring);
        if (p.G == 1) {
          // ---------------- one input row at a time (weights resident, or a single-row weight pass)
          // The barrier probes of row t+1 (its first A slot, the accumulator block it initialises) are issued right
          // after the waits of row t: their ~100-200 cycle latency then overlaps the UMMAs of row t instead of idling
          // the tensor pipe between rows; a failed probe falls back to the blocking wait.
          bool probed = false, probe_a = false, probe_e = false;
          for (int t = t_first; t <= t_last; ++t, ++it) {
            const bool tr = p.trace && blockIdx.x == 0 && it < 64;
            if (tr) p.trace[it * 8 + 0] = clock64();
            int r_lo = t + 2 - sg.y1; if (r_lo < rm_lo) r_lo = rm_lo;     // taps landing on rows of the segment
            int r_hi = t + 1 - sg.y0; if (r_hi > rm_hi) r_hi = rm_hi;
            int blk = blk_top + r_lo; uint32_t use = use_top;
            if (blk >= ring) { blk -= ring; use -= 1u; }                  // block / use of the row served by tap r_lo
            const uint4* row_plan;
            if (r_lo == rm_lo && r_hi == rm_hi) {
              row_plan = table + blk * 8;
            } else {
              build_plan(plan, blk, use, r_lo, r_hi);
              row_plan = plan;
            }
            const uint4 hdr = row_plan[0];
            uint4 pn[3], pf[4];
#pragma unroll
            for (int i = 0; i < 3; ++i) pn[i] = row_plan[1 + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) pf[i] = row_plan[4 + i];
            const bool initialises = (r_lo == rm_lo);
            if (tr) p.trace[it * 8 + 1] = clock64();                                      // plan in registers
            for (int c = 0; c < n_chunks; ++c) {
              int cc = c;
              if (SPLIT) { while (cc >= p.nb) cc -= p.nb; }
              const int ksteps = (cc == p.nb - 1) ? p.ksteps_last : 4;
              const int bc = (SPLIT && c >= 2 * p.nb) ? c - p.nb : cc;                     // resident weight chunk of A chunk c
              if (!(c == 0 && probed && probe_a)) mbar_wait(&a_full[a_slot], a_phase);
              if (c == 0 && initialises && !(probed && probe_e))
                mbar_wait(&acc_empty[blk], (use & 1u) ^ 1u);                              // previous user drained
              if (tr && c == 0) p.trace[it * 8 + 7] = clock64() | (static_cast<unsigned long long>((probed && probe_a) ? 1 : 0) << 62) |
                                                      (static_cast<unsigned long long>((probed && probe_e) ? 1 : 0) << 61);
              tc_fence_after();
              if (tr && c == 0) p.trace[it * 8 + 2] = clock64();
              if (c == 0) {
                probed = false;
                if (t < t_last) {
                  int nslot = a_slot + n_chunks; uint32_t nphase = a_phase;
                  while (nslot >= a_slots) { nslot -= a_slots; nphase ^= 1; }
                  int nr_lo = t + 3 - sg.y1; if (nr_lo < rm_lo) nr_lo = rm_lo;
                  int nbt = blk_top - 1; uint32_t nut = use_top;
                  if (nbt < 0) { nbt = ring - 1; ++nut; }
                  int nblk = nbt + nr_lo;
                  if (nblk >= ring) { nblk -= ring; nut -= 1u; }
                  probe_a = mbar_test_wait(&a_full[nslot], nphase);
                  probe_e = (nr_lo != rm_lo) || mbar_test_wait(&acc_empty[nblk], (nut & 1u) ^ 1u);
                  probed = true;
                }
              }
              const uint32_t a_lo0 = d_lo + ((a_base + static_cast<uint32_t>(a_slot) * a_slot_bytes) >> 4);
              for (int si = 0; si < n_s; ++si) {
                uint32_t b_tile;
                if (p.w_resident) {
                  b_tile = b_base + static_cast<uint32_t>(bc * n_s + si) * b_tile_bytes;
                } else {
                  mbar_wait(&b_full[b_stage], b_phase);
                  tc_fence_after();
                  b_tile = b_base + static_cast<uint32_t>(b_stage) * b_tile_bytes;
                }
                const uint32_t s_off = (si == 0) ? s_off0 : ((si == 1) ? s_off1 : s_off2);
                issue(hdr, pn, pf, initialises && c == 0 && si == 0, a_lo0 + s_off, d_lo + (b_tile >> 4), ksteps);
                if (!p.w_resident) {
                  umma_commit(&b_empty[b_stage]);
                  if (++b_stage == p.b_stages) { b_stage = 0; b_phase ^= 1; }
                }
              }
              umma_commit(&a_empty[a_slot]);
              if (++a_slot == a_slots) { a_slot = 0; a_phase ^= 1; }
            }
            // output row completed by this input row: the one served by tap rm_hi, if it belongs to the segment
            {
              const int j = t + 1 - rm_hi;
              if (j >= sg.y0 && j < sg.y1) {
                int blk_done = blk_top + rm_hi; if (blk_done >= ring) blk_done -= ring;
                umma_commit(&acc_full[blk_done]);
              }
            }
            if (--blk_top < 0) { blk_top = ring - 1; ++use_top; }
            if (tr) p.trace[it * 8 + 3] = clock64();
          }
        } else {
          // ---------------- groups of G input rows per weight pass (streamed weights)
          for (int tg = t_first; tg <= t_last; tg += p.G, ++it) {
            const int gcur = (t_last + 1 - tg < p.G) ? (t_last + 1 - tg) : p.G;
            const bool tr = p.trace && blockIdx.x == 0 && it < 64;
            if (tr) p.trace[it * 8 + 0] = clock64();
            {
              int bt = blk_top; uint32_t ut = use_top;
              for (int g = 0; g < gcur; ++g) {
                const int t = tg + g;
                int r_lo = t + 2 - sg.y1; if (r_lo < rm_lo) r_lo = rm_lo;
                int r_hi = t + 1 - sg.y0; if (r_hi > rm_hi) r_hi = rm_hi;
                int blk = bt + r_lo; uint32_t use = ut;
                if (blk >= ring) { blk -= ring; use -= 1u; }
                if (r_lo == rm_lo && r_hi == rm_hi) {
                  // interior row: the tabulated plan of this block index, with the wait parity of this use patched in
                  // (building a plan costs ~500 cycles of this thread, 8 % of a 96-channel group)
                  uint4* dst = plan + g * 8;
                  const uint4* src = table + blk * 8;
#pragma unroll
                  for (int i = 0; i < 8; ++i) dst[i] = src[i];
                  dst[0].w = (use & 1u) ^ 1u;
                } else {
                  build_plan(plan + g * 8, blk, use, r_lo, r_hi);
                }
                if (--bt < 0) { bt = ring - 1; ++ut; }
              }
            }
            const int a_slot0 = a_slot; const uint32_t a_phase0 = a_phase;
            if (tr) p.trace[it * 8 + 1] = clock64();                                      // plans built
            long long wait_b = 0, wait_a = 0;
            for (int c = 0; c < n_chunks; ++c) {
              int cc = c;
              if (SPLIT) { while (cc >= p.nb) cc -= p.nb; }
              const int ksteps = (cc == p.nb - 1) ? p.ksteps_last : 4;
              for (int si = 0; si < n_s; ++si) {
                const long long tb0 = tr ? clock64() : 0;
                mbar_wait(&b_full[b_stage], b_phase);
                tc_fence_after();
                if (tr) wait_b += clock64() - tb0;
                const uint32_t b_lo0 = d_lo + ((b_base + static_cast<uint32_t>(b_stage) * b_tile_bytes) >> 4);
                const uint32_t s_off = (si == 0) ? s_off0 : ((si == 1) ? s_off1 : s_off2);
                int lin_slot = a_slot0 + c * gcur;      // slots of this chunk's rows: a_slot0 + c * gcur + g (mod a_slots)
                uint32_t ph = a_phase0;
                while (lin_slot >= a_slots) { lin_slot -= a_slots; ph ^= 1; }
                for (int g = 0; g < gcur; ++g) {
                  const uint4* row_plan = plan + g * 8;
                  const uint4 hdr = row_plan[0];
                  const bool first_pass = (c == 0 && si == 0 && hdr.y != 0);
                  uint4 pn[3], pf[4];
#pragma unroll
                  for (int i = 0; i < 3; ++i) pn[i] = row_plan[1 + i];
#pragma unroll
                  for (int i = 0; i < 4; ++i) pf[i] = row_plan[4 + i];
                  if (si == 0) {
                    const long long ta0 = tr ? clock64() : 0;
                    mbar_wait(&a_full[lin_slot], ph);
                    if (first_pass) mbar_wait(&acc_empty[hdr.z], hdr.w);    // previous user of the new row's block drained
                    tc_fence_after();
                    if (tr) wait_a += clock64() - ta0;
                    if (tr && c == 0 && g == 0) p.trace[it * 8 + 2] = clock64();
                  }
                  const uint32_t a_lo = d_lo + ((a_base + static_cast<uint32_t>(lin_slot) * a_slot_bytes) >> 4) + s_off;
                  issue(hdr, pn, pf, first_pass, a_lo, b_lo0, ksteps);
                  if (si == n_s - 1) umma_commit(&a_empty[lin_slot]);
                  if (++lin_slot == a_slots) { lin_slot = 0; ph ^= 1; }
                }
                umma_commit(&b_empty[b_stage]);
                if (++b_stage == p.b_stages) { b_stage = 0; b_phase ^= 1; }
              }
            }
            // advance the A ring past this group
            a_slot = a_slot0 + n_chunks * gcur; a_phase = a_phase0;
            while (a_slot >= a_slots) { a_slot -= a_slots; a_phase ^= 1; }
            // output rows completed by this group: tap rm_hi of each input row, if that row belongs to the segment
            for (int g = 0; g < gcur; ++g) {
              const int j = tg + g + 1 - rm_hi;
              if (j >= sg.y0 && j < sg.y1) {
                int blk_done = blk_top + rm_hi; if (blk_done >= ring) blk_done -= ring;
                umma_commit(&acc_full[blk_done]);
              }
              if (--blk_top < 0) { blk_top = ring - 1; ++use_top; }
            }
            if (tr) { p.trace[it * 8 + 3] = clock64(); p.trace[512 + it] = wait_b; p.trace[576 + it] = wait_a; }
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
    const uint32_t n_sets = (p.ring < kRowsEpiSets) ? static_cast<uint32_t>(p.ring) : static_cast<uint32_t>(kRowsEpiSets);
    uint8_t* stage = st_smem + static_cast<size_t>(warp) * 4096;   // one 4 KB staging row set per warp
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    RowsWalker walk(p);
    RowsSegment sg;
    uint32_t q = 0;
    int it = 0;
    const uint32_t ring = static_cast<uint32_t>(p.ring);
    while (walk.next(sg)) {
      for (int j = sg.y0; j < sg.y1; ++j, ++q, ++it) {
        if (q % n_sets != static_cast<uint32_t>(eset)) continue;
        const int blk = p.ring - 1 - static_cast<int>(q % ring);
        const uint32_t use = q / ring;
        const bool tr = p.trace && blockIdx.x == 0 && it < 64 && wq == 0 && lane == 0;
        if (tr) p.trace[it * 8 + 4] = clock64();
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
        mbar_wait(&acc_full[blk], use & 1u);
        tc_fence_after();
        if (tr) p.trace[it * 8 + 5] = clock64();
        const uint32_t t_blk = t_lane + static_cast<uint32_t>(blk * p.cpad);
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
            }
          }
        }
        __syncwarp();
        if (tr) p.trace[it * 8 + 6] = clock64();
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kRowsWarpTmem) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dd
