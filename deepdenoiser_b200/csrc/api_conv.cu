// C-ABI: context, weight packing and the convolution entry points (tcgen05 fp16 path + exact fp32 SIMT path).
#include <string.h>

#include <vector>

#include "conv_rows.cuh"
#include "dd_internal.h"

namespace dd {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// Exact path: direct convolution on CUDA cores, fp32 accumulate, fp32 or fp16 storage.
// One thread = one output pixel x 8 output channels; weights [tap][cout][cin] fp32.
// Used when activations are DD_F32 (parity mode) — the fp16 tensor-core path is the fast one.
struct ConvSimtParams {
  View x, y, res, yrelu;
  const float* w;      // [taps][cout][cin]
  const float* bias;
  int cout, ksize, relu, has_res, has_yrelu;
  int ups, ay, ax;     // pixel shuffle for transposed 2x2 (ups == 2)
  int x_ups, x_ay, x_ax;  // input read through a stride-2 sub-pixel lattice (input gradient of the transposed 2x2 conv)
  int oh, ow;          // output grid (== input grid unless x_ups == 2)
};

constexpr int kSimtCob = 8;

__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvSimtParams p) {
  const int cgroups = (p.cout + kSimtCob - 1) / kSimtCob;
  const size_t npix = static_cast<size_t>(p.x.n) * p.oh * p.ow;
  const size_t total = npix * cgroups;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // consecutive threads -> consecutive pixels (same channel group => weight loads are warp-uniform)
  const int cg = static_cast<int>(idx / npix);
  const size_t pixel = idx % npix;
  const int x0 = static_cast<int>(pixel % p.ow);
  const int y0 = static_cast<int>((pixel / p.ow) % p.oh);
  const int n0 = static_cast<int>(pixel / (static_cast<size_t>(p.ow) * p.oh));
  const int co0 = cg * kSimtCob;
  const int cin = p.x.c;
  const int pad = (p.ksize - 1) / 2;

  float acc[kSimtCob];
#pragma unroll
  for (int i = 0; i < kSimtCob; ++i) acc[i] = 0.f;

  for (int r = 0; r < p.ksize; ++r) {
    int yy = y0 + r - pad;
    if (yy < 0 || yy >= p.oh) continue;
    if (p.x_ups == 2) { yy = 2 * yy + p.x_ay; if (yy >= p.x.h) continue; }   // tap 2 of a 3x3 stride-2 lattice leaves the image
    for (int s = 0; s < p.ksize; ++s) {
      int xx = x0 + s - pad;
      if (xx < 0 || xx >= p.ow) continue;
      if (p.x_ups == 2) { xx = 2 * xx + p.x_ax; if (xx >= p.x.w) continue; }
      const size_t ipix = p.x.pix(n0, yy, xx);
      const float* wt = p.w + (static_cast<size_t>(r * p.ksize + s) * p.cout + co0) * cin;
      for (int c = 0; c < cin; ++c) {
        const float xv = p.x.load(ipix, c);
#pragma unroll
        for (int i = 0; i < kSimtCob; ++i) {
          if (co0 + i < p.cout) acc[i] = fmaf(xv, __ldg(wt + static_cast<size_t>(i) * cin + c), acc[i]);
        }
      }
    }
  }
  const int oy = (p.ups == 2) ? 2 * y0 + p.ay : y0;
  const int ox = (p.ups == 2) ? 2 * x0 + p.ax : x0;
  const size_t opix = p.y.pix(n0, oy, ox);
#pragma unroll
  for (int i = 0; i < kSimtCob; ++i) {
    const int co = co0 + i;
    if (co >= p.cout) break;
    float v = acc[i] + (p.bias ? p.bias[co] : 0.f);
    if (p.has_res) v += p.res.load(opix, co);
    if (p.has_yrelu) p.yrelu.store(opix, co, fmaxf(v, 0.f));
    if (p.relu) v = fmaxf(v, 0.f);
    p.y.store(opix, co, v);
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(dd_ctx* ctx, CUtensorMap* map, CUtensorMapDataType dt, void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapL2promotion promo) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, dt, static_cast<cuuint32_t>(rank), base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu %llu strides %llu %llu box %u %u %u)",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)dims[2], (unsigned long long)dims[3], (unsigned long long)strides[0],
              (unsigned long long)strides[1], box[0], box[1], box[2]);
    return DD_ERR_CUDA;
  }
  return DD_OK;
}

// Packed fp16 weight tensor of the row-pipeline kernel: [chunk][shift s][tap r][cpad_total][64 ch], zero padded.
//   3x3: s, r in 0..2 (tap = r*3 + s);  1x1: one tile per chunk;  transposed 2x2: 1x1 with rows (sub-pixel, cpad)
static int packed_cpad(int cout) { return round_up(cout, 32); }

struct TcLaunch {
  const dd_tensor* x;
  const void* w_packed;
  int ksize;                // 3 or 1
  int rows_total;           // cpad rows per (chunk, s, r) in the packed tensor
  int row0;                 // first row (output channel) used by this launch
  int cpad;                 // rows used == accumulator block width
  int ngroups, group_c;     // column groups inside the block (sub-pixels of conv2d_transpose 2x2)
  int cout;                 // channels stored per group
  const float* bias;
  int bias_count;
  uint32_t flags;
  const dd_tensor* residual;
  const dd_tensor* y;       // for ngroups > 1 / ups == 2: the full-resolution output
  const dd_tensor* y_relu;
  int ups, sp0;             // ups == 2: group g -> sub-pixel sp0 + g -> (ay, ax) = (sp >> 1, sp & 1)
  int s_mask, rm_lo, rm_hi; // tap subset (3x3 only); s_mask == 0 -> all
  const dd_tensor* x_lo;    // split-fp16 mode: low halves of the input (same shape as x) / of the 16-bit output
  const dd_tensor* y_lo;
  float* colsum;            // optional column sums of the output (device, fp32, accumulated), see ConvRowsParams
};

// view of `t` seen through a stride-`ups` sub-pixel lattice starting at (ay, ax)
static int encode_out_map(dd_ctx* ctx, CUtensorMap* map, const dd_tensor* t, int c_first, int channels, int in_h, int in_w,
                          int ups, int ay, int ax) {
  const size_t es = elem_size(t->dtype);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(channels), static_cast<cuuint64_t>(in_w), static_cast<cuuint64_t>(in_h),
                        static_cast<cuuint64_t>(t->n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ups) * t->cstride * es,
                           static_cast<cuuint64_t>(ups) * t->w * t->cstride * es,
                           static_cast<cuuint64_t>(t->h) * t->w * t->cstride * es};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(128 / es), 32, 1, 1};
  uint8_t* base = reinterpret_cast<uint8_t*>(t->ptr) +
                  ((static_cast<size_t>(ay) * t->w + ax) * t->cstride + t->coff + c_first) * es;
  return encode_map(ctx, map, t->dtype == DD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 :
                    (t->dtype == DD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), base,
                    4, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_NONE);
}

static int launch_conv_rows(dd_ctx* ctx, const TcLaunch& L, cudaStream_t stream) {
  const dd_tensor* x = L.x;
  const dd_tensor* y = L.y;
  DD_CHECK_ARG(ctx->encode_tiled, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  DD_CHECK_ARG(is_half_type(x->dtype), "tensor-core conv needs fp16 / bf16 input");
  DD_CHECK_ARG(y->dtype == DD_F32 || y->dtype == x->dtype, "tensor-core conv writes fp32 or the input's 16-bit type");
  DD_CHECK_ARG(x->coff % 8 == 0 && x->cstride % 8 == 0, "conv input view must be 16-byte aligned (coff %d cstride %d)",
               x->coff, x->cstride);
  DD_CHECK_ARG(L.cpad % 32 == 0 && L.cpad >= 32 && L.cpad <= 256, "accumulator block width %d unsupported", L.cpad);
  const int align_out = is_half_type(y->dtype) ? 8 : 4;
  DD_CHECK_ARG(y->coff % align_out == 0 && y->cstride % align_out == 0, "conv output view misaligned");

  ConvRowsParams p;
  memset(&p, 0, sizeof(p));
  p.N = x->n; p.H = x->h; p.W = x->w;
  p.strips = (p.W + kRowsTileW - 1) / kRowsTileW;
  p.total_rows = static_cast<long long>(p.N) * p.strips * p.H;
  const int cin16 = round_up(x->c, 16);
  p.nb = (cin16 + 63) / 64;
  p.ksteps_last = (cin16 - 64 * (p.nb - 1)) / 16;
  p.split = L.x_lo ? 1 : 0;
  p.n_chunks = p.split ? 3 * p.nb : p.nb;
  const int w_chunks = p.split ? 2 * p.nb : p.nb;          // chunks of the packed weight tensor ([W_hi | W_lo] when split)
  if (p.split) {
    DD_CHECK_ARG(x->dtype == DD_F16 && L.x_lo->dtype == DD_F16 && L.x_lo->c == x->c && L.x_lo->n == x->n && L.x_lo->h == x->h &&
                     L.x_lo->w == x->w && L.x_lo->coff % 8 == 0 && L.x_lo->cstride % 8 == 0, "split conv: bad low-half input view");
    DD_CHECK_ARG(!L.residual && !L.y_relu, "split conv: no residual / relu copy");
  }
  if (L.y_lo) {
    DD_CHECK_ARG(p.split && y->dtype == DD_F16 && L.y_lo->dtype == DD_F16 && L.y_lo->c == y->c && L.y_lo->coff % 8 == 0 &&
                     L.y_lo->cstride % 8 == 0, "split conv: bad low-half output view");
    p.split_out = 1;
  }
  const bool k3 = (L.ksize == 3);
  p.halo = k3 ? 1 : 0;
  if (k3) {
    const int mask = L.s_mask ? L.s_mask : 7;
    for (int s = 0; s < 3; ++s)
      if (mask & (1 << s)) p.s_list[p.n_s++] = s;
    p.rm_lo = L.rm_lo; p.rm_hi = L.rm_hi;
    p.tiles_per_chunk = 3;
    p.b_r0 = L.rm_lo;
  } else {
    p.n_s = 1; p.s_list[0] = 0;
    p.rm_lo = p.rm_hi = 1;
    p.tiles_per_chunk = 1;
    p.b_r0 = 0;
  }
  const int n_r = p.rm_hi - p.rm_lo + 1;
  p.cpad = L.cpad;
  p.b_row0 = L.row0;
  p.ring = 512 / p.cpad; if (p.ring > kRowsMaxRing) p.ring = kRowsMaxRing;
  p.max_stack = 256 / p.cpad;
  DD_CHECK_ARG(p.ring >= n_r + 1 || n_r == 1, "accumulator ring too small (cpad %d)", p.cpad);
  DD_CHECK_ARG(n_r <= 2 * p.max_stack, "taps of a row need more than two UMMA pieces (cpad %d)", p.cpad);

  // shared memory plan
  const uint32_t a_box_w = k3 ? kRowsTileW + 2 : kRowsTileW;
  p.a_tx_bytes = a_box_w * 128u;
  p.a_slot_bytes = static_cast<uint32_t>(round_up(static_cast<int>(p.a_tx_bytes), 1024));
  p.b_tile_bytes = static_cast<uint32_t>(n_r * p.cpad) * 128u;
  p.b_tx_bytes = p.b_tile_bytes;
  const size_t stage_bytes = static_cast<size_t>(kRowsEpiWarps) * 4096;   // one 4 KB staging row set per epilogue warp
  const size_t colsum_bytes = L.colsum ? static_cast<size_t>(kRowsEpiWarps) * 128 * sizeof(float) : 0;   // per-warp column sums
  const size_t fixed = stage_bytes + 1024 /*bias*/ + 3072 /*barriers + MMA plans*/ + colsum_bytes;
  const size_t avail = ctx->max_smem_optin - 1024 /*alignment slack*/ - fixed;
  const size_t w_total = static_cast<size_t>(w_chunks) * p.n_s * p.b_tile_bytes;
  size_t b_bytes;
  if (!ctx->conv_force_stream && w_total + 4ull * p.a_slot_bytes <= avail) {
    p.w_resident = 1; p.b_stages = 1; p.G = 1;
    b_bytes = w_total;
    p.a_slots = static_cast<int>((avail - w_total) / p.a_slot_bytes);
    if (p.a_slots > 8) p.a_slots = 8;
  } else {
    p.w_resident = 0;
    // measured (tools/probe_conv_stream.py, round 2): L2 feeds the weight stream easily, what matters is that the epilogue
    // overlaps the next rows - rows of a group complete together, so small groups + a deeper weight pipeline win when the
    // TMEM ring is short (128-column blocks: 3 stages, 1 row per pass: 128->128 850 -> 942 TFLOP/s; 96-column blocks: 2 rows
    // per pass: 96->96 814 -> 868)
    p.b_stages = (ctx->conv_b_stages >= 2 && ctx->conv_b_stages <= 6) ? ctx->conv_b_stages : (p.cpad >= 128 ? 3 : 2);
    b_bytes = static_cast<size_t>(p.b_stages) * p.b_tile_bytes;
    if (!ctx->conv_b_stages && b_bytes + 2ull * p.a_slot_bytes > avail) { p.b_stages = 2; b_bytes = 2ull * p.b_tile_bytes; }
    DD_CHECK_ARG(b_bytes + 2ull * p.a_slot_bytes <= avail, "weight tile too large for shared memory (cpad %d)", p.cpad);
    p.a_slots = static_cast<int>((avail - b_bytes) / p.a_slot_bytes);
    if (p.a_slots > 12) p.a_slots = 12;
    int g = p.ring - n_r;                    // live blocks of a group: G + (n_r - 1), one more being drained
    if (n_r == 1) g = p.ring - 1;
    if (g > p.a_slots / 2) g = p.a_slots / 2;
    if (g < 1) g = 1;
    if (g > 8) g = 8;
    if (ctx->conv_rows > 0 && ctx->conv_rows < g) g = ctx->conv_rows;
    p.G = g;
    if (!ctx->conv_b_stages && p.b_stages == 2 &&
        b_bytes + static_cast<size_t>(p.a_slots) * p.a_slot_bytes + p.b_tile_bytes <= avail) {
      p.b_stages = 3; b_bytes += p.b_tile_bytes;
    }
  }
  DD_CHECK_ARG(p.a_slots >= p.G && p.a_slots >= 2, "no shared-memory configuration fits (cpad %d chunks %d)", p.cpad, p.n_chunks);
  p.a_off = 0;
  p.b_off = static_cast<uint32_t>(p.a_slots) * p.a_slot_bytes;
  p.stage_off = p.b_off + static_cast<uint32_t>(round_up(static_cast<int>(b_bytes), 1024));
  p.bias_off = p.stage_off + static_cast<uint32_t>(stage_bytes);
  p.bar_off = p.bias_off + 1024;
  const size_t smem = 1024 + p.bar_off + 3072 + colsum_bytes;

  // epilogue
  p.ngroups = L.ngroups; p.group_c = L.group_c; p.cout_store = L.cout; p.ups = L.ups;
  p.relu = (L.flags & DD_CONV_RELU) ? 1 : 0;
  p.out_f32 = (y->dtype == DD_F32);
  p.bf16 = (x->dtype == DD_BF16);
  p.bias = L.bias; p.bias_count = L.bias_count;
  p.trace = ctx->conv_trace;
  p.dbg = ctx->conv_dbg;
  if (L.residual) {
    DD_CHECK_ARG(L.ups == 1 && L.ngroups == 1, "residual needs a plain convolution");
    DD_CHECK_ARG(L.residual->dtype == x->dtype && L.residual->coff % 8 == 0 && L.residual->cstride % 8 == 0 &&
                     L.cout % 8 == 0, "residual must be aligned, of the input's 16-bit type, with a multiple of 8 channels");
    p.residual = reinterpret_cast<const __half*>(L.residual->ptr);
    p.res_cstride = L.residual->cstride; p.res_coff = L.residual->coff;
    p.res_is_mask = (L.flags & DD_CONV_RESIDUAL_MASK) ? 1 : 0;
  }

  ConvRowsMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(x->c), static_cast<cuuint64_t>(x->w), static_cast<cuuint64_t>(x->h),
                          static_cast<cuuint64_t>(x->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(x->cstride) * 2, static_cast<cuuint64_t>(x->w) * x->cstride * 2,
                             static_cast<cuuint64_t>(x->h) * x->w * x->cstride * 2};
    cuuint32_t box[4] = {64, a_box_w, 1, 1};
    void* base = reinterpret_cast<__half*>(x->ptr) + x->coff;
    int rc = encode_map(ctx, &maps.a, x->dtype == DD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, 4, dims, strides, box,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
    if (p.split) {
      const dd_tensor* xl = L.x_lo;
      cuuint64_t lstrides[3] = {static_cast<cuuint64_t>(xl->cstride) * 2, static_cast<cuuint64_t>(xl->w) * xl->cstride * 2,
                                static_cast<cuuint64_t>(xl->h) * xl->w * xl->cstride * 2};
      rc = encode_map(ctx, &maps.a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, reinterpret_cast<__half*>(xl->ptr) + xl->coff, 4, dims,
                      lstrides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
      if (rc) return rc;
    }
  }
  {
    const int r_total = k3 ? 3 : 1;
    cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(L.rows_total), static_cast<cuuint64_t>(r_total),
                          static_cast<cuuint64_t>(w_chunks * p.tiles_per_chunk)};
    cuuint64_t strides[3] = {128, static_cast<cuuint64_t>(L.rows_total) * 128,
                             static_cast<cuuint64_t>(L.rows_total) * 128 * r_total};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.cpad), static_cast<cuuint32_t>(n_r), 1};
    int rc = encode_map(ctx, &maps.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, const_cast<void*>(L.w_packed), 4, dims, strides, box,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  for (int g = 0; g < L.ngroups; ++g) {
    const int sp = L.sp0 + g;
    int rc = encode_out_map(ctx, &maps.out[g], y, 0, L.cout, x->h, x->w, L.ups, L.ups == 2 ? (sp >> 1) : 0,
                            L.ups == 2 ? (sp & 1) : 0);
    if (rc) return rc;
    if (p.split_out) {
      rc = encode_out_map(ctx, &maps.out_lo[g], L.y_lo, 0, L.cout, x->h, x->w, L.ups, L.ups == 2 ? (sp >> 1) : 0,
                          L.ups == 2 ? (sp & 1) : 0);
      if (rc) return rc;
    }
  }
  if (L.y_relu) {
    DD_CHECK_ARG(L.y_relu->dtype == x->dtype && L.y_relu->coff % 8 == 0 && L.y_relu->cstride % 8 == 0 && L.ngroups == 1,
                 "relu-copy output must be aligned and of the input's 16-bit type");
    p.has_relu_copy = 1;
    int rc = encode_out_map(ctx, &maps.out_relu, L.y_relu, 0, L.cout, x->h, x->w, L.ups, L.ups == 2 ? (L.sp0 >> 1) : 0,
                            L.ups == 2 ? (L.sp0 & 1) : 0);
    if (rc) return rc;
  }

  if (L.colsum) {
    DD_CHECK_ARG(L.ngroups == 1 && L.cout <= 128 && !p.split, "column sums: plain convolution slices of at most 128 channels");
    p.colsum = L.colsum;
  }
  p.epi_plain = (!p.out_f32 && !p.residual && !p.has_relu_copy && !p.split_out && !p.colsum) ? 1 : 0;

  int grid = ctx->sm_count;
  if (p.total_rows < grid) grid = static_cast<int>(p.total_rows);
  p.rows_per_cta = static_cast<int>((p.total_rows + grid - 1) / grid);
  grid = static_cast<int>((p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta);
  if (p.split) {
    DD_CUDA(cudaFuncSetAttribute(conv_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conv_rows_kernel<true><<<grid, kRowsThreads, smem, stream>>>(maps, p);
  } else {
    DD_CUDA(cudaFuncSetAttribute(conv_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    conv_rows_kernel<false><<<grid, kRowsThreads, smem, stream>>>(maps, p);
  }
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

static int launch_conv_simt(dd_ctx* ctx, const dd_tensor* x, const float* w, const float* bias, int ksize, int cout,
                            uint32_t flags, const dd_tensor* residual, const dd_tensor* y, const dd_tensor* y_relu,
                            int ups, int ay, int ax, cudaStream_t stream) {
  ConvSimtParams p;
  memset(&p, 0, sizeof(p));
  p.x = make_view(x); p.y = make_view(y);
  if (residual) { p.res = make_view(residual); p.has_res = 1; }
  if (y_relu) { p.yrelu = make_view(y_relu); p.has_yrelu = 1; }
  p.w = w; p.bias = bias; p.cout = cout; p.ksize = ksize;
  p.relu = (flags & DD_CONV_RELU) ? 1 : 0;
  p.ups = ups; p.ay = ay; p.ax = ax;
  p.x_ups = 1; p.oh = x->h; p.ow = x->w;
  if (ups == -2) {   // input read through sub-pixel (ay, ax) of a 2x finer grid; output on the coarse grid
    p.ups = 1; p.x_ups = 2; p.x_ay = ay; p.x_ax = ax; p.oh = x->h / 2; p.ow = x->w / 2;
  }
  const size_t total = static_cast<size_t>(x->n) * p.oh * p.ow * ((cout + kSimtCob - 1) / kSimtCob);
  const unsigned blocks = static_cast<unsigned>((total + 127) / 128);
  conv_simt_kernel<<<blocks, 128, 0, stream>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_abi_version(void) { return DD_B200_ABI_VERSION; }
const char* dd_last_error(void) { return g_err; }

int dd_ctx_create(int device, dd_ctx** out) {
  DD_CHECK_ARG(out, "out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("no CUDA device (%s) - libdd_b200 has no CPU fallback", cudaGetErrorString(e));
    return DD_ERR_NO_DEVICE;
  }
  DD_CHECK_ARG(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
  DD_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libdd_b200 is built for sm_100a only", device, prop.major, prop.minor);
    return DD_ERR_UNSUPPORTED;
  }
  dd_ctx* ctx = new dd_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = prop.sharedMemPerBlockOptin;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess) {
    ctx->encode_tiled = fn;
  }
  *out = ctx;
  return DD_OK;
}

int dd_ctx_destroy(dd_ctx* ctx) {
  delete ctx;
  return DD_OK;
}

int dd_ctx_sm_count(const dd_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int64_t dd_ctx_launch_count(const dd_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int dd_ctx_set_trace_buffer(dd_ctx* ctx, void* device_buffer) {
  DD_CHECK_ARG(ctx, "NULL ctx");
  ctx->conv_trace = reinterpret_cast<unsigned long long*>(device_buffer);
  return DD_OK;
}

int dd_ctx_set_option(dd_ctx* ctx, const char* name, int value) {
  DD_CHECK_ARG(ctx && name, "NULL argument");
  if (!strcmp(name, "conv_rows")) { ctx->conv_rows = value; return DD_OK; }               // cap on G (rows per weight pass)
  if (!strcmp(name, "conv_b_stages")) { ctx->conv_b_stages = value; return DD_OK; }       // weight stages when streaming (0 = auto)
  if (!strcmp(name, "std_generic")) { ctx->std_generic = value; return DD_OK; }           // generic standardisation tile kernel (A/B)
  if (!strcmp(name, "conv_dbg")) { ctx->conv_dbg = value; return DD_OK; }                 // ablation switches (wrong results)
  if (!strcmp(name, "conv_force_stream")) { ctx->conv_force_stream = value; return DD_OK; } // never keep weights resident
  set_error("unknown option '%s'", name);
  return DD_ERR_INVALID;
}

size_t dd_conv2d_packed_bytes(int ksize, int cin, int cout, int dtype, int transposed) {
  const int k2 = ksize * ksize;
  if (is_half_type(dtype) || dtype == DD_F16X2) {
    const int chunks = (round_up(cin, 16) + 63) / 64 * (dtype == DD_F16X2 ? 2 : 1);
    const int rows = transposed ? k2 * packed_cpad(cout) : packed_cpad(cout);
    const int tiles = transposed ? 1 : k2;
    return static_cast<size_t>(chunks) * tiles * rows * 64 * 2;
  }
  return static_cast<size_t>(k2) * cout * cin * 4;
}

int dd_conv2d_pack_weights(dd_ctx* ctx, const float* w, int ksize, int cin, int cout, int dtype, int transposed,
                           void* packed_dev, void* stream) {
  DD_CHECK_ARG(ctx && w && packed_dev, "NULL argument");
  DD_CHECK_ARG(ksize >= 1 && ksize <= 3 && cin > 0 && cout > 0, "bad conv shape");
  DD_CHECK_ARG(!transposed || ksize == 2, "only 2x2 transposed convolutions are packed here");
  DD_CHECK_ARG(transposed || ksize != 2, "2x2 kernels exist only as transposed convolutions");
  const size_t bytes = dd_conv2d_packed_bytes(ksize, cin, cout, dtype, transposed);
  const int k2 = ksize * ksize;
  std::vector<uint8_t> host(bytes, 0);
  if (is_half_type(dtype) || dtype == DD_F16X2) {
    const int cpad = packed_cpad(cout);
    uint16_t* dst = reinterpret_cast<uint16_t*>(host.data());
    auto cvt = [dtype](float v) -> uint16_t {
      if (dtype == DD_BF16) { const __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<const uint16_t*>(&b); }
      const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h);
    };
    // DD_F16X2: W = W_hi + W_lo, W_hi = fp16(W), W_lo = fp16(W - W_hi); the chunks of W_lo follow those of W_hi
    const int nb = (round_up(cin, 16) + 63) / 64;
    auto lo_of = [](float v) -> float { return v - __half2float(__float2half_rn(v)); };
    if (!transposed) {
      // TF [kh,kw,cin,cout] -> [chunk][s][r][cpad][64]   (tap (r,s): r = row offset, s = column offset)
      for (int r = 0; r < ksize; ++r)
        for (int s = 0; s < ksize; ++s)
          for (int c = 0; c < cin; ++c)
            for (int o = 0; o < cout; ++o) {
              const float v = w[(static_cast<size_t>(r * ksize + s) * cin + c) * cout + o];
              const size_t tile = static_cast<size_t>(c / 64) * ksize + s;
              dst[((tile * ksize + r) * cpad + o) * 64 + (c % 64)] = cvt(v);
              if (dtype == DD_F16X2) {
                const size_t tile_lo = static_cast<size_t>(nb + c / 64) * ksize + s;
                dst[((tile_lo * ksize + r) * cpad + o) * 64 + (c % 64)] = cvt(lo_of(v));
              }
            }
    } else {
      // TF transpose layout [kh,kw,cout,cin] -> [chunk][sub-pixel][cpad][64]: one 1x1 GEMM, rows (sub-pixel, cout)
      for (int sp = 0; sp < k2; ++sp)
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c) {
            const float v = w[(static_cast<size_t>(sp) * cout + o) * cin + c];
            dst[((static_cast<size_t>(c / 64) * k2 + sp) * cpad + o) * 64 + (c % 64)] = cvt(v);
            if (dtype == DD_F16X2)
              dst[((static_cast<size_t>(nb + c / 64) * k2 + sp) * cpad + o) * 64 + (c % 64)] = cvt(lo_of(v));
          }
    }
  } else {
    float* dst = reinterpret_cast<float*>(host.data());
    if (!transposed) {
      for (int t = 0; t < k2; ++t)
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c)
            dst[(static_cast<size_t>(t) * cout + o) * cin + c] = w[(static_cast<size_t>(t) * cin + c) * cout + o];
    } else {
      memcpy(dst, w, bytes);  // [sp][cout][cin] already
    }
  }
  DD_CUDA(cudaMemcpyAsync(packed_dev, host.data(), bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  DD_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));  // staging buffer dies with this call
  return DD_OK;
}

static int conv2d_fwd_impl(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* x_lo, const void* w_packed, const float* bias,
                           int ksize, uint32_t flags, const dd_tensor* residual, const dd_tensor* y, const dd_tensor* y_lo,
                           const dd_tensor* y_relu, void* stream, float* colsum = nullptr) {
  DD_CHECK_ARG(ctx && w_packed, "NULL argument");
  DD_CHECK_ARG(tensor_ok(x) && tensor_ok(y), "bad tensor descriptor");
  DD_CHECK_ARG(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
  DD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w, "conv2d_fwd: spatial dims differ");
  DD_CHECK_ARG(!residual || (tensor_ok(residual) && residual->c == y->c), "bad residual");
  DD_CHECK_ARG(!y_relu || (tensor_ok(y_relu) && y_relu->c == y->c), "bad y_relu");
  DD_CHECK_ARG(((flags & DD_CONV_RELU_COPY) != 0) == (y_relu != nullptr), "DD_CONV_RELU_COPY needs y_relu");
  DD_CHECK_ARG(!(flags & DD_CONV_RESIDUAL_MASK) || (residual && is_half_type(x->dtype)), "DD_CONV_RESIDUAL_MASK: 16-bit path with a mask tensor");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x->dtype == DD_F32) {
    DD_CHECK_ARG(y->dtype == DD_F32, "exact path writes fp32");
    DD_CHECK_ARG(!colsum, "column sums are fused on the 16-bit tensor-core path only");
    return launch_conv_simt(ctx, x, reinterpret_cast<const float*>(w_packed), bias, ksize, y->c, flags, residual, y,
                            y_relu, 1, 0, 0, s);
  }
  // output channels are processed in slices: at most 256 (one UMMA N) for 1x1, at most 128 for 3x3 (the TMEM ring must hold
  // the three output rows an input row feeds plus one being drained: 4 blocks of <= 128 columns); each slice is a launch
  const int cpad_total = packed_cpad(y->c);
  const int max_slice = (ksize == 3) ? 128 : 256;
  const int n_slices = (cpad_total + max_slice - 1) / max_slice;
  const int slice = round_up((cpad_total + n_slices - 1) / n_slices, 32);
  for (int row0 = 0; row0 < cpad_total; row0 += slice) {
    const int cpad = (cpad_total - row0 < slice) ? (cpad_total - row0) : slice;
    const int cout = (y->c - row0 < cpad) ? (y->c - row0) : cpad;
    dd_tensor ys = *y, yr, rs, yl;
    ys.coff += row0; ys.c = cout;
    if (y_lo) { yl = *y_lo; yl.coff += row0; yl.c = cout; }
    if (y_relu) { yr = *y_relu; yr.coff += row0; yr.c = cout; }
    if (residual) { rs = *residual; rs.coff += row0; rs.c = cout; }
    TcLaunch L;
    memset(&L, 0, sizeof(L));
    L.x = x; L.w_packed = w_packed; L.ksize = ksize; L.rows_total = cpad_total; L.row0 = row0; L.cpad = cpad;
    L.ngroups = 1; L.group_c = cpad; L.cout = cout;
    L.bias = bias ? bias + row0 : nullptr;
    L.bias_count = round_up(y->c, 16) - row0 < 256 ? round_up(y->c, 16) - row0 : 256;
    L.flags = flags; L.residual = residual ? &rs : nullptr; L.y = &ys; L.y_relu = y_relu ? &yr : nullptr;
    L.ups = 1; L.sp0 = 0; L.s_mask = 0; L.rm_lo = 0; L.rm_hi = 2;
    L.x_lo = x_lo; L.y_lo = y_lo ? &yl : nullptr;
    L.colsum = colsum ? colsum + row0 : nullptr;
    int rc = launch_conv_rows(ctx, L, s);
    if (rc) return rc;
  }
  return DD_OK;
}

int dd_conv2d_fwd(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias, int ksize, uint32_t flags,
                  const dd_tensor* residual, const dd_tensor* y, const dd_tensor* y_relu, void* stream) {
  return conv2d_fwd_impl(ctx, x, nullptr, w_packed, bias, ksize, flags, residual, y, nullptr, y_relu, stream);
}

int dd_conv2d_fwd_colsum(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias, int ksize, uint32_t flags,
                         const dd_tensor* residual, const dd_tensor* y, float* colsum_dev, void* stream) {
  DD_CHECK_ARG(colsum_dev, "colsum_dev is NULL");
  return conv2d_fwd_impl(ctx, x, nullptr, w_packed, bias, ksize, flags, residual, y, nullptr, nullptr, stream, colsum_dev);
}

int dd_conv2d_fwd_split(dd_ctx* ctx, const dd_tensor* x_hi, const dd_tensor* x_lo, const void* w_packed, const float* bias,
                        int ksize, uint32_t flags, const dd_tensor* y_hi, const dd_tensor* y_lo, void* stream) {
  DD_CHECK_ARG(tensor_ok(x_hi) && tensor_ok(x_lo) && x_hi->dtype == DD_F16, "split conv: fp16 (hi, lo) input pair expected");
  DD_CHECK_ARG(tensor_ok(y_hi) && ((y_hi->dtype == DD_F32 && !y_lo) || (y_hi->dtype == DD_F16 && tensor_ok(y_lo))),
               "split conv: output is an fp16 (hi, lo) pair or one fp32 tensor");
  return conv2d_fwd_impl(ctx, x_hi, x_lo, w_packed, bias, ksize, flags, nullptr, y_hi, y_lo, nullptr, stream);
}

static int transpose2x2_impl(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* x_lo, const void* w_packed, const float* bias,
                             uint32_t flags, const dd_tensor* y, const dd_tensor* y_lo, void* stream) {
  DD_CHECK_ARG(ctx && w_packed, "NULL argument");
  DD_CHECK_ARG(tensor_ok(x) && tensor_ok(y), "bad tensor descriptor");
  DD_CHECK_ARG(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w, "transpose2x2: output must be 2x input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cout = y->c;
  if (x->dtype == DD_F32) {
    DD_CHECK_ARG(y->dtype == DD_F32, "exact path writes fp32");
    for (int sp = 0; sp < 4; ++sp) {
      const float* w = reinterpret_cast<const float*>(w_packed) + static_cast<size_t>(sp) * cout * x->c;
      int rc = launch_conv_simt(ctx, x, w, bias, 1, cout, flags, nullptr, y, nullptr, 2, sp >> 1, sp & 1, s);
      if (rc) return rc;
    }
    return DD_OK;
  }
  const int cpad1 = packed_cpad(cout);
  DD_CHECK_ARG(cpad1 <= 256, "cout %d > 256", cout);
  // sub-pixels per launch: as many column groups as fit in one UMMA (N <= 256)
  int per = 4;
  while (per * cpad1 > 256) per >>= 1;
  for (int sp0 = 0; sp0 < 4; sp0 += per) {
    TcLaunch L;
    memset(&L, 0, sizeof(L));
    L.x = x; L.w_packed = w_packed; L.ksize = 1; L.rows_total = 4 * cpad1; L.row0 = sp0 * cpad1; L.cpad = per * cpad1;
    L.ngroups = per; L.group_c = cpad1; L.cout = cout;
    L.bias = bias; L.bias_count = round_up(cout, 16);
    L.flags = flags; L.residual = nullptr; L.y = y; L.y_relu = nullptr;
    L.ups = 2; L.sp0 = sp0;
    L.x_lo = x_lo; L.y_lo = y_lo;
    int rc = launch_conv_rows(ctx, L, s);
    if (rc) return rc;
  }
  return DD_OK;
}

int dd_conv2d_transpose2x2_fwd(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias,
                               uint32_t flags, const dd_tensor* y, void* stream) {
  return transpose2x2_impl(ctx, x, nullptr, w_packed, bias, flags, y, nullptr, stream);
}

int dd_conv2d_transpose2x2_fwd_split(dd_ctx* ctx, const dd_tensor* x_hi, const dd_tensor* x_lo, const void* w_packed,
                                     const float* bias, uint32_t flags, const dd_tensor* y_hi, const dd_tensor* y_lo, void* stream) {
  DD_CHECK_ARG(tensor_ok(x_hi) && tensor_ok(x_lo) && tensor_ok(y_hi) && tensor_ok(y_lo) && x_hi->dtype == DD_F16 &&
                   y_hi->dtype == DD_F16, "split transposed conv: fp16 (hi, lo) pairs expected");
  return transpose2x2_impl(ctx, x_hi, x_lo, w_packed, bias, flags, y_hi, y_lo, stream);
}

int dd_conv2d_transpose3x3_fwd(dd_ctx* ctx, const dd_tensor* x, const void* const* w_phase, const float* bias,
                               uint32_t flags, const dd_tensor* y, const dd_tensor* y_relu, void* stream) {
  DD_CHECK_ARG(ctx && w_phase && w_phase[0] && w_phase[1] && w_phase[2] && w_phase[3], "NULL argument");
  DD_CHECK_ARG(tensor_ok(x) && tensor_ok(y), "bad tensor descriptor");
  DD_CHECK_ARG(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w, "transpose3x3: output must be 2x input");
  DD_CHECK_ARG(!y_relu || (tensor_ok(y_relu) && y_relu->c == y->c && y_relu->h == y->h && y_relu->w == y->w),
               "bad y_relu");
  DD_CHECK_ARG(((flags & DD_CONV_RELU_COPY) != 0) == (y_relu != nullptr), "DD_CONV_RELU_COPY needs y_relu");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cout = y->c;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    if (x->dtype == DD_F32) {
      DD_CHECK_ARG(y->dtype == DD_F32, "exact path writes fp32");
      int rc = launch_conv_simt(ctx, x, reinterpret_cast<const float*>(w_phase[ph]), bias, 3, cout, flags, nullptr, y,
                                y_relu, 2, py, px, s);
      if (rc) return rc;
      continue;
    }
    // taps of the phase kernel sit at (r, s) = (dy+1, dx+1), dy,dx in {0,-1}; W index py-2dy (px-2dx) must be <= 2:
    // phase 0 uses offsets {0,-1}, phase 1 only offset 0
    const int cpad = packed_cpad(cout);
    DD_CHECK_ARG(cpad <= 256, "cout %d > 256", cout);
    TcLaunch L;
    memset(&L, 0, sizeof(L));
    L.x = x; L.w_packed = w_phase[ph]; L.ksize = 3; L.rows_total = cpad; L.row0 = 0; L.cpad = cpad;
    L.ngroups = 1; L.group_c = cpad; L.cout = cout;
    L.bias = bias; L.bias_count = round_up(cout, 16);
    L.flags = flags; L.residual = nullptr; L.y = y; L.y_relu = y_relu;
    L.ups = 2; L.sp0 = ph;
    L.s_mask = (px == 0) ? 3 : 2;                 // s = dx + 1: {0,1} or {1}
    L.rm_lo = (py == 0) ? 0 : 1; L.rm_hi = 1;     // r = dy + 1
    int rc = launch_conv_rows(ctx, L, s);
    if (rc) return rc;
  }
  return DD_OK;
}


/* dx = conv2d_transpose_2x2_s2^T (dz): dx[i,j,c] = sum_{a,b,o} dz[2i+a,2j+b,o] W[a,b,o,c] - exact fp32 path.
 * w_dgrad: [sub-pixel][cin][cout] fp32 (dd_conv2d_repack_f32 with transposed = 1). */
int dd_conv2d_transpose2x2_dgrad(dd_ctx* ctx, const dd_tensor* dz, const float* w_dgrad, const dd_tensor* dx, void* stream) {
  DD_CHECK_ARG(ctx && w_dgrad && tensor_ok(dz) && tensor_ok(dx), "bad argument");
  DD_CHECK_ARG(dz->dtype == DD_F32 && dx->dtype == DD_F32, "exact path only");
  DD_CHECK_ARG(dz->n == dx->n && dz->h == 2 * dx->h && dz->w == 2 * dx->w, "transpose2x2_dgrad: dz must be 2x dx");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int sp = 0; sp < 4; ++sp) {
    const float* w = w_dgrad + static_cast<size_t>(sp) * dx->c * dz->c;
    int rc = launch_conv_simt(ctx, dz, w, nullptr, 1, dx->c, 0, sp == 0 ? nullptr : dx, dx, nullptr, -2, sp >> 1, sp & 1, s);
    if (rc) return rc;
  }
  return DD_OK;
}

/* dx = conv2d_transpose_3x3_s2^T (dz) (Tiramisu.py:62-64): dx[i,j,c] = sum_{r,s,o} dz[2i+r,2j+s,o] W[r,s,o,c], taps that
 * leave the image contribute nothing (TF 'SAME': the transposed conv keeps rows/cols [0, 2n) of the 2n+1 it produces).
 * w_dgrad: [tap][cin][cout] fp32 (dd_conv2d_repack_f32 with transposed = 1, ksize = 3). */
int dd_conv2d_transpose3x3_dgrad(dd_ctx* ctx, const dd_tensor* dz, const float* w_dgrad, const dd_tensor* dx, void* stream) {
  DD_CHECK_ARG(ctx && w_dgrad && tensor_ok(dz) && tensor_ok(dx), "bad argument");
  DD_CHECK_ARG(dz->dtype == DD_F32 && dx->dtype == DD_F32, "exact path only");
  DD_CHECK_ARG(dz->n == dx->n && dz->h == 2 * dx->h && dz->w == 2 * dx->w, "transpose3x3_dgrad: dz must be 2x dx");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int tap = 0; tap < 9; ++tap) {
    const float* w = w_dgrad + static_cast<size_t>(tap) * dx->c * dz->c;
    int rc = launch_conv_simt(ctx, dz, w, nullptr, 1, dx->c, 0, tap == 0 ? nullptr : dx, dx, nullptr, -2, tap / 3, tap % 3, s);
    if (rc) return rc;
  }
  return DD_OK;
}

}  // extern "C"
