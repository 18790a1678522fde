// C-ABI: context, weight packing and the convolution entry points (tcgen05 fp16 path + exact fp32 SIMT path).
#include <string.h>

#include <vector>

#include "conv_tc.cuh"
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
};

constexpr int kSimtCob = 8;

__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvSimtParams p) {
  const int cgroups = (p.cout + kSimtCob - 1) / kSimtCob;
  const size_t total = static_cast<size_t>(p.x.n) * p.x.h * p.x.w * cgroups;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // consecutive threads -> consecutive pixels (same channel group => weight loads are warp-uniform)
  const size_t npix = static_cast<size_t>(p.x.n) * p.x.h * p.x.w;
  const int cg = static_cast<int>(idx / npix);
  const size_t pixel = idx % npix;
  const int x0 = static_cast<int>(pixel % p.x.w);
  const int y0 = static_cast<int>((pixel / p.x.w) % p.x.h);
  const int n0 = static_cast<int>(pixel / (static_cast<size_t>(p.x.w) * p.x.h));
  const int co0 = cg * kSimtCob;
  const int cin = p.x.c;
  const int pad = (p.ksize - 1) / 2;

  float acc[kSimtCob];
#pragma unroll
  for (int i = 0; i < kSimtCob; ++i) acc[i] = 0.f;

  for (int r = 0; r < p.ksize; ++r) {
    const int yy = y0 + r - pad;
    if (yy < 0 || yy >= p.x.h) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int xx = x0 + s - pad;
      if (xx < 0 || xx >= p.x.w) continue;
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

static int encode_map(dd_ctx* ctx, CUtensorMap* map, void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapL2promotion promo) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), base, dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu box %u %u %u)",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)dims[2], box[0], box[1], box[2]);
    return DD_ERR_CUDA;
  }
  return DD_OK;
}

struct TcLaunch {
  const dd_tensor* x;
  const void* w_packed;     // [taps][rows_total][cin64] fp16
  int rows_total;           // rows in the packed weight matrix
  int row0;                 // first row used by this launch
  int n_umma;               // rows used
  int ngroups, group_c, cout_store;
  int ksize;
  const float* bias;
  uint32_t flags;
  const dd_tensor* residual;
  const dd_tensor* y;
  const dd_tensor* y_relu;
  int ups, sp0;
  uint32_t tap_mask;        // 0 = all taps
};

static int launch_conv_tc(dd_ctx* ctx, const TcLaunch& L, cudaStream_t stream) {
  const dd_tensor* x = L.x;
  const dd_tensor* y = L.y;
  DD_CHECK_ARG(ctx->encode_tiled, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  DD_CHECK_ARG(x->dtype == DD_F16, "tensor-core conv needs fp16 input");
  DD_CHECK_ARG(x->coff % 8 == 0 && x->cstride % 8 == 0, "conv input view must be 16-byte aligned (coff %d cstride %d)",
               x->coff, x->cstride);
  DD_CHECK_ARG(L.n_umma % 16 == 0 && L.n_umma >= 16 && L.n_umma <= 256, "UMMA N %d unsupported", L.n_umma);
  const int align_out = (y->dtype == DD_F16) ? 8 : 4;
  DD_CHECK_ARG(y->coff % align_out == 0 && y->cstride % align_out == 0, "conv output view misaligned");
  DD_CHECK_ARG(y->coff + L.cout_store <= y->cstride, "conv output buffer too narrow for padded store (%d+%d>%d)",
               y->coff, L.cout_store, y->cstride);

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = x->n; p.H = x->h; p.W = x->w;
  p.Cin = round_up(x->c, 16);
  p.n_umma = L.n_umma;
  p.acc_stride = round_up(L.n_umma, 32);
  p.taps = L.ksize * L.ksize;
  p.tap_mask = L.tap_mask ? L.tap_mask : (p.taps == 9 ? 0x1FFu : 1u);
  p.n_chunks = (p.Cin + kConvCH - 1) / kConvCH;
  p.shift_mode = (p.taps == 9) ? ctx->conv_shift_mode : 0;

  // rows per tile: as many as TMEM (2 sets) and shared memory allow
  const size_t smem_cap = ctx->max_smem_optin - 1024 /*align*/ - 512 /*barriers*/;
  const uint32_t b_stage = static_cast<uint32_t>(L.n_umma) * 128u;
  int R = 0, a_stages = 2, b_stages = 0;
  uint32_t a_stage = 0;
  const int cand[3] = {4, 2, 1};
  for (int ci = 0; ci < 3; ++ci) {
    int r = cand[ci];
    if (ctx->conv_rows > 0) r = ctx->conv_rows;
    if (2 * r * p.acc_stride > 512) { if (ctx->conv_rows > 0) break; continue; }
    const int box_w = (p.taps == 9 && p.shift_mode != 2) ? kConvTileW + 2 : kConvTileW;
    const int rows = (p.taps == 9) ? r + 2 : r;
    const uint32_t as = static_cast<uint32_t>(round_up(rows * box_w * 128, 1024));
    const size_t left = (smem_cap > 2ull * as) ? smem_cap - 2ull * as : 0;
    int bs = static_cast<int>(left / b_stage);
    if (bs > 8) bs = 8;
    if (bs >= 2) { R = r; a_stage = as; b_stages = bs; break; }
    if (ctx->conv_rows > 0) break;
  }
  DD_CHECK_ARG(R > 0, "no tile configuration fits (n_umma %d taps %d)", L.n_umma, p.taps);
  if (ctx->conv_b_stages > 0 && ctx->conv_b_stages < b_stages) b_stages = ctx->conv_b_stages;
  p.R = R;
  p.a_box_w = (p.taps == 9 && p.shift_mode != 2) ? kConvTileW + 2 : kConvTileW;
  const int a_rows = (p.taps == 9) ? R + 2 : R;
  p.a_stages = a_stages; p.b_stages = b_stages;
  p.a_stage_bytes = a_stage; p.b_stage_bytes = b_stage;
  p.a_tx_bytes = static_cast<uint32_t>(a_rows * p.a_box_w * 128);
  p.b_tx_bytes = b_stage;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * R * p.acc_stride)) cols <<= 1;
  p.tmem_cols = cols;
  p.strips = (p.W + kConvTileW - 1) / kConvTileW;
  p.bands = (p.H + R - 1) / R;
  p.num_tiles = p.strips * p.bands * p.N;

  p.ngroups = L.ngroups; p.group_c = L.group_c; p.cout_store = L.cout_store;
  p.ups = L.ups; p.sp0 = L.sp0;
  p.OH = y->h; p.OW = y->w;
  p.bias = L.bias;
  p.relu = (L.flags & DD_CONV_RELU) ? 1 : 0;
  p.out_f32 = (y->dtype == DD_F32);
  p.out = y->ptr; p.out_cstride = y->cstride; p.out_coff = y->coff;
  if (L.y_relu) {
    DD_CHECK_ARG(L.y_relu->dtype == DD_F16 && L.y_relu->coff % 8 == 0 && L.y_relu->cstride % 8 == 0,
                 "relu-copy output must be aligned fp16");
    p.out_relu = reinterpret_cast<__half*>(L.y_relu->ptr);
    p.out_relu_cstride = L.y_relu->cstride; p.out_relu_coff = L.y_relu->coff;
  }
  if (L.residual) {
    DD_CHECK_ARG(L.residual->dtype == DD_F16 && L.residual->coff % 8 == 0 && L.residual->cstride % 8 == 0,
                 "residual must be aligned fp16");
    p.residual = reinterpret_cast<const __half*>(L.residual->ptr);
    p.res_cstride = L.residual->cstride; p.res_coff = L.residual->coff;
  }

  // tensor maps
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(x->c), static_cast<cuuint64_t>(x->w),
                          static_cast<cuuint64_t>(x->h), static_cast<cuuint64_t>(x->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(x->cstride) * 2,
                             static_cast<cuuint64_t>(x->w) * x->cstride * 2,
                             static_cast<cuuint64_t>(x->h) * x->w * x->cstride * 2};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(kConvCH), static_cast<cuuint32_t>(p.a_box_w),
                         static_cast<cuuint32_t>(a_rows), 1};
    void* base = reinterpret_cast<__half*>(x->ptr) + x->coff;
    int rc = encode_map(ctx, &tmA, base, 4, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
  }
  {
    const int cin64 = round_up(x->c, 64);
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cin64), static_cast<cuuint64_t>(L.rows_total - L.row0),
                          static_cast<cuuint64_t>(p.taps)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(cin64) * 2, static_cast<cuuint64_t>(L.rows_total) * cin64 * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(kConvCH), static_cast<cuuint32_t>(L.n_umma), 1};
    void* base = reinterpret_cast<__half*>(const_cast<void*>(L.w_packed)) + static_cast<size_t>(L.row0) * cin64;
    int rc = encode_map(ctx, &tmB, base, 3, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }

  const size_t smem = 1024 + static_cast<size_t>(a_stages) * a_stage + static_cast<size_t>(b_stages) * b_stage + 512;
  DD_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
  conv_tc_kernel<<<grid, kConvThreads, smem, stream>>>(tmA, tmB, p);
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
  const size_t total = static_cast<size_t>(x->n) * x->h * x->w * ((cout + kSimtCob - 1) / kSimtCob);
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

int dd_ctx_set_option(dd_ctx* ctx, const char* name, int value) {
  DD_CHECK_ARG(ctx && name, "NULL argument");
  if (!strcmp(name, "conv_shift_mode")) { DD_CHECK_ARG(value >= 0 && value <= 2, "bad shift mode"); ctx->conv_shift_mode = value; return DD_OK; }
  if (!strcmp(name, "conv_rows")) { ctx->conv_rows = value; return DD_OK; }
  if (!strcmp(name, "conv_b_stages")) { ctx->conv_b_stages = value; return DD_OK; }
  set_error("unknown option '%s'", name);
  return DD_ERR_INVALID;
}

size_t dd_conv2d_packed_bytes(int ksize, int cin, int cout, int dtype, int transposed) {
  const int taps = transposed ? 1 : ksize * ksize;
  const int groups = transposed ? ksize * ksize : 1;
  if (dtype == DD_F16) {
    const int rows = transposed ? groups * round_up(cout, 16) : round_up(cout, 16);
    return static_cast<size_t>(taps) * rows * round_up(cin, 64) * 2;
  }
  return static_cast<size_t>(taps) * groups * cout * cin * 4;
}

int dd_conv2d_pack_weights(dd_ctx* ctx, const float* w, int ksize, int cin, int cout, int dtype, int transposed,
                           void* packed_dev, void* stream) {
  DD_CHECK_ARG(ctx && w && packed_dev, "NULL argument");
  DD_CHECK_ARG(ksize >= 1 && ksize <= 3 && cin > 0 && cout > 0, "bad conv shape");
  DD_CHECK_ARG(!transposed || ksize == 2, "only 2x2 transposed convolutions are packed here");
  const size_t bytes = dd_conv2d_packed_bytes(ksize, cin, cout, dtype, transposed);
  const int k2 = ksize * ksize;
  std::vector<uint8_t> host(bytes, 0);
  if (dtype == DD_F16) {
    const int cin64 = round_up(cin, 64);
    const int c16 = round_up(cout, 16);
    __half* dst = reinterpret_cast<__half*>(host.data());
    if (!transposed) {
      // TF [kh,kw,cin,cout] -> [tap][cout16][cin64]
      for (int t = 0; t < k2; ++t)
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c)
            dst[(static_cast<size_t>(t) * c16 + o) * cin64 + c] = __float2half_rn(w[(static_cast<size_t>(t) * cin + c) * cout + o]);
    } else {
      // TF transpose layout [kh,kw,cout,cin] -> one 1x1 GEMM with rows (sub-pixel, cout16)
      for (int sp = 0; sp < k2; ++sp)
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c)
            dst[(static_cast<size_t>(sp) * c16 + o) * cin64 + c] = __float2half_rn(w[(static_cast<size_t>(sp) * cout + o) * cin + c]);
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

int dd_conv2d_fwd(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias, int ksize, uint32_t flags,
                  const dd_tensor* residual, const dd_tensor* y, const dd_tensor* y_relu, void* stream) {
  DD_CHECK_ARG(ctx && w_packed, "NULL argument");
  DD_CHECK_ARG(tensor_ok(x) && tensor_ok(y), "bad tensor descriptor");
  DD_CHECK_ARG(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
  DD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w, "conv2d_fwd: spatial dims differ");
  DD_CHECK_ARG(!residual || (tensor_ok(residual) && residual->c == y->c), "bad residual");
  DD_CHECK_ARG(!y_relu || (tensor_ok(y_relu) && y_relu->c == y->c), "bad y_relu");
  DD_CHECK_ARG(((flags & DD_CONV_RELU_COPY) != 0) == (y_relu != nullptr), "DD_CONV_RELU_COPY needs y_relu");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x->dtype == DD_F32) {
    DD_CHECK_ARG(y->dtype == DD_F32, "exact path writes fp32");
    return launch_conv_simt(ctx, x, reinterpret_cast<const float*>(w_packed), bias, ksize, y->c, flags, residual, y,
                            y_relu, 1, 0, 0, s);
  }
  const int c16 = round_up(y->c, 16);
  DD_CHECK_ARG(c16 <= 256, "cout %d > 256: split the layer", y->c);
  TcLaunch L;
  memset(&L, 0, sizeof(L));
  L.x = x; L.w_packed = w_packed; L.rows_total = c16; L.row0 = 0; L.n_umma = c16;
  L.ngroups = 1; L.group_c = c16; L.cout_store = round_up(y->c, 8);
  L.ksize = ksize; L.bias = bias; L.flags = flags; L.residual = residual; L.y = y; L.y_relu = y_relu;
  L.ups = 1; L.sp0 = 0;
  return launch_conv_tc(ctx, L, s);
}

int dd_conv2d_transpose2x2_fwd(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias,
                               uint32_t flags, const dd_tensor* y, void* stream) {
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
  const int c16 = round_up(cout, 16);
  DD_CHECK_ARG(c16 == round_up(cout, 8), "transpose2x2 tensor path needs cout %% 16 in {0, 9..15}");
  // sub-pixels per launch: as many column groups as fit in one UMMA (N <= 256)
  int per = 4;
  while (per * c16 > 256) per >>= 1;
  DD_CHECK_ARG(per >= 1, "cout too large for transpose2x2");
  for (int sp0 = 0; sp0 < 4; sp0 += per) {
    TcLaunch L;
    memset(&L, 0, sizeof(L));
    L.x = x; L.w_packed = w_packed; L.rows_total = 4 * c16; L.row0 = sp0 * c16; L.n_umma = per * c16;
    L.ngroups = per; L.group_c = c16; L.cout_store = round_up(cout, 8);
    L.ksize = 1; L.bias = bias; L.flags = flags; L.residual = nullptr; L.y = y; L.y_relu = nullptr;
    L.ups = 2; L.sp0 = sp0;
    int rc = launch_conv_tc(ctx, L, s);
    if (rc) return rc;
  }
  return DD_OK;
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
    // taps of the phase kernel sit at slab offsets (dy+1, dx+1), dy,dx in {0,-1}; W index py-2dy must be <= 2
    uint32_t mask = 0;
    for (int dy = 0; dy >= -1; --dy)
      for (int dx = 0; dx >= -1; --dx)
        if (py - 2 * dy <= 2 && px - 2 * dx <= 2) mask |= 1u << ((dy + 1) * 3 + (dx + 1));
    const int c16 = round_up(cout, 16);
    DD_CHECK_ARG(c16 <= 256, "cout %d > 256", cout);
    TcLaunch L;
    memset(&L, 0, sizeof(L));
    L.x = x; L.w_packed = w_phase[ph]; L.rows_total = c16; L.row0 = 0; L.n_umma = c16;
    L.ngroups = 1; L.group_c = c16; L.cout_store = round_up(cout, 8);
    L.ksize = 3; L.bias = bias; L.flags = flags; L.residual = nullptr; L.y = y; L.y_relu = y_relu;
    L.ups = 2; L.sp0 = ph; L.tap_mask = mask;
    DD_CHECK_ARG(ctx->conv_shift_mode == 0, "tap subsets need conv_shift_mode 0");
    int rc = launch_conv_tc(ctx, L, s);
    if (rc) return rc;
  }
  return DD_OK;
}

}  // extern "C"
