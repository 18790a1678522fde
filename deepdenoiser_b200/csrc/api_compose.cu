// C-ABI: MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93) as ONE tcgen05 kernel (compose_rows.cuh).
#include <string.h>

#include "compose_rows.cuh"

using namespace dd;

extern "C" {

size_t dd_compose_weights_bytes(void) { return kCrWBytes; }
size_t dd_compose_params_floats(void) { return kCrFloats; }

/* Host-side packing of the four 3x3 24->24 layers (TF kernels [3,3,24,24] = [kh,kw,cin,cout], biases [24]) into the operand
 * layout of compose_rows_kernel: per layer three tiles [kw = s][8-channel chunk j][row n = kh * 32 + cout][8 cin] 16-bit, zero
 * padded; the centre tile (s = 1) has a 4th chunk whose K indices 24 / 25 hold the bias (hi / lo halves) in the rows of the
 * centre tap (kh = 1); then the 24x24 identity of the residual connections and 96 rows of zeros. */
int dd_compose_pack_weights(const float* const* conv_w, const float* const* conv_b, int dtype, void* blob_host) {
  DD_CHECK_ARG(conv_w && conv_b && blob_host && (dtype == DD_F16 || dtype == DD_BF16), "bad argument");
  uint16_t* dst = reinterpret_cast<uint16_t*>(blob_host);
  memset(dst, 0, kCrWBytes);
  auto cvt = [dtype](float v) -> uint16_t {
    if (dtype == DD_BF16) { const __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<const uint16_t*>(&b); }
    const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h);
  };
  auto back = [dtype](uint16_t bits) -> float {
    if (dtype == DD_BF16) return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&bits));
    return __half2float(*reinterpret_cast<const __half*>(&bits));
  };
  for (int l = 0; l < 4; ++l) {
    for (int s = 0; s < 3; ++s) {
      const size_t tile = (static_cast<size_t>(l) * kCrWLayer + s * kCrWTile + (s == 2 ? kCrWChunk : 0)) / 2;   // in 16-bit elements
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < kCrC; ++c)
          for (int o = 0; o < kCrC; ++o)
            dst[tile + (static_cast<size_t>(c / 8) * 96 + r * 32 + o) * 8 + (c % 8)] = cvt(conv_w[l][((r * 3 + s) * kCrC + c) * kCrC + o]);
      if (s == 1)
        for (int o = 0; o < kCrC; ++o) {
          const uint16_t hi = cvt(conv_b[l][o]);
          dst[tile + (static_cast<size_t>(3) * 96 + 32 + o) * 8 + 0] = hi;
          dst[tile + (static_cast<size_t>(3) * 96 + 32 + o) * 8 + 1] = cvt(conv_b[l][o] - back(hi));
        }
    }
  }
  for (int o = 0; o < kCrC; ++o)
    dst[kCrOffIdent / 2 + (static_cast<size_t>(o / 8) * 32 + o) * 8 + (o % 8)] = cvt(1.f);
  return DD_OK;
}

int dd_compose_scales_fwd(dd_ctx* ctx, const dd_tensor* small, const dd_tensor* large, const void* packed_dev,
                          const float* params_host, int dtype, const dd_invert_params* inv, const dd_tensor* out, void* stream) {
  DD_CHECK_ARG(ctx && packed_dev && params_host && tensor_ok(small) && tensor_ok(large) && tensor_ok(out), "bad argument");
  DD_CHECK_ARG(dtype == DD_F16 || dtype == DD_BF16, "compose_scales: operand type must be fp16 or bf16");
  DD_CHECK_ARG(small->c == 3 && large->c == 3 && out->c == 3 && small->dtype == DD_F32 && large->dtype == DD_F32 &&
                   out->dtype == DD_F32, "compose_scales: small / large / out must be fp32 rgb");
  DD_CHECK_ARG(large->h == 2 * small->h && large->w == 2 * small->w && small->n == large->n && large->n == out->n &&
                   large->h == out->h && large->w == out->w, "compose_scales: dims");
  DD_CHECK_ARG(reinterpret_cast<uintptr_t>(packed_dev) % 16 == 0, "compose_scales: packed weights must be 16-byte aligned");
  ComposeRowsParams p;
  memset(&p, 0, sizeof(p));
  p.small = reinterpret_cast<const float*>(small->ptr); p.small_cs = small->cstride; p.small_co = small->coff;
  p.large = reinterpret_cast<const float*>(large->ptr); p.large_cs = large->cstride; p.large_co = large->coff;
  p.out = reinterpret_cast<float*>(out->ptr); p.out_cs = out->cstride; p.out_co = out->coff;
  p.wblob = static_cast<const uint8_t*>(packed_dev);
  p.N = large->n; p.H = large->h; p.W = large->w;
  p.strips = (p.W + kCrValid - 1) / kCrValid;
  p.total_rows = static_cast<long long>(p.N) * p.strips * p.H;
  p.bf16 = (dtype == DD_BF16);
  if (inv) { p.has_inv = 1; p.inv = *inv; p.sqrt_var = sqrtf(inv->variance); }
  memcpy(p.fl, params_host, sizeof(float) * kCrFloats);
  p.trace = ctx->conv_trace;
  long long grid = ctx->sm_count;
  if (p.total_rows < grid) grid = p.total_rows;
  p.rows_per_cta = static_cast<int>((p.total_rows + grid - 1) / grid);
  grid = (p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  if (p.bf16) {
    DD_CUDA(cudaFuncSetAttribute(compose_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCrSmem));
    compose_rows_kernel<true><<<static_cast<unsigned>(grid), kCrThreads, kCrSmem, static_cast<cudaStream_t>(stream)>>>(p);
  } else {
    DD_CUDA(cudaFuncSetAttribute(compose_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCrSmem));
    compose_rows_kernel<false><<<static_cast<unsigned>(grid), kCrThreads, kCrSmem, static_cast<cudaStream_t>(stream)>>>(p);
  }
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // extern "C"
