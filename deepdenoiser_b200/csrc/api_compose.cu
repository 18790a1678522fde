// C-ABI: MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93) as ONE tcgen05 kernel (compose_rows.cuh).
#include <string.h>

#include "compose_rows.cuh"

using namespace dd;

extern "C" {

size_t dd_compose_weights_bytes(void) { return kCrWBytes; }
size_t dd_compose_params_floats(void) { return kCrFloats; }

/* Host-side packing of the four 3x3 24->24 kernels (TF layout [3,3,24,24] = [kh,kw,cin,cout]) into the operand layout of
 * compose_rows_kernel: [layer][kw = s][8-channel chunk j][row n = kh * 32 + cout][8 cin] 16-bit, zero padded. */
int dd_compose_pack_weights(const float* const* conv_w, int dtype, void* blob_host) {
  DD_CHECK_ARG(conv_w && blob_host && (dtype == DD_F16 || dtype == DD_BF16), "bad argument");
  uint16_t* dst = reinterpret_cast<uint16_t*>(blob_host);
  memset(dst, 0, kCrWBytes);
  for (int l = 0; l < 4; ++l)
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s)
        for (int c = 0; c < kCrC; ++c)
          for (int o = 0; o < kCrC; ++o) {
            const float v = conv_w[l][((r * 3 + s) * kCrC + c) * kCrC + o];
            uint16_t bits;
            if (dtype == DD_BF16) { const __nv_bfloat16 b = __float2bfloat16_rn(v); bits = *reinterpret_cast<const uint16_t*>(&b); }
            else { const __half h = __float2half_rn(v); bits = *reinterpret_cast<const uint16_t*>(&h); }
            const size_t tile = static_cast<size_t>(l * 3 + s) * (kCrWTile / 2);
            dst[tile + (static_cast<size_t>(c / 8) * 96 + r * 32 + o) * 8 + (c % 8)] = bits;
          }
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
