// Internal helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/dd_b200.h"

struct dd_ctx {
  int device = 0;
  int sm_count = 0;
  size_t max_smem_optin = 0;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved through the runtime
  int conv_rows = 0;             // cap on input rows per weight pass (0 = auto), see conv_rows.cuh
  int conv_b_stages = 0;         // debug: weight stages of the streaming configuration (0 = auto)
  int std_generic = 0;           // dd_standardize_variance_batch: force the generic tile kernel (A/B switch of the fp32 fast path)
  int conv_dbg = 0;              // debug: ablation switches of conv_rows_kernel (ConvRowsParams::dbg)
  int conv_force_stream = 0;     // debug: stream weights even when they would fit in shared memory
  unsigned long long* conv_trace = nullptr;  // debug: device buffer [64][8] of clock64 stamps (CTA 0)
  std::atomic<int64_t> launches{0};
};

namespace dd {

void set_error(const char* fmt, ...);
// api_wgrad.cu: tensor-core weight gradient (fp16 operands), dw fp32 accumulated; layout 0 [tap][cin][cout], 1 [tap][cout][cin]
int launch_wgrad_rows(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* dz, int ksize, int layout, float* dw, float scale,
                      cudaStream_t stream);

#define DD_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::dd::set_error(__VA_ARGS__);             \
      return DD_ERR_INVALID;                    \
    }                                           \
  } while (0)

#define DD_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::dd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DD_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define DD_LAUNCH_CHECK(ctx)                                                                  \
  do {                                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) {                                                                 \
      ::dd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DD_ERR_CUDA;                                                                     \
    }                                                                                         \
    (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                                  \
  } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline size_t elem_size(int dtype) { return (dtype == DD_F16 || dtype == DD_BF16) ? 2 : 4; }
inline bool is_half_type(int dtype) { return dtype == DD_F16 || dtype == DD_BF16; }

inline bool tensor_ok(const dd_tensor* t) {
  return t && t->ptr && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && t->coff >= 0 &&
         t->coff + t->c <= t->cstride && (t->dtype == DD_F32 || t->dtype == DD_F16 || t->dtype == DD_BF16);
}

// Device-side view with typed element access (fp16 or fp32 storage, fp32 math).
struct View {
  void* ptr;
  int f16;     // storage is fp16 (the 16-byte vectorised fast paths key on this)
  int bf16;    // storage is bfloat16
  int n, h, w, c, cstride, coff;
  __host__ __device__ size_t pix(int in, int y, int x) const {
    return (static_cast<size_t>(in) * h + y) * w + x;
  }
  __device__ float load(size_t pixel, int ch) const {
    const size_t i = pixel * cstride + coff + ch;
    if (f16) return __half2float(reinterpret_cast<const __half*>(ptr)[i]);
    if (bf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ptr)[i]);
    return reinterpret_cast<const float*>(ptr)[i];
  }
  __device__ void store(size_t pixel, int ch, float v) const {
    const size_t i = pixel * cstride + coff + ch;
    if (f16) reinterpret_cast<__half*>(ptr)[i] = __float2half_rn(v);
    else if (bf16) reinterpret_cast<__nv_bfloat16*>(ptr)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(ptr)[i] = v;
  }
};

// 8 consecutive 16-bit channels (one 16-byte access) <-> fp32, for the vectorised elementwise kernels
__device__ __forceinline__ void unpack8(const uint4& v, int bf16, float (&f)[8]) {
  if (bf16) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int bf16) {
  uint4 v;
  if (bf16) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  } else {
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  }
  return v;
}
// a view the 16-byte vectorised kernels can take: 16-bit storage, 8-channel granularity
__host__ __device__ inline bool vec16_ok(const View& v) { return (v.f16 || v.bf16) && v.c % 8 == 0 && v.coff % 8 == 0 && v.cstride % 8 == 0; }

inline View make_view(const dd_tensor* t) {
  View v;
  v.ptr = t->ptr; v.f16 = (t->dtype == DD_F16); v.bf16 = (t->dtype == DD_BF16);
  v.n = t->n; v.h = t->h; v.w = t->w; v.c = t->c; v.cstride = t->cstride; v.coff = t->coff;
  return v;
}

}  // namespace dd
