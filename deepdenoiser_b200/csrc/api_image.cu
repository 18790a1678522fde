// C-ABI: HBM-bound kernels around the convolution stack — pooling, source standardisation + local
// variance, network-input assembly, the kernel-prediction apply and the multi-scale composition.
#include <string.h>

#include <vector>

#include "dd_internal.h"

namespace dd {

__device__ __forceinline__ float signed_log1p(float v) { return copysignf(log1pf(fabsf(v)), v) * (v != 0.f); }
__device__ __forceinline__ float signed_expm1(float v) { return copysignf(expm1f(fabsf(v)), v) * (v != 0.f); }
// np.pad(mode='symmetric'): mirror including the edge sample (Conv2dUtilities.py:77-95)
__device__ __forceinline__ int sym_index(int i, int n) {
  // valid for -n <= i < 2n, iterate for tiny images
  while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - i - 1);
  return i;
}

// ------------------------------------------------------------------------------------------------
// max pooling, stride 2, TF 'SAME' (pad_total = max((out-1)*2 + k - n, 0), pad_before = pad_total / 2)
struct PoolParams {
  View x, y;
  int ksize, pad_y, pad_x;
};

__global__ void __launch_bounds__(256) maxpool_generic_kernel(const PoolParams p) {
  const size_t total = static_cast<size_t>(p.y.n) * p.y.h * p.y.w * p.y.c;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % p.y.c);
  const size_t opix = idx / p.y.c;
  const int ox = static_cast<int>(opix % p.y.w);
  const int oy = static_cast<int>((opix / p.y.w) % p.y.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.y.w) * p.y.h));
  float m = -INFINITY;
  for (int r = 0; r < p.ksize; ++r) {
    const int yy = 2 * oy - p.pad_y + r;
    if (yy < 0 || yy >= p.x.h) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int xx = 2 * ox - p.pad_x + s;
      if (xx < 0 || xx >= p.x.w) continue;
      m = fmaxf(m, p.x.load(p.x.pix(n, yy, xx), c));
    }
  }
  p.y.store(opix, c, m);
}

// 16-bit storage (fp16 or bf16), 8 channels (16 bytes) per thread
template <typename T2>
__device__ __forceinline__ T2 pool_ninf();
template <> __device__ __forceinline__ __half2 pool_ninf<__half2>() { return __float2half2_rn(-INFINITY); }
template <> __device__ __forceinline__ __nv_bfloat162 pool_ninf<__nv_bfloat162>() { return __float2bfloat162_rn(-INFINITY); }

template <typename T2>
__global__ void __launch_bounds__(256) maxpool_h8_kernel(const PoolParams p) {
  const int cv = p.y.c / 8;
  const size_t total = static_cast<size_t>(p.y.n) * p.y.h * p.y.w * cv;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % cv) * 8;
  const size_t opix = idx / cv;
  const int ox = static_cast<int>(opix % p.y.w);
  const int oy = static_cast<int>((opix / p.y.w) % p.y.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.y.w) * p.y.h));
  const uint16_t* xin = reinterpret_cast<const uint16_t*>(p.x.ptr);
  T2 m[4];
  const T2 ninf = pool_ninf<T2>();
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = ninf;
  for (int r = 0; r < p.ksize; ++r) {
    const int yy = 2 * oy - p.pad_y + r;
    if (yy < 0 || yy >= p.x.h) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int xx = 2 * ox - p.pad_x + s;
      if (xx < 0 || xx >= p.x.w) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(xin + p.x.pix(n, yy, xx) * p.x.cstride + p.x.coff + c));
      const T2* h = reinterpret_cast<const T2*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], h[i]);
    }
  }
  uint4 o;
  T2* oh = reinterpret_cast<T2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) oh[i] = m[i];
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y.ptr) + opix * p.y.cstride + p.y.coff + c) = o;
}

// Two vertically adjacent outputs per thread (3x3 / stride 2: input rows 2oy .. 2oy+4, the middle row is shared): all 15 16-byte
// loads of the thread are issued before the first maximum is taken - 7.5 loads per output instead of 9 and more bytes in flight.
template <typename T2>
__global__ void __launch_bounds__(256) maxpool3_h8x2_kernel(const PoolParams p) {
  const int cv = p.y.c / 8;
  const int oh2 = p.y.h >> 1;                                      // output row pairs (p.y.h is even)
  const size_t total = static_cast<size_t>(p.y.n) * oh2 * p.y.w * cv;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % cv) * 8;
  const size_t q = idx / cv;
  const int ox = static_cast<int>(q % p.y.w);
  const int oyp = static_cast<int>((q / p.y.w) % oh2);
  const int n = static_cast<int>(q / (static_cast<size_t>(p.y.w) * oh2));
  const uint16_t* xin = reinterpret_cast<const uint16_t*>(p.x.ptr);
  const T2 ninf = pool_ninf<T2>();
  uint4 v[5][3];
  const int y0 = 4 * oyp - p.pad_y, x0 = 2 * ox - p.pad_x;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int yy = y0 + r, xx = x0 + s;
      T2* h = reinterpret_cast<T2*>(&v[r][s]);
      if (yy >= 0 && yy < p.x.h && xx >= 0 && xx < p.x.w) {
        v[r][s] = __ldg(reinterpret_cast<const uint4*>(xin + p.x.pix(n, yy, xx) * p.x.cstride + p.x.coff + c));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = ninf;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    T2 m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = ninf;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const T2* h = reinterpret_cast<const T2*>(&v[2 * o + r][s]);
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], h[i]);
      }
    }
    uint4 out;
    T2* oh = reinterpret_cast<T2*>(&out);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = m[i];
    const size_t opix = p.y.pix(n, 2 * oyp + o, ox);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y.ptr) + opix * p.y.cstride + p.y.coff + c) = out;
  }
}

// Training form: also records WHICH element of the window is the (first) maximum - one byte per (window, channel), r * k + s -
// so that the backward pass is a pure gather (maxpool_bwd_index_kernel in api_train.cu) instead of re-deriving the first maximum
// of up to four windows per input pixel.
__global__ void __launch_bounds__(256) maxpool_h8_index_kernel(const PoolParams p, uint8_t* __restrict__ index) {
  const int cv = p.y.c / 8;
  const size_t total = static_cast<size_t>(p.y.n) * p.y.h * p.y.w * cv;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % cv) * 8;
  const size_t opix = idx / cv;
  const int ox = static_cast<int>(opix % p.y.w);
  const int oy = static_cast<int>((opix / p.y.w) % p.y.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.y.w) * p.y.h));
  const uint16_t* xin = reinterpret_cast<const uint16_t*>(p.x.ptr);
  float m[8]; uint32_t arg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; arg[i] = 255u; }
  for (int r = 0; r < p.ksize; ++r) {
    const int yy = 2 * oy - p.pad_y + r;
    if (yy < 0 || yy >= p.x.h) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int xx = 2 * ox - p.pad_x + s;
      if (xx < 0 || xx >= p.x.w) continue;
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(xin + p.x.pix(n, yy, xx) * p.x.cstride + p.x.coff + c)), p.x.bf16, v);
      const uint32_t code = static_cast<uint32_t>(r * p.ksize + s);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (arg[i] == 255u || v[i] > m[i]) { m[i] = v[i]; arg[i] = code; }     // first maximum in scan order (TF MaxPoolGrad)
    }
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y.ptr) + opix * p.y.cstride + p.y.coff + c) = pack8(m, p.y.bf16);
  uint2 packed;
  packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
  packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
  *reinterpret_cast<uint2*>(index + opix * p.y.c + c) = packed;
}

// split-fp16 pairs (x = hi + lo, the "float16x2" mode): the maximum is taken on the fp32 sums and re-split exactly
__global__ void __launch_bounds__(256) maxpool_split_kernel(const PoolParams p, const View xlo, const View ylo) {
  const int cv = p.y.c / 8;
  const size_t total = static_cast<size_t>(p.y.n) * p.y.h * p.y.w * cv;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % cv) * 8;
  const size_t opix = idx / cv;
  const int ox = static_cast<int>(opix % p.y.w);
  const int oy = static_cast<int>((opix / p.y.w) % p.y.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.y.w) * p.y.h));
  const uint16_t* xh = reinterpret_cast<const uint16_t*>(p.x.ptr);
  const uint16_t* xl = reinterpret_cast<const uint16_t*>(xlo.ptr);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  for (int r = 0; r < p.ksize; ++r) {
    const int yy = 2 * oy - p.pad_y + r;
    if (yy < 0 || yy >= p.x.h) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int xx = 2 * ox - p.pad_x + s;
      if (xx < 0 || xx >= p.x.w) continue;
      const size_t ip = p.x.pix(n, yy, xx);
      float fh[8], fl[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(xh + ip * p.x.cstride + p.x.coff + c)), 0, fh);
      unpack8(__ldg(reinterpret_cast<const uint4*>(xl + ip * xlo.cstride + xlo.coff + c)), 0, fl);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], fh[i] + fl[i]);
    }
  }
  float hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { hi[i] = __half2float(__float2half_rn(m[i])); lo[i] = m[i] - hi[i]; }
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y.ptr) + opix * p.y.cstride + p.y.coff + c) = pack8(hi, 0);
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ylo.ptr) + opix * ylo.cstride + ylo.coff + c) = pack8(lo, 0);
}

// average pooling factor x factor, stride factor, TF 'SAME' (padded cells excluded from the divisor)
struct AvgPoolParams {
  View x, y;
  int f, pad_y, pad_x;
};
__global__ void __launch_bounds__(256) avgpool_kernel(const AvgPoolParams p) {
  const size_t total = static_cast<size_t>(p.y.n) * p.y.h * p.y.w;
  const size_t opix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (opix >= total) return;
  const int ox = static_cast<int>(opix % p.y.w);
  const int oy = static_cast<int>((opix / p.y.w) % p.y.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.y.w) * p.y.h));
  for (int c = 0; c < p.y.c; ++c) {
    float acc = 0.f;
    int cnt = 0;
    for (int r = 0; r < p.f; ++r) {
      const int yy = oy * p.f - p.pad_y + r;
      if (yy < 0 || yy >= p.x.h) continue;
      for (int s = 0; s < p.f; ++s) {
        const int xx = ox * p.f - p.pad_x + s;
        if (xx < 0 || xx >= p.x.w) continue;
        acc += p.x.load(p.x.pix(n, yy, xx), c);
        ++cnt;
      }
    }
    p.y.store(opix, c, acc / static_cast<float>(cnt));
  }
}

// ------------------------------------------------------------------------------------------------
struct StdParams {
  View src, sout, vout;
  dd_standardize_params q;
  int has_s, has_v;
  float inv_sqrt_var;
};

__device__ __forceinline__ float standardize_value(float v, const dd_standardize_params& q, float inv_sqrt_var) {
  if (q.use_log1p) v = signed_log1p(v);
  if (q.mean != 0.f) v -= q.mean;
  if (q.variance != 1.f) v *= inv_sqrt_var;
  return v;
}

// One 32 x 8 pixel tile per block.  The (standardised, unless compute_before_standardization) values of the tile and
// its one-pixel symmetric halo are computed ONCE into shared memory (log1p is the expensive part), then every thread
// forms the 3x3 / plus-shaped local mean and second moment of its pixel from shared memory.
constexpr int kStdTileW = 32, kStdTileH = 8, kStdHaloW = kStdTileW + 2, kStdHaloH = kStdTileH + 2;
__device__ __forceinline__ void standardize_variance_tile(const StdParams& p, int n, float (*s_val)[kStdHaloH * kStdHaloW]);

__global__ void __launch_bounds__(256) standardize_variance_kernel(const StdParams p) {
  __shared__ float s_val[3][kStdHaloH * kStdHaloW];
  standardize_variance_tile(p, blockIdx.z, s_val);
}

// every pass of a predict() call in ONE launch: job = blockIdx.z / n_images (44 launches of ~58 MB each ran at a third of the
// HBM rate: a 1080p pass is too small to fill the machine between its ramp-up and its tail)
__global__ void __launch_bounds__(256) standardize_variance_batch_kernel(const StdParams* __restrict__ jobs, int n_images) {
  __shared__ float s_val[3][kStdHaloH * kStdHaloW];
  __shared__ StdParams job;
  const int j = blockIdx.z / n_images;
  if (threadIdx.x < sizeof(StdParams) / 4)
    reinterpret_cast<uint32_t*>(&job)[threadIdx.x] = reinterpret_cast<const uint32_t*>(jobs + j)[threadIdx.x];
  __syncthreads();
  standardize_variance_tile(job, blockIdx.z - j * n_images, s_val);
}

__device__ __forceinline__ void standardize_variance_tile(const StdParams& p, int n, float (*s_val)[kStdHaloH * kStdHaloW]) {
  const int ty0 = blockIdx.y * kStdTileH, tx0 = blockIdx.x * kStdTileW;
  const int h = p.src.h, w = p.src.w, C = p.src.c;
  if (p.has_v) {
    for (int i = threadIdx.x; i < kStdHaloH * kStdHaloW; i += 256) {
      const int ly = i / kStdHaloW, lx = i - ly * kStdHaloW;
      const int yy = sym_index(ty0 + ly - 1, h), xx = sym_index(tx0 + lx - 1, w);
      const size_t sp = p.src.pix(n, yy, xx);
      for (int c = 0; c < C; ++c) {
        float v = p.src.load(sp, c);
        if (!p.q.compute_before_standardization) v = standardize_value(v, p.q, p.inv_sqrt_var);
        s_val[c][i] = v;
      }
    }
    __syncthreads();
  }
  const int ly = threadIdx.x >> 5, lx = threadIdx.x & 31;
  const int y0 = ty0 + ly, x0 = tx0 + lx;
  if (y0 >= h || x0 >= w) return;
  const size_t pixel = p.src.pix(n, y0, x0);
  float var_sum = 0.f;
  for (int c = 0; c < C; ++c) {
    if (p.has_s) {
      const float centre = (p.has_v && !p.q.compute_before_standardization)
                               ? s_val[c][(ly + 1) * kStdHaloW + lx + 1]
                               : standardize_value(p.src.load(pixel, c), p.q, p.inv_sqrt_var);
      if (C == 1) { p.sout.store(pixel, 0, centre); p.sout.store(pixel, 1, centre); p.sout.store(pixel, 2, centre); }
      else p.sout.store(pixel, c, centre);
    }
    if (p.has_v) {
      float m = 0.f, m2 = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          if (p.q.variance_mode == 1 && r != 1 && s != 1) continue;  // 'neighbor': plus-shaped stencil
          const float v = s_val[c][(ly + r) * kStdHaloW + lx + s];
          m += v;
          m2 += v * v;
        }
      }
      const float inv_cnt = (p.q.variance_mode == 1) ? (1.f / 5.f) : (1.f / 9.f);
      m *= inv_cnt; m2 *= inv_cnt;
      const float msq = m * m;
      float var = m2 - msq;
      if (p.q.relative_variance) var = var / fmaxf(msq, p.q.epsilon);
      if (p.q.compress_to_one_channel) var_sum += var;
      else p.vout.store(pixel, c, var);
    }
  }
  if (p.has_v && p.q.compress_to_one_channel) p.vout.store(pixel, 0, var_sum / static_cast<float>(C));
}

// fp32 fast path of the batched launch (the frame's 22 passes): a 64 x 16 pixel tile per block (halo overhead 1.16 instead of
// 1.33, a quarter of the blocks), four vertically adjacent pixels per thread so that the six row sums a thread forms serve four
// 3 x 3 windows (4.5 shared-memory loads per window instead of 9), fp32 pointers and 32-bit offsets inside an image.
// Measured (22 passes of 1080p, L2 flushed): 0.905 -> 0.805 ms.  The kernel is bound by log1pf (one per halo element), not by
// memory: a variant with fully coalesced row loads / stores (thread k moves float k of a 198-float halo row) was SLOWER
// (1.09 ms) and is not kept.
constexpr int kStdFastW = 64, kStdFastH = 16, kStdFastPitch = kStdFastW + 4;
__global__ void __launch_bounds__(256) standardize_variance_fast_kernel(const StdParams* __restrict__ jobs, int n_images) {
  __shared__ float s_u[3][kStdFastH + 2][kStdFastPitch];
  __shared__ StdParams job;
  const int j = blockIdx.z / n_images, n = blockIdx.z - j * n_images;
  if (threadIdx.x < sizeof(StdParams) / 4)
    reinterpret_cast<uint32_t*>(&job)[threadIdx.x] = reinterpret_cast<const uint32_t*>(jobs + j)[threadIdx.x];
  __syncthreads();
  const int h = job.src.h, w = job.src.w, C = job.src.c;
  const int ty0 = blockIdx.y * kStdFastH, tx0 = blockIdx.x * kStdFastW;
  const dd_standardize_params q = job.q;
  const float inv_sqrt_var = job.inv_sqrt_var;
  const bool before = q.compute_before_standardization != 0;
  const int s_cs = job.src.cstride;
  const float* src = reinterpret_cast<const float*>(job.src.ptr) + static_cast<size_t>(n) * h * w * s_cs + job.src.coff;
  constexpr int HW = kStdFastW + 2, HH = kStdFastH + 2;
  const int o_cs = job.sout.cstride, v_cs = job.vout.cstride;
  float* sout = job.has_s ? reinterpret_cast<float*>(job.sout.ptr) + static_cast<size_t>(n) * h * w * o_cs + job.sout.coff : nullptr;
  for (int i = threadIdx.x; i < HH * HW; i += 256) {
    const int ly = i / HW, lx = i - ly * HW;
    const int yy = sym_index(ty0 + ly - 1, h), xx = sym_index(tx0 + lx - 1, w);
    const float* sp = src + (yy * w + xx) * s_cs;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(sp + c);
      s_u[c][ly][lx] = before ? v : standardize_value(v, q, inv_sqrt_var);
    }
  }
  __syncthreads();
  const int lx = threadIdx.x & (kStdFastW - 1), ly0 = (threadIdx.x >> 6) * 4;
  const int x = tx0 + lx;
  if (x >= w) return;
  const bool plus = q.variance_mode == 1;
  const float inv_cnt = plus ? (1.f / 5.f) : (1.f / 9.f);
  float* vout = job.has_v ? reinterpret_cast<float*>(job.vout.ptr) + static_cast<size_t>(n) * h * w * v_cs + job.vout.coff : nullptr;
  float var_sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) {
    float rs[6], rs2[6], cen[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const float a = s_u[c][ly0 + r][lx], b = s_u[c][ly0 + r][lx + 1], d = s_u[c][ly0 + r][lx + 2];
      cen[r] = b;
      rs[r] = a + b + d;
      rs2[r] = a * a + b * b + d * d;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int y = ty0 + ly0 + i;
      if (y >= h) break;
      const int pix = y * w + x;
      if (sout) {
        const float centre = before ? standardize_value(cen[i + 1], q, inv_sqrt_var) : cen[i + 1];
        if (C == 1) { sout[pix * o_cs] = centre; sout[pix * o_cs + 1] = centre; sout[pix * o_cs + 2] = centre; }
        else sout[pix * o_cs + c] = centre;
      }
      if (vout) {
        float m, m2;
        if (plus) {
          m = cen[i] + rs[i + 1] + cen[i + 2];
          m2 = cen[i] * cen[i] + rs2[i + 1] + cen[i + 2] * cen[i + 2];
        } else {
          m = rs[i] + rs[i + 1] + rs[i + 2];
          m2 = rs2[i] + rs2[i + 1] + rs2[i + 2];
        }
        m *= inv_cnt; m2 *= inv_cnt;
        const float msq = m * m;
        float var = m2 - msq;
        if (q.relative_variance) var = var / fmaxf(msq, q.epsilon);
        if (q.compress_to_one_channel) var_sum[i] += var;
        else vout[pix * v_cs + c] = var;
      }
    }
  }
  if (vout && q.compress_to_one_channel) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int y = ty0 + ly0 + i;
      if (y < h) vout[(y * w + x) * v_cs] = var_sum[i] / static_cast<float>(C);
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct AssembleParams {
  const dd_gather_entry* table;
  View out;
  int tuples, n;
  int split;       // float16x2 mode: out = fp16(v), out_lo = fp16(v - out)
  View out_lo;
};
__global__ void __launch_bounds__(256) assemble_kernel(const AssembleParams p) {
  const size_t img_pix = static_cast<size_t>(p.out.h) * p.out.w;
  const size_t total = static_cast<size_t>(p.tuples) * p.n * img_pix;
  const size_t opix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (opix >= total) return;
  const size_t b = opix / img_pix;         // output image = t * n + i
  const int t = static_cast<int>(b / p.n);
  const size_t spix = (b % p.n) * img_pix + (opix % img_pix);  // pixel index inside the [n,h,w] source
  const dd_gather_entry* row = p.table + static_cast<size_t>(t) * p.out.c;
  if (vec16_ok(p.out)) {
    uint16_t* o = reinterpret_cast<uint16_t*>(p.out.ptr) + opix * p.out.cstride + p.out.coff;
    for (int c0 = 0; c0 < p.out.c; c0 += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const dd_gather_entry e = row[c0 + i];
        f[i] = e.ptr ? __ldg(e.ptr + spix * e.cstride + e.cidx) : e.constant;
      }
      *reinterpret_cast<uint4*>(o + c0) = pack8(f, p.out.bf16);
      if (p.split) {
        float lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) lo[i] = f[i] - __half2float(__float2half_rn(f[i]));
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out_lo.ptr) + opix * p.out_lo.cstride + p.out_lo.coff + c0) = pack8(lo, 0);
      }
    }
  } else {
    for (int c = 0; c < p.out.c; ++c) {
      const dd_gather_entry e = row[c];
      p.out.store(opix, c, e.ptr ? __ldg(e.ptr + spix * e.cstride + e.cidx) : e.constant);
    }
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int kComposeMaxC = 32;
struct ComposeHeadParams {
  View small, large, y;
  float w[6 * kComposeMaxC];
  float b[kComposeMaxC];
  int c_mid;
};
__global__ void __launch_bounds__(256) compose_head_kernel(const ComposeHeadParams p) {
  const size_t total = static_cast<size_t>(p.large.n) * p.large.h * p.large.w;
  const size_t pixel = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pixel >= total) return;
  const int x0 = static_cast<int>(pixel % p.large.w);
  const int y0 = static_cast<int>((pixel / p.large.w) % p.large.h);
  const int n = static_cast<int>(pixel / (static_cast<size_t>(p.large.w) * p.large.h));
  const size_t spix = p.small.pix(n, y0 >> 1, x0 >> 1);
  float in[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) { in[c] = p.small.load(spix, c); in[3 + c] = p.large.load(pixel, c); }
  for (int c = 0; c < p.y.c; ++c) {
    float v = 0.f;
    if (c < p.c_mid) {
      v = p.b[c];
#pragma unroll
      for (int k = 0; k < 6; ++k) v = fmaf(in[k], p.w[k * p.c_mid + c], v);
      v = fmaxf(v, 0.f);
    }
    p.y.store(pixel, c, v);
  }
}

struct ComposeTailParams {
  View t, small, large, out;
  float w[kComposeMaxC];
  float b;
  int c_mid;
  int has_inv;
  dd_invert_params inv;
  float sqrt_var;
};
__device__ __forceinline__ float invert_value(float v, const dd_invert_params& q, float sqrt_var) {
  if (q.variance != 1.f) v *= sqrt_var;
  if (q.mean != 0.f) v += q.mean;
  if (q.use_log1p) v = signed_expm1(v);
  return v;
}
__global__ void __launch_bounds__(256) compose_tail_kernel(const ComposeTailParams p) {
  const size_t total = static_cast<size_t>(p.large.n) * p.large.h * p.large.w;
  const size_t pixel = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pixel >= total) return;
  const int x0 = static_cast<int>(pixel % p.large.w);
  const int y0 = static_cast<int>((pixel / p.large.w) % p.large.h);
  const int n = static_cast<int>(pixel / (static_cast<size_t>(p.large.w) * p.large.h));
  float a = p.b;
  for (int c = 0; c < p.c_mid; ++c) a = fmaf(p.t.load(pixel, c), p.w[c], a);
  a = fmaxf(a, 0.f);
  const float wgt = 1.f / (1.f + __expf(-a));
  const size_t spix = p.small.pix(n, y0 >> 1, x0 >> 1);
  const int yb = y0 & ~1, xb = x0 & ~1;
  for (int c = 0; c < 3; ++c) {
    const float low = 0.25f * (p.large.load(p.large.pix(n, yb, xb), c) + p.large.load(p.large.pix(n, yb, xb + 1), c) +
                               p.large.load(p.large.pix(n, yb + 1, xb), c) + p.large.load(p.large.pix(n, yb + 1, xb + 1), c));
    float v = p.large.load(pixel, c) - wgt * low + wgt * p.small.load(spix, c);
    if (p.has_inv) v = invert_value(v, p.inv, p.sqrt_var);
    p.out.store(pixel, c, v);
  }
}

struct InvertParams {
  View x, y;
  dd_invert_params inv;
  float sqrt_var;
};
__global__ void __launch_bounds__(256) invert_kernel(const InvertParams p) {
  const size_t total = static_cast<size_t>(p.x.n) * p.x.h * p.x.w;
  const size_t pixel = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pixel >= total) return;
  for (int c = 0; c < p.x.c; ++c) p.y.store(pixel, c, invert_value(p.x.load(pixel, c), p.inv, p.sqrt_var));
}

struct CastParams { View x, y; };
__global__ void __launch_bounds__(256) cast_copy_kernel(const CastParams p) {
  const size_t total = static_cast<size_t>(p.x.n) * p.x.h * p.x.w * p.x.c;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % p.x.c);
  const size_t pixel = idx / p.x.c;
  p.y.store(pixel, c, p.x.load(pixel, c));
}

__global__ void __launch_bounds__(256) l2_flush_kernel(uint4* buf, size_t n16, uint32_t tag) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride)
    buf[i] = make_uint4(tag, tag + 1, tag + 2, static_cast<uint32_t>(i));
}

inline unsigned blocks_for(size_t total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }
inline bool same_spatial(const dd_tensor* a, const dd_tensor* b) { return a->n == b->n && a->h == b->h && a->w == b->w; }

}  // namespace dd

using namespace dd;

extern "C" {

int dd_maxpool_s2_fwd(dd_ctx* ctx, const dd_tensor* x, int ksize, const dd_tensor* y, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y), "bad argument");
  DD_CHECK_ARG(ksize == 2 || ksize == 3, "maxpool ksize must be 2 or 3");
  const int oh = (x->h + 1) / 2, ow = (x->w + 1) / 2;
  DD_CHECK_ARG(y->n == x->n && y->h == oh && y->w == ow && y->c == x->c, "maxpool: bad output dims");
  PoolParams p;
  p.x = make_view(x); p.y = make_view(y); p.ksize = ksize;
  const int pty = (oh - 1) * 2 + ksize - x->h, ptx = (ow - 1) * 2 + ksize - x->w;
  p.pad_y = (pty > 0 ? pty : 0) / 2; p.pad_x = (ptx > 0 ? ptx : 0) / 2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = is_half_type(x->dtype) && y->dtype == x->dtype && x->c % 8 == 0 && x->coff % 8 == 0 &&
                   x->cstride % 8 == 0 && y->coff % 8 == 0 && y->cstride % 8 == 0;
  if (vec && ksize == 3 && (y->h & 1) == 0) {
    const size_t total = static_cast<size_t>(y->n) * (y->h / 2) * y->w * (y->c / 8);
    if (x->dtype == DD_BF16) maxpool3_h8x2_kernel<__nv_bfloat162><<<blocks_for(total, 256), 256, 0, s>>>(p);
    else maxpool3_h8x2_kernel<__half2><<<blocks_for(total, 256), 256, 0, s>>>(p);
  } else if (vec) {
    const size_t total = static_cast<size_t>(y->n) * y->h * y->w * (y->c / 8);
    if (x->dtype == DD_BF16) maxpool_h8_kernel<__nv_bfloat162><<<blocks_for(total, 256), 256, 0, s>>>(p);
    else maxpool_h8_kernel<__half2><<<blocks_for(total, 256), 256, 0, s>>>(p);
  } else {
    const size_t total = static_cast<size_t>(y->n) * y->h * y->w * y->c;
    maxpool_generic_kernel<<<blocks_for(total, 256), 256, 0, s>>>(p);
  }
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_maxpool_s2_fwd_index(dd_ctx* ctx, const dd_tensor* x, int ksize, const dd_tensor* y, uint8_t* index_dev, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y) && index_dev, "bad argument");
  DD_CHECK_ARG(ksize == 2 || ksize == 3, "maxpool ksize must be 2 or 3");
  const int oh = (x->h + 1) / 2, ow = (x->w + 1) / 2;
  DD_CHECK_ARG(y->n == x->n && y->h == oh && y->w == ow && y->c == x->c, "maxpool: bad output dims");
  DD_CHECK_ARG(is_half_type(x->dtype) && y->dtype == x->dtype && x->c % 8 == 0 && x->coff % 8 == 0 && x->cstride % 8 == 0 &&
                   y->coff % 8 == 0 && y->cstride % 8 == 0, "maxpool_index: fp16 / bf16 views with multiples of 8 channels expected");
  PoolParams p;
  p.x = make_view(x); p.y = make_view(y); p.ksize = ksize;
  const int pty = (oh - 1) * 2 + ksize - x->h, ptx = (ow - 1) * 2 + ksize - x->w;
  p.pad_y = (pty > 0 ? pty : 0) / 2; p.pad_x = (ptx > 0 ? ptx : 0) / 2;
  const size_t total = static_cast<size_t>(y->n) * y->h * y->w * (y->c / 8);
  maxpool_h8_index_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, index_dev);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_maxpool_s2_fwd_split(dd_ctx* ctx, const dd_tensor* x_hi, const dd_tensor* x_lo, int ksize, const dd_tensor* y_hi,
                            const dd_tensor* y_lo, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x_hi) && tensor_ok(x_lo) && tensor_ok(y_hi) && tensor_ok(y_lo), "bad argument");
  DD_CHECK_ARG(ksize == 2 || ksize == 3, "maxpool ksize must be 2 or 3");
  const dd_tensor* all[4] = {x_hi, x_lo, y_hi, y_lo};
  for (const dd_tensor* t : all)
    DD_CHECK_ARG(t->dtype == DD_F16 && t->c == x_hi->c && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0,
                 "maxpool_split: fp16 views with multiples of 8 channels expected");
  const int oh = (x_hi->h + 1) / 2, ow = (x_hi->w + 1) / 2;
  DD_CHECK_ARG(y_hi->n == x_hi->n && y_hi->h == oh && y_hi->w == ow && y_lo->h == oh && y_lo->w == ow && x_lo->h == x_hi->h &&
                   x_lo->w == x_hi->w, "maxpool_split: bad dims");
  PoolParams p;
  p.x = make_view(x_hi); p.y = make_view(y_hi); p.ksize = ksize;
  const int pty = (oh - 1) * 2 + ksize - x_hi->h, ptx = (ow - 1) * 2 + ksize - x_hi->w;
  p.pad_y = (pty > 0 ? pty : 0) / 2; p.pad_x = (ptx > 0 ? ptx : 0) / 2;
  const size_t total = static_cast<size_t>(y_hi->n) * oh * ow * (y_hi->c / 8);
  maxpool_split_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, make_view(x_lo), make_view(y_lo));
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_avgpool_fwd(dd_ctx* ctx, const dd_tensor* x, int factor, const dd_tensor* y, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y), "bad argument");
  DD_CHECK_ARG(factor >= 1 && factor <= 16, "avgpool factor out of range");
  const int oh = (x->h + factor - 1) / factor, ow = (x->w + factor - 1) / factor;
  DD_CHECK_ARG(y->n == x->n && y->h == oh && y->w == ow && y->c == x->c, "avgpool: bad output dims");
  AvgPoolParams p;
  p.x = make_view(x); p.y = make_view(y); p.f = factor;
  p.pad_y = (oh * factor - x->h) / 2; p.pad_x = (ow * factor - x->w) / 2;
  const size_t total = static_cast<size_t>(y->n) * y->h * y->w;
  avgpool_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

static int make_std_params(const dd_tensor* src, const dd_standardize_params* prm, const dd_tensor* std_out,
                           const dd_tensor* var_out, StdParams& p);

int dd_standardize_variance_batch(dd_ctx* ctx, int count, const dd_tensor* const* src, const dd_standardize_params* prm,
                                  const dd_tensor* const* std_out, const dd_tensor* const* var_out, void* table_dev,
                                  size_t table_bytes, void* stream) {
  DD_CHECK_ARG(ctx && count > 0 && src && prm && std_out && var_out && table_dev, "bad argument");
  DD_CHECK_ARG(table_bytes >= static_cast<size_t>(count) * sizeof(StdParams), "table_dev too small (%zu bytes per job)", sizeof(StdParams));
  static_assert(sizeof(StdParams) % 4 == 0 && sizeof(StdParams) / 4 <= 256, "job record is copied by one block");
  std::vector<StdParams> jobs(static_cast<size_t>(count));
  for (int i = 0; i < count; ++i) {
    int rc = make_std_params(src[i], prm + i, std_out[i], var_out[i], jobs[i]);
    if (rc) return rc;
    DD_CHECK_ARG(src[i]->n == src[0]->n && src[i]->h == src[0]->h && src[i]->w == src[0]->w, "batched passes must share [n,h,w]");
  }
  const dd_tensor* s0 = src[0];
  DD_CHECK_ARG(static_cast<long long>(s0->n) * count <= 65535 && (s0->h + kStdTileH - 1) / kStdTileH <= 65535, "standardize: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // pageable source: the runtime stages it before returning, the vector may die with this call
  DD_CUDA(cudaMemcpyAsync(table_dev, jobs.data(), jobs.size() * sizeof(StdParams), cudaMemcpyHostToDevice, s));
  bool fast = static_cast<long long>(s0->h) * s0->w * 4 < (1ll << 31);          // 32-bit offsets inside an image
  for (const StdParams& jb : jobs)
    fast = fast && jb.src.cstride <= 4 && (!jb.has_s || jb.sout.cstride <= 4) && (!jb.has_v || jb.vout.cstride <= 4) && !jb.src.f16 && !jb.src.bf16 && (!jb.has_s || (!jb.sout.f16 && !jb.sout.bf16)) && (!jb.has_v || (!jb.vout.f16 && !jb.vout.bf16));
  if (fast && !ctx->std_generic) {
    dim3 grid((s0->w + kStdFastW - 1) / kStdFastW, (s0->h + kStdFastH - 1) / kStdFastH, s0->n * count);
    standardize_variance_fast_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const StdParams*>(table_dev), s0->n);
  } else {
    dim3 grid((s0->w + kStdTileW - 1) / kStdTileW, (s0->h + kStdTileH - 1) / kStdTileH, s0->n * count);
    standardize_variance_batch_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const StdParams*>(table_dev), s0->n);
  }
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

size_t dd_standardize_variance_job_bytes(void) { return sizeof(StdParams); }

static int make_std_params(const dd_tensor* src, const dd_standardize_params* prm, const dd_tensor* std_out,
                           const dd_tensor* var_out, StdParams& p) {
  DD_CHECK_ARG(prm && src && tensor_ok(src), "bad argument");
  DD_CHECK_ARG(src->c == 1 || src->c == 3, "source must have 1 or 3 channels");
  DD_CHECK_ARG(std_out || var_out, "nothing to compute");
  memset(&p, 0, sizeof(p));
  p.src = make_view(src); p.q = *prm;
  p.inv_sqrt_var = 1.f / sqrtf(prm->variance);
  if (std_out) {
    DD_CHECK_ARG(tensor_ok(std_out) && same_spatial(src, std_out) && std_out->c == 3, "std_out must be [n,h,w,3]");
    p.sout = make_view(std_out); p.has_s = 1;
  }
  if (var_out && prm->use_variance) {
    const int vc = prm->compress_to_one_channel ? 1 : src->c;
    DD_CHECK_ARG(tensor_ok(var_out) && same_spatial(src, var_out) && var_out->c == vc, "var_out has wrong dims");
    p.vout = make_view(var_out); p.has_v = 1;
  }
  return DD_OK;
}

int dd_standardize_variance(dd_ctx* ctx, const dd_tensor* src, const dd_standardize_params* prm,
                            const dd_tensor* std_out, const dd_tensor* var_out, void* stream) {
  DD_CHECK_ARG(ctx && prm && tensor_ok(src), "bad argument");
  DD_CHECK_ARG(src->c == 1 || src->c == 3, "source must have 1 or 3 channels");
  DD_CHECK_ARG(std_out || var_out, "nothing to compute");
  StdParams p;
  memset(&p, 0, sizeof(p));
  p.src = make_view(src); p.q = *prm;
  p.inv_sqrt_var = 1.f / sqrtf(prm->variance);
  if (std_out) {
    DD_CHECK_ARG(tensor_ok(std_out) && same_spatial(src, std_out) && std_out->c == 3, "std_out must be [n,h,w,3]");
    p.sout = make_view(std_out); p.has_s = 1;
  }
  if (var_out && prm->use_variance) {
    const int vc = prm->compress_to_one_channel ? 1 : src->c;
    DD_CHECK_ARG(tensor_ok(var_out) && same_spatial(src, var_out) && var_out->c == vc, "var_out has wrong dims");
    p.vout = make_view(var_out); p.has_v = 1;
  }
  DD_CHECK_ARG(src->n <= 65535 && (src->h + kStdTileH - 1) / kStdTileH <= 65535, "standardize: grid too large");
  dim3 grid((src->w + kStdTileW - 1) / kStdTileW, (src->h + kStdTileH - 1) / kStdTileH, src->n);
  standardize_variance_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_assemble_input(dd_ctx* ctx, const dd_gather_entry* table_dev, int tuples, int n, const dd_tensor* out,
                      void* stream) {
  DD_CHECK_ARG(ctx && table_dev && tensor_ok(out), "bad argument");
  DD_CHECK_ARG(tuples > 0 && n > 0 && out->n == tuples * n, "assemble: out.n must be tuples*n");
  AssembleParams p;
  memset(&p, 0, sizeof(p));
  p.table = table_dev; p.out = make_view(out); p.tuples = tuples; p.n = n;
  const size_t total = static_cast<size_t>(out->n) * out->h * out->w;
  assemble_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_assemble_input_split(dd_ctx* ctx, const dd_gather_entry* table_dev, int tuples, int n, const dd_tensor* out_hi,
                            const dd_tensor* out_lo, void* stream) {
  DD_CHECK_ARG(ctx && table_dev && tensor_ok(out_hi) && tensor_ok(out_lo), "bad argument");
  DD_CHECK_ARG(tuples > 0 && n > 0 && out_hi->n == tuples * n, "assemble: out.n must be tuples*n");
  AssembleParams p;
  memset(&p, 0, sizeof(p));
  p.table = table_dev; p.out = make_view(out_hi); p.out_lo = make_view(out_lo); p.tuples = tuples; p.n = n; p.split = 1;
  DD_CHECK_ARG(out_hi->dtype == DD_F16 && out_lo->dtype == DD_F16 && vec16_ok(p.out) && vec16_ok(p.out_lo) && out_lo->c == out_hi->c &&
                   out_lo->h == out_hi->h && out_lo->w == out_hi->w, "assemble_split: fp16 views with multiples of 8 channels expected");
  const size_t total = static_cast<size_t>(out_hi->n) * out_hi->h * out_hi->w;
  assemble_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_compose_head_fwd(dd_ctx* ctx, const dd_tensor* small, const dd_tensor* large, const float* w, const float* b,
                        int c_mid, const dd_tensor* y, void* stream) {
  DD_CHECK_ARG(ctx && w && b && tensor_ok(small) && tensor_ok(large) && tensor_ok(y), "bad argument");
  DD_CHECK_ARG(c_mid > 0 && c_mid <= kComposeMaxC && y->c >= c_mid, "compose width unsupported");
  DD_CHECK_ARG(small->c == 3 && large->c == 3 && large->h == 2 * small->h && large->w == 2 * small->w &&
                   small->n == large->n && same_spatial(large, y), "compose_head: dims");
  ComposeHeadParams p;
  memset(&p, 0, sizeof(p));
  p.small = make_view(small); p.large = make_view(large); p.y = make_view(y); p.c_mid = c_mid;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // weights are tiny: stage them through the kernel parameter block (host pointers expected)
  memcpy(p.w, w, sizeof(float) * 6 * c_mid);
  memcpy(p.b, b, sizeof(float) * c_mid);
  const size_t total = static_cast<size_t>(large->n) * large->h * large->w;
  compose_head_kernel<<<blocks_for(total, 256), 256, 0, s>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_compose_tail_fwd(dd_ctx* ctx, const dd_tensor* t, const float* w, const float* b, int c_mid,
                        const dd_tensor* small, const dd_tensor* large, const dd_invert_params* inv,
                        const dd_tensor* out, void* stream) {
  DD_CHECK_ARG(ctx && w && b && tensor_ok(t) && tensor_ok(small) && tensor_ok(large) && tensor_ok(out), "bad argument");
  DD_CHECK_ARG(c_mid > 0 && c_mid <= kComposeMaxC && t->c >= c_mid, "compose width unsupported");
  DD_CHECK_ARG(small->c == 3 && large->c == 3 && out->c == 3 && large->h == 2 * small->h && large->w == 2 * small->w &&
                   small->n == large->n && same_spatial(large, t) && same_spatial(large, out), "compose_tail: dims");
  ComposeTailParams p;
  memset(&p, 0, sizeof(p));
  p.t = make_view(t); p.small = make_view(small); p.large = make_view(large); p.out = make_view(out);
  p.c_mid = c_mid;
  memcpy(p.w, w, sizeof(float) * c_mid);
  p.b = b[0];
  if (inv) { p.has_inv = 1; p.inv = *inv; p.sqrt_var = sqrtf(inv->variance); }
  const size_t total = static_cast<size_t>(large->n) * large->h * large->w;
  compose_tail_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_invert_standardization(dd_ctx* ctx, const dd_tensor* x, const dd_invert_params* inv, const dd_tensor* y,
                              void* stream) {
  DD_CHECK_ARG(ctx && inv && tensor_ok(x) && tensor_ok(y) && same_spatial(x, y) && x->c == y->c, "bad argument");
  InvertParams p;
  p.x = make_view(x); p.y = make_view(y); p.inv = *inv; p.sqrt_var = sqrtf(inv->variance);
  const size_t total = static_cast<size_t>(x->n) * x->h * x->w;
  invert_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_cast_copy(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y) && same_spatial(x, y) && x->c == y->c, "bad argument");
  CastParams p;
  p.x = make_view(x); p.y = make_view(y);
  const size_t total = static_cast<size_t>(x->n) * x->h * x->w * x->c;
  cast_copy_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_l2_flush(dd_ctx* ctx, void* scratch, size_t bytes, void* stream) {
  DD_CHECK_ARG(ctx && scratch && bytes >= 16, "bad argument");
  static uint32_t tag = 0;
  l2_flush_kernel<<<ctx->sm_count * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<uint4*>(scratch),
                                                                                    bytes / 16, ++tag);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // extern "C"
