// C-ABI: training-side kernels — loss (LossDifference.py:15-36 + Training.py:116-243), the backward of every forward
// op (what tf.train.AdamOptimizer.minimize derives by autodiff, Training.py:700-702) and the TF-form Adam update.
// This is the EXACT (fp32 accumulate, CUDA-core) training path of round 1: correctness first, every kernel is
// parity-tested against torch-autograd of the oracle; the tensor-core wgrad is future work.
#include <string.h>

#include "dd_internal.h"

namespace dd {

inline unsigned nblocks(size_t total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }
inline bool same_dims(const dd_tensor* a, const dd_tensor* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}

// ------------------------------------------------------------------------------------------------ elementwise
struct Ewise { View a, b, c, out, out2, out3; int op; float alpha; };
enum { EW_RELU_MASK = 0, EW_MULADD = 1, EW_MULADD_BWD = 2, EW_AXPY = 3, EW_FILL = 4, EW_INVERT_BWD = 5, EW_RELU_MASK_ACC = 6 };
struct EwiseInv { int use_log1p; float mean, variance, sqrt_var; };

__global__ void __launch_bounds__(256) ewise_kernel(const Ewise p, const EwiseInv q) {
  const size_t total = static_cast<size_t>(p.out.n) * p.out.h * p.out.w * p.out.c;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = static_cast<int>(idx % p.out.c);
  const size_t pix = idx / p.out.c;
  switch (p.op) {
    case EW_RELU_MASK:      // out = a * [b > 0]        (dz = dy . relu'(y))
      p.out.store(pix, ch, p.b.load(pix, ch) > 0.f ? p.a.load(pix, ch) : 0.f);
      break;
    case EW_RELU_MASK_ACC:  // out += a * [b > 0]       (dense-block concat: several consumers add into one gradient)
      if (p.b.load(pix, ch) > 0.f) p.out.store(pix, ch, p.out.load(pix, ch) + p.a.load(pix, ch));
      break;
    case EW_MULADD:         // out = a * (b + c)        (lighting = color * (direct + indirect), Training.py:420-428)
      p.out.store(pix, ch, p.a.load(pix, ch) * (p.b.load(pix, ch) + p.c.load(pix, ch)));
      break;
    case EW_MULADD_BWD: {   // g = out: a' += g (b + c), b' += g a, c' += g a   (accumulated into out2/out3 and `out` is g)
      // here: a,b,c forward operands; out = g (read); out2 = da (accumulate), out3 = db (accumulate), alpha unused;
      // dc is accumulated by the caller with a second AXPY of db's increment -> we write the same increment to both
      const float g = p.out.load(pix, ch);
      const float av = p.a.load(pix, ch), bv = p.b.load(pix, ch), cv = p.c.load(pix, ch);
      p.out2.store(pix, ch, p.out2.load(pix, ch) + g * (bv + cv));
      p.out3.store(pix, ch, g * av);          // increment for BOTH direct and indirect
      break;
    }
    case EW_AXPY:           // out += alpha * a
      p.out.store(pix, ch, p.out.load(pix, ch) + p.alpha * p.a.load(pix, ch));
      break;
    case EW_FILL:
      p.out.store(pix, ch, p.alpha);
      break;
    case EW_INVERT_BWD: {   // y = signed_expm1(x*sqrt(var)+mean): out = a * dy/dx at x = b      (Architecture.py:48-55)
      float x = p.b.load(pix, ch);
      float d = 1.f;
      if (q.variance != 1.f) { x *= q.sqrt_var; d *= q.sqrt_var; }
      if (q.mean != 0.f) x += q.mean;
      if (q.use_log1p) d *= expf(fabsf(x));
      p.out.store(pix, ch, p.a.load(pix, ch) * d);
      break;
    }
  }
}

// 16-byte vectorised variant for aligned fp16 tensors (8 channels per thread): the activation-gradient traffic of the
// tensor-core training path (ReLU masks through concatenations, gradient accumulation).
__global__ void __launch_bounds__(256) ewise_vec_kernel(const Ewise p) {
  const int G = p.out.c >> 3;
  const size_t total = static_cast<size_t>(p.out.n) * p.out.h * p.out.w * G;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = static_cast<int>(idx % G);
  const size_t pix = idx / G;
  auto at = [&](const View& v) { return reinterpret_cast<uint16_t*>(v.ptr) + pix * v.cstride + v.coff + g * 8; };
  float a[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(at(p.a)), p.a.bf16, a);
  if (p.op == EW_AXPY) {
    unpack8(*reinterpret_cast<const uint4*>(at(p.out)), p.out.bf16, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(p.alpha, a[i], o[i]);
  } else {
    float b[8];
    unpack8(*reinterpret_cast<const uint4*>(at(p.b)), p.b.bf16, b);
    if (p.op == EW_RELU_MASK_ACC) {
      unpack8(*reinterpret_cast<const uint4*>(at(p.out)), p.out.bf16, o);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (b[i] > 0.f) ? o[i] + a[i] : o[i];
  }
  *reinterpret_cast<uint4*>(at(p.out)) = pack8(o, p.out.bf16);
}

// Flat fp32 form: every operand is a dense fp32 tensor (no channel window), so the op runs over a flat array with 32-bit
// indices and four elements per thread.  The image-space gradient plumbing of a training step (axpy / lighting products /
// inverse-standardisation backward on [B,h,w,3] fp32 tensors) is ~140 launches per step; the generic kernel spends a 64-bit
// division and a modulo per element on them.
struct EwiseFlat { const float* a; const float* b; const float* c; float* out; float* out2; float* out3; int op; float alpha; unsigned n4; };
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 o) { return make_float4(fmaf(s, a.x, o.x), fmaf(s, a.y, o.y), fmaf(s, a.z, o.z), fmaf(s, a.w, o.w)); }
__global__ void __launch_bounds__(256) ewise_flat_kernel(const EwiseFlat p, const EwiseInv q) {
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= p.n4) return;
  const float4* a4 = reinterpret_cast<const float4*>(p.a);
  const float4* b4 = reinterpret_cast<const float4*>(p.b);
  const float4* c4 = reinterpret_cast<const float4*>(p.c);
  float4* o4 = reinterpret_cast<float4*>(p.out);
  switch (p.op) {
    case EW_AXPY: o4[i] = f4_fma(p.alpha, a4[i], o4[i]); break;
    case EW_FILL: o4[i] = make_float4(p.alpha, p.alpha, p.alpha, p.alpha); break;
    case EW_MULADD: {
      const float4 a = a4[i], b = b4[i], c = c4[i];
      o4[i] = make_float4(a.x * (b.x + c.x), a.y * (b.y + c.y), a.z * (b.z + c.z), a.w * (b.w + c.w));
      break;
    }
    case EW_MULADD_BWD: {
      const float4 g = o4[i], a = a4[i], b = b4[i], c = c4[i];
      float4* d2 = reinterpret_cast<float4*>(p.out2);
      const float4 o = d2[i];
      d2[i] = make_float4(o.x + g.x * (b.x + c.x), o.y + g.y * (b.y + c.y), o.z + g.z * (b.z + c.z), o.w + g.w * (b.w + c.w));
      reinterpret_cast<float4*>(p.out3)[i] = make_float4(g.x * a.x, g.y * a.y, g.z * a.z, g.w * a.w);
      break;
    }
    case EW_RELU_MASK: {
      const float4 a = a4[i], b = b4[i];
      o4[i] = make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f);
      break;
    }
    case EW_INVERT_BWD: {
      const float4 a = a4[i], b = b4[i];
      const float xs[4] = {b.x, b.y, b.z, b.w};
      float d[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float x = xs[k]; d[k] = 1.f;
        if (q.variance != 1.f) { x *= q.sqrt_var; d[k] *= q.sqrt_var; }
        if (q.mean != 0.f) x += q.mean;
        if (q.use_log1p) d[k] *= expf(fabsf(x));
      }
      o4[i] = make_float4(a.x * d[0], a.y * d[1], a.z * d[2], a.w * d[3]);
      break;
    }
    default: break;
  }
}
inline bool flat_ok(const View& v) { return v.ptr && !v.f16 && !v.bf16 && v.coff == 0 && v.cstride == v.c && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0; }
// launches the flat kernel when the op's operands allow it (mask bits: 1 a, 2 b, 4 c, 8 out2, 16 out3; `out` always)
inline bool try_ewise_flat(const Ewise& p, const EwiseInv& q, unsigned uses, cudaStream_t s) {
  if (p.op == EW_RELU_MASK_ACC) return false;
  const size_t total = static_cast<size_t>(p.out.n) * p.out.h * p.out.w * p.out.c;
  if ((total & 3) || total / 4 >= 0xffffff00u || !flat_ok(p.out)) return false;
  if (((uses & 1) && !flat_ok(p.a)) || ((uses & 2) && !flat_ok(p.b)) || ((uses & 4) && !flat_ok(p.c)) ||
      ((uses & 8) && !flat_ok(p.out2)) || ((uses & 16) && !flat_ok(p.out3))) return false;
  EwiseFlat f;
  f.a = reinterpret_cast<const float*>(p.a.ptr); f.b = reinterpret_cast<const float*>(p.b.ptr); f.c = reinterpret_cast<const float*>(p.c.ptr);
  f.out = reinterpret_cast<float*>(p.out.ptr); f.out2 = reinterpret_cast<float*>(p.out2.ptr); f.out3 = reinterpret_cast<float*>(p.out3.ptr);
  f.op = p.op; f.alpha = p.alpha; f.n4 = static_cast<unsigned>(total / 4);
  ewise_flat_kernel<<<(f.n4 + 255) / 256, 256, 0, s>>>(f, q);
  return true;
}

inline bool vec_ok(const View& v) { return vec16_ok(v); }
// launches the vectorised kernel when every operand allows it; returns false otherwise
inline bool try_ewise_vec(const Ewise& p, bool uses_b, cudaStream_t s) {
  if (!(vec_ok(p.a) && vec_ok(p.out) && (!uses_b || vec_ok(p.b)))) return false;
  const size_t total = static_cast<size_t>(p.out.n) * p.out.h * p.out.w * (p.out.c / 8);
  ewise_vec_kernel<<<nblocks(total, 256), 256, 0, s>>>(p);
  return true;
}

// ------------------------------------------------------------------------------------------------ loss
struct LossParams {
  View pred, target, dpred;
  int kind;          // 0 DIFFERENCE 1 ABSOLUTE 2 SMOOTH_ABSOLUTE 3 SQUARED 4 SMAPE
  float weight;      // loss_weight * scale_factor / (N*h*w): d(loss)/d(channel-summed difference of one pixel)
  float epsilon;
  int accumulate;    // dpred += instead of =
  float* loss;       // scalar accumulator (atomicAdd of weight * sum)
};

__device__ __forceinline__ void loss_elem(int kind, float p, float t, float eps, float& val, float& grad) {
  const float d = p - t;
  switch (kind) {
    case 0: val = d; grad = 1.f; break;
    case 1: val = fabsf(d); grad = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); break;
    case 2: {
      const float a = fabsf(d);
      if (a < 1.f) { val = 0.5f * a * a; grad = d; } else { val = a - 0.5f; grad = (d > 0.f) ? 1.f : -1.f; }
      break;
    }
    case 3: val = d * d; grad = 2.f * d; break;
    default: {  // SMAPE: |p-t| / (|p| + |t| + eps)
      const float a = fabsf(d), den = fabsf(p) + fabsf(t) + eps;
      val = a / den;
      const float sd = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
      const float sp = (p > 0.f) ? 1.f : ((p < 0.f) ? -1.f : 0.f);
      // d/dp = sd/den - a sp/den^2.  Written without den^2 (overflows fp32 from |p| ~ 1.8e19: direct predictions reach
      // exp(46)) and, where p lies beyond t on its own side (sd == sp), without the cancellation den - a:
      // den - a = |t| + sp t + eps exactly.
      if (sd == sp && sd != 0.f) grad = sp * ((fabsf(t) + sp * t + eps) / den) / den;
      else grad = (sd - val * sp) / den;
      break;
    }
  }
}

__global__ void __launch_bounds__(256) loss_kernel(const LossParams p) {
  const size_t total = static_cast<size_t>(p.pred.n) * p.pred.h * p.pred.w;
  const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float sum = 0.f;
  if (pix < total) {
    for (int c = 0; c < p.pred.c; ++c) {
      float v, g;
      loss_elem(p.kind, p.pred.load(pix, c), p.target.load(pix, c), p.epsilon, v, g);
      sum += v;
      if (p.dpred.ptr) {
        const float prev = p.accumulate ? p.dpred.load(pix, c) : 0.f;
        p.dpred.store(pix, c, prev + p.weight * g);
      }
    }
  }
  // block reduction -> one atomic per block
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(p.loss, p.weight * v);
  }
}

// ------------------------------------------------------------------------------------------------ variation / masked losses
// BaseFeatureTraining.variation_mean (Training.py:139-186, 304-346): LossDifference.difference of the horizontal
// (x[:, :, 1:] - x[:, :, :-1]) and vertical forward differences of prediction and target, mean over both sets.  One thread
// per pixel; the gradient is GATHERED (each pixel takes part in up to four difference terms), so no atomics.
struct VarLossParams { View pred, target, dpred; int kind; float weight, epsilon; float* loss; };
__global__ void __launch_bounds__(256) variation_loss_kernel(const VarLossParams p) {
  const int h = p.pred.h, w = p.pred.w;
  const size_t total = static_cast<size_t>(p.pred.n) * h * w;
  const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float sum = 0.f;
  if (pix < total) {
    const int x = static_cast<int>(pix % w), y = static_cast<int>((pix / w) % h);
    for (int c = 0; c < p.pred.c; ++c) {
      const float pc = p.pred.load(pix, c), tc = p.target.load(pix, c);
      float g = 0.f, v, gr;
      if (x + 1 < w) { loss_elem(p.kind, p.pred.load(pix + 1, c) - pc, p.target.load(pix + 1, c) - tc, p.epsilon, v, gr); sum += v; g -= gr; }
      if (x > 0) { loss_elem(p.kind, pc - p.pred.load(pix - 1, c), tc - p.target.load(pix - 1, c), p.epsilon, v, gr); g += gr; }
      if (y + 1 < h) { loss_elem(p.kind, p.pred.load(pix + w, c) - pc, p.target.load(pix + w, c) - tc, p.epsilon, v, gr); sum += v; g -= gr; }
      if (y > 0) { loss_elem(p.kind, pc - p.pred.load(pix - w, c), tc - p.target.load(pix - w, c), p.epsilon, v, gr); g += gr; }
      if (p.dpred.ptr) p.dpred.store(pix, c, p.dpred.load(pix, c) + p.weight * g);
    }
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(p.loss, p.weight * v);
  }
}

// BaseFeatureTraining.masked_mean (Training.py:131-137): sum(difference * mask) / sum(mask) (0 when the mask is empty),
// mask = Conv2dUtilities.non_zero_mask of the corresponding colour target (Conv2dUtilities.py:69-74: sign(sum_c |x|)).
struct MaskedLossParams { View pred, target, mask, dpred; int kind; float weight, epsilon; float* loss; const float* mask_sum; };
__global__ void __launch_bounds__(256) mask_sum_kernel(const View mask, float* out) {
  const size_t total = static_cast<size_t>(mask.n) * mask.h * mask.w;
  const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float m = 0.f;
  if (pix < total) {
    float a = 0.f;
    for (int c = 0; c < mask.c; ++c) a += fabsf(mask.load(pix, c));
    m = (a > 0.f) ? 1.f : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
  if ((threadIdx.x & 31) == 0 && m != 0.f) atomicAdd(out, m);
}
__global__ void __launch_bounds__(256) masked_loss_kernel(const MaskedLossParams p) {
  const size_t total = static_cast<size_t>(p.pred.n) * p.pred.h * p.pred.w;
  const size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const float ms = *p.mask_sum;
  const float wgt = (ms > 0.f) ? p.weight / ms : 0.f;
  float sum = 0.f;
  if (pix < total && wgt != 0.f) {
    float a = 0.f;
    for (int c = 0; c < p.mask.c; ++c) a += fabsf(p.mask.load(pix, c));
    if (a > 0.f) {
      for (int c = 0; c < p.pred.c; ++c) {
        float v, g;
        loss_elem(p.kind, p.pred.load(pix, c), p.target.load(pix, c), p.epsilon, v, g);
        sum += v;
        if (p.dpred.ptr) p.dpred.store(pix, c, p.dpred.load(pix, c) + wgt * g);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0 && sum != 0.f) atomicAdd(p.loss, wgt * sum);
}

// ------------------------------------------------------------------------------------------------ conv wgrad (exact)
// dW[r,s,c,o] += sum_{n,y,x} x[n, y+r-pad, x+s-pad, c] * dz[n,y,x,o]      (TF layout [kh,kw,cin,cout])
// db[o]       += sum dz[n,y,x,o]
// transposed 2x2 (ups == 2): dz pixel of (x-grid pixel (i,j), sub-pixel (ay,ax)) is (2i+ay, 2j+ax); layout [a,b,cout,cin]
// One block = one 8x16 pixel tile; thread (c_i, o_i) owns a CB x OB patch of (cin, cout) pairs for all taps.
constexpr int kWgTileH = 8, kWgTileW = 16, kWgThreads = 256, kWgMaxC = 32;
struct WgradParams {
  View x, dz;
  float* dw; float* db;
  int ksize, cin, cout, c0, o0;    // this launch covers cin [c0, c0+32) x cout [o0, o0+32)
  int ups, ay, ax, transposed_layout;
  int tiles_x, tiles_y;
};

__global__ void __launch_bounds__(kWgThreads) wgrad_kernel(const WgradParams p) {
  extern __shared__ float sm[];
  const int pad = (p.ksize - 1) / 2;
  const int TH = kWgTileH + 2 * pad, TW = kWgTileW + 2 * pad;
  float* sx = sm;                                   // [TH][TW][32]
  float* sdz = sm + TH * TW * kWgMaxC;              // [8][16][32]
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kWgTileH, x0 = blockIdx.x * kWgTileW;
  const int cin_n = min(kWgMaxC, p.cin - p.c0), cout_n = min(kWgMaxC, p.cout - p.o0);
  for (int i = threadIdx.x; i < TH * TW * kWgMaxC; i += kWgThreads) {
    const int c = i % kWgMaxC, px = (i / kWgMaxC) % TW, py = i / (kWgMaxC * TW);
    const int yy = y0 + py - pad, xx = x0 + px - pad;
    float v = 0.f;
    if (c < cin_n && yy >= 0 && yy < p.x.h && xx >= 0 && xx < p.x.w) v = p.x.load(p.x.pix(n, yy, xx), p.c0 + c);
    sx[i] = v;
  }
  for (int i = threadIdx.x; i < kWgTileH * kWgTileW * kWgMaxC; i += kWgThreads) {
    const int o = i % kWgMaxC, px = (i / kWgMaxC) % kWgTileW, py = i / (kWgMaxC * kWgTileW);
    const int yy = y0 + py, xx = x0 + px;
    float v = 0.f;
    if (o < cout_n && yy < p.x.h && xx < p.x.w) {
      const int zy = (p.ups == 2) ? 2 * yy + p.ay : yy, zx = (p.ups == 2) ? 2 * xx + p.ax : xx;
      if (zy < p.dz.h && zx < p.dz.w) v = p.dz.load(p.dz.pix(n, zy, zx), p.o0 + o);   // 3x3 stride-2 taps may leave the image
    }
    sdz[i] = v;
  }
  __syncthreads();
  // thread -> (c, o-quad): 32 c x 8 groups of 4 o
  const int c = threadIdx.x & 31, og = threadIdx.x >> 5;
  for (int r = 0; r < p.ksize; ++r) {
    for (int s = 0; s < p.ksize; ++s) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int py = 0; py < kWgTileH; ++py) {
        for (int px = 0; px < kWgTileW; ++px) {
          const float xv = sx[((py + r) * TW + px + s) * kWgMaxC + c];
          const float4 dv = *reinterpret_cast<const float4*>(&sdz[(py * kWgTileW + px) * kWgMaxC + og * 4]);
          acc[0] = fmaf(xv, dv.x, acc[0]); acc[1] = fmaf(xv, dv.y, acc[1]);
          acc[2] = fmaf(xv, dv.z, acc[2]); acc[3] = fmaf(xv, dv.w, acc[3]);
        }
      }
      if (c < cin_n) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int o = og * 4 + i;
          if (o >= cout_n) continue;
          const size_t tap = static_cast<size_t>(r) * p.ksize + s;
          const size_t idx = p.transposed_layout ? (static_cast<size_t>(p.o0 + o)) * p.cin + p.c0 + c   // [cout][cin] of one tap
                                                 : (tap * p.cin + p.c0 + c) * p.cout + p.o0 + o;
          atomicAdd(p.dw + idx, acc[i]);
        }
      }
    }
  }
  if (p.db && p.c0 == 0 && threadIdx.x < kWgMaxC && threadIdx.x < cout_n) {
    float s = 0.f;
    for (int i = 0; i < kWgTileH * kWgTileW; ++i) s += sdz[i * kWgMaxC + threadIdx.x];
    atomicAdd(p.db + p.o0 + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------------------------------------ max-pool backward
struct PoolBwdParams { View x, y, dy, dx; int ksize, pad_y, pad_x; };
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const PoolBwdParams p) {
  // output-centric: route dy to the FIRST maximum of the window (TF MaxPoolGrad semantics), atomics because 3x3/s2 windows overlap
  const size_t total = static_cast<size_t>(p.y.n) * p.y.h * p.y.w * p.y.c;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % p.y.c);
  const size_t opix = idx / p.y.c;
  const int ox = static_cast<int>(opix % p.y.w);
  const int oy = static_cast<int>((opix / p.y.w) % p.y.h);
  const int n = static_cast<int>(opix / (static_cast<size_t>(p.y.w) * p.y.h));
  float m = -INFINITY; size_t arg = 0; bool found = false;
  for (int r = 0; r < p.ksize; ++r) {
    const int yy = 2 * oy - p.pad_y + r;
    if (yy < 0 || yy >= p.x.h) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int xx = 2 * ox - p.pad_x + s;
      if (xx < 0 || xx >= p.x.w) continue;
      const size_t ip = p.x.pix(n, yy, xx);
      const float v = p.x.load(ip, c);
      if (!found || v > m) { m = v; arg = ip; found = true; }
    }
  }
  if (found) {
    float* dst = reinterpret_cast<float*>(p.dx.ptr) + arg * p.dx.cstride + p.dx.coff + c;
    atomicAdd(dst, p.dy.load(opix, c));
  }
}

// Gather form for 16-bit tensors (the tensor-core training path): one thread per (INPUT pixel, 8 channels) visits the <= 4
// windows that contain the pixel, and takes a window's gradient when the pixel is the window's FIRST maximum (TF MaxPoolGrad;
// decided by comparing with the stored pooled value and, on a match, with the window elements that precede the pixel).  No
// atomics, no fp32 scratch tensor, no zero fill, and the result is ADDED to the 16-bit gradient in place - the scatter form
// above moved ~7 GB per full-resolution level of a cfg5 step for 1.1 GB of activations.
__global__ void __launch_bounds__(256) maxpool_bwd_gather_kernel(const PoolBwdParams p) {
  const int cv = p.x.c / 8;
  const size_t total = static_cast<size_t>(p.x.n) * p.x.h * p.x.w * cv;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % cv) * 8;
  const size_t ipix = idx / cv;
  const int xx = static_cast<int>(ipix % p.x.w);
  const int yy = static_cast<int>((ipix / p.x.w) % p.x.h);
  const int n = static_cast<int>(ipix / (static_cast<size_t>(p.x.w) * p.x.h));
  const int bf = p.x.bf16, k = p.ksize;
  const uint16_t* xin = reinterpret_cast<const uint16_t*>(p.x.ptr);
  const uint16_t* yin = reinterpret_cast<const uint16_t*>(p.y.ptr);
  const uint16_t* gin = reinterpret_cast<const uint16_t*>(p.dy.ptr);
  float xv[8], acc[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(xin + ipix * p.x.cstride + p.x.coff + c)), bf, xv);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  // windows oy with 2 oy - pad <= yy <= 2 oy - pad + k - 1
  int oy_lo = yy + p.pad_y - k + 1; oy_lo = oy_lo <= 0 ? 0 : (oy_lo + 1) >> 1;
  int oy_hi = (yy + p.pad_y) >> 1; if (oy_hi > p.y.h - 1) oy_hi = p.y.h - 1;
  int ox_lo = xx + p.pad_x - k + 1; ox_lo = ox_lo <= 0 ? 0 : (ox_lo + 1) >> 1;
  int ox_hi = (xx + p.pad_x) >> 1; if (ox_hi > p.y.w - 1) ox_hi = p.y.w - 1;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const size_t op = p.y.pix(n, oy, ox);
      float m[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(yin + op * p.y.cstride + p.y.coff + c)), bf, m);
      unsigned eq = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) eq |= (xv[i] == m[i]) ? (1u << i) : 0u;
      if (!eq) continue;
      // earlier elements of the window (row-major scan) that already hold the maximum take the gradient instead
      const int y0 = 2 * oy - p.pad_y, x0 = 2 * ox - p.pad_x;
      bool reached = false;
      for (int r = 0; r < k && !reached; ++r) {
        const int y2 = y0 + r;
        if (y2 < 0 || y2 >= p.x.h) continue;
        for (int s = 0; s < k; ++s) {
          const int x2 = x0 + s;
          if (x2 < 0 || x2 >= p.x.w) continue;
          if (y2 == yy && x2 == xx) { reached = true; break; }
          float e[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(xin + p.x.pix(n, y2, x2) * p.x.cstride + p.x.coff + c)), bf, e);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (e[i] == m[i]) eq &= ~(1u << i);
        }
      }
      if (!eq) continue;
      float g[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gin + op * p.dy.cstride + p.dy.coff + c)), bf, g);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (eq & (1u << i)) acc[i] += g[i];
    }
  }
  uint16_t* dst = reinterpret_cast<uint16_t*>(p.dx.ptr) + ipix * p.dx.cstride + p.dx.coff + c;
  float d[8];
  unpack8(*reinterpret_cast<const uint4*>(dst), bf, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] += acc[i];
  *reinterpret_cast<uint4*>(dst) = pack8(d, bf);
}

// Index form: the forward pass (dd_maxpool_s2_fwd_index) recorded the first maximum of every window, so an input pixel only
// compares its position inside each of its <= 4 windows with the recorded byte: 8 + 16 bytes per window instead of up to nine
// 16-byte activation loads (the gather form above spent 3.2 ms per full-resolution level of a cfg5 step on its tie-break scans).
__global__ void __launch_bounds__(256) maxpool_bwd_index_kernel(const PoolBwdParams p, const uint8_t* __restrict__ index) {
  const int cv = p.dx.c / 8;
  const size_t total = static_cast<size_t>(p.dx.n) * p.dx.h * p.dx.w * cv;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % cv) * 8;
  const size_t ipix = idx / cv;
  const int xx = static_cast<int>(ipix % p.dx.w);
  const int yy = static_cast<int>((ipix / p.dx.w) % p.dx.h);
  const int n = static_cast<int>(ipix / (static_cast<size_t>(p.dx.w) * p.dx.h));
  const int bf = p.dx.bf16, k = p.ksize;
  const uint16_t* gin = reinterpret_cast<const uint16_t*>(p.dy.ptr);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  int oy_lo = yy + p.pad_y - k + 1; oy_lo = oy_lo <= 0 ? 0 : (oy_lo + 1) >> 1;
  int oy_hi = (yy + p.pad_y) >> 1; if (oy_hi > p.dy.h - 1) oy_hi = p.dy.h - 1;
  int ox_lo = xx + p.pad_x - k + 1; ox_lo = ox_lo <= 0 ? 0 : (ox_lo + 1) >> 1;
  int ox_hi = (xx + p.pad_x) >> 1; if (ox_hi > p.dy.w - 1) ox_hi = p.dy.w - 1;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const size_t op = p.dy.pix(n, oy, ox);
      const uint32_t code = static_cast<uint32_t>((yy - (2 * oy - p.pad_y)) * k + (xx - (2 * ox - p.pad_x)));
      const uint2 a = __ldg(reinterpret_cast<const uint2*>(index + op * p.dy.c + c));
      const uint32_t want = code * 0x01010101u;
      const uint32_t e0 = a.x ^ want, e1 = a.y ^ want;               // a zero byte = this pixel is the window's first maximum
      if ((((e0 - 0x01010101u) & ~e0) | ((e1 - 0x01010101u) & ~e1)) & 0x80808080u) {
        float g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(gin + op * p.dy.cstride + p.dy.coff + c)), bf, g);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (((e0 >> (8 * i)) & 255u) == 0u) acc[i] += g[i];
          if (((e1 >> (8 * i)) & 255u) == 0u) acc[4 + i] += g[4 + i];
        }
      }
    }
  }
  uint16_t* dst = reinterpret_cast<uint16_t*>(p.dx.ptr) + ipix * p.dx.cstride + p.dx.coff + c;
  float d[8];
  unpack8(*reinterpret_cast<const uint4*>(dst), bf, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] += acc[i];
  *reinterpret_cast<uint4*>(dst) = pack8(d, bf);
}

// ------------------------------------------------------------------------------------------------ kernel prediction backward
// dlogits_k = p_k (G_k - sum_c g_c out_c),  G_k = sum_c g_c S_c[k],  p = softmax(logits)        (no gradient to the source: it is data)
struct KpBwdParams { View src, logits, dout, dlogits; int K, F, ipt; };
__device__ __forceinline__ int sym_idx(int i, int n) {
  while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - i - 1);
  return i;
}
// LANES threads per (pixel, feature), lanes stride over the K*K taps, so the logit loads and the gradient stores of a pixel
// are contiguous (one thread per pixel walks K*K channels: at K = 21 every access of a warp touches 32 different cache
// lines - measured 15 ms per Tiramisu training step).  LANES = 8 for small kernels, 32 from K*K >= 64.
template <int LANES>
__global__ void __launch_bounds__(256) kernel_predict_bwd_coop_kernel(const KpBwdParams p) {
  const size_t per_img = static_cast<size_t>(p.src.h) * p.src.w;
  const size_t total = static_cast<size_t>(p.logits.n) * p.F * per_img;
  const size_t idx = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LANES;
  const int sub = threadIdx.x % LANES;
  const bool valid = idx < total;                    // uniform across the LANES of a pixel
  const int K = p.K, K2 = K * K, pad = (K - 1) / 2;
  // consecutive groups -> features of one pixel first (their logits are adjacent), then pixels
  const int f = valid ? static_cast<int>(idx % p.F) : 0;
  const size_t pb = valid ? idx / p.F : 0;
  const size_t pi = pb % per_img;
  const int b = static_cast<int>(pb / per_img);
  const int img = ((b / p.ipt) * p.F + f) * p.ipt + (b % p.ipt);
  const int y = static_cast<int>(pi / p.src.w), x = static_cast<int>(pi % p.src.w);
  const size_t lpix = valid ? p.logits.pix(b, y, x) : 0;
  const size_t opix = valid ? p.src.pix(img, y, x) : 0;
  const int coff = f * K2;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (valid) { g0 = p.dout.load(opix, 0); g1 = p.dout.load(opix, 1); g2 = p.dout.load(opix, 2); }
  float mx = -INFINITY;
  if (valid) for (int t = sub; t < K2; t += LANES) mx = fmaxf(mx, p.logits.load(lpix, coff + t));
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f, dot = 0.f;
  if (valid) {
    for (int t = sub; t < K2; t += LANES) {
      const float e = __expf(p.logits.load(lpix, coff + t) - mx);
      const int i = t / K, j = t - i * K;
      const size_t sp = p.src.pix(img, sym_idx(y + i - pad, p.src.h), sym_idx(x + j - pad, p.src.w));
      const float G = g0 * p.src.load(sp, 0) + g1 * p.src.load(sp, 1) + g2 * p.src.load(sp, 2);
      sum += e; dot += e * G;
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  if (!valid) return;
  const float inv = 1.f / sum;
  dot *= inv;
  for (int t = sub; t < K2; t += LANES) {
    const float pk = __expf(p.logits.load(lpix, coff + t) - mx) * inv;
    const int i = t / K, j = t - i * K;
    const size_t sp = p.src.pix(img, sym_idx(y + i - pad, p.src.h), sym_idx(x + j - pad, p.src.w));
    const float G = g0 * p.src.load(sp, 0) + g1 * p.src.load(sp, 1) + g2 * p.src.load(sp, 2);
    p.dlogits.store(lpix, coff + t, pk * (G - dot));
  }
}

// Tiled form for one feature per tuple and K in {3, 5} (the U-Net KPCN configurations): a block owns a 32 x 8 pixel tile, stages
// the symmetric-padded source halo in shared memory once, and ONE thread per pixel keeps the K*K logits in registers (16-byte
// loads of its row of the fp32 logits tensor), so every logit is read once and exponentiated once.  The cooperative kernel above
// reads each logit three times and recomputes the tap addresses twice: 3.3 ms per full-resolution launch of a cfg5 step for
// 1.7 GB of traffic.
template <int K>
__global__ void __launch_bounds__(256) kernel_predict_bwd_tile_kernel(const KpBwdParams p) {
  constexpr int K2 = K * K, PAD = (K - 1) / 2, TW = 32, TH = 8, HW = TW + 2 * PAD, HH = TH + 2 * PAD, NV = (K2 + 3) / 4;
  __shared__ float s_src[HH * HW * 3];
  const int h = p.src.h, w = p.src.w;
  const int b = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const float* src = reinterpret_cast<const float*>(p.src.ptr);
  for (int i = threadIdx.x; i < HH * HW; i += 256) {
    const int ly = i / HW, lx = i - ly * HW;
    const size_t sp = p.src.pix(b, sym_idx(y0 + ly - PAD, h), sym_idx(x0 + lx - PAD, w)) * p.src.cstride + p.src.coff;
    s_src[3 * i] = __ldg(src + sp); s_src[3 * i + 1] = __ldg(src + sp + 1); s_src[3 * i + 2] = __ldg(src + sp + 2);
  }
  __syncthreads();
  const int ly = threadIdx.x >> 5, lx = threadIdx.x & 31;
  const int y = y0 + ly, x = x0 + lx;
  if (y >= h || x >= w) return;
  const size_t pix = p.src.pix(b, y, x);
  const float* gp = reinterpret_cast<const float*>(p.dout.ptr) + pix * p.dout.cstride + p.dout.coff;
  const float g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
  float l[NV * 4];
  const float4* lp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.logits.ptr) + pix * p.logits.cstride + p.logits.coff);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 v = __ldg(lp + i);
    l[4 * i] = v.x; l[4 * i + 1] = v.y; l[4 * i + 2] = v.z; l[4 * i + 3] = v.w;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < K2; ++t) mx = fmaxf(mx, l[t]);
  float G[K2];
  float sum = 0.f, dot = 0.f;
#pragma unroll
  for (int t = 0; t < K2; ++t) {
    const int i = t / K, j = t % K;
    const float* sv = s_src + ((ly + i) * HW + lx + j) * 3;
    G[t] = g0 * sv[0] + g1 * sv[1] + g2 * sv[2];
    l[t] = __expf(l[t] - mx);
    sum += l[t]; dot += l[t] * G[t];
  }
  const float inv = 1.f / sum;
  dot *= inv;
  float out[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) out[t] = (t < K2) ? l[t] * inv * (G[t] - dot) : 0.f;
  if (p.dlogits.f16 || p.dlogits.bf16) {
    // 16-bit gradient rows padded to a multiple of 8 channels (the padding is never read: tensor maps clip at c)
    uint16_t* dp = reinterpret_cast<uint16_t*>(p.dlogits.ptr) + pix * p.dlogits.cstride + p.dlogits.coff;
#pragma unroll
    for (int v = 0; v < (K2 + 7) / 8; ++v) {
      float f8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f8[i] = out[8 * v + i];
      *reinterpret_cast<uint4*>(dp + 8 * v) = pack8(f8, p.dlogits.bf16);
    }
  } else {
    float* dp = reinterpret_cast<float*>(p.dlogits.ptr) + pix * p.dlogits.cstride + p.dlogits.coff;
#pragma unroll
    for (int t = 0; t < K2; ++t) dp[t] = out[t];
  }
}

// ------------------------------------------------------------------------------------------------ compose backward
constexpr int kCmpC = 32;
struct ComposeTailBwdParams {
  View t, small, large, dout, dt, dsmall, dlarge;   // dsmall / dlarge are ACCUMULATED (atomicAdd, fp32)
  float w[kCmpC]; float b; int c_mid;
  float* dw; float* db;                              // [c_mid], [1] accumulated
};
__global__ void __launch_bounds__(256) compose_tail_bwd_kernel(const ComposeTailBwdParams p) {
  // grid-stride over pixels: the parameter gradients are accumulated per thread and flushed once per warp at the end
  const size_t total = static_cast<size_t>(p.large.n) * p.large.h * p.large.w;
  float dwl[kCmpC];
  float dbl = 0.f;
#pragma unroll
  for (int c = 0; c < kCmpC; ++c) dwl[c] = 0.f;
  for (size_t pixel = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; pixel < total;
       pixel += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x0 = static_cast<int>(pixel % p.large.w);
    const int y0 = static_cast<int>((pixel / p.large.w) % p.large.h);
    const int n = static_cast<int>(pixel / (static_cast<size_t>(p.large.w) * p.large.h));
    float a = p.b;
    for (int c = 0; c < p.c_mid; ++c) a = fmaf(p.t.load(pixel, c), p.w[c], a);
    const float ar = fmaxf(a, 0.f);
    const float wgt = 1.f / (1.f + expf(-ar));
    const size_t spix = p.small.pix(n, y0 >> 1, x0 >> 1);
    const int yb = y0 & ~1, xb = x0 & ~1;
    float dwgt = 0.f;
    float* dl = reinterpret_cast<float*>(p.dlarge.ptr);
    float* ds = reinterpret_cast<float*>(p.dsmall.ptr);
    for (int c = 0; c < 3; ++c) {
      const float g = p.dout.load(pixel, c);
      const float low = 0.25f * (p.large.load(p.large.pix(n, yb, xb), c) + p.large.load(p.large.pix(n, yb, xb + 1), c) +
                                 p.large.load(p.large.pix(n, yb + 1, xb), c) + p.large.load(p.large.pix(n, yb + 1, xb + 1), c));
      const float su = p.small.load(spix, c);
      dwgt += g * (su - low);
      // out = large - wgt*low + wgt*su
      atomicAdd(dl + pixel * p.dlarge.cstride + p.dlarge.coff + c, g);
      const float dlow = -wgt * g * 0.25f;
      atomicAdd(dl + p.large.pix(n, yb, xb) * p.dlarge.cstride + p.dlarge.coff + c, dlow);
      atomicAdd(dl + p.large.pix(n, yb, xb + 1) * p.dlarge.cstride + p.dlarge.coff + c, dlow);
      atomicAdd(dl + p.large.pix(n, yb + 1, xb) * p.dlarge.cstride + p.dlarge.coff + c, dlow);
      atomicAdd(dl + p.large.pix(n, yb + 1, xb + 1) * p.dlarge.cstride + p.dlarge.coff + c, dlow);
      atomicAdd(ds + spix * p.dsmall.cstride + p.dsmall.coff + c, wgt * g);
    }
    // sigmoid'(ar) = e / (1 + e)^2 with e = exp(-ar): `wgt * (1 - wgt)` cancels catastrophically once the gate saturates (ar > ~10
    // in fp32), which direct predictions (no kernel prediction: unbounded inputs of the compose net) reach routinely
    const float en = expf(-ar);
    const float da = (a > 0.f) ? dwgt * en / ((1.f + en) * (1.f + en)) : 0.f;
#pragma unroll
    for (int c = 0; c < kCmpC; ++c) {
      if (c < p.c_mid) {
        const float tv = p.t.load(pixel, c);
        p.dt.store(pixel, c, da * p.w[c]);
        dwl[c] = fmaf(da, tv, dwl[c]);
      }
    }
    dbl += da;
  }
  __shared__ float s_acc[kCmpC + 1];
  for (int i = threadIdx.x; i <= kCmpC; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kCmpC; ++c) {
    float v = dwl[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && c < p.c_mid) atomicAdd(&s_acc[c], v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dbl += __shfl_xor_sync(0xffffffffu, dbl, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[kCmpC], dbl);
  __syncthreads();
  if (threadIdx.x < p.c_mid) atomicAdd(p.dw + threadIdx.x, s_acc[threadIdx.x]);
  if (threadIdx.x == 0) atomicAdd(p.db, s_acc[kCmpC]);
}

struct ComposeHeadBwdParams {
  View small, large, y, dy, dsmall, dlarge;    // y = relu output of the head (mask), dy its gradient; dsmall/dlarge accumulated
  float w[6 * kCmpC]; int c_mid;
  float* dw; float* db;                        // [6][c_mid], [c_mid] accumulated
};
__global__ void __launch_bounds__(256) compose_head_bwd_kernel(const ComposeHeadBwdParams p) {
  // grid-stride over 256-pixel groups; dW [6][c_mid] and db [c_mid] are reduced per warp into shared memory and flushed with
  // one global atomic per entry and block (every warp hitting the same 7 * c_mid global addresses serialised the old version)
  __shared__ float s_acc[7 * kCmpC];
  for (int i = threadIdx.x; i < 7 * kCmpC; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const size_t total = static_cast<size_t>(p.large.n) * p.large.h * p.large.w;
  for (size_t base = static_cast<size_t>(blockIdx.x) * blockDim.x; base < total; base += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pixel = base + threadIdx.x;
    const bool valid = pixel < total;
    float in[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float din[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    size_t spix = 0;
    if (valid) {
      const int x0 = static_cast<int>(pixel % p.large.w);
      const int y0 = static_cast<int>((pixel / p.large.w) % p.large.h);
      const int n = static_cast<int>(pixel / (static_cast<size_t>(p.large.w) * p.large.h));
      spix = p.small.pix(n, y0 >> 1, x0 >> 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) { in[c] = p.small.load(spix, c); in[3 + c] = p.large.load(pixel, c); }
    }
    for (int c = 0; c < p.c_mid; ++c) {
      float dz = 0.f;
      if (valid && p.y.load(pixel, c) > 0.f) dz = p.dy.load(pixel, c);
      float v[7];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        din[k] = fmaf(dz, p.w[k * p.c_mid + c], din[k]);
        v[k] = dz * in[k];
      }
      v[6] = dz;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      }
      if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) atomicAdd(&s_acc[k * kCmpC + c], v[k]);
      }
    }
    if (valid) {
      float* dl = reinterpret_cast<float*>(p.dlarge.ptr);
      float* ds = reinterpret_cast<float*>(p.dsmall.ptr);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        atomicAdd(ds + spix * p.dsmall.cstride + p.dsmall.coff + c, din[c]);
        atomicAdd(dl + pixel * p.dlarge.cstride + p.dlarge.coff + c, din[3 + c]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 7 * p.c_mid; i += blockDim.x) {
    const int k = i / p.c_mid, c = i % p.c_mid;
    if (k < 6) atomicAdd(p.dw + k * p.c_mid + c, s_acc[k * kCmpC + c]);
    else atomicAdd(p.db + c, s_acc[6 * kCmpC + c]);
  }
}

// ------------------------------------------------------------------------------------------------ reductions / optimizer
struct ChanSumParams { View x; float* out; int images_per_group; int groups; };
// out[g][c] += sum over the images of group g and all pixels of x[.., c]   (embedding-row gradient: SourceEncoder broadcast bwd)
__global__ void __launch_bounds__(256) channel_sum_kernel(const ChanSumParams p) {
  const size_t per_group = static_cast<size_t>(p.images_per_group) * p.x.h * p.x.w;
  const int g = blockIdx.y;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_group; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = static_cast<size_t>(g) * per_group + i;
    for (int c = 0; c < p.x.c; ++c) acc[c] += p.x.load(pix, c);
  }
  for (int c = 0; c < p.x.c; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(p.out + g * p.x.c + c, v);
  }
}

// tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t); theta -= lr_t m / (sqrt(v) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* w, const float* g, float* m, float* v, size_t n, float lr_t, float b1,
                                                   float b2, float eps, float gscale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

// ---- overflow-guarded optimizer step (fp16 training): no host synchronisation anywhere
// state (device, int32[4]): [0] steps skipped so far, [1] "this step saw a non-finite gradient", [2] steps applied,
//                           [3] bit pattern of the lr_t (float) of the step being applied
__global__ void __launch_bounds__(256) nonfinite_kernel(const float* g, size_t n, int* state) {
  bool bad = false;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = g[i];
    bad |= !(fabsf(v) <= 3.0e38f);        // inf and NaN both fail
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(state + 1, 1);
}
__global__ void adam_decide_kernel(int* state, float lr, float b1, float b2) {
  if (state[1]) {
    state[0] += 1;
  } else {
    const int t = ++state[2];
    const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), static_cast<double>(t))) /
                        (1.0 - pow(static_cast<double>(b1), static_cast<double>(t)));
    state[3] = __float_as_int(static_cast<float>(lr_t));
  }
}
__global__ void __launch_bounds__(256) adam_guarded_kernel(float* w, const float* g, float* m, float* v, size_t n, const int* state,
                                                           float b1, float b2, float eps, float gscale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || state[1]) return;
  const float lr_t = __int_as_float(state[3]);
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

// device-side weight repack after an optimizer step (fp32 exact path): TF [kh,kw,cin,cout] -> [tap][cout][cin] (forward)
// and [flipped tap][cin][cout] (input-gradient convolution: a conv cout -> cin with the spatially flipped kernel)
__global__ void __launch_bounds__(256) repack_f32_kernel(const float* w, float* fwd, float* bwd, int k2, int cin, int cout,
                                                         int transposed) {
  const size_t total = static_cast<size_t>(k2) * cin * cout;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (transposed) {   // TF conv2d_transpose layout [sub-pixel][cout][cin]: forward uses it as is, dgrad needs [sub-pixel][cin][cout]
    const int c = static_cast<int>(i % cin);
    const int o = static_cast<int>((i / cin) % cout);
    const int sp = static_cast<int>(i / (static_cast<size_t>(cout) * cin));
    if (fwd) fwd[i] = w[i];
    if (bwd) bwd[(static_cast<size_t>(sp) * cin + c) * cout + o] = w[i];
    return;
  }
  const int o = static_cast<int>(i % cout);
  const int c = static_cast<int>((i / cout) % cin);
  const int t = static_cast<int>(i / (static_cast<size_t>(cout) * cin));
  const float v = w[i];
  if (fwd) fwd[(static_cast<size_t>(t) * cout + o) * cin + c] = v;
  if (bwd) bwd[(static_cast<size_t>(k2 - 1 - t) * cin + c) * cout + o] = v;
}

}  // namespace dd

using namespace dd;

extern "C" {

int dd_relu_bwd(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* dz, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(dy) && tensor_ok(y) && tensor_ok(dz) && same_dims(dy, y) && same_dims(dy, dz), "bad argument");
  Ewise p; memset(&p, 0, sizeof(p)); EwiseInv q; memset(&q, 0, sizeof(q));
  p.a = make_view(dy); p.b = make_view(y); p.out = make_view(dz); p.op = EW_RELU_MASK;
  const size_t total = static_cast<size_t>(dy->n) * dy->h * dy->w * dy->c;
  if (!try_ewise_vec(p, true, static_cast<cudaStream_t>(stream)) && !try_ewise_flat(p, q, 3, static_cast<cudaStream_t>(stream)))
    ewise_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, q);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_relu_bwd_acc(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* dz_acc, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(dy) && tensor_ok(y) && tensor_ok(dz_acc) && same_dims(dy, y) && same_dims(dy, dz_acc), "bad argument");
  Ewise p; memset(&p, 0, sizeof(p)); EwiseInv q; memset(&q, 0, sizeof(q));
  p.a = make_view(dy); p.b = make_view(y); p.out = make_view(dz_acc); p.op = EW_RELU_MASK_ACC;
  const size_t total = static_cast<size_t>(dy->n) * dy->h * dy->w * dy->c;
  if (!try_ewise_vec(p, true, static_cast<cudaStream_t>(stream)))
    ewise_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, q);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_muladd_fwd(dd_ctx* ctx, const dd_tensor* a, const dd_tensor* b, const dd_tensor* c, const dd_tensor* out, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(a) && tensor_ok(b) && tensor_ok(c) && tensor_ok(out) && same_dims(a, b) && same_dims(a, c) &&
                   same_dims(a, out), "bad argument");
  Ewise p; memset(&p, 0, sizeof(p)); EwiseInv q; memset(&q, 0, sizeof(q));
  p.a = make_view(a); p.b = make_view(b); p.c = make_view(c); p.out = make_view(out); p.op = EW_MULADD;
  const size_t total = static_cast<size_t>(a->n) * a->h * a->w * a->c;
  if (!try_ewise_flat(p, q, 7, static_cast<cudaStream_t>(stream)))
    ewise_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, q);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_muladd_bwd(dd_ctx* ctx, const dd_tensor* a, const dd_tensor* b, const dd_tensor* c, const dd_tensor* g,
                  const dd_tensor* da_acc, const dd_tensor* dbc_inc, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(a) && tensor_ok(b) && tensor_ok(c) && tensor_ok(g) && tensor_ok(da_acc) && tensor_ok(dbc_inc) &&
                   same_dims(a, b) && same_dims(a, c) && same_dims(a, g) && same_dims(a, da_acc) && same_dims(a, dbc_inc),
               "bad argument");
  Ewise p; memset(&p, 0, sizeof(p)); EwiseInv q; memset(&q, 0, sizeof(q));
  p.a = make_view(a); p.b = make_view(b); p.c = make_view(c); p.out = make_view(g); p.out2 = make_view(da_acc);
  p.out3 = make_view(dbc_inc); p.op = EW_MULADD_BWD;
  const size_t total = static_cast<size_t>(a->n) * a->h * a->w * a->c;
  if (!try_ewise_flat(p, q, 31, static_cast<cudaStream_t>(stream)))
    ewise_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, q);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_axpy(dd_ctx* ctx, float alpha, const dd_tensor* x, const dd_tensor* y, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y) && same_dims(x, y), "bad argument");
  Ewise p; memset(&p, 0, sizeof(p)); EwiseInv q; memset(&q, 0, sizeof(q));
  p.a = make_view(x); p.out = make_view(y); p.op = EW_AXPY; p.alpha = alpha;
  const size_t total = static_cast<size_t>(x->n) * x->h * x->w * x->c;
  if (!try_ewise_vec(p, false, static_cast<cudaStream_t>(stream)) && !try_ewise_flat(p, q, 1, static_cast<cudaStream_t>(stream)))
    ewise_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, q);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_fill(dd_ctx* ctx, float value, const dd_tensor* y, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(y), "bad argument");
  Ewise p; memset(&p, 0, sizeof(p)); EwiseInv q; memset(&q, 0, sizeof(q));
  p.out = make_view(y); p.op = EW_FILL; p.alpha = value;
  const size_t total = static_cast<size_t>(y->n) * y->h * y->w * y->c;
  if (!try_ewise_flat(p, q, 0, static_cast<cudaStream_t>(stream)))
    ewise_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, q);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_invert_standardization_bwd(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* x, const dd_invert_params* inv,
                                  const dd_tensor* dx, void* stream) {
  DD_CHECK_ARG(ctx && inv && tensor_ok(dy) && tensor_ok(x) && tensor_ok(dx) && same_dims(dy, x) && same_dims(dy, dx), "bad argument");
  Ewise p; memset(&p, 0, sizeof(p)); EwiseInv q;
  q.use_log1p = inv->use_log1p; q.mean = inv->mean; q.variance = inv->variance; q.sqrt_var = sqrtf(inv->variance);
  p.a = make_view(dy); p.b = make_view(x); p.out = make_view(dx); p.op = EW_INVERT_BWD;
  const size_t total = static_cast<size_t>(dy->n) * dy->h * dy->w * dy->c;
  if (!try_ewise_flat(p, q, 3, static_cast<cudaStream_t>(stream)))
    ewise_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, q);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_loss_fwd_bwd(dd_ctx* ctx, const dd_tensor* pred, const dd_tensor* target, int kind, float weight, float epsilon,
                    float* loss_dev, const dd_tensor* dpred, int accumulate, void* stream) {
  DD_CHECK_ARG(ctx && loss_dev && tensor_ok(pred) && tensor_ok(target) && same_dims(pred, target), "bad argument");
  DD_CHECK_ARG(kind >= 0 && kind <= 4, "unknown loss difference %d", kind);
  DD_CHECK_ARG(!dpred || (tensor_ok(dpred) && same_dims(pred, dpred)), "bad dpred");
  LossParams p; memset(&p, 0, sizeof(p));
  p.pred = make_view(pred); p.target = make_view(target);
  if (dpred) p.dpred = make_view(dpred);
  p.kind = kind; p.weight = weight; p.epsilon = epsilon; p.accumulate = accumulate; p.loss = loss_dev;
  const size_t total = static_cast<size_t>(pred->n) * pred->h * pred->w;
  loss_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_loss_variation_fwd_bwd(dd_ctx* ctx, const dd_tensor* pred, const dd_tensor* target, int kind, float weight, float epsilon,
                              float* loss_dev, const dd_tensor* dpred_acc, void* stream) {
  DD_CHECK_ARG(ctx && loss_dev && tensor_ok(pred) && tensor_ok(target) && same_dims(pred, target), "bad argument");
  DD_CHECK_ARG(!dpred_acc || (tensor_ok(dpred_acc) && same_dims(pred, dpred_acc)), "bad gradient tensor");
  DD_CHECK_ARG(kind >= 0 && kind <= 4, "unknown loss difference");
  VarLossParams p; memset(&p, 0, sizeof(p));
  p.pred = make_view(pred); p.target = make_view(target);
  if (dpred_acc) p.dpred = make_view(dpred_acc);
  p.kind = kind; p.weight = weight; p.epsilon = epsilon; p.loss = loss_dev;
  const size_t total = static_cast<size_t>(pred->n) * pred->h * pred->w;
  variation_loss_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_mask_sum(dd_ctx* ctx, const dd_tensor* mask_src, float* sum_dev, void* stream) {
  DD_CHECK_ARG(ctx && sum_dev && tensor_ok(mask_src), "bad argument");
  const size_t total = static_cast<size_t>(mask_src->n) * mask_src->h * mask_src->w;
  mask_sum_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(make_view(mask_src), sum_dev);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_loss_masked_fwd_bwd(dd_ctx* ctx, const dd_tensor* pred, const dd_tensor* target, const dd_tensor* mask_src,
                           const float* mask_sum_dev, int kind, float weight, float epsilon, float* loss_dev,
                           const dd_tensor* dpred_acc, void* stream) {
  DD_CHECK_ARG(ctx && loss_dev && mask_sum_dev && tensor_ok(pred) && tensor_ok(target) && tensor_ok(mask_src) &&
                   same_dims(pred, target), "bad argument");
  DD_CHECK_ARG(mask_src->n == pred->n && mask_src->h == pred->h && mask_src->w == pred->w, "mask dims differ");
  DD_CHECK_ARG(!dpred_acc || (tensor_ok(dpred_acc) && same_dims(pred, dpred_acc)), "bad gradient tensor");
  DD_CHECK_ARG(kind >= 0 && kind <= 4, "unknown loss difference");
  MaskedLossParams p; memset(&p, 0, sizeof(p));
  p.pred = make_view(pred); p.target = make_view(target); p.mask = make_view(mask_src);
  if (dpred_acc) p.dpred = make_view(dpred_acc);
  p.kind = kind; p.weight = weight; p.epsilon = epsilon; p.loss = loss_dev; p.mask_sum = mask_sum_dev;
  const size_t total = static_cast<size_t>(pred->n) * pred->h * pred->w;
  masked_loss_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_conv2d_wgrad(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* dz, int ksize, int transposed, float* dw, float* db,
                    void* stream) {
  DD_CHECK_ARG(ctx && dw && tensor_ok(x) && tensor_ok(dz), "bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int cin = x->c, cout = dz->c;
  if (x->dtype == DD_F16 && dz->dtype == DD_F16) {
    // mixed-precision path: tcgen05 wgrad (api_wgrad.cu); the bias gradient comes from dd_relu_bwd_bias
    DD_CHECK_ARG(!transposed && !db, "tensor-core wgrad: stride-1 convolutions only, bias gradient via dd_relu_bwd_bias");
    DD_CHECK_ARG(x->n == dz->n && x->h == dz->h && x->w == dz->w, "wgrad: spatial dims differ");
    return launch_wgrad_rows(ctx, x, dz, ksize, 0, dw, 1.f, s);
  }
  if (!transposed) {
    DD_CHECK_ARG(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
    DD_CHECK_ARG(x->n == dz->n && x->h == dz->h && x->w == dz->w, "wgrad: spatial dims differ");
  } else {
    DD_CHECK_ARG((ksize == 2 || ksize == 3) && dz->n == x->n && dz->h == 2 * x->h && dz->w == 2 * x->w,
                 "transposed wgrad: ksize 2 or 3, dz must be 2x");
  }
  const int kk = transposed ? 1 : ksize;
  const int pad = (kk - 1) / 2;
  const size_t smem = (static_cast<size_t>(kWgTileH + 2 * pad) * (kWgTileW + 2 * pad) + kWgTileH * kWgTileW) * kWgMaxC * sizeof(float);
  DD_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  WgradParams p; memset(&p, 0, sizeof(p));
  p.x = make_view(x); p.dz = make_view(dz); p.ksize = kk; p.cin = cin; p.cout = cout;
  p.tiles_x = (x->w + kWgTileW - 1) / kWgTileW; p.tiles_y = (x->h + kWgTileH - 1) / kWgTileH;
  dim3 grid(p.tiles_x, p.tiles_y, x->n);
  DD_CHECK_ARG(x->n <= 65535 && p.tiles_y <= 65535, "wgrad grid too large");
  // transposed: one launch per tap (a, b) of the k x k stride-2 kernel: dW[a,b,o,c] = sum x[i,j,c] dz[2i+a,2j+b,o]
  const int subs = transposed ? ksize * ksize : 1;
  for (int sp = 0; sp < subs; ++sp) {
    const int tap_a = transposed ? sp / ksize : 0, tap_b = transposed ? sp % ksize : 0;
    for (int c0 = 0; c0 < cin; c0 += kWgMaxC) {
      for (int o0 = 0; o0 < cout; o0 += kWgMaxC) {
        p.c0 = c0; p.o0 = o0;
        p.ups = transposed ? 2 : 1; p.ay = tap_a; p.ax = tap_b; p.transposed_layout = transposed;
        p.dw = transposed ? dw + static_cast<size_t>(sp) * cout * cin : dw;
        // bias gradient = sum over ALL dz pixels: taps (a, b) in {0,1}^2 visit every pixel exactly once
        p.db = (!transposed || (tap_a < 2 && tap_b < 2)) ? db : nullptr;
        wgrad_kernel<<<grid, kWgThreads, smem, s>>>(p);
        DD_LAUNCH_CHECK(ctx);
      }
    }
  }
  return DD_OK;
}

int dd_maxpool_s2_bwd(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* dy, int ksize, const dd_tensor* dx,
                      void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y) && tensor_ok(dy) && tensor_ok(dx) && same_dims(y, dy) && same_dims(x, dx),
               "bad argument");
  DD_CHECK_ARG(dx->dtype == DD_F32, "maxpool_bwd accumulates into an fp32 gradient");
  DD_CHECK_ARG(ksize == 2 || ksize == 3, "maxpool ksize must be 2 or 3");
  PoolBwdParams p;
  p.x = make_view(x); p.y = make_view(y); p.dy = make_view(dy); p.dx = make_view(dx); p.ksize = ksize;
  const int oh = (x->h + 1) / 2, ow = (x->w + 1) / 2;
  const int pty = (oh - 1) * 2 + ksize - x->h, ptx = (ow - 1) * 2 + ksize - x->w;
  p.pad_y = (pty > 0 ? pty : 0) / 2; p.pad_x = (ptx > 0 ? ptx : 0) / 2;
  const size_t total = static_cast<size_t>(y->n) * y->h * y->w * y->c;
  maxpool_bwd_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_maxpool_s2_bwd_acc(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* dy, int ksize, const dd_tensor* dx,
                          void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y) && tensor_ok(dy) && tensor_ok(dx) && same_dims(y, dy) && same_dims(x, dx),
               "bad argument");
  DD_CHECK_ARG(ksize == 2 || ksize == 3, "maxpool ksize must be 2 or 3");
  const dd_tensor* all[4] = {x, y, dy, dx};
  for (const dd_tensor* t : all)
    DD_CHECK_ARG(t->dtype == x->dtype && (t->dtype == DD_F16 || t->dtype == DD_BF16) && t->c % 8 == 0 && t->coff % 8 == 0 &&
                     t->cstride % 8 == 0, "maxpool_bwd_acc: fp16 / bf16 views with multiples of 8 channels expected");
  PoolBwdParams p;
  p.x = make_view(x); p.y = make_view(y); p.dy = make_view(dy); p.dx = make_view(dx); p.ksize = ksize;
  const int oh = (x->h + 1) / 2, ow = (x->w + 1) / 2;
  DD_CHECK_ARG(y->h == oh && y->w == ow && y->n == x->n && y->c == x->c, "maxpool_bwd_acc: pooled dims");
  const int pty = (oh - 1) * 2 + ksize - x->h, ptx = (ow - 1) * 2 + ksize - x->w;
  p.pad_y = (pty > 0 ? pty : 0) / 2; p.pad_x = (ptx > 0 ? ptx : 0) / 2;
  const size_t total = static_cast<size_t>(x->n) * x->h * x->w * (x->c / 8);
  maxpool_bwd_gather_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_maxpool_s2_bwd_index(dd_ctx* ctx, const uint8_t* index_dev, const dd_tensor* dy, int ksize, const dd_tensor* dx, void* stream) {
  DD_CHECK_ARG(ctx && index_dev && tensor_ok(dy) && tensor_ok(dx), "bad argument");
  DD_CHECK_ARG(ksize == 2 || ksize == 3, "maxpool ksize must be 2 or 3");
  const dd_tensor* both[2] = {dy, dx};
  for (const dd_tensor* t : both)
    DD_CHECK_ARG(t->dtype == dx->dtype && is_half_type(t->dtype) && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0,
                 "maxpool_bwd_index: fp16 / bf16 views with multiples of 8 channels expected");
  const int oh = (dx->h + 1) / 2, ow = (dx->w + 1) / 2;
  DD_CHECK_ARG(dy->h == oh && dy->w == ow && dy->n == dx->n && dy->c == dx->c, "maxpool_bwd_index: pooled dims");
  PoolBwdParams p;
  p.x = make_view(dx); p.y = make_view(dy); p.dy = make_view(dy); p.dx = make_view(dx); p.ksize = ksize;
  const int pty = (oh - 1) * 2 + ksize - dx->h, ptx = (ow - 1) * 2 + ksize - dx->w;
  p.pad_y = (pty > 0 ? pty : 0) / 2; p.pad_x = (ptx > 0 ? ptx : 0) / 2;
  const size_t total = static_cast<size_t>(dx->n) * dx->h * dx->w * (dx->c / 8);
  maxpool_bwd_index_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, index_dev);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_kernel_predict_bwd(dd_ctx* ctx, const dd_tensor* src, const dd_tensor* logits, const dd_tensor* dout, int ksize,
                          int features, int images_per_tuple, const dd_tensor* dlogits, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(src) && tensor_ok(logits) && tensor_ok(dout) && tensor_ok(dlogits), "bad argument");
  DD_CHECK_ARG(ksize >= 1 && (ksize & 1) && features >= 1 && logits->c == features * ksize * ksize, "bad kernel size / channels");
  DD_CHECK_ARG(src->c == 3 && same_dims(src, dout) && src->n == logits->n * features && same_dims(logits, dlogits), "dims");
  DD_CHECK_ARG(images_per_tuple >= 1 && logits->n % images_per_tuple == 0, "logits.n must be a multiple of images_per_tuple");
  KpBwdParams p;
  p.src = make_view(src); p.logits = make_view(logits); p.dout = make_view(dout); p.dlogits = make_view(dlogits);
  p.K = ksize; p.F = features; p.ipt = images_per_tuple;
  // tiled kernel: one feature per tuple, K = 3 / 5, fp32 logits / sources / output gradient in 16-byte aligned rows that are wide
  // enough for the vector accesses (the rows of these tensors are padded to multiples of 8 channels by their producers)
  const int k2 = ksize * ksize, lv = (k2 + 3) / 4 * 4, dv = (k2 + 7) / 8 * 8;
  const bool dl16 = is_half_type(dlogits->dtype);
  if (features == 1 && (ksize == 3 || ksize == 5) && logits->dtype == DD_F32 && src->dtype == DD_F32 && dout->dtype == DD_F32 &&
      logits->coff % 4 == 0 && logits->cstride % 4 == 0 && logits->coff + lv <= logits->cstride &&
      (dlogits->dtype == DD_F32 || (dl16 && dlogits->coff % 8 == 0 && dlogits->cstride % 8 == 0 && dlogits->coff + dv <= dlogits->cstride)) &&
      src->n <= 65535 && (src->h + 7) / 8 <= 65535) {
    dim3 grid((src->w + 31) / 32, (src->h + 7) / 8, src->n);
    if (ksize == 5) kernel_predict_bwd_tile_kernel<5><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else kernel_predict_bwd_tile_kernel<3><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    DD_LAUNCH_CHECK(ctx);
    return DD_OK;
  }
  const size_t total = static_cast<size_t>(logits->n) * features * src->h * src->w;
  if (ksize * ksize >= 64)
    kernel_predict_bwd_coop_kernel<32><<<nblocks(total * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    kernel_predict_bwd_coop_kernel<8><<<nblocks(total * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_compose_tail_bwd(dd_ctx* ctx, const dd_tensor* t, const float* w, const float* b, int c_mid, const dd_tensor* small,
                        const dd_tensor* large, const dd_tensor* dout, const dd_tensor* dt, const dd_tensor* dsmall,
                        const dd_tensor* dlarge, float* dw_dev, float* db_dev, void* stream) {
  DD_CHECK_ARG(ctx && w && b && dw_dev && db_dev && tensor_ok(t) && tensor_ok(small) && tensor_ok(large) && tensor_ok(dout) &&
                   tensor_ok(dt) && tensor_ok(dsmall) && tensor_ok(dlarge), "bad argument");
  DD_CHECK_ARG(c_mid > 0 && c_mid <= kCmpC && t->c >= c_mid && dt->c >= c_mid, "compose width unsupported");
  DD_CHECK_ARG(dsmall->dtype == DD_F32 && dlarge->dtype == DD_F32 && same_dims(small, dsmall) && same_dims(large, dlarge) &&
                   same_dims(large, dout), "gradient dims / dtype");
  ComposeTailBwdParams p; memset(&p, 0, sizeof(p));
  p.t = make_view(t); p.small = make_view(small); p.large = make_view(large); p.dout = make_view(dout); p.dt = make_view(dt);
  p.dsmall = make_view(dsmall); p.dlarge = make_view(dlarge);
  memcpy(p.w, w, sizeof(float) * c_mid); p.b = b[0]; p.c_mid = c_mid; p.dw = dw_dev; p.db = db_dev;
  const size_t total = static_cast<size_t>(large->n) * large->h * large->w;
  unsigned blocks = nblocks(total, 256);
  if (blocks > static_cast<unsigned>(ctx->sm_count) * 8) blocks = static_cast<unsigned>(ctx->sm_count) * 8;
  compose_tail_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_compose_head_bwd(dd_ctx* ctx, const dd_tensor* small, const dd_tensor* large, const float* w, int c_mid,
                        const dd_tensor* y, const dd_tensor* dy, const dd_tensor* dsmall, const dd_tensor* dlarge,
                        float* dw_dev, float* db_dev, void* stream) {
  DD_CHECK_ARG(ctx && w && dw_dev && db_dev && tensor_ok(small) && tensor_ok(large) && tensor_ok(y) && tensor_ok(dy) &&
                   tensor_ok(dsmall) && tensor_ok(dlarge), "bad argument");
  DD_CHECK_ARG(c_mid > 0 && c_mid <= kCmpC && y->c >= c_mid && dy->c >= c_mid, "compose width unsupported");
  DD_CHECK_ARG(dsmall->dtype == DD_F32 && dlarge->dtype == DD_F32 && same_dims(small, dsmall) && same_dims(large, dlarge),
               "gradient dims / dtype");
  ComposeHeadBwdParams p; memset(&p, 0, sizeof(p));
  p.small = make_view(small); p.large = make_view(large); p.y = make_view(y); p.dy = make_view(dy);
  p.dsmall = make_view(dsmall); p.dlarge = make_view(dlarge);
  memcpy(p.w, w, sizeof(float) * 6 * c_mid); p.c_mid = c_mid; p.dw = dw_dev; p.db = db_dev;
  const size_t total = static_cast<size_t>(large->n) * large->h * large->w;
  unsigned blocks = nblocks(total, 256);
  if (blocks > static_cast<unsigned>(ctx->sm_count) * 8) blocks = static_cast<unsigned>(ctx->sm_count) * 8;
  compose_head_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_channel_sum(dd_ctx* ctx, const dd_tensor* x, int groups, float* out_dev, void* stream) {
  DD_CHECK_ARG(ctx && out_dev && tensor_ok(x) && groups > 0 && x->n % groups == 0 && x->c <= 16, "bad argument");
  ChanSumParams p;
  p.x = make_view(x); p.out = out_dev; p.groups = groups; p.images_per_group = x->n / groups;
  dim3 grid(64, groups);
  channel_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_adam_step(dd_ctx* ctx, float* w, const float* g, float* m, float* v, size_t count, float lr, float beta1, float beta2,
                 float epsilon, int64_t step, float grad_scale, void* stream) {
  DD_CHECK_ARG(ctx && w && g && m && v && count > 0 && step >= 1, "bad argument");
  const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))) /
                      (1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
  adam_kernel<<<nblocks(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, g, m, v, count, static_cast<float>(lr_t),
                                                                                  beta1, beta2, epsilon, grad_scale);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_adam_step_guarded(dd_ctx* ctx, float* w, const float* g, float* m, float* v, size_t count, float lr, float beta1,
                         float beta2, float epsilon, float grad_scale, int32_t* state_dev, void* stream) {
  DD_CHECK_ARG(ctx && w && g && m && v && count > 0 && state_dev, "bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DD_CUDA(cudaMemsetAsync(state_dev + 1, 0, sizeof(int32_t), s));
  unsigned blocks = nblocks(count, 256);
  if (blocks > 1184u) blocks = 1184u;
  nonfinite_kernel<<<blocks, 256, 0, s>>>(g, count, state_dev);
  DD_LAUNCH_CHECK(ctx);
  adam_decide_kernel<<<1, 1, 0, s>>>(state_dev, lr, beta1, beta2);
  DD_LAUNCH_CHECK(ctx);
  adam_guarded_kernel<<<nblocks(count, 256), 256, 0, s>>>(w, g, m, v, count, state_dev, beta1, beta2, epsilon, grad_scale);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_conv2d_repack_f32(dd_ctx* ctx, const float* w_dev, int ksize, int cin, int cout, int transposed, float* fwd_packed,
                         float* dgrad_packed, void* stream) {
  DD_CHECK_ARG(ctx && w_dev && (fwd_packed || dgrad_packed) && ksize >= 1 && ksize <= 3 && cin > 0 && cout > 0, "bad argument");
  const size_t total = static_cast<size_t>(ksize) * ksize * cin * cout;
  repack_f32_kernel<<<nblocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w_dev, fwd_packed, dgrad_packed,
                                                                                       ksize * ksize, cin, cout, transposed);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // extern "C"
