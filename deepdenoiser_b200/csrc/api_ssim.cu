// C-ABI: multi-scale SSIM loss term (BaseFeatureTraining.ms_ssim, Training.py:188-204: 1 - mean(tf.image.ssim_multiscale(
// predicted, target, max_val = 1, power_factors = (0.0448, 0.2856, 0.3001)))) - forward statistics and the hand-derived
// backward.  tf.image.ssim_multiscale is TensorFlow code outside the reference [external]: 11x11 Gaussian (sigma 1.5, softmax
// normalised) applied as a VALID depthwise filter, c1 = (0.01 max)^2, c2 = (0.03 max)^2,
//   l  = (2 mx my + c1) / (mx^2 + my^2 + c1)            cs = (2 sxy - 2 mx my + c2) / (sxx + syy - mx^2 - my^2 + c2)
// per level k: cs_k = mean(cs), ssim_k = mean(l cs) over the filtered pixels, per (image, channel); levels are 2x2 average
// poolings of the previous one; ms = prod_k relu(v_k)^w_k with v = cs for all but the last level (ssim there); mean over
// channels, then over the batch.  The tiny per-(image, channel, level) algebra runs on the host side of the ABI
// (deepdenoiser_b200/training.py); these kernels do everything that touches pixels.
#include <math.h>
#include <string.h>

#include "dd_internal.h"

namespace dd {

constexpr int kSsimF = 11;
struct SsimGauss { float g[kSsimF]; };     // separable: the 2D kernel is the outer product of the normalised 1D one

static SsimGauss make_gauss() {
  SsimGauss k;
  double sum = 0.0, v[kSsimF];
  for (int i = 0; i < kSsimF; ++i) { const double c = i - (kSsimF - 1) / 2.0; v[i] = exp(-0.5 * c * c / (1.5 * 1.5)); sum += v[i]; }
  for (int i = 0; i < kSsimF; ++i) k.g[i] = static_cast<float>(v[i] / sum);
  return k;
}

// stats[n, y, x, {mx, my, sxy, sxx+syy} x C] over the VALID region (h - 10) x (w - 10)
struct SsimStatsParams { View x, y, stats; SsimGauss k; };
__global__ void __launch_bounds__(256) ssim_stats_kernel(const SsimStatsParams p) {
  const int C = p.x.c, ho = p.stats.h, wo = p.stats.w;
  const size_t total = static_cast<size_t>(p.stats.n) * ho * wo * C;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const size_t opix = idx / C;
  const int ox = static_cast<int>(opix % wo), oy = static_cast<int>((opix / wo) % ho);
  const int n = static_cast<int>(opix / (static_cast<size_t>(wo) * ho));
  float mx = 0.f, my = 0.f, sxy = 0.f, e = 0.f;
  for (int i = 0; i < kSsimF; ++i) {
    float rx = 0.f, ry = 0.f, rxy = 0.f, re = 0.f;
    for (int j = 0; j < kSsimF; ++j) {
      const size_t ip = p.x.pix(n, oy + i, ox + j);
      const float a = p.x.load(ip, c), b = p.y.load(ip, c), g = p.k.g[j];
      rx = fmaf(g, a, rx); ry = fmaf(g, b, ry); rxy = fmaf(g, a * b, rxy); re = fmaf(g, a * a + b * b, re);
    }
    const float g = p.k.g[i];
    mx = fmaf(g, rx, mx); my = fmaf(g, ry, my); sxy = fmaf(g, rxy, sxy); e = fmaf(g, re, e);
  }
  p.stats.store(opix, c, mx); p.stats.store(opix, C + c, my); p.stats.store(opix, 2 * C + c, sxy); p.stats.store(opix, 3 * C + c, e);
}

__device__ __forceinline__ void ssim_terms(float mx, float my, float sxy, float e, float c1, float c2, float& l, float& cs,
                                           float& nl, float& dl, float& nc, float& dc) {
  nl = 2.f * mx * my + c1; dl = mx * mx + my * my + c1;
  nc = 2.f * sxy - 2.f * mx * my + c2; dc = e - mx * mx - my * my + c2;
  l = nl / dl; cs = nc / dc;
}

// sums[n][c][0] += sum cs, sums[n][c][1] += sum l * cs over the filtered pixels
struct SsimReduceParams { View stats; float* sums; int C; float c1, c2; };
__global__ void __launch_bounds__(256) ssim_reduce_kernel(const SsimReduceParams p) {
  const int C = p.C, ho = p.stats.h, wo = p.stats.w;
  const int n = blockIdx.y;
  const size_t per = static_cast<size_t>(ho) * wo;
  for (int c = 0; c < C; ++c) {
    float a = 0.f, b = 0.f;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t pix = static_cast<size_t>(n) * per + i;
      float l, cs, nl, dl, nc, dc;
      ssim_terms(p.stats.load(pix, c), p.stats.load(pix, C + c), p.stats.load(pix, 2 * C + c), p.stats.load(pix, 3 * C + c), p.c1, p.c2,
                 l, cs, nl, dl, nc, dc);
      a += cs; b += l * cs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.sums + (n * C + c) * 2, a); atomicAdd(p.sums + (n * C + c) * 2 + 1, b); }
  }
}

// Given d objective / d mean(cs) = coef[n][c][0] and d objective / d mean(l cs) = coef[n][c][1]: the gradient with respect to
// the prediction x of this level, accumulated:  dx[p] += sum_q g(p - q) (dmx[q] + y[p] dsxy[q] + 2 x[p] de[q]).
struct SsimBwdParams { View x, y, stats, dx; const float* coef; SsimGauss k; float c1, c2; };
__global__ void __launch_bounds__(256) ssim_bwd_kernel(const SsimBwdParams p) {
  const int C = p.x.c, h = p.x.h, w = p.x.w, ho = p.stats.h, wo = p.stats.w;
  const size_t total = static_cast<size_t>(p.x.n) * h * w * C;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const size_t pix = idx / C;
  const int px = static_cast<int>(pix % w), py = static_cast<int>((pix / w) % h);
  const int n = static_cast<int>(pix / (static_cast<size_t>(w) * h));
  const float inv_count = 1.f / static_cast<float>(ho * wo);
  const float A = p.coef[(n * C + c) * 2] * inv_count, B = p.coef[(n * C + c) * 2 + 1] * inv_count;
  const float xv = p.x.load(pix, c), yv = p.y.load(pix, c);
  float acc = 0.f;
  for (int i = 0; i < kSsimF; ++i) {
    const int qy = py - i;
    if (qy < 0 || qy >= ho) continue;
    for (int j = 0; j < kSsimF; ++j) {
      const int qx = px - j;
      if (qx < 0 || qx >= wo) continue;
      const size_t q = p.stats.pix(n, qy, qx);
      const float mx = p.stats.load(q, c), my = p.stats.load(q, C + c);
      float l, cs, nl, dl, nc, dc;
      ssim_terms(mx, my, p.stats.load(q, 2 * C + c), p.stats.load(q, 3 * C + c), p.c1, p.c2, l, cs, nl, dl, nc, dc);
      const float d_cs = A + B * l, d_l = B * cs;
      const float dl_dmx = (2.f * my * dl - nl * 2.f * mx) / (dl * dl);
      const float dcs_dmx = (-2.f * my * dc + 2.f * mx * nc) / (dc * dc);
      const float dmx = d_l * dl_dmx + d_cs * dcs_dmx;
      const float dsxy = d_cs * 2.f / dc;
      const float de = -d_cs * nc / (dc * dc);
      acc = fmaf(p.k.g[i] * p.k.g[j], dmx + yv * dsxy + 2.f * xv * de, acc);
    }
  }
  p.dx.store(pix, c, p.dx.load(pix, c) + acc);
}

// dx_fine[2i+a, 2j+b] += 0.25 * dx_coarse[i, j]: adjoint of the 2x2 average pooling between levels
struct PoolAdjParams { View coarse, fine; };
__global__ void __launch_bounds__(256) avgpool2_adjoint_kernel(const PoolAdjParams p) {
  const int C = p.fine.c;
  const size_t total = static_cast<size_t>(p.fine.n) * p.fine.h * p.fine.w * C;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const size_t pix = idx / C;
  const int x = static_cast<int>(pix % p.fine.w), y = static_cast<int>((pix / p.fine.w) % p.fine.h);
  const int n = static_cast<int>(pix / (static_cast<size_t>(p.fine.w) * p.fine.h));
  p.fine.store(pix, c, p.fine.load(pix, c) + 0.25f * p.coarse.load(p.coarse.pix(n, y >> 1, x >> 1), c));
}

inline unsigned ssim_blocks(size_t total) { return static_cast<unsigned>((total + 255) / 256); }

}  // namespace dd

using namespace dd;

extern "C" {

int dd_ssim_stats(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* stats, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y) && tensor_ok(stats), "bad argument");
  DD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, "ssim: x and y differ");
  DD_CHECK_ARG(x->h >= kSsimF && x->w >= kSsimF, "ssim: the image must be at least 11 x 11 (filter size of tf.image.ssim)");
  DD_CHECK_ARG(stats->n == x->n && stats->h == x->h - (kSsimF - 1) && stats->w == x->w - (kSsimF - 1) && stats->c == 4 * x->c,
               "ssim: stats must be [n, h-10, w-10, 4c]");
  SsimStatsParams p;
  p.x = make_view(x); p.y = make_view(y); p.stats = make_view(stats); p.k = make_gauss();
  const size_t total = static_cast<size_t>(stats->n) * stats->h * stats->w * x->c;
  ssim_stats_kernel<<<ssim_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_ssim_reduce(dd_ctx* ctx, const dd_tensor* stats, int channels, float max_val, float* sums_dev, void* stream) {
  DD_CHECK_ARG(ctx && sums_dev && tensor_ok(stats) && channels > 0 && stats->c == 4 * channels && stats->n <= 65535, "bad argument");
  SsimReduceParams p;
  p.stats = make_view(stats); p.sums = sums_dev; p.C = channels;
  p.c1 = (0.01f * max_val) * (0.01f * max_val); p.c2 = (0.03f * max_val) * (0.03f * max_val);
  dim3 grid(32, stats->n);
  ssim_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_ssim_bwd(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* stats, const float* coef_dev, float max_val,
                const dd_tensor* dx_acc, void* stream) {
  DD_CHECK_ARG(ctx && coef_dev && tensor_ok(x) && tensor_ok(y) && tensor_ok(stats) && tensor_ok(dx_acc), "bad argument");
  DD_CHECK_ARG(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c && stats->c == 4 * x->c &&
                   stats->h == x->h - (kSsimF - 1) && stats->w == x->w - (kSsimF - 1) && dx_acc->n == x->n && dx_acc->h == x->h &&
                   dx_acc->w == x->w && dx_acc->c == x->c, "ssim_bwd: dims");
  SsimBwdParams p;
  p.x = make_view(x); p.y = make_view(y); p.stats = make_view(stats); p.dx = make_view(dx_acc); p.coef = coef_dev; p.k = make_gauss();
  p.c1 = (0.01f * max_val) * (0.01f * max_val); p.c2 = (0.03f * max_val) * (0.03f * max_val);
  const size_t total = static_cast<size_t>(x->n) * x->h * x->w * x->c;
  ssim_bwd_kernel<<<ssim_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

int dd_avgpool2_adjoint(dd_ctx* ctx, const dd_tensor* dcoarse, const dd_tensor* dfine_acc, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(dcoarse) && tensor_ok(dfine_acc), "bad argument");
  DD_CHECK_ARG(dfine_acc->n == dcoarse->n && dfine_acc->h == 2 * dcoarse->h && dfine_acc->w == 2 * dcoarse->w &&
                   dfine_acc->c == dcoarse->c, "avgpool2_adjoint: fine must be 2x coarse");
  PoolAdjParams p;
  p.coarse = make_view(dcoarse); p.fine = make_view(dfine_acc);
  const size_t total = static_cast<size_t>(dfine_acc->n) * dfine_acc->h * dfine_acc->w * dfine_acc->c;
  avgpool2_adjoint_kernel<<<ssim_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

}  // extern "C"
