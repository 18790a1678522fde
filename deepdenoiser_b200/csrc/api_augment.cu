// C-ABI: on-device data augmentation of training tiles.  One gather kernel applies, per example, the
// reference's chain DataAugmentation.flip_left_right -> rotate_90 -> permute_rgb -> rotate_normal
// (DataAugmentation.py:10-200 in the order of Training.py:803-815) so a batch never returns to the host between the
// TFRecord decode and the network input.
#include <string.h>

#include "dd_internal.h"

namespace dd {

struct AugParams {
  const float* x; float* y;
  int E, S, C, xcs, xoff, ycs, yoff, kind;
  const int32_t* flip; const int32_t* rot; const int32_t* perm; const float* rotation;
};

__constant__ int kPermTable[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};   // DataAugmentation.py:117-123

__global__ void __launch_bounds__(256) augment_kernel(const AugParams p) {
  const size_t total = static_cast<size_t>(p.E) * p.S * p.S;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int S = p.S;
  const int oj = static_cast<int>(idx % S), oi = static_cast<int>((idx / S) % S), e = static_cast<int>(idx / (static_cast<size_t>(S) * S));
  const int flip = p.flip ? (p.flip[e] > 0) : 0;
  const int k = p.rot ? (p.rot[e] & 3) : 0;
  // tf.image.rot90 (counter-clockwise) of the flipped image F: out[i][j] = F[j][S-1-i] (k=1), F[S-1-i][S-1-j] (k=2),
  // F[S-1-j][i] (k=3); F[i][j] = in[i][S-1-j] when flipped.
  int fi, fj;
  switch (k) {
    case 1: fi = oj; fj = S - 1 - oi; break;
    case 2: fi = S - 1 - oi; fj = S - 1 - oj; break;
    case 3: fi = S - 1 - oj; fj = oi; break;
    default: fi = oi; fj = oj; break;
  }
  if (flip) fj = S - 1 - fj;
  const float* src = p.x + ((static_cast<size_t>(e) * S + fi) * S + fj) * p.xcs + p.xoff;
  float* dst = p.y + idx * p.ycs + p.yoff;
  if (p.C != 3 || p.kind == DD_AUG_PLAIN) {
    for (int c = 0; c < p.C; ++c) dst[c] = src[c];
    return;
  }
  float v0 = src[0], v1 = src[1], v2 = src[2];
  if (p.kind == DD_AUG_SCREEN_SPACE_NORMAL) {
    if (flip) v0 = -v0;                                           // :31-45
    if (k == 1) { const float t = v0; v0 = -v1; v1 = t; }         // x -> -y, y -> x   (:76-83)
    else if (k == 2) { v0 = -v0; v1 = -v1; }                      // :85-90
    else if (k == 3) { const float t = v1; v1 = -v0; v0 = t; }    // x -> y, y -> -x   (:92-99)
  } else if (p.kind == DD_AUG_COLOR) {
    const int q = p.perm ? p.perm[e] : 0;
    if (q > 0 && q < 6) {
      const float in[3] = {v0, v1, v2};
      v0 = in[kPermTable[q][0]]; v1 = in[kPermTable[q][1]]; v2 = in[kPermTable[q][2]];
    }
  } else if (p.kind == DD_AUG_NORMAL && p.rotation) {
    const float* R = p.rotation + static_cast<size_t>(e) * 9;     // out = in . R  (tf.matmul(inputs, rotation_matrix), :193)
    const float a = v0, b = v1, c = v2;
    v0 = a * R[0] + b * R[3] + c * R[6];
    v1 = a * R[1] + b * R[4] + c * R[7];
    v2 = a * R[2] + b * R[5] + c * R[8];
  }
  dst[0] = v0; dst[1] = v1; dst[2] = v2;
}


// ------------------------------------------------------------------------------------------------ inference tiling
// Prediction.py:282-310 cuts the frame into overlapping tiles on the host and :384-441 pastes the kept part of every
// predicted tile back with numpy slicing; here both directions are one launch over a device table of
// {y, x, crop_y0, crop_y1, crop_x0, crop_x1} per tile.
struct TileParams { View image, tiles; const int32_t* table; int size; };

__global__ void __launch_bounds__(256) tiles_gather_kernel(const TileParams p) {
  const size_t total = static_cast<size_t>(p.tiles.n) * p.size * p.size;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % p.size), py = static_cast<int>((idx / p.size) % p.size);
  const int t = static_cast<int>(idx / (static_cast<size_t>(p.size) * p.size));
  const int32_t* e = p.table + t * 6;
  const size_t ipix = p.image.pix(0, e[0] + py, e[1] + px);
  for (int c = 0; c < p.tiles.c; ++c) p.tiles.store(idx, c, p.image.load(ipix, c));
}

__global__ void __launch_bounds__(256) tiles_scatter_kernel(const TileParams p) {
  const size_t total = static_cast<size_t>(p.tiles.n) * p.size * p.size;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % p.size), py = static_cast<int>((idx / p.size) % p.size);
  const int t = static_cast<int>(idx / (static_cast<size_t>(p.size) * p.size));
  const int32_t* e = p.table + t * 6;
  if (py < e[2] || py >= e[3] || px < e[4] || px >= e[5]) return;       // outside the part of the tile that is kept
  const size_t ipix = p.image.pix(0, e[0] + py, e[1] + px);
  for (int c = 0; c < p.tiles.c; ++c) p.image.store(ipix, c, p.tiles.load(idx, c));
}

}  // namespace dd

using namespace dd;

extern "C" int dd_augment_tiles(dd_ctx* ctx, const dd_tensor* x, int kind, const int32_t* flip_dev, const int32_t* rot_dev,
                                const int32_t* perm_dev, const float* rotation_dev, const dd_tensor* y, void* stream) {
  DD_CHECK_ARG(ctx && tensor_ok(x) && tensor_ok(y), "bad argument");
  DD_CHECK_ARG(x->dtype == DD_F32 && y->dtype == DD_F32, "augmentation works on fp32 tiles");
  DD_CHECK_ARG(x->h == x->w && y->n == x->n && y->h == x->h && y->w == x->w && y->c == x->c, "augment: square tiles of equal shape");
  DD_CHECK_ARG(x->ptr != y->ptr, "augment: out of place only");
  DD_CHECK_ARG(kind >= DD_AUG_PLAIN && kind <= DD_AUG_NORMAL, "augment: unknown kind");
  DD_CHECK_ARG(kind == DD_AUG_PLAIN || x->c == 3, "augment: colour / normal passes have 3 channels");
  AugParams p;
  memset(&p, 0, sizeof(p));
  p.x = reinterpret_cast<const float*>(x->ptr); p.y = reinterpret_cast<float*>(y->ptr);
  p.E = x->n; p.S = x->h; p.C = x->c; p.xcs = x->cstride; p.xoff = x->coff; p.ycs = y->cstride; p.yoff = y->coff;
  p.kind = kind; p.flip = flip_dev; p.rot = rot_dev; p.perm = perm_dev; p.rotation = rotation_dev;
  const size_t total = static_cast<size_t>(p.E) * p.S * p.S;
  augment_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

static int tiles_common(dd_ctx* ctx, const dd_tensor* image, const int32_t* table_dev, const dd_tensor* tiles, TileParams* p) {
  DD_CHECK_ARG(ctx && table_dev && tensor_ok(image) && tensor_ok(tiles), "bad argument");
  DD_CHECK_ARG(image->n == 1 && tiles->h == tiles->w && tiles->c == image->c && tiles->h <= image->h && tiles->w <= image->w,
               "tiles: image [1,H,W,C], tiles [T,S,S,C] with S <= H, W");
  p->image = make_view(image); p->tiles = make_view(tiles); p->table = table_dev; p->size = tiles->h;
  return DD_OK;
}

extern "C" int dd_tiles_gather(dd_ctx* ctx, const dd_tensor* image, const int32_t* table_dev, const dd_tensor* tiles, void* stream) {
  TileParams p;
  int rc = tiles_common(ctx, image, table_dev, tiles, &p);
  if (rc) return rc;
  const size_t total = static_cast<size_t>(tiles->n) * tiles->h * tiles->w;
  tiles_gather_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}

extern "C" int dd_tiles_scatter(dd_ctx* ctx, const dd_tensor* tiles, const int32_t* table_dev, const dd_tensor* image, void* stream) {
  TileParams p;
  int rc = tiles_common(ctx, image, table_dev, tiles, &p);
  if (rc) return rc;
  const size_t total = static_cast<size_t>(tiles->n) * tiles->h * tiles->w;
  tiles_scatter_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DD_LAUNCH_CHECK(ctx);
  return DD_OK;
}
