"""GPU parity of the mixed-precision (tensor-core) training kernels against torch fp32 references evaluated on the SAME
fp16-rounded operands: tcgen05 weight gradient (dd_conv2d_wgrad_tc), device-side weight packing for the forward and the
input-gradient convolution (dd_conv2d_pack_weights_dev), space-to-depth + ReLU mask and the fused ReLU-backward /
bias-gradient pass.  Tolerance: 2e-3 of the result scale (fp32 accumulation order; inputs are identical fp16 values)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deepdenoiser_b200 import _lib

pytestmark = pytest.mark.gpu
_b = ctypes.byref


def _fp(t):
  return ctypes.c_void_p(t.data_ptr())


def rel_err(got, want):
  return float((got.double() - want.double()).abs().max()) / max(1e-6, float(want.double().abs().max()))


WGRAD_CASES = [
    # ks, cin, cout, n, h, w
    (3, 64, 64, 1, 9, 128),
    (3, 64, 64, 2, 21, 150),
    (3, 32, 64, 1, 16, 300),
    (3, 96, 96, 1, 10, 130),
    (3, 192, 96, 1, 10, 64),
    (3, 128, 128, 2, 9, 140),
    (3, 24, 24, 1, 12, 40),
    (3, 64, 64, 3, 40, 256),     # more rows than CTAs x minimum: several segments per CTA
    (1, 64, 25, 2, 9, 150),
    (1, 25, 25, 1, 7, 33),
    (1, 384, 128, 1, 8, 64),
]


@pytest.mark.parametrize("ks,cin,cout,n,h,w", WGRAD_CASES)
def test_wgrad_tc(ctx, ks, cin, cout, n, h, w):
  g = torch.Generator(device="cuda").manual_seed(ks * 1000 + cin + cout + h)
  cs_in, cs_out = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
  x = torch.zeros(n, h, w, cs_in, device="cuda", dtype=torch.float16)
  dz = torch.zeros(n, h, w, cs_out, device="cuda", dtype=torch.float16)
  x[..., :cin] = torch.randn(n, h, w, cin, device="cuda", generator=g).half()
  dz[..., :cout] = torch.randn(n, h, w, cout, device="cuda", generator=g).half()
  # padding channels hold garbage on purpose: the kernel must ignore them
  x[..., cin:] = 7.0
  dz[..., cout:] = -3.0
  dw = torch.full((ks, ks, cin, cout), 0.5, device="cuda")
  ctx.call("dd_conv2d_wgrad_tc", _b(_lib.desc(x, cin, 0)), _b(_lib.desc(dz, cout, 0)), ks, 0, _fp(dw), ctypes.c_float(0.25))
  xf = x[..., :cin].float().permute(0, 3, 1, 2).requires_grad_(False)
  wt = torch.zeros(cout, cin, ks, ks, device="cuda", requires_grad=True)
  y = F.conv2d(xf, wt, padding=ks // 2)
  y.backward(dz[..., :cout].float().permute(0, 3, 1, 2))
  want = 0.5 + 0.25 * wt.grad.permute(2, 3, 1, 0)            # [kh,kw,cin,cout]
  err = rel_err(dw, want)
  assert err <= 2e-3, "wgrad %s: relative error %.3e" % ((ks, cin, cout, n, h, w), err)


def test_wgrad_tc_transposed_layout(ctx):
  n, h, w, cin, cout = 1, 6, 70, 64, 48
  g = torch.Generator(device="cuda").manual_seed(5)
  x = torch.randn(n, h, w, cin, device="cuda", generator=g).half()
  dz = torch.randn(n, h, w, cout, device="cuda", generator=g).half()
  dw = torch.zeros(cout, cin, device="cuda")
  ctx.call("dd_conv2d_wgrad_tc", _b(_lib.desc(x)), _b(_lib.desc(dz)), 1, 1, _fp(dw), ctypes.c_float(1.0))
  want = torch.einsum("nhwc,nhwo->oc", x.float(), dz.float())
  assert rel_err(dw, want) <= 2e-3


@pytest.mark.parametrize("ks,cin,cout", [(3, 64, 64), (3, 32, 96), (1, 64, 25), (3, 192, 96)])
def test_pack_weights_dev_forward_matches_host(ctx, ks, cin, cout):
  wt = torch.randn(ks, ks, cin, cout) * 0.1
  host = ctx.pack_conv_weights(wt, torch.float16)
  devp = torch.zeros_like(host)
  ctx.call("dd_conv2d_pack_weights_dev", _fp(wt.cuda()), ks, cin, cout, 0, _fp(devp))
  assert torch.equal(host, devp)


def test_pack_weights_dev_transposed_matches_host(ctx):
  cin, cout = 96, 64
  wt = torch.randn(2, 2, cout, cin) * 0.1
  host = ctx.pack_conv_weights(wt, torch.float16, transposed=True)
  devp = torch.zeros_like(host)
  ctx.call("dd_conv2d_pack_weights_dev", _fp(wt.cuda()), 2, cin, cout, 2, _fp(devp))
  assert torch.equal(host, devp)


@pytest.mark.parametrize("ks,cin,cout", [(3, 64, 64), (3, 32, 64), (1, 64, 25), (3, 96, 128), (3, 192, 96)])
def test_dgrad_through_forward_kernel(ctx, ks, cin, cout):
  """dx = conv(dz, flipped / channel-swapped W) on the tcgen05 forward kernel == autograd input gradient."""
  n, h, w = 1, 12, 140
  g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout)
  wt = (torch.randn(ks, ks, cin, cout, device="cuda", generator=g) * 0.1)
  cs_out = (cout + 7) // 8 * 8
  dz = torch.zeros(n, h, w, cs_out, device="cuda", dtype=torch.float16)
  dz[..., :cout] = torch.randn(n, h, w, cout, device="cuda", generator=g).half()
  nbytes = ctx.lib.dd_conv2d_packed_bytes(ks, cout, cin, _lib.DD_F16, 0)
  packed = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
  ctx.call("dd_conv2d_pack_weights_dev", _fp(wt), ks, cin, cout, 1, _fp(packed))
  cs_in = (cin + 7) // 8 * 8
  dx = torch.zeros(n, h, w, cs_in, device="cuda", dtype=torch.float16)
  ctx.conv2d(_lib.desc(dz, cout, 0), packed, None, ks, _lib.desc(dx, cin, 0))
  xin = torch.zeros(n, cin, h, w, device="cuda", requires_grad=True)
  w16 = wt.half().float().permute(3, 2, 0, 1)
  y = F.conv2d(xin, w16, padding=ks // 2)
  y.backward(dz[..., :cout].float().permute(0, 3, 1, 2))
  want = xin.grad.permute(0, 2, 3, 1)
  assert rel_err(dx[..., :cin].float(), want) <= 3e-3


def test_space_to_depth_mask(ctx):
  n, h, w, c = 2, 6, 10, 16
  dy = torch.randn(n, 2 * h, 2 * w, c, device="cuda").half()
  y = torch.randn(n, 2 * h, 2 * w, c, device="cuda").half()
  out = torch.empty(n, h, w, 4 * c, device="cuda", dtype=torch.float16)
  ctx.call("dd_space_to_depth2_mask", _b(_lib.desc(dy)), _b(_lib.desc(y)), _b(_lib.desc(out)))
  m = dy * (y > 0)
  want = torch.cat([m[:, ay::2, ax::2, :] for ay in (0, 1) for ax in (0, 1)], dim=3)
  assert torch.equal(out, want)
  ctx.call("dd_space_to_depth2_mask", _b(_lib.desc(dy)), None, _b(_lib.desc(out)))
  want = torch.cat([dy[:, ay::2, ax::2, :] for ay in (0, 1) for ax in (0, 1)], dim=3)
  assert torch.equal(out, want)


@pytest.mark.parametrize("c,cstride,dtype", [(64, 64, torch.float16), (96, 200, torch.float16), (25, 32, torch.float16),
                                             (24, 24, torch.float32)])
def test_relu_bwd_bias(ctx, c, cstride, dtype):
  n, h, w = 2, 13, 37
  dy = torch.randn(n, h, w, cstride, device="cuda").to(dtype)
  y = torch.randn(n, h, w, cstride, device="cuda").to(dtype)
  dz = torch.full((n, h, w, cstride), 9.0, device="cuda", dtype=dtype)
  db = torch.full((c,), 1.0, device="cuda")
  coff = 8 if cstride >= c + 8 else 0
  ctx.call("dd_relu_bwd_bias", _b(_lib.desc(dy, c, coff)), _b(_lib.desc(y, c, coff)), _b(_lib.desc(dz, c, coff)), _fp(db),
           ctypes.c_float(0.5))
  want = (dy * (y > 0))[..., coff:coff + c]
  assert torch.equal(dz[..., coff:coff + c], want)
  untouched = torch.ones(cstride, dtype=torch.bool)
  untouched[coff:coff + c] = False
  assert bool((dz[..., untouched] == 9.0).all())
  want_db = 1.0 + 0.5 * want.float().sum(dim=(0, 1, 2))
  assert rel_err(db, want_db) <= 1e-3
  # bias gradient only (no mask, no store)
  db2 = torch.zeros(c, device="cuda")
  ctx.call("dd_relu_bwd_bias", _b(_lib.desc(dy, c, coff)), None, None, _fp(db2), ctypes.c_float(1.0))
  assert rel_err(db2, dy[..., coff:coff + c].float().sum(dim=(0, 1, 2))) <= 1e-3


def test_vectorised_elementwise_fp16_windows(ctx):
  """dd_relu_bwd / dd_relu_bwd_acc / dd_axpy on aligned fp16 channel windows (16-byte vectorised path): bit exact against
  torch in fp32 with one fp16 rounding, and channels outside the window stay untouched."""
  n, h, w, cs, c, coff = 2, 9, 21, 48, 24, 16
  dy = torch.randn(n, h, w, cs, device="cuda").half()
  y = torch.randn(n, h, w, cs, device="cuda").half()
  dz = torch.full((n, h, w, cs), 3.0, device="cuda", dtype=torch.float16)
  win = slice(coff, coff + c)
  ctx.call("dd_relu_bwd", _b(_lib.desc(dy, c, coff)), _b(_lib.desc(y, c, coff)), _b(_lib.desc(dz, c, coff)))
  assert torch.equal(dz[..., win], (dy * (y > 0))[..., win])
  assert bool((dz[..., :coff] == 3.0).all()) and bool((dz[..., coff + c:] == 3.0).all())
  acc = torch.randn(n, h, w, cs, device="cuda").half()
  want = acc.clone()
  want[..., win] = (acc.float() + (dy * (y > 0)).float())[..., win].half()
  ctx.call("dd_relu_bwd_acc", _b(_lib.desc(dy, c, coff)), _b(_lib.desc(y, c, coff)), _b(_lib.desc(acc, c, coff)))
  assert torch.equal(acc, want)
  yy = torch.randn(n, h, w, cs, device="cuda").half()
  want = yy.clone()
  want[..., win] = torch.addcmul(yy.float(), dy.float(), torch.tensor(0.5, device="cuda"))[..., win].half()
  ctx.call("dd_axpy", ctypes.c_float(0.5), _b(_lib.desc(dy, c, coff)), _b(_lib.desc(yy, c, coff)))
  assert torch.equal(yy, want)


@pytest.mark.parametrize("cin,cout", [(64, 64), (96, 24)])
def test_conv_epilogue_relu_mask(ctx, cin, cout):
  """DD_CONV_RESIDUAL_MASK: y = conv(x) * [mask > 0] - the ReLU backward of the previous layer fused into the input-gradient
  convolution.  Must equal the unfused pair (conv, then dd_relu_bwd) bit for bit."""
  n, h, w = 1, 11, 150
  x = torch.randn(n, h, w, cin, device="cuda").half()
  mask = torch.randn(n, h, w, cout, device="cuda").half()
  wp = ctx.pack_conv_weights(torch.randn(3, 3, cin, cout) * 0.1, torch.float16)
  plain = torch.empty(n, h, w, cout, device="cuda", dtype=torch.float16)
  fused = torch.empty_like(plain)
  ctx.conv2d(_lib.desc(x), wp, None, 3, _lib.desc(plain))
  ctx.conv2d(_lib.desc(x), wp, None, 3, _lib.desc(fused), residual=_lib.desc(mask), residual_is_mask=True)
  assert torch.equal(fused, plain * (mask > 0))


@pytest.mark.parametrize("k,features,ipt,h,w,ldtype,gdtype", [
    (5, 1, 1, 9, 14, torch.float32, None), (21, 1, 2, 12, 10, torch.float32, None), (3, 3, 1, 8, 11, torch.float32, None),
    (5, 3, 2, 7, 9, torch.float16, None),
    # the tiled kernel (one feature per tuple, K = 3 / 5, fp32 logits): tiles cut by the image border, 16-bit gradient rows
    (5, 1, 2, 19, 45, torch.float32, None), (5, 1, 1, 19, 45, torch.float32, torch.bfloat16),
    (3, 1, 1, 10, 70, torch.float32, torch.float16)])
def test_kernel_predict_bwd_matches_autograd(ctx, k, features, ipt, h, w, ldtype, gdtype):
  """dd_kernel_predict_bwd against torch autograd of softmax + symmetric-padded KxK gather (KernelPrediction.py:11-63)."""
  tuples = 2
  b, k2, pad = tuples * ipt, k * k, (k - 1) // 2
  g = torch.Generator(device="cuda").manual_seed(k * 10 + features)
  cs = (features * k2 + 7) // 8 * 8
  logits = torch.zeros(b, h, w, cs, device="cuda")
  logits[..., :features * k2] = torch.randn(b, h, w, features * k2, device="cuda", generator=g) * 2
  logits = logits.to(ldtype)
  src = torch.randn(features * b, h, w, 3, device="cuda", generator=g)
  dout = torch.randn(features * b, h, w, 3, device="cuda", generator=g)
  dl = torch.zeros(b, h, w, cs, device="cuda", dtype=gdtype or ldtype)
  ctx.call("dd_kernel_predict_bwd", _b(_lib.desc(src)), _b(_lib.desc(logits, features * k2, 0)), _b(_lib.desc(dout)), k, features,
           ipt, _b(_lib.desc(dl, features * k2, 0)))
  lg = logits.float()[..., :features * k2].clone().requires_grad_(True)
  idx_y = torch.arange(-pad, h + pad, device="cuda")
  idx_x = torch.arange(-pad, w + pad, device="cuda")
  sym = lambda i, n: torch.where(i < 0, -i - 1, torch.where(i >= n, 2 * n - i - 1, i))   # noqa: E731
  loss = 0.0
  for bi in range(b):
    t, n = divmod(bi, ipt)
    for f in range(features):
      o = (t * features + f) * ipt + n
      padded = src[o][sym(idx_y, h)][:, sym(idx_x, w)]                       # [h+2p, w+2p, 3], symmetric padding
      wts = torch.softmax(lg[bi, :, :, f * k2:(f + 1) * k2], dim=-1)
      out = sum(wts[..., i * k + j, None] * padded[i:i + h, j:j + w] for i in range(k) for j in range(k))
      loss = loss + (out * dout[o]).sum()
  loss.backward()
  tol = {torch.float16: 2e-3, torch.bfloat16: 1e-2, torch.float32: 2e-5}[gdtype or ldtype]
  assert rel_err(dl.float()[..., :features * k2], lg.grad) <= tol


# ------------------------------------------------------------------------------------------------ bfloat16 storage
def test_conv_and_wgrad_bfloat16(ctx):
  """The tensor-core kernels with bf16 operands (kind::f16, A/B format 1): forward conv, input gradient with fused ReLU mask
  and weight gradient against torch on the same bf16-rounded operands."""
  n, h, w, cin, cout = 1, 12, 140, 64, 96
  g = torch.Generator(device="cuda").manual_seed(3)
  x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
  wt = torch.randn(3, 3, cin, cout, device="cuda", generator=g) * 0.1
  bias = torch.randn(96, device="cuda", generator=g)
  nbytes = ctx.lib.dd_conv2d_packed_bytes(3, cin, cout, _lib.DD_BF16, 0)
  packed = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
  ctx.call("dd_conv2d_pack_weights_dev", _fp(wt), 3, cin, cout, 0 | _lib.DD_PACK_BF16, _fp(packed))
  host_packed = ctx.pack_conv_weights(wt.cpu(), torch.bfloat16)
  assert torch.equal(packed, host_packed)
  y = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
  ctx.conv2d(_lib.desc(x), packed, bias, 3, _lib.desc(y), relu=True)
  want = torch.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float().permute(3, 2, 0, 1), bias[:cout], padding=1))
  assert rel_err(y.float(), want.permute(0, 2, 3, 1)) <= 1e-2            # one bf16 rounding of the result (2^-8)
  # fp32 output of the same kernel is exact up to accumulation order
  y32 = torch.empty(n, h, w, cout, device="cuda")
  ctx.conv2d(_lib.desc(x), packed, bias, 3, _lib.desc(y32), relu=True)
  assert rel_err(y32, want.permute(0, 2, 3, 1)) <= 1e-5
  # weight gradient
  dz = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
  dw = torch.zeros(3, 3, cin, cout, device="cuda")
  ctx.call("dd_conv2d_wgrad_tc", _b(_lib.desc(x)), _b(_lib.desc(dz)), 3, 0, _fp(dw), ctypes.c_float(1.0))
  wz = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
  F.conv2d(x.float().permute(0, 3, 1, 2), wz, padding=1).backward(dz.float().permute(0, 3, 1, 2))
  assert rel_err(dw, wz.grad.permute(2, 3, 1, 0)) <= 2e-3
  # input gradient with the fused ReLU mask
  pb = torch.zeros(ctx.lib.dd_conv2d_packed_bytes(3, cout, cin, _lib.DD_BF16, 0), dtype=torch.uint8, device="cuda")
  ctx.call("dd_conv2d_pack_weights_dev", _fp(wt), 3, cin, cout, 1 | _lib.DD_PACK_BF16, _fp(pb))
  dx = torch.empty(n, h, w, cin, device="cuda", dtype=torch.bfloat16)
  ctx.conv2d(_lib.desc(dz), pb, None, 3, _lib.desc(dx), residual=_lib.desc(x), residual_is_mask=True)
  xin = torch.zeros(n, cin, h, w, device="cuda", requires_grad=True)
  F.conv2d(xin, wt.bfloat16().float().permute(3, 2, 0, 1), padding=1).backward(dz.float().permute(0, 3, 1, 2))
  want_dx = xin.grad.permute(0, 2, 3, 1) * (x.float() > 0)
  assert rel_err(dx.float(), want_dx) <= 1e-2


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_maxpool_16bit_vector_path(ctx, dtype):
  """3x3 stride-2 TF-'SAME' max pooling (tail padding) on 16-bit tensors, 16-byte vectorised kernel: bit exact."""
  n, h, w, c = 2, 12, 20, 24
  x = torch.randn(n, h, w, c, device="cuda").to(dtype)
  y = torch.empty(n, h // 2, w // 2, c, device="cuda", dtype=dtype)
  ctx.maxpool_s2(_lib.desc(x), 3, _lib.desc(y))
  xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1), value=float("-inf"))
  want = F.max_pool2d(xp, 3, 2).permute(0, 2, 3, 1).to(dtype)
  assert torch.equal(y, want)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("k,h,w", [(3, 12, 20), (3, 11, 17), (2, 12, 20), (2, 9, 15)])
def test_maxpool_backward_gather_matches_the_scatter_form(ctx, dtype, k, h, w):
  """dd_maxpool_s2_bwd_acc (gather, 16-bit, accumulating) == dd_maxpool_s2_bwd (fp32 atomics) on data FULL of ties (ReLU
  zeros and coarsely quantised positives): the gradient goes to the first maximum of every window (TF MaxPoolGrad)."""
  n, c = 2, 16
  x = torch.clamp(torch.round(torch.randn(n, h, w, c, device="cuda") * 2.0) / 2.0, min=0.0).to(dtype)
  oh, ow = (h + 1) // 2, (w + 1) // 2
  y = torch.empty(n, oh, ow, c, device="cuda", dtype=dtype)
  ctx.maxpool_s2(_lib.desc(x), k, _lib.desc(y))
  dy = (torch.round(torch.randn(n, oh, ow, c, device="cuda") * 8.0) / 8.0).to(dtype)      # exactly representable sums
  want = torch.zeros(n, h, w, c, device="cuda", dtype=torch.float32)
  ctx.call("dd_maxpool_s2_bwd", _b(_lib.desc(x)), _b(_lib.desc(y)), _b(_lib.desc(dy)), k, _b(_lib.desc(want)))
  base = (torch.round(torch.randn(n, h, w, c, device="cuda") * 4.0) / 4.0).to(dtype)
  got = base.clone()
  ctx.call("dd_maxpool_s2_bwd_acc", _b(_lib.desc(x)), _b(_lib.desc(y)), _b(_lib.desc(dy)), k, _b(_lib.desc(got)))
  torch.cuda.synchronize()
  assert torch.equal(got.float(), base.float() + want)


@pytest.mark.parametrize("cin,cout,n,h,w", [(64, 64, 2, 13, 300), (96, 128, 1, 9, 130), (128, 24, 1, 10, 64)])
def test_dgrad_conv_with_fused_mask_and_bias_gradient(ctx, cin, cout, n, h, w):
  """dd_conv2d_fwd_colsum with DD_CONV_RESIDUAL_MASK: the masked output equals the plain fused-mask launch bit for bit and the
  accumulated column sums equal the sum over all pixels of that output (the BiasAddGrad of the layer below)."""
  x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
  mask = torch.randn(n, h, w, cout, device="cuda").bfloat16()
  wt = torch.randn(3, 3, cin, cout, device="cuda") * 0.1
  pb = torch.zeros(ctx.lib.dd_conv2d_packed_bytes(3, cin, cout, _lib.DD_BF16, 0), dtype=torch.uint8, device="cuda")
  ctx.call("dd_conv2d_pack_weights_dev", _fp(wt), 3, cin, cout, 0 | _lib.DD_PACK_BF16, _fp(pb))
  plain = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
  fused = torch.empty_like(plain)
  ctx.conv2d(_lib.desc(x), pb, None, 3, _lib.desc(plain), residual=_lib.desc(mask), residual_is_mask=True)
  base = torch.randn(cout, device="cuda")
  db = base.clone()
  ctx.conv2d(_lib.desc(x), pb, None, 3, _lib.desc(fused), residual=_lib.desc(mask), residual_is_mask=True, colsum=db)
  torch.cuda.synchronize()
  assert torch.equal(fused, plain)
  # reference sum from fp32 arithmetic on the same bf16 operands (the kernel sums before the 16-bit rounding)
  ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
  want = (ref * (mask.float() > 0)).sum(dim=(0, 1, 2))
  assert rel_err(db - base, want) <= 2e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("k,h,w", [(3, 12, 20), (3, 11, 17), (2, 12, 20), (2, 9, 15)])
def test_maxpool_index_pair_matches_the_scatter_form(ctx, dtype, k, h, w):
  """dd_maxpool_s2_fwd_index / dd_maxpool_s2_bwd_index (recorded first maximum, pure-gather backward) == dd_maxpool_s2_fwd +
  dd_maxpool_s2_bwd (fp32 atomics) on tie-heavy data."""
  n, c = 2, 16
  x = torch.clamp(torch.round(torch.randn(n, h, w, c, device="cuda") * 2.0) / 2.0, min=0.0).to(dtype)
  oh, ow = (h + 1) // 2, (w + 1) // 2
  y_ref = torch.empty(n, oh, ow, c, device="cuda", dtype=dtype)
  ctx.maxpool_s2(_lib.desc(x), k, _lib.desc(y_ref))
  y = torch.empty_like(y_ref)
  idx = torch.empty(n, oh, ow, c, device="cuda", dtype=torch.uint8)
  ctx.call("dd_maxpool_s2_fwd_index", _b(_lib.desc(x)), k, _b(_lib.desc(y)), _fp(idx))
  assert torch.equal(y, y_ref)
  dy = (torch.round(torch.randn(n, oh, ow, c, device="cuda") * 8.0) / 8.0).to(dtype)
  want = torch.zeros(n, h, w, c, device="cuda", dtype=torch.float32)
  ctx.call("dd_maxpool_s2_bwd", _b(_lib.desc(x)), _b(_lib.desc(y_ref)), _b(_lib.desc(dy)), k, _b(_lib.desc(want)))
  base = (torch.round(torch.randn(n, h, w, c, device="cuda") * 4.0) / 4.0).to(dtype)
  got = base.clone()
  ctx.call("dd_maxpool_s2_bwd_index", _fp(idx), _b(_lib.desc(dy)), k, _b(_lib.desc(got)))
  torch.cuda.synchronize()
  assert torch.equal(got.float(), base.float() + want)
