"""GPU parity of every libdd_b200 entry point against the oracle operators (oracle/np_ops.py), called through
the C ABI (ctypes) exactly like the product does.  fp32 paths: <= 2e-5 relative to the output scale (fp32
reassociation); fp16 tensor-core convolution: compared with the oracle run on the SAME fp16-rounded operands,
tolerance 2e-3 * output scale (one fp16 rounding of the result + fp32 accumulation order)."""
import os

import numpy as np
import pytest
import torch

from deepdenoiser_b200 import _lib
from oracle import np_ops

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(11)


def dev(x, dtype=torch.float32):
  return torch.from_numpy(np.ascontiguousarray(x)).to("cuda", dtype)


def close(got, want, tol, what=""):
  got = got.detach().float().cpu().numpy().astype(np.float64)
  scale = max(1.0, float(np.abs(want).max()))
  err = float(np.abs(got - want).max())
  assert err <= tol * scale, "%s: max err %.3e > %.1e * %.3g" % (what, err, tol, scale)


def pad_bias(b):
  out = torch.zeros((b.shape[0] + 15) // 16 * 16)
  out[:b.shape[0]] = torch.from_numpy(b)
  return out.cuda()


# ------------------------------------------------------------------------------------------------ convolution
@pytest.mark.parametrize("ks,cin,cout,h,w", [(3, 12, 10, 17, 23), (1, 7, 5, 9, 31), (3, 32, 64, 8, 130)])
def test_conv2d_exact_fp32(ctx, ks, cin, cout, h, w):
  x = RNG.standard_normal((2, h, w, cin)).astype(np.float32)
  k = (RNG.standard_normal((ks, ks, cin, cout)) * 0.2).astype(np.float32)
  b = RNG.standard_normal(cout).astype(np.float32)
  res = RNG.standard_normal((2, h, w, cout)).astype(np.float32)
  wp = ctx.pack_conv_weights(torch.from_numpy(k), torch.float32)
  y = torch.empty(2, h, w, cout, device="cuda")
  yr = torch.empty(2, h, w, cout, device="cuda")
  ctx.conv2d(_lib.desc(dev(x)), wp, dev(b), ks, _lib.desc(y), relu=False, residual=_lib.desc(dev(res)),
             y_relu=_lib.desc(yr))
  want = np_ops.conv2d_same(x.astype(np.float64), k, b) + res
  close(y, want, 2e-5, "conv")
  close(yr, np.maximum(want, 0), 2e-5, "relu copy")


CONV_TC_CASES = [
    # ks, cin, cout, n, h, w, cstride_in, coff_in
    (3, 64, 64, 2, 21, 150, 64, 0),
    (3, 32, 64, 1, 16, 300, 32, 0),
    (3, 16, 64, 1, 16, 128, 16, 0),
    (3, 9, 64, 1, 12, 64, 16, 0),            # cfg1: 9 source channels in a 16-wide buffer
    (3, 96, 96, 1, 10, 130, 96, 0),
    (3, 192, 96, 1, 10, 130, 200, 8),        # concat buffer window
    (3, 128, 128, 1, 9, 140, 128, 0),
    (3, 96, 192, 1, 9, 140, 96, 0),          # wide output: two 96-column slices (input gradient of the 192->96 layer)
    (3, 24, 24, 1, 12, 40, 24, 0),           # compose residual convs
    (1, 64, 25, 2, 9, 150, 64, 0),           # post-process K=5
    (1, 25, 25, 1, 9, 150, 32, 0),
    (1, 128, 75, 1, 7, 40, 128, 0),          # COMBINED post-process
    (1, 320, 200, 1, 6, 129, 320, 0),        # wide 1x1 (Tiramisu transition, split over chunks)
]


@pytest.mark.parametrize("ks,cin,cout,n,h,w,cs,coff", CONV_TC_CASES)
def test_conv2d_tensor_core_fp16(ctx, ks, cin, cout, n, h, w, cs, coff):
  xfull = (RNG.standard_normal((n, h, w, cs)) * 0.5).astype(np.float16)
  k = (RNG.standard_normal((ks, ks, cin, cout)) / np.sqrt(ks * ks * cin)).astype(np.float32)
  b = (RNG.standard_normal(cout) * 0.1).astype(np.float32)
  wp = ctx.pack_conv_weights(torch.from_numpy(k), torch.float16)
  c8 = (cout + 7) // 8 * 8
  xd = dev(xfull, torch.float16)
  for out_dtype in (torch.float16, torch.float32):
    y = torch.full((n, h, w, c8 + 8), float("nan"), dtype=out_dtype, device="cuda")
    ctx.conv2d(_lib.desc(xd, cin, coff), wp, pad_bias(b), ks, _lib.desc(y, cout, 8 if c8 + 8 >= cout + 8 else 0),
               relu=True)
    x64 = xfull[..., coff:coff + cin].astype(np.float64)
    k64 = k.astype(np.float16).astype(np.float64)
    want = np_ops.conv2d_same(x64, k64, b.astype(np.float64), relu=True)
    close(y[..., 8:8 + cout], want, 2e-3 if out_dtype == torch.float16 else 2e-4, "tc conv %s" % out_dtype)
    assert torch.isnan(y[..., :8].float()).all(), "wrote outside the channel window"


def _split16(x):
  """x (float32) -> fp16 pair (hi, lo) with x ~= hi + lo to ~22 bits."""
  hi = x.astype(np.float16)
  lo = (x - hi.astype(np.float32)).astype(np.float16)
  return hi, lo


SPLIT_CASES = [
    # ks, cin, cout, n, h, w
    (3, 64, 64, 2, 21, 150),
    (3, 32, 64, 1, 16, 300),
    (3, 96, 96, 1, 10, 130),       # 1.5 chunks per part
    (3, 192, 96, 1, 10, 130),      # streamed weights
    (3, 128, 128, 1, 9, 140),
    (1, 64, 25, 2, 9, 150),        # post-process 1x1, fp32 output
    (1, 25, 25, 1, 9, 150),
]


@pytest.mark.parametrize("ks,cin,cout,n,h,w", SPLIT_CASES)
def test_conv2d_split_fp16x2(ctx, ks, cin, cout, n, h, w):
  """dd_conv2d_fwd_split (float16x2 mode: x = hi + lo, W = W_hi + W_lo, three MMA passes) against the float64 convolution of
  the ORIGINAL float32 operands: 2e-5 of the output scale - the floor is the fp32 accumulation of 3 * K products in TMEM
  (measured 7e-6 at K = 864), the dropped lo.lo term and the fp16 rounding of the lo halves are ~2^-22 relative; plain fp16
  operands give ~5e-4 here."""
  x = (RNG.standard_normal((n, h, w, cin)) * 0.5).astype(np.float32)
  k = (RNG.standard_normal((ks, ks, cin, cout)) / np.sqrt(ks * ks * cin)).astype(np.float32)
  b = (RNG.standard_normal(cout) * 0.1).astype(np.float32)
  wp = ctx.pack_conv_weights(torch.from_numpy(k), "float16x2")
  ci8, co8 = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
  xbuf = torch.zeros(n, h, w, 2 * ci8, dtype=torch.float16, device="cuda")
  hi, lo = _split16(x)
  xbuf[..., :cin] = dev(hi, torch.float16)
  xbuf[..., ci8:ci8 + cin] = dev(lo, torch.float16)
  want = np_ops.conv2d_same(x.astype(np.float64), k.astype(np.float64), b.astype(np.float64), relu=True)
  # (a) fp16 pair out
  ybuf = torch.full((n, h, w, 2 * co8), float("nan"), dtype=torch.float16, device="cuda")
  ctx.conv2d_split(_lib.desc(xbuf, cin, 0), _lib.desc(xbuf, cin, ci8), wp, pad_bias(b), ks, _lib.desc(ybuf, cout, 0),
                   _lib.desc(ybuf, cout, co8), relu=True)
  got = ybuf[..., :cout].float() + ybuf[..., co8:co8 + cout].float()
  close(got, want, 2e-5, "split conv, pair out")
  # (b) fp32 out
  y32 = torch.full((n, h, w, co8), float("nan"), device="cuda")
  ctx.conv2d_split(_lib.desc(xbuf, cin, 0), _lib.desc(xbuf, cin, ci8), wp, pad_bias(b), ks, _lib.desc(y32, cout, 0), None, relu=True)
  close(y32[..., :cout], want, 2e-5, "split conv, fp32 out")


def test_split_transpose2x2_maxpool_assemble(ctx):
  n, h, w, cin, cout = 2, 6, 70, 128, 96
  x = (RNG.standard_normal((n, h, w, cin)) * 0.5).astype(np.float32)
  k = (RNG.standard_normal((2, 2, cout, cin)) / np.sqrt(cin)).astype(np.float32)
  b = (RNG.standard_normal(cout) * 0.1).astype(np.float32)
  wp = ctx.pack_conv_weights(torch.from_numpy(k), "float16x2", transposed=True)
  hi, lo = _split16(x)
  xbuf = torch.cat([dev(hi, torch.float16), dev(lo, torch.float16)], dim=3).contiguous()
  # written into the second half of a skip-concat pair buffer [hi(2c) | lo(2c)]
  ybuf = torch.full((n, 2 * h, 2 * w, 4 * cout), float("nan"), dtype=torch.float16, device="cuda")
  ctx.conv2d_transpose2x2_split(_lib.desc(xbuf, cin, 0), _lib.desc(xbuf, cin, cin), wp, pad_bias(b),
                                _lib.desc(ybuf, cout, cout), _lib.desc(ybuf, cout, 3 * cout), relu=True)
  want = np_ops.conv2d_transpose_same_s2(x.astype(np.float64), k.astype(np.float64), b.astype(np.float64), relu=True)
  close(ybuf[..., cout:2 * cout].float() + ybuf[..., 3 * cout:].float(), want, 2e-5, "split convT2x2")
  assert torch.isnan(ybuf[..., :cout].float()).all() and torch.isnan(ybuf[..., 2 * cout:3 * cout].float()).all()
  # max-pool 3x3 s2 on pairs: exact max of hi + lo, re-split
  c = 64
  v = (RNG.standard_normal((n, 10, 14, c)) * 2).astype(np.float32)
  vh, vl = _split16(v)
  vbuf = torch.cat([dev(vh, torch.float16), dev(vl, torch.float16)], dim=3).contiguous()
  pbuf = torch.empty(n, 5, 7, 2 * c, dtype=torch.float16, device="cuda")
  ctx.maxpool_s2_split(_lib.desc(vbuf, c, 0), _lib.desc(vbuf, c, c), 3, _lib.desc(pbuf, c, 0), _lib.desc(pbuf, c, c))
  pair = vh.astype(np.float32) + vl.astype(np.float32)
  close(pbuf[..., :c].float() + pbuf[..., c:].float(), np_ops.max_pool_same_s2(pair.astype(np.float64), 3), 1e-6, "split maxpool")


def test_conv2d_tensor_core_residual_and_relu_copy(ctx):
  n, h, w, c = 1, 11, 70, 24
  x = (RNG.standard_normal((n, h, w, c)) * 0.5).astype(np.float16)
  res = (RNG.standard_normal((n, h, w, c)) * 0.5).astype(np.float16)
  k = (RNG.standard_normal((3, 3, c, c)) / np.sqrt(9 * c)).astype(np.float32)
  b = (RNG.standard_normal(c) * 0.1).astype(np.float32)
  wp = ctx.pack_conv_weights(torch.from_numpy(k), torch.float16)
  y = torch.empty(n, h, w, c, dtype=torch.float16, device="cuda")
  yr = torch.empty(n, h, w, c, dtype=torch.float16, device="cuda")
  ctx.conv2d(_lib.desc(dev(x, torch.float16)), wp, pad_bias(b), 3, _lib.desc(y), relu=False,
             residual=_lib.desc(dev(res, torch.float16)), y_relu=_lib.desc(yr))
  want = np_ops.conv2d_same(x.astype(np.float64), k.astype(np.float16).astype(np.float64), b.astype(np.float64))
  want = want + res.astype(np.float64)
  close(y, want, 2e-3, "residual")
  close(yr, np.maximum(want, 0), 2e-3, "relu copy")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("cin,cout", [(128, 96), (96, 64), (24, 16)])
def test_conv2d_transpose_2x2(ctx, dtype, cin, cout):
  n, h, w = 2, 6, 70
  x = (RNG.standard_normal((n, h, w, cin)) * 0.5).astype(np.float16)
  k = (RNG.standard_normal((2, 2, cout, cin)) / np.sqrt(cin)).astype(np.float32)
  b = (RNG.standard_normal(cout) * 0.1).astype(np.float32)
  wp = ctx.pack_conv_weights(torch.from_numpy(k), dtype, transposed=True)
  y = torch.full((n, 2 * h, 2 * w, 2 * cout), float("nan"), dtype=dtype, device="cuda")
  ctx.conv2d_transpose2x2(_lib.desc(dev(x, dtype)), wp, pad_bias(b), _lib.desc(y, cout, cout), relu=True)
  k64 = k.astype(np.float16).astype(np.float64) if dtype == torch.float16 else k.astype(np.float64)
  want = np_ops.conv2d_transpose_same_s2(x.astype(np.float64), k64, b.astype(np.float64), relu=True)
  close(y[..., cout:], want, 2e-3 if dtype == torch.float16 else 2e-5, "convT2x2")
  assert torch.isnan(y[..., :cout].float()).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_conv2d_transpose_3x3_tf_same(ctx, dtype):
  n, h, w, cin, cout = 1, 7, 66, 40, 32
  x = (RNG.standard_normal((n, h, w, cin)) * 0.5).astype(np.float16)
  k = (RNG.standard_normal((3, 3, cout, cin)) / np.sqrt(4 * cin)).astype(np.float32)
  b = (RNG.standard_normal(cout) * 0.1).astype(np.float32)
  phases = []
  for py in range(2):
    for px in range(2):
      wz = np.zeros((3, 3, cin, cout), dtype=np.float32)
      for dy in (0, -1):
        for dx in (0, -1):
          r, s = py - 2 * dy, px - 2 * dx
          if r <= 2 and s <= 2:
            wz[dy + 1, dx + 1] = k[r, s].T
      phases.append(ctx.pack_conv_weights(torch.from_numpy(wz), dtype))
  y = torch.full((n, 2 * h, 2 * w, cout), float("nan"), dtype=dtype, device="cuda")
  ya = torch.full((n, 2 * h, 2 * w, cout), float("nan"), dtype=dtype, device="cuda")
  ctx.conv2d_transpose3x3(_lib.desc(dev(x, dtype)), phases, pad_bias(b), _lib.desc(y), _lib.desc(ya), relu=True)
  k64 = k.astype(np.float16).astype(np.float64) if dtype == torch.float16 else k.astype(np.float64)
  want = np_ops.conv2d_transpose_same_s2(x.astype(np.float64), k64, b.astype(np.float64), relu=True)
  close(y, want, 2e-3 if dtype == torch.float16 else 2e-5, "convT3x3")
  close(ya, want, 2e-3 if dtype == torch.float16 else 2e-5, "convT3x3 relu copy")


# ------------------------------------------------------------------------------------------------ pooling
@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("dtype,c,h,w", [(torch.float16, 64, 12, 18), (torch.float16, 24, 7, 9), (torch.float32, 5, 9, 6)])
def test_maxpool_tf_same(ctx, k, dtype, c, h, w):
  x = RNG.standard_normal((2, h, w, c)).astype(np.float16)
  y = torch.empty(2, (h + 1) // 2, (w + 1) // 2, c, dtype=dtype, device="cuda")
  ctx.maxpool_s2(_lib.desc(dev(x, dtype)), k, _lib.desc(y))
  close(y, np_ops.max_pool_same_s2(x.astype(np.float64), k), 0.0, "maxpool")      # exact: a max of inputs


@pytest.mark.parametrize("f,h,w", [(2, 12, 20), (4, 12, 20), (2, 7, 9)])
def test_avgpool_tf_same(ctx, f, h, w):
  x = RNG.standard_normal((3, h, w, 3)).astype(np.float32)
  y = torch.empty(3, -(-h // f), -(-w // f), 3, device="cuda")
  ctx.avgpool(_lib.desc(dev(x)), f, _lib.desc(y))
  close(y, np_ops.avg_pool_same(x.astype(np.float64), f), 2e-6, "avgpool")


# ------------------------------------------------------------------------------------------------ source encoder
@pytest.mark.parametrize("c", [1, 3])
@pytest.mark.parametrize("mode,rel,before,comp,log1p,mean,var",
                         [("uniform", True, False, True, True, 0.0, 1.0),
                          ("neighbor", False, False, False, True, 0.25, 2.0),
                          ("uniform", True, True, True, False, 0.0, 1.0),
                          ("neighbor", True, True, False, True, -0.5, 0.5)])
def test_standardize_variance(ctx, c, mode, rel, before, comp, log1p, mean, var):
  x = (RNG.standard_normal((2, 9, 13, c)) * 3).astype(np.float32)
  x[0, :3, :4] = 0.0
  prm = _lib.dd_standardize_params(int(log1p), mean, var, 1, 0 if mode == "uniform" else 1, int(rel), int(before),
                                   int(comp), 1e-4)
  s = torch.empty(2, 9, 13, 3, device="cuda")
  vc = 1 if comp else c
  v = torch.empty(2, 9, 13, vc, device="cuda")
  ctx.standardize_variance(_lib.desc(dev(x)), prm, _lib.desc(s), _lib.desc(v))
  x64 = x.astype(np.float64)
  std = np_ops.signed_log1p(x64) if log1p else x64
  std = (std - mean) / np.sqrt(var)
  want_v = np_ops.variance_feature(x64 if before else std, mode, rel, comp)
  close(s, np.repeat(std, 3, axis=3) if c == 1 else std, 2e-6, "standardize")
  # E[x^2]-E[x]^2 cancels catastrophically in fp32; relative variance divides by >= 1e-4
  tol = 3e-2 if rel else 2e-5
  close(v, want_v, tol, "variance")


@pytest.mark.parametrize("n,h,w", [(1, 8, 8), (2, 37, 71), (1, 16, 130)])
@pytest.mark.parametrize("generic", [0, 1])
def test_standardize_variance_batch_all_variants_in_one_launch(ctx, n, h, w, generic):
  """dd_standardize_variance_batch: every parametrisation as one job of ONE launch (fp32 fast path: 64 x 16 tiles, four pixels per
  thread; generic=1 forces the generic tile kernel), 1- and 3-channel sources, tiles cut by the image border on both axes."""
  jobs, wants = [], []
  keep = []
  for c in (3, 1):
    for mode in ("uniform", "neighbor"):
      for rel in (False, True):
        for before in (False, True):
          for comp in (False, True):
            log1p, mean, var = (not before), (0.3 if rel else 0.0), (1.7 if comp else 1.0)
            x = (RNG.standard_normal((n, h, w, c)) * 3).astype(np.float32)
            x[0, :3, :4] = 0.0
            prm = _lib.dd_standardize_params(int(log1p), mean, var, 1, 0 if mode == "uniform" else 1, int(rel), int(before), int(comp), 1e-4)
            sd = torch.full((n, h, w, 3), float("nan"), device="cuda")
            vd = torch.full((n, h, w, 1 if comp else c), float("nan"), device="cuda")
            xd = dev(x)
            keep.append((xd, sd, vd))
            jobs.append((_lib.desc(xd), prm, _lib.desc(sd), _lib.desc(vd)))
            x64 = x.astype(np.float64)
            std = np_ops.signed_log1p(x64) if log1p else x64
            std = (std - mean) / np.sqrt(var)
            wants.append((np.repeat(std, 3, axis=3) if c == 1 else std, np_ops.variance_feature(x64 if before else std, mode, rel, comp), rel))
  table = torch.empty(len(jobs) * int(ctx.lib.dd_standardize_variance_job_bytes()), dtype=torch.uint8, device="cuda")
  ctx.set_option("std_generic", generic)
  try:
    ctx.standardize_variance_batch(jobs, table)
  finally:
    ctx.set_option("std_generic", 0)
  for (xd, sd, vd), (want_s, want_v, rel) in zip(keep, wants):
    close(sd, want_s, 2e-6, "standardize (batch)")
    close(vd, want_v, 3e-2 if rel else 2e-5, "variance (batch)")


def test_assemble_input_gathers_channels(ctx):
  n, h, w, tuples, c = 2, 5, 7, 3, 16
  a = RNG.standard_normal((n, h, w, 3)).astype(np.float32)
  b = RNG.standard_normal((n, h, w, 1)).astype(np.float32)
  emb = RNG.standard_normal((tuples, 4)).astype(np.float32)
  ad, bd, ed = dev(a), dev(b), dev(emb)
  table = np.zeros((tuples, c), dtype=np.dtype([("ptr", "<u8"), ("cstride", "<i4"), ("cidx", "<i4"),
                                                ("constant", "<f4"), ("pad", "<i4")]))
  want = np.zeros((tuples * n, h, w, c))
  for t in range(tuples):
    for ch in range(3):
      table[t, ch] = (ad.data_ptr(), 3, (ch + t) % 3, 0, 0)
      want[t * n:(t + 1) * n, ..., ch] = a[..., (ch + t) % 3]
    table[t, 3] = (bd.data_ptr(), 1, 0, 0, 0)
    want[t * n:(t + 1) * n, ..., 3] = b[..., 0]
    for j in range(4):
      table[t, 4 + j] = (ed.data_ptr() + (t * 4 + j) * 4, 0, 0, 0, 0)
      want[t * n:(t + 1) * n, ..., 4 + j] = emb[t, j]
    table[t, 8] = (0, 0, 0, 1.5, 0)
    want[t * n:(t + 1) * n, ..., 8] = 1.5
  td = torch.from_numpy(table.view(np.uint8).reshape(-1)).cuda()
  for dtype, tol in ((torch.float32, 0.0), (torch.float16, 1e-3)):
    out = torch.empty(tuples * n, h, w, c, dtype=dtype, device="cuda")
    ctx.assemble_input(td, tuples, n, _lib.desc(out))
    close(out, want, tol, "assemble %s" % dtype)


# ------------------------------------------------------------------------------------------------ kernel prediction
@pytest.mark.parametrize("k,features,ipt,h,w,ldtype", [(5, 1, 1, 19, 37, torch.float32), (5, 3, 2, 16, 40, torch.float16),
                                                       (21, 1, 2, 33, 45, torch.float32), (21, 1, 1, 12, 10, torch.float16),
                                                       (3, 3, 1, 8, 8, torch.float32), (7, 2, 1, 40, 33, torch.float32)])
def test_kernel_predict(ctx, k, features, ipt, h, w, ldtype):
  tuples = 2
  b = tuples * ipt
  k2 = k * k
  cpad = (features * k2 + 7) // 8 * 8
  logits = np.zeros((b, h, w, cpad), dtype=np.float32)
  logits[..., :features * k2] = RNG.standard_normal((b, h, w, features * k2)) * 2
  logits = logits.astype(np.float16).astype(np.float32) if ldtype == torch.float16 else logits
  src = (RNG.standard_normal((features * b, h, w, 3)) * 2).astype(np.float32)
  out = torch.empty(features * b, h, w, 3, device="cuda")
  ctx.kernel_predict(_lib.desc(dev(src)), _lib.desc(dev(logits, ldtype), features * k2, 0), k, features, ipt,
                     _lib.desc(out))
  want = np.zeros((features * b, h, w, 3))
  for bi in range(b):
    t, n = divmod(bi, ipt)
    for f in range(features):
      o = (t * features + f) * ipt + n
      want[o] = np_ops.kernel_prediction(src[o:o + 1].astype(np.float64),
                                         logits[bi:bi + 1, ..., f * k2:(f + 1) * k2].astype(np.float64), k)[0]
  close(out, want, 5e-6, "kernel prediction")


def test_kernel_predict_known_answers(ctx):
  k, h, w = 5, 12, 13
  src = RNG.standard_normal((1, h, w, 3)).astype(np.float32)
  # equal logits -> symmetric-border box filter; one-hot -> shift by (i-p, j-p)
  out = torch.empty(1, h, w, 3, device="cuda")
  ctx.kernel_predict(_lib.desc(dev(src)), _lib.desc(dev(np.zeros((1, h, w, 32), np.float32)), 25, 0), k, 1, 1, _lib.desc(out))
  xp = np.pad(src.astype(np.float64), ((0, 0), (2, 2), (2, 2), (0, 0)), mode="symmetric")
  close(out, sum(xp[:, i:i + h, j:j + w] for i in range(5) for j in range(5)) / 25, 1e-6, "box")
  logits = np.full((1, h, w, 32), -1e4, np.float32)
  logits[..., 1 * 5 + 4] = 0
  ctx.kernel_predict(_lib.desc(dev(src)), _lib.desc(dev(logits), 25, 0), k, 1, 1, _lib.desc(out))
  close(out, xp[:, 1:1 + h, 4:4 + w], 1e-7, "shift")


# ------------------------------------------------------------------------------------------------ multi-scale
def test_compose_head_tail_and_invert(ctx):
  n, h, w = 2, 10, 14
  small = RNG.standard_normal((n, h // 2, w // 2, 3)).astype(np.float32)
  large = RNG.standard_normal((n, h, w, 3)).astype(np.float32)
  hw_ = (RNG.standard_normal((6, 24)) * 0.4).astype(np.float32)
  hb = (RNG.standard_normal(24) * 0.1).astype(np.float32)
  y = torch.empty(n, h, w, 24, device="cuda")
  ctx.compose_head(_lib.desc(dev(small)), _lib.desc(dev(large)), torch.from_numpy(hw_), torch.from_numpy(hb), 24, _lib.desc(y))
  up = np_ops.resize_nearest_x2(small.astype(np.float64))
  x6 = np.concatenate([up, large.astype(np.float64)], axis=3)
  want_head = np.maximum(x6 @ hw_.astype(np.float64) + hb, 0)
  close(y, want_head, 2e-6, "compose head")
  t = RNG.standard_normal((n, h, w, 24)).astype(np.float32)
  tw = (RNG.standard_normal(24) * 0.3).astype(np.float32)
  tb = np.array([0.1], np.float32)
  out = torch.empty(n, h, w, 3, device="cuda")
  for inv in (None, _lib.dd_invert_params(1, 0.25, 2.0)):
    ctx.compose_tail(_lib.desc(dev(t)), torch.from_numpy(tw), torch.from_numpy(tb), 24, _lib.desc(dev(small)),
                     _lib.desc(dev(large)), inv, _lib.desc(out))
    wgt = np_ops.sigmoid(np.maximum(t.astype(np.float64) @ tw.astype(np.float64) + 0.1, 0))[..., None]
    assert wgt.min() >= 0.5
    low = np_ops.resize_nearest_x2(np_ops.avg_pool_same(large.astype(np.float64), 2))
    want = large - wgt * low + wgt * up
    if inv is not None:
      want = np_ops.signed_expm1(want * np.sqrt(2.0) + 0.25)
    close(out, want, 5e-6, "compose tail")
  x = (RNG.standard_normal((n, h, w, 3)) * 2).astype(np.float32)
  xd = dev(x)
  ctx.invert_standardization(_lib.desc(xd), _lib.dd_invert_params(1, 0.0, 1.0), _lib.desc(xd))   # in place
  close(xd, np_ops.signed_expm1(x.astype(np.float64)), 2e-6, "invert")


def _compose_case(n, h, w, seed):
  rng = np.random.default_rng(seed)
  small = rng.standard_normal((n, h // 2, w // 2, 3)).astype(np.float32)
  large = (small.repeat(2, axis=1).repeat(2, axis=2) + 0.3 * rng.standard_normal((n, h, w, 3))).astype(np.float32)
  head_w = (rng.standard_normal((1, 1, 6, 24)) * 0.4).astype(np.float32)
  head_b = (rng.standard_normal(24) * 0.1).astype(np.float32)
  conv_w = [(rng.standard_normal((3, 3, 24, 24)) * 0.08).astype(np.float32) for _ in range(4)]
  conv_b = [(rng.standard_normal(24) * 0.1).astype(np.float32) for _ in range(4)]
  tail_w = (rng.standard_normal((1, 1, 24, 1)) * 0.3).astype(np.float32)
  tail_b = np.array([0.05], np.float32)
  return small, large, head_w, head_b, conv_w, conv_b, tail_w, tail_b


def _compose_oracle(small, large, head_w, head_b, conv_w, conv_b, tail_w, tail_b, inv, rows=None):
  """float64 restatement of MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93); `rows` = (y0, y1) evaluates
  a row band of a tall image only (the four 3x3 layers need 4 rows of context on each side)."""
  f64 = lambda a: np.asarray(a, dtype=np.float64)  # noqa: E731
  h = large.shape[1]
  a, b = (0, h) if rows is None else (max(0, (rows[0] - 4) & ~1), min(h, (rows[1] + 4 + 1) & ~1))
  small, large = small[:, a // 2:b // 2], large[:, a:b]
  up = np_ops.resize_nearest_x2(f64(small))
  x = np_ops.conv2d_same(np.concatenate([up, f64(large)], axis=3), f64(head_w), f64(head_b), relu=True)
  for blk in range(2):
    r = x
    for i in range(2):
      r = np_ops.conv2d_same(np.maximum(r, 0), f64(conv_w[2 * blk + i]), f64(conv_b[2 * blk + i]), relu=False)
    x = x + r
  wgt = np_ops.sigmoid(np_ops.conv2d_same(x, f64(tail_w), f64(tail_b), relu=True))
  low = np_ops.resize_nearest_x2(np_ops.avg_pool_same(f64(large), 2))
  want = f64(large) - wgt * low + wgt * up
  if inv:
    want = np_ops.signed_expm1(want * np.sqrt(2.0) + 0.25)
  if rows is not None:
    want = want[:, rows[0] - a:rows[1] - a]
  return want


@pytest.mark.parametrize("n,h,w,inv,dtype", [(1, 16, 32, False, "f16"), (2, 20, 44, True, "f16"), (1, 50, 70, False, "f16"),
                                            (3, 6, 10, True, "f16"), (1, 34, 250, False, "f16"), (2, 300, 130, True, "f16"),
                                            (1, 2, 2, False, "f16"), (2, 20, 44, True, "bf16")])
def test_compose_scales_fused(ctx, n, h, w, inv, dtype):
  """dd_compose_scales_fwd (one tcgen05 launch: 16-bit activations in shared memory, fp32 accumulation in TMEM) against the
  float64 restatement of MultiScalePrediction.compose_scales with the SAME weights; several strips (w > 122), several CTA row
  ranges with recomputed halos, image borders and degenerate sizes all occur.
  Bound: fp16 activations between the layers perturb the blend weight by ~1e-4 relative: 2e-3 of the output scale
  (measured <= 6e-4); bf16 (8-bit significand) 2e-2."""
  case = _compose_case(n, h, w, seed=n * 1000 + h + w)
  small, large = case[0], case[1]
  code = _lib.DD_F16 if dtype == "f16" else _lib.DD_BF16
  blob, floats, code = _lib.pack_compose_weights(*case[2:], dtype=code)
  packed = (torch.from_numpy(blob).cuda(), floats, code)
  out = torch.full((n, h, w, 3), float("nan"), device="cuda")
  ip = _lib.dd_invert_params(1, 0.25, 2.0) if inv else None
  ctx.compose_scales(_lib.desc(dev(small)), _lib.desc(dev(large)), packed, ip, _lib.desc(out))
  want = _compose_oracle(*case, inv)
  err = np.abs(out.cpu().numpy() - want)
  print("compose %s %dx%dx%d: max %.2e mean %.2e" % (dtype, n, h, w, err.max() / max(1.0, np.abs(want).max()), err.mean()))
  close(out, want, 2e-3 if dtype == "f16" else 2e-2, "fused compose")
  assert err.mean() < (2e-4 if dtype == "f16" else 2e-3)


def _reference_components():
  """(inputs, outputs) of tests/golden/refshim_components.npz: the reference's own building blocks executed over
  oracle/tf_shim by tests/golden/make_reference_golden.py (inputs are a closed form, regenerated here)."""
  import importlib.util
  golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
  spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(golden, "make_reference_golden.py"))
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m.component_inputs(), np.load(os.path.join(golden, "refshim_components.npz"))


@pytest.mark.parametrize("k", [3, 5, 7, 21])
def test_kernel_predict_matches_the_reference_code_fixture(ctx, k):
  """dd_kernel_predict_fwd (fp32 logits) against KernelPrediction.kernel_prediction of the reference itself."""
  inp, z = _reference_components()
  src, logits = inp["kp%d|src" % k].astype(np.float32), inp["kp%d|logits" % k].astype(np.float32)
  k2 = k * k
  cpad = (k2 + 7) // 8 * 8
  padded = np.zeros(logits.shape[:3] + (cpad,), np.float32)
  padded[..., :k2] = logits
  out = torch.empty(src.shape, device="cuda")
  ctx.kernel_predict(_lib.desc(dev(src)), _lib.desc(dev(padded), k2, 0), k, 1, src.shape[0], _lib.desc(out))
  close(out, z["kp%d|out" % k], 5e-6, "kernel prediction vs reference code")


def test_compose_scales_matches_the_reference_code_fixture(ctx):
  """dd_compose_scales_fwd against MultiScalePrediction.compose_scales of the reference itself (fp16 activations between the
  layers: the bound of test_compose_scales_fused)."""
  inp, z = _reference_components()
  f32 = lambda key: inp["compose|" + key].astype(np.float32)      # noqa: E731
  names = ["conv2d"] + ["conv2d_%d" % i for i in range(1, 6)]
  blob, floats, code = _lib.pack_compose_weights(f32(names[0] + "/kernel"), f32(names[0] + "/bias"),
                                                 [f32(n + "/kernel") for n in names[1:5]], [f32(n + "/bias") for n in names[1:5]],
                                                 f32(names[5] + "/kernel"), f32(names[5] + "/bias"))
  out = torch.full(inp["compose|large"].shape, float("nan"), device="cuda")
  ctx.compose_scales(_lib.desc(dev(f32("small"))), _lib.desc(dev(f32("large"))), (torch.from_numpy(blob).cuda(), floats, code), None,
                     _lib.desc(out))
  close(out, z["compose|out"], 2e-3, "compose vs reference code")


def test_variance_and_standardisation_match_the_reference_code_fixture(ctx):
  """dd_standardize_variance against FeatureEngineering.variance / Utilities.signed_log1p of the reference itself."""
  inp, z = _reference_components()
  x = inp["var|x"].astype(np.float32)
  for mode in ("uniform", "neighbor"):
    for rel in (False, True):
      for one in (False, True):
        want = z["var|%s|%d|%d" % (mode, rel, one)]
        std = torch.empty(x.shape, device="cuda")
        var = torch.empty(want.shape, device="cuda")
        params = _lib.dd_standardize_params(0, 0.0, 1.0, 1, 1 if mode == "neighbor" else 0, int(rel), 0, int(one), 1e-4)
        ctx.standardize_variance(_lib.desc(dev(x)), params, _lib.desc(std), _lib.desc(var))
        tol = 3e-2 if rel else 1e-5       # relative variance divides by a squared mean that is clamped at 1e-4
        close(var, want, tol, "variance %s rel %d one %d" % (mode, rel, one))
  u = inp["util|x"].astype(np.float32).reshape(1, 8, 8, 1)
  std = torch.empty(1, 8, 8, 3, device="cuda")            # a 1-channel pass is replicated to 3 channels
  ctx.standardize_variance(_lib.desc(dev(u)), _lib.dd_standardize_params(1, 0.0, 1.0, 0, 0, 0, 0, 0, 1e-4), _lib.desc(std), None)
  close(std, np.repeat(z["util|log1p"].reshape(1, 8, 8, 1), 3, axis=3), 2e-6, "signed_log1p vs reference code")


def test_compose_scales_at_benchmark_shape(ctx):
  """8 x 1080 x 1920 (every CTA walks several (image, strip) segments): sampled row bands against the float64 oracle."""
  n, h, w = 8, 1080, 1920
  rng = np.random.default_rng(3)
  case = _compose_case(1, 4, 4, seed=5)
  g = torch.Generator(device="cuda").manual_seed(9)
  small = torch.randn(n, h // 2, w // 2, 3, device="cuda", generator=g)
  large = small.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2) + 0.3 * torch.randn(n, h, w, 3, device="cuda", generator=g)
  blob, floats, code = _lib.pack_compose_weights(*case[2:])
  out = torch.full((n, h, w, 3), float("nan"), device="cuda")
  ctx.compose_scales(_lib.desc(small), _lib.desc(large), (torch.from_numpy(blob).cuda(), floats, code), None, _lib.desc(out))
  torch.cuda.synchronize()
  assert bool(torch.isfinite(out).all())
  strips = (w + 121) // 122
  rows_per_cta = -(-(n * strips * h) // ctx.sm_count())
  bands = [(0, 0), (n - 1, h - 8), (3, 536)]
  for k in rng.choice(np.arange(1, ctx.sm_count()), size=5, replace=False):
    col, y = divmod(int(k) * rows_per_cta, h)
    bands.append((col // strips, max(0, min(h - 8, y - 4))))
  worst = 0.0
  for i, y0 in bands:
    want = _compose_oracle(small[i:i + 1].cpu().numpy(), large[i:i + 1].cpu().numpy(), *case[2:], False, rows=(y0, y0 + 8))
    got = out[i:i + 1, y0:y0 + 8].cpu().numpy()
    err = float(np.abs(got - want).max()) / max(1.0, float(np.abs(want).max()))
    worst = max(worst, err)
    assert err <= 2e-3, (i, y0, err)
  print("compose 8x1080x1920: worst band error %.2e" % worst)


def test_bad_arguments_are_reported_not_crashed(ctx):
  x = torch.zeros(1, 4, 4, 3, device="cuda")
  with pytest.raises(_lib.DDError, match="kernel size"):
    ctx.kernel_predict(_lib.desc(x), _lib.desc(torch.zeros(1, 4, 4, 16, device="cuda")), 4, 1, 1, _lib.desc(x))
  with pytest.raises(_lib.DDError):
    ctx.maxpool_s2(_lib.desc(x), 5, _lib.desc(x))
  with pytest.raises(_lib.DDError, match="unknown option"):
    ctx.set_option("nope", 1)
  n0 = ctx.launch_count()
  ctx.avgpool(_lib.desc(x), 2, _lib.desc(torch.zeros(1, 2, 2, 3, device="cuda")))
  assert ctx.launch_count() == n0 + 1


# ------------------------------------------------------------------------------------------------ inference tiling
def test_device_tiling_matches_host_slicing(ctx):
  """dd_tiles_gather / dd_tiles_scatter == the reference's numpy slicing (Prediction.py:282-310, 384-441), bit exact."""
  from deepdenoiser_b200 import prediction
  h, w = 150, 333
  tiles, size, overlap = prediction.tile_grid(h, w, 64, 7)
  image = torch.from_numpy(RNG.standard_normal((h, w, 3)).astype(np.float32))
  host_tiles = prediction.cut_tiles(image, tiles)
  dev_tiles = prediction.cut_tiles(image.cuda(), tiles, ctx)
  assert torch.equal(dev_tiles.cpu(), host_tiles)
  pred = torch.from_numpy(RNG.standard_normal(tuple(host_tiles.shape)).astype(np.float32))
  host_image = prediction.stitch_tiles(pred, tiles, h, w)
  dev_image = prediction.stitch_tiles(pred.cuda(), tiles, h, w, ctx)
  assert torch.equal(dev_image.cpu(), host_image)
  # every pixel is covered exactly by the kept crops: stitching the cut tiles gives the image back
  assert torch.equal(prediction.stitch_tiles(dev_tiles, tiles, h, w, ctx).cpu(), image)


# ------------------------------------------------------------------------------------------------ fused output head
@pytest.mark.parametrize("k,features,ipt,cin,cstride,coff,h,w", [
    (5, 1, 1, 64, 64, 0, 21, 37), (5, 3, 2, 64, 64, 0, 9, 50), (3, 1, 2, 16, 16, 0, 16, 24), (3, 3, 1, 24, 40, 8, 10, 33),
    (5, 1, 1, 96, 96, 0, 8, 16), (5, 1, 3, 128, 128, 0, 17, 19)])
def test_post_kp_fused(ctx, k, features, ipt, cin, cstride, coff, h, w):
  """dd_post_kp_fwd == conv1x1+ReLU -> conv1x1 -> softmax kernel prediction of the oracle on the same fp16-rounded
  operands (x, weights; the hidden layer is rounded to fp16 like the unfused path stores it): 2e-3 of the output scale."""
  tuples = 2
  b = tuples * ipt
  o = features * k * k
  x = np.zeros((b, h, w, cstride), dtype=np.float32)
  x[..., coff:coff + cin] = np.abs(RNG.standard_normal((b, h, w, cin)))          # core outputs are ReLU'd
  x = x.astype(np.float16)
  w1 = (RNG.standard_normal((1, 1, cin, o)) / np.sqrt(cin)).astype(np.float32)
  b1 = (RNG.standard_normal(o) * 0.1).astype(np.float32)
  w2 = (RNG.standard_normal((1, 1, o, o)) * 1.5 / np.sqrt(o)).astype(np.float32)
  b2 = (RNG.standard_normal(o) * 0.1).astype(np.float32)
  src = (RNG.standard_normal((features * b, h, w, 3)) * 2).astype(np.float32)
  blob = torch.from_numpy(_lib.pack_post_kp_weights(w1, b1, w2, b2, k, features)).cuda()
  out = torch.full((features * b, h, w, 3), 7.0, device="cuda")
  ctx.post_kp(_lib.desc(torch.from_numpy(x).cuda(), cin, coff), blob, _lib.desc(dev(src)), k, features, ipt, _lib.desc(out))
  xr = x[..., coff:coff + cin].astype(np.float64)
  w1r, w2r = w1.astype(np.float16).astype(np.float64), w2.astype(np.float16).astype(np.float64)
  hidden = np.maximum(np_ops.conv2d_same(xr, w1r, b1), 0).astype(np.float16).astype(np.float64)
  logits = np_ops.conv2d_same(hidden, w2r, b2)
  want = np.zeros((features * b, h, w, 3))
  k2 = k * k
  for bi in range(b):
    t, n = divmod(bi, ipt)
    for f in range(features):
      oi = (t * features + f) * ipt + n
      want[oi] = np_ops.kernel_prediction(src[oi:oi + 1].astype(np.float64), logits[bi:bi + 1, ..., f * k2:(f + 1) * k2], k)[0]
  close(out, want, 2e-3, "fused post + kernel prediction")
