"""Pins the oracle's operators: known-answer / property tests derived from the reference's math (SURVEY.md
section 4) and agreement of the two independent restatements (numpy slicing vs torch F.* formulations)."""
import numpy as np
import pytest
import torch

from oracle import np_ops, torch_ops

RNG = np.random.default_rng(7)


def _t(x):
  return torch.from_numpy(np.ascontiguousarray(x))


# ------------------------------------------------------------------ kernel prediction (KernelPrediction.py:11-63)
@pytest.mark.parametrize("k", [3, 5, 21])
def test_kp_equal_logits_is_symmetric_box_filter(k):
  x = RNG.standard_normal((1, 24, 26, 3))
  logits = np.full((1, 24, 26, k * k), 0.3)
  out = np_ops.kernel_prediction(x, logits, k)
  p = (k - 1) // 2
  xp = np.pad(x, ((0, 0), (p, p), (p, p), (0, 0)), mode="symmetric")
  box = sum(xp[:, i:i + 24, j:j + 26] for i in range(k) for j in range(k)) / (k * k)
  np.testing.assert_allclose(out, box, atol=1e-12)


def test_kp_constant_image_unchanged():
  x = np.full((2, 9, 11, 3), 1.75)
  logits = RNG.standard_normal((2, 9, 11, 25))
  np.testing.assert_allclose(np_ops.kernel_prediction(x, logits, 5), x, atol=1e-12)


@pytest.mark.parametrize("i,j", [(0, 0), (2, 2), (1, 4), (4, 0)])
def test_kp_one_hot_logit_shifts_by_tap_offset(i, j):
  k, p = 5, 2
  x = RNG.standard_normal((1, 12, 13, 3))
  logits = np.full((1, 12, 13, 25), -1e4)
  logits[..., i * k + j] = 0.0      # stack index = i*K + j, i = row offset (KernelPrediction.py:32-39)
  out = np_ops.kernel_prediction(x, logits, k)
  xp = np.pad(x, ((0, 0), (p, p), (p, p), (0, 0)), mode="symmetric")
  np.testing.assert_allclose(out, xp[:, i:i + 12, j:j + 13], atol=1e-12)


def test_kp_backends_agree():
  x = RNG.standard_normal((2, 10, 12, 3)).astype(np.float32)
  logits = RNG.standard_normal((2, 10, 12, 49)).astype(np.float32)
  a = np_ops.kernel_prediction(x.astype(np.float64), logits.astype(np.float64), 7)
  b = torch_ops.kernel_prediction(_t(x), _t(logits), 7).numpy()
  np.testing.assert_allclose(a, b, atol=2e-6)


# ------------------------------------------------------------------ symmetric pad / variance
def test_symmetric_pad_includes_edge():
  x = np.arange(5, dtype=np.float64).reshape(1, 1, 5, 1)
  x = np.repeat(x, 3, axis=1)
  got = np_ops.pad_symmetric(x, 2)[0, 2, :, 0]
  np.testing.assert_array_equal(got, [1, 0, 0, 1, 2, 3, 4, 4, 3])
  np.testing.assert_array_equal(torch_ops.pad_symmetric(_t(x), 2)[0, 2, :, 0].numpy(), got)


@pytest.mark.parametrize("mode", ["uniform", "neighbor"])
def test_variance_of_constant_is_zero_and_backends_agree(mode):
  const = np.full((1, 8, 8, 3), 2.5)
  assert np.abs(np_ops.variance_feature(const, mode, True, True)).max() < 1e-12
  x = RNG.standard_normal((2, 9, 7, 3))
  for rel in (False, True):
    for comp in (False, True):
      a = np_ops.variance_feature(x, mode, rel, comp)
      b = torch_ops.variance_feature(_t(x), mode, rel, comp).numpy()
      assert a.shape == (2, 9, 7, 1 if comp else 3)
      np.testing.assert_allclose(a, b, atol=1e-10)


def test_variance_uniform_matches_direct_formula():
  x = RNG.standard_normal((1, 6, 6, 1))
  xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)), mode="symmetric")
  win = np.stack([xp[0, i:i + 6, j:j + 6, 0] for i in range(3) for j in range(3)], -1)
  np.testing.assert_allclose(np_ops.variance_feature(x)[0, ..., 0], win.var(-1), atol=1e-12)


# ------------------------------------------------------------------ elementwise
def test_signed_log1p_roundtrip():
  x = np.concatenate([RNG.standard_normal(100) * 30, [0.0, -0.0, 1e-8, -1e4]])
  np.testing.assert_allclose(np_ops.signed_expm1(np_ops.signed_log1p(x)), x, rtol=1e-12, atol=1e-15)
  np.testing.assert_allclose(torch_ops.signed_log1p(_t(x)).numpy(), np_ops.signed_log1p(x), atol=1e-12)
  assert np_ops.signed_log1p(np.array([-3.0]))[0] < 0


# ------------------------------------------------------------------ convolutions
def test_conv2d_same_identity_and_shift():
  x = RNG.standard_normal((1, 5, 6, 2))
  k = np.zeros((3, 3, 2, 2))
  k[1, 1] = np.eye(2)
  np.testing.assert_allclose(np_ops.conv2d_same(x, k), x)
  k = np.zeros((3, 3, 2, 2))
  k[0, 2] = np.eye(2)          # cross-correlation: out[y,x] = in[y-1, x+1], zero outside
  y = np_ops.conv2d_same(x, k)
  np.testing.assert_allclose(y[0, 1:, :-1], x[0, :-1, 1:])
  assert np.all(y[0, 0] == 0) and np.all(y[0, :, -1] == 0)


@pytest.mark.parametrize("ks", [1, 3])
def test_conv2d_backends_agree(ks):
  x = RNG.standard_normal((2, 7, 9, 5))
  k = RNG.standard_normal((ks, ks, 5, 4))
  b = RNG.standard_normal(4)
  a = np_ops.conv2d_same(x, k, b, relu=True)
  t = torch_ops.conv2d_same(_t(x), k, b, relu=True).numpy()
  np.testing.assert_allclose(a, t, atol=1e-10)
  assert a.min() >= 0


@pytest.mark.parametrize("ks", [2, 3])
def test_conv2d_transpose_is_adjoint_of_strided_same_conv(ks):
  """conv2d_transpose(SAME, s2) is defined as the input-gradient of the SAME stride-2 conv:
  <conv_s2(u), v> == <u, conv_transpose(v)>."""
  cin, cout, h, w = 3, 4, 6, 8          # transpose maps cin -> cout, (h, w) -> (2h, 2w)
  k = RNG.standard_normal((ks, ks, cout, cin))
  v = RNG.standard_normal((1, h, w, cin))
  u = RNG.standard_normal((1, 2 * h, 2 * w, cout))
  # forward SAME stride-2 conv of u with kernel [kh,kw,cout(in),cin(out)]: pad tail only (even size)
  tot = max((h - 1) * 2 + ks - 2 * h, 0)
  pb = tot // 2
  up = np.pad(u, ((0, 0), (pb, tot - pb), (pb, tot - pb), (0, 0)))
  fwd = np.zeros((1, h, w, cin))
  for r in range(ks):
    for s in range(ks):
      fwd += up[:, r:r + 2 * h:2, s:s + 2 * w:2, :] @ k[r, s]
  lhs = (fwd * v).sum()
  rhs = (u * np_ops.conv2d_transpose_same_s2(v, k)).sum()
  np.testing.assert_allclose(lhs, rhs, rtol=1e-10)
  np.testing.assert_allclose(np_ops.conv2d_transpose_same_s2(v, k, relu=True),
                             torch_ops.conv2d_transpose_same_s2(_t(v), k, relu=True).numpy(), atol=1e-10)


def test_conv2d_transpose_2x2_is_pixel_shuffle_of_1x1():
  x = RNG.standard_normal((1, 3, 4, 5))
  k = RNG.standard_normal((2, 2, 6, 5))
  y = np_ops.conv2d_transpose_same_s2(x, k)
  for a in range(2):
    for b in range(2):
      np.testing.assert_allclose(y[:, a::2, b::2], x @ k[a, b].T, atol=1e-12)


# ------------------------------------------------------------------ pooling / resampling
def test_maxpool3_same_pads_tail_only():
  x = np.arange(36, dtype=np.float64).reshape(1, 6, 6, 1)
  y = np_ops.max_pool_same_s2(x, 3)
  assert y.shape == (1, 3, 3, 1)
  # window of output i covers rows 2i..2i+2 clipped: top-left output is max of rows 0-2, cols 0-2
  assert y[0, 0, 0, 0] == 14 and y[0, 2, 2, 0] == 35 and y[0, 1, 0, 0] == 26
  np.testing.assert_array_equal(y, torch_ops.max_pool_same_s2(_t(x), 3).numpy())
  # not torch's max_pool2d(3, 2, padding=1)
  alt = torch.nn.functional.max_pool2d(_t(x).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).numpy()
  assert not np.array_equal(y, alt)


def test_maxpool_odd_sizes_and_k2():
  x = RNG.standard_normal((2, 7, 5, 3))
  for k in (2, 3):
    a = np_ops.max_pool_same_s2(x, k)
    assert a.shape == (2, 4, 3, 3)
    np.testing.assert_array_equal(a, torch_ops.max_pool_same_s2(_t(x), k).numpy())


def test_scale_down_of_scale_up_is_identity():
  x = RNG.standard_normal((1, 5, 6, 3))
  np.testing.assert_allclose(np_ops.avg_pool_same(np_ops.resize_nearest_x2(x), 2), x, atol=1e-12)
  np.testing.assert_allclose(torch_ops.avg_pool_same(torch_ops.resize_nearest_x2(_t(x)), 2).numpy(), x, atol=1e-12)


def test_avgpool_same_excludes_padding_from_divisor():
  x = np.ones((1, 5, 7, 1))
  for f in (2, 4):
    np.testing.assert_allclose(np_ops.avg_pool_same(x, f), 1.0)
    np.testing.assert_allclose(torch_ops.avg_pool_same(_t(x), f).numpy(), 1.0)


# ------------------------------------------------------------------ loss (LossDifference.py:15-36)
@pytest.mark.parametrize("kind", ["DIFFERENCE", "ABSOLUTE", "SMOOTH_ABSOLUTE", "SQUARED", "SMAPE"])
def test_loss_difference_backends_agree(kind):
  p = RNG.standard_normal((2, 4, 5, 3)) * 2
  t = RNG.standard_normal((2, 4, 5, 3)) * 2
  a = np_ops.loss_difference(p, t, kind)
  assert a.shape == (2, 4, 5)
  np.testing.assert_allclose(a, torch_ops.loss_difference(_t(p), _t(t), kind).numpy(), atol=1e-12)
  if kind == "SMAPE":
    assert np.all(a <= 3.0) and np.all(a >= 0)
    np.testing.assert_allclose(np_ops.loss_difference(p, p, kind), 0)


def test_ms_ssim_oracle_known_answers():
  """tf.image.ssim_multiscale restated (oracle/reference_loss.py): identical images score exactly 1, the Gaussian is
  normalised with centre weight (1 / sum_i exp(-i^2 / 4.5))^2, the score is symmetric in its arguments, falls with noise, and a
  constant offset only touches the luminance term of the LAST scale (cs is offset invariant)."""
  import torch
  from oracle import reference_loss as rl
  g = rl._fspecial_gauss(11, 1.5, torch.float64)
  c = torch.arange(11, dtype=torch.float64) - 5
  assert abs(float(g.sum()) - 1.0) < 1e-12 and abs(float(g[5, 5]) - float(1.0 / torch.exp(-c ** 2 / 4.5).sum()) ** 2) < 1e-12
  torch.manual_seed(0)
  x = torch.rand(2, 48, 52, 3, dtype=torch.float64)
  assert float((rl.ssim_multiscale(x, x) - 1).abs().max()) < 1e-12 and abs(float(rl.ms_ssim_loss(x, x))) < 1e-12
  y1, y2 = x + 0.05 * torch.randn_like(x), x + 0.2 * torch.randn_like(x)
  s1, s2 = rl.ssim_multiscale(x, y1), rl.ssim_multiscale(x, y2)
  assert bool((s1 > s2).all()) and bool((s1 < 1).all())
  assert float((rl.ssim_multiscale(x, y1) - rl.ssim_multiscale(y1, x)).abs().max()) < 1e-12
  # offset: structure terms unchanged, so only the luminance factor of the last scale (power 0.3001) can lower the score
  off = rl.ssim_multiscale(x, x + 0.3)
  assert bool((off < 1).all()) and bool((off > 0.5).all())


def test_two_restatements_agree_on_random_shapes():
  """Property: the numpy (im2col / loops) and the torch (F.conv2d / pooling) restatements of the TF operators agree to 1e-10
  in float64 on random shapes, including odd sizes where TF 'SAME' pads only the tail."""
  import torch
  from hypothesis import given, settings, strategies as st
  from oracle import torch_ops

  @settings(max_examples=40, deadline=None)
  @given(n=st.integers(1, 2), h=st.integers(2, 13), w=st.integers(2, 13), cin=st.integers(1, 5), cout=st.integers(1, 5),
         ks=st.sampled_from([1, 3]), seed=st.integers(0, 10 ** 6))
  def check(n, h, w, cin, cout, ks, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, h, w, cin))
    k = rng.standard_normal((ks, ks, cin, cout))
    b = rng.standard_normal(cout)
    tx = torch.from_numpy(x)
    close = lambda a, t: np.abs(a - t.numpy()).max() <= 1e-10   # noqa: E731
    assert close(np_ops.conv2d_same(x, k, b, relu=True), torch_ops.conv2d_same(tx, torch.from_numpy(k), torch.from_numpy(b), relu=True))
    for pk in (2, 3):
      assert close(np_ops.max_pool_same_s2(x, pk), torch_ops.max_pool_same_s2(tx, pk))
    assert close(np_ops.avg_pool_same(x, 2), torch_ops.avg_pool_same(tx, 2))
    kt = rng.standard_normal((3, 3, cout, cin))
    assert close(np_ops.conv2d_transpose_same_s2(x, kt, b), torch_ops.conv2d_transpose_same_s2(tx, torch.from_numpy(kt), torch.from_numpy(b)))
    ksz = 3
    logits = rng.standard_normal((n, h, w, ksz * ksz))
    src = rng.standard_normal((n, h, w, 3))
    if h >= 2 and w >= 2:
      assert close(np_ops.kernel_prediction(src, logits, ksz), torch_ops.kernel_prediction(torch.from_numpy(src), torch.from_numpy(logits), ksz))
    assert close(np_ops.variance_feature(src), torch_ops.variance_feature(torch.from_numpy(src)))

  check()
