"""ORACLE (test infrastructure, not product code) - numpy restatement of the TensorFlow-1.x operators the
DeepDenoiser hot path calls.

PARITY: the functions that restate REFERENCE code (kernel_prediction, variance_feature, loss_difference,
signed_log1p / signed_expm1, and - through oracle/reference_model.py - compose_scales and the whole predict path)
are pinned against that code itself, executed in this container over oracle/tf_shim
(tests/golden/refshim_components.npz, tests/test_reference_golden.py: 1e-12).  The functions that restate
TENSORFLOW kernels (conv2d_same, conv2d_transpose_same_s2, max_pool_same_s2, avg_pool_same, pad_symmetric,
resize_nearest_x2) remain PARITY UNPINNED: the reference ships no tests / golden vectors and TensorFlow 1.x cannot
be installed here, so they are held only by (a) the TF semantics written down in SURVEY.md Appendix A, (b) the
known-answer / property tests in tests/test_oracle_ops.py, (c) agreement with the independent torch-CPU formulation
in oracle/torch_ops.py and with the shim's third formulation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  All tensors are NHWC numpy arrays; `dtype` of the inputs is preserved (float64 for the oracle
proper, float32 to mimic the reference's arithmetic type).
Every function cites the reference call site (file:line under /root/reference/TensorFlow) it restates.
"""
import numpy as np


# ---------------------------------------------------------------------------------------------- padding
def same_padding(n, k, stride):
  """TF 'SAME' padding for one spatial dim: (out, pad_before, pad_after)."""
  out = -(-n // stride)
  total = max((out - 1) * stride + k - n, 0)
  return out, total // 2, total - total // 2


def pad_symmetric(x, pad):
  """Conv2dUtilities.pad_equally(mode='symmetric') (Conv2dUtilities.py:77-95): mirror incl. the edge."""
  return np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0)), mode="symmetric")


# ---------------------------------------------------------------------------------------------- conv
def _im2col(xp, kh, kw, oh, ow):
  n, _, _, c = xp.shape
  cols = np.empty((n, oh, ow, kh, kw, c), dtype=xp.dtype)
  for r in range(kh):
    for s in range(kw):
      cols[:, :, :, r, s, :] = xp[:, r:r + oh, s:s + ow, :]
  return cols.reshape(n * oh * ow, kh * kw * c)


def conv2d_same(x, kernel, bias=None, relu=False):
  """tf.layers.conv2d(padding='same', strides=1) (UNet.py:29-31, Tiramisu.py:35-37, Architecture.py:238-243).
  kernel: TF layout [kh, kw, cin, cout]; cross-correlation; bias then activation."""
  kh, kw, cin, cout = kernel.shape
  n, h, w, c = x.shape
  assert c == cin, (c, cin)
  _, pt, pb = same_padding(h, kh, 1)
  _, pl, pr = same_padding(w, kw, 1)
  xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
  y = _im2col(xp, kh, kw, h, w) @ kernel.reshape(kh * kw * cin, cout).astype(x.dtype)
  y = y.reshape(n, h, w, cout)
  if bias is not None:
    y = y + bias.astype(x.dtype)
  if relu:
    y = np.maximum(y, 0)
  return y


def conv2d_transpose_same_s2(x, kernel, bias=None, relu=False):
  """tf.layers.conv2d_transpose(strides=2, padding='same') (UNet.py:56-58 k=2, Tiramisu.py:62-64 k=3).
  kernel: TF layout [kh, kw, cout, cin].  Defined as the input-gradient of the SAME stride-2 conv:
  out[2i+r, 2j+s, o] += sum_c x[i,j,c] * kernel[r,s,o,c], full size cropped at the TAIL to 2n."""
  kh, kw, cout, cin = kernel.shape
  n, h, w, c = x.shape
  assert c == cin
  full = np.zeros((n, 2 * h + kh, 2 * w + kw, cout), dtype=x.dtype)
  for r in range(kh):
    for s in range(kw):
      contrib = x.reshape(-1, cin) @ kernel[r, s].T.astype(x.dtype)  # [pixels, cout]
      full[:, r:r + 2 * h:2, s:s + 2 * w:2, :] += contrib.reshape(n, h, w, cout)
  y = full[:, :2 * h, :2 * w, :]
  if bias is not None:
    y = y + bias.astype(x.dtype)
  if relu:
    y = np.maximum(y, 0)
  return y


# ---------------------------------------------------------------------------------------------- pooling
def max_pool_same_s2(x, k):
  """tf.layers.max_pooling2d(pool_size=k, strides=2, padding='same') (UNet.py:42-44 k=3, Tiramisu.py:55-57 k=2).
  Padding never wins the max (-inf)."""
  n, h, w, c = x.shape
  oh, pt, pb = same_padding(h, k, 2)
  ow, pl, pr = same_padding(w, k, 2)
  xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)), constant_values=-np.inf)
  y = np.full((n, oh, ow, c), -np.inf, dtype=x.dtype)
  for r in range(k):
    for s in range(k):
      y = np.maximum(y, xp[:, r:r + 2 * oh:2, s:s + 2 * ow:2, :][:, :oh, :ow, :])
  return y


def avg_pool_same(x, f):
  """MultiScalePrediction.scale_down (MultiScalePrediction.py:11-13): average_pooling2d(f, f, 'same');
  padded cells are excluded from the divisor."""
  n, h, w, c = x.shape
  oh, pt, pb = same_padding(h, f, f)
  ow, pl, pr = same_padding(w, f, f)
  xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
  ones = np.pad(np.ones((1, h, w, 1), dtype=x.dtype), ((0, 0), (pt, pb), (pl, pr), (0, 0)))
  s = xp.reshape(n, oh, f, ow, f, c).sum(axis=(2, 4))
  cnt = ones.reshape(1, oh, f, ow, f, 1).sum(axis=(2, 4))
  return s / cnt


def resize_nearest_x2(x):
  """MultiScalePrediction.scale_up (MultiScalePrediction.py:16-33): NEAREST_NEIGHBOR, align_corners=False."""
  return np.repeat(np.repeat(x, 2, axis=1), 2, axis=2)


# ---------------------------------------------------------------------------------------------- elementwise
def signed_log1p(x):
  """Utilities.signed_log1p (Utilities.py:3-4)."""
  return np.sign(x) * np.log1p(np.abs(x))


def signed_expm1(x):
  """Utilities.signed_expm1 (Utilities.py:6-7)."""
  return np.sign(x) * np.expm1(np.abs(x))


def softmax_channels(x):
  """tf.nn.softmax(axis=channel) (KernelPrediction.py:22-23)."""
  m = x.max(axis=-1, keepdims=True)
  e = np.exp(x - m)
  return e / e.sum(axis=-1, keepdims=True)


# ---------------------------------------------------------------------------------------------- kernel prediction
def kernel_prediction(inputs, kernel_inputs, kernel_size, use_softmax=True):
  """KernelPrediction.kernel_prediction (KernelPrediction.py:11-63), written the way the reference does:
  softmax, symmetric pad, K*K shifted slices stacked in (i row-major, j) order, multiply, reduce_sum,
  the same kernel for every colour channel."""
  n, h, w, c = inputs.shape
  assert kernel_inputs.shape == (n, h, w, kernel_size ** 2)
  pad = (kernel_size - 1) // 2
  weights = softmax_channels(kernel_inputs) if use_softmax else kernel_inputs
  padded = pad_symmetric(inputs, pad)
  out = np.zeros_like(inputs)
  for i in range(kernel_size):
    for j in range(kernel_size):
      out += padded[:, i:i + h, j:j + w, :] * weights[:, :, :, i * kernel_size + j][..., None]
  return out


# ---------------------------------------------------------------------------------------------- variance feature
def local_mean(x, variance_mode="uniform"):
  """FeatureEngineering._local_mean (FeatureEngineering.py:11-55): 3x3 symmetric-padded box ('uniform')
  or plus-shaped ('neighbor') mean, per channel."""
  n, h, w, c = x.shape
  if variance_mode == "uniform":
    filt = np.ones((3, 3))
  else:
    assert variance_mode == "neighbor"
    filt = np.array([[0., 1., 0.], [1., 1., 1.], [0., 1., 0.]])
  filt = filt / filt.sum()
  xp = pad_symmetric(x, 1)
  out = np.zeros_like(x)
  for r in range(3):
    for s in range(3):
      if filt[r, s] != 0:
        out += xp[:, r:r + h, s:s + w, :] * x.dtype.type(filt[r, s])
  return out


def variance_feature(x, variance_mode="uniform", relative_variance=False, compress_to_one_channel=False,
                     epsilon=1e-4):
  """FeatureEngineering.variance (FeatureEngineering.py:57-70)."""
  mean = local_mean(x, variance_mode)
  sq_mean = mean * mean
  result = local_mean(x * x, variance_mode) - sq_mean
  if relative_variance:
    result = result / np.maximum(sq_mean, x.dtype.type(epsilon))
  if compress_to_one_channel:
    result = result.mean(axis=-1, keepdims=True)
  return result


# ---------------------------------------------------------------------------------------------- loss
def loss_difference(predicted, target, kind, epsilon=1e-2):
  """LossDifference.difference (LossDifference.py:15-36): element difference then reduce_sum over channels."""
  d = predicted - target
  if kind == "DIFFERENCE":
    r = d
  elif kind == "ABSOLUTE":
    r = np.abs(d)
  elif kind == "SMOOTH_ABSOLUTE":
    a = np.abs(d)
    r = np.where(a < 1, 0.5 * a * a, a - 0.5)
  elif kind == "SQUARED":
    r = d * d
  elif kind == "SMAPE":
    r = np.abs(d) / (np.abs(predicted) + np.abs(target) + epsilon)
  else:
    raise ValueError(kind)
  return r.sum(axis=3)


# ---------------------------------------------------------------------------------------------- backend helpers
# (oracle/reference_model.py is written against this small surface so it can also run on oracle/torch_ops.py)
def asarray(x, dtype):
  return np.asarray(x, dtype=dtype)


def to_numpy(x):
  return np.asarray(x)


def concat(xs, axis=-1):
  return np.concatenate(list(xs), axis=axis)


def relu(x):
  return np.maximum(x, 0)


def sigmoid(x):
  return 1.0 / (1.0 + np.exp(-x))


def tile_hw(row, n, h, w):
  """row [C] -> [n,h,w,C] (FeatureFlags.feature_flags tiling, FeatureFlags.py:59-66)."""
  return np.broadcast_to(row.reshape(1, 1, 1, -1), (n, h, w, row.shape[0]))


def full(shape, value, dtype):
  return np.full(tuple(shape), value, dtype=dtype)
