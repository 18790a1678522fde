"""ORACLE (test infrastructure, not product code) - torch-CPU restatement of the TensorFlow-1.x operators on the
DeepDenoiser hot path, written independently of oracle/np_ops.py (F.conv2d / F.max_pool2d / unfold
formulations instead of im2col + slicing).  Same function names and NHWC convention as np_ops, so
oracle/reference_model.py can run on either backend:

  * np_ops  + float64  -> the oracle proper
  * torch_ops + float32 -> the "restated reference on CPU" that bench.py times as cpu_baseline / --impl reference
    (oneDNN kernels on all host threads; the closest thing to the reference's TF-CPU path that can run here)

PARITY: see np_ops.py (reference code pinned through oracle/tf_shim; TensorFlow's kernels unpinned).  tests/test_oracle_ops.py requires the two backends to agree.
Reference call sites are cited per function (file:line under /root/reference/TensorFlow).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _nchw(x):
  return x.permute(0, 3, 1, 2)


def _nhwc(x):
  return x.permute(0, 2, 3, 1).contiguous()


def asarray(x, dtype):
  if isinstance(x, torch.Tensor):
    return x.to(dtype)          # keeps the autograd graph (gradient oracle for the training path)
  return torch.as_tensor(np.asarray(x), dtype=dtype)


def to_numpy(x):
  return x.detach().cpu().numpy()


def same_padding(n, k, stride):
  out = -(-n // stride)
  total = max((out - 1) * stride + k - n, 0)
  return out, total // 2, total - total // 2


def pad_symmetric(x, pad):
  """Conv2dUtilities.pad_equally(mode='symmetric') (Conv2dUtilities.py:77-95).  torch has no edge-including
  mirror mode, so it is assembled from flipped border strips."""
  if pad == 0:
    return x
  h, w = x.shape[1], x.shape[2]
  assert pad <= h and pad <= w
  top = torch.flip(x[:, :pad], dims=[1])
  bottom = torch.flip(x[:, h - pad:], dims=[1])
  x = torch.cat([top, x, bottom], dim=1)
  left = torch.flip(x[:, :, :pad], dims=[2])
  right = torch.flip(x[:, :, w - pad:], dims=[2])
  return torch.cat([left, x, right], dim=2)


def conv2d_same(x, kernel, bias=None, relu=False):
  """tf.layers.conv2d(padding='same', strides=1) (UNet.py:29-31, Tiramisu.py:35-37,50-52,77-79,
  Architecture.py:238-243, MultiScalePrediction.py:64-66,73-75,88-90); kernel TF layout [kh,kw,cin,cout]."""
  kernel = torch.as_tensor(kernel, dtype=x.dtype)
  kh, kw = kernel.shape[0], kernel.shape[1]
  w = kernel.permute(3, 2, 0, 1).contiguous()
  b = None if bias is None else torch.as_tensor(bias, dtype=x.dtype)
  y = F.conv2d(_nchw(x), w, b, padding=(kh // 2, kw // 2))
  if relu:
    y = F.relu(y)
  return _nhwc(y)


def conv2d_transpose_same_s2(x, kernel, bias=None, relu=False):
  """tf.layers.conv2d_transpose(strides=2, padding='same') (UNet.py:56-58 k=2, Tiramisu.py:62-64 k=3);
  kernel TF layout [kh,kw,cout,cin]; full output cropped at the tail."""
  kernel = torch.as_tensor(kernel, dtype=x.dtype)
  n, h, w_, _ = x.shape
  wt = kernel.permute(3, 2, 0, 1).contiguous()  # torch conv_transpose weight: [cin, cout, kh, kw]
  b = None if bias is None else torch.as_tensor(bias, dtype=x.dtype)
  y = F.conv_transpose2d(_nchw(x), wt, b, stride=2, padding=0)[:, :, :2 * h, :2 * w_]
  if relu:
    y = F.relu(y)
  return _nhwc(y)


def max_pool_same_s2(x, k):
  """tf.layers.max_pooling2d(k, 2, 'same') (UNet.py:42-44, Tiramisu.py:55-57): -inf padding, tail-heavy."""
  _, h, w, _ = x.shape
  _, pt, pb = same_padding(h, k, 2)
  _, pl, pr = same_padding(w, k, 2)
  xp = F.pad(_nchw(x), (pl, pr, pt, pb), value=float("-inf"))
  return _nhwc(F.max_pool2d(xp, k, 2))


def avg_pool_same(x, f):
  """MultiScalePrediction.scale_down (MultiScalePrediction.py:11-13)."""
  _, h, w, _ = x.shape
  _, pt, pb = same_padding(h, f, f)
  _, pl, pr = same_padding(w, f, f)
  xp = F.pad(_nchw(x), (pl, pr, pt, pb))
  ones = F.pad(torch.ones(1, 1, h, w, dtype=x.dtype), (pl, pr, pt, pb))
  return _nhwc(F.avg_pool2d(xp, f, f) / F.avg_pool2d(ones, f, f))


def resize_nearest_x2(x):
  """MultiScalePrediction.scale_up (MultiScalePrediction.py:16-33)."""
  return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def signed_log1p(x):
  """Utilities.signed_log1p (Utilities.py:3-4)."""
  return torch.sign(x) * torch.log1p(torch.abs(x))


def signed_expm1(x):
  """Utilities.signed_expm1 (Utilities.py:6-7)."""
  return torch.sign(x) * torch.expm1(torch.abs(x))


def softmax_channels(x):
  return torch.softmax(x, dim=-1)


def kernel_prediction(inputs, kernel_inputs, kernel_size, use_softmax=True):
  """KernelPrediction.kernel_prediction (KernelPrediction.py:11-63) via F.unfold: patches are laid out
  (channel, i, j) with i the row offset, matching the reference's i*K+j stacking order."""
  n, h, w, c = inputs.shape
  k2 = kernel_size * kernel_size
  assert tuple(kernel_inputs.shape) == (n, h, w, k2)
  pad = (kernel_size - 1) // 2
  weights = softmax_channels(kernel_inputs) if use_softmax else kernel_inputs
  padded = _nchw(pad_symmetric(inputs, pad))
  patches = F.unfold(padded, kernel_size).reshape(n, c, k2, h, w)
  out = (patches * weights.permute(0, 3, 1, 2).unsqueeze(1)).sum(dim=2)
  return _nhwc(out)


def local_mean(x, variance_mode="uniform"):
  """FeatureEngineering._local_mean (FeatureEngineering.py:11-55)."""
  c = x.shape[3]
  if variance_mode == "uniform":
    filt = torch.ones(3, 3, dtype=x.dtype)
  else:
    assert variance_mode == "neighbor"
    filt = torch.tensor([[0., 1., 0.], [1., 1., 1.], [0., 1., 0.]], dtype=x.dtype)
  filt = filt / filt.sum()
  w = filt.reshape(1, 1, 3, 3).repeat(c, 1, 1, 1)
  return _nhwc(F.conv2d(_nchw(pad_symmetric(x, 1)), w, groups=c))


def variance_feature(x, variance_mode="uniform", relative_variance=False, compress_to_one_channel=False,
                     epsilon=1e-4):
  """FeatureEngineering.variance (FeatureEngineering.py:57-70)."""
  mean = local_mean(x, variance_mode)
  sq_mean = mean * mean
  result = local_mean(x * x, variance_mode) - sq_mean
  if relative_variance:
    result = result / torch.clamp(sq_mean, min=epsilon)
  if compress_to_one_channel:
    result = result.mean(dim=-1, keepdim=True)
  return result


def loss_difference(predicted, target, kind, epsilon=1e-2):
  """LossDifference.difference (LossDifference.py:15-36)."""
  d = predicted - target
  if kind == "DIFFERENCE":
    r = d
  elif kind == "ABSOLUTE":
    r = d.abs()
  elif kind == "SMOOTH_ABSOLUTE":
    a = d.abs()
    r = torch.where(a < 1, 0.5 * a * a, a - 0.5)
  elif kind == "SQUARED":
    r = d * d
  elif kind == "SMAPE":
    r = d.abs() / (predicted.abs() + target.abs() + epsilon)
  else:
    raise ValueError(kind)
  return r.sum(dim=3)


# -- small tensor helpers reference_model.py needs from a backend
def concat(xs, axis=-1):
  return torch.cat(list(xs), dim=axis)


def relu(x):
  return F.relu(x)


def sigmoid(x):
  return torch.sigmoid(x)


def sqrt_scalar(v, like):
  return float(np.sqrt(np.asarray(v, dtype=np.float64))) if like.dtype == torch.float64 else float(np.sqrt(np.float32(v)))


def tile_hw(row, n, h, w):
  """row [C] -> [n,h,w,C] (FeatureFlags.feature_flags tiling, FeatureFlags.py:59-66)."""
  return row.reshape(1, 1, 1, -1).expand(n, h, w, -1)


def full(shape, value, dtype):
  return torch.full(tuple(shape), float(value), dtype=dtype)
