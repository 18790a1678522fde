"""TEST INFRASTRUCTURE - not part of the product, never imported by deepdenoiser_b200/.

A minimal EAGER stand-in for the ~100 TensorFlow 1.x symbols that the reference's prediction, loss, augmentation and tiling
code touches (/root/reference/TensorFlow/{Architecture,UNet,Tiramisu,KernelPrediction,MultiScalePrediction,FeatureEngineering,
Conv2dUtilities,SourceEncoder,FeatureFlags,Utilities,LossDifference,DataAugmentation}.py, Training.py up to its Estimator and its
model_fn, Prediction.py's main()), backed by torch CPU tensors.  With this directory first on sys.path, `import tensorflow as tf` inside the UNMODIFIED reference
modules resolves to this file, so the reference's own Python - its control flow, slicing arithmetic, scope / variable
naming, data-format conversions - executes here and produces the vectors of tests/golden/refshim_*.npz
(tests/golden/make_reference_golden.py).  TensorFlow itself cannot be installed in this image.

What this pins and what it does not: everything the reference WRITES is executed as written; what TensorFlow's kernels DO
(SAME padding of conv / pool / transposed conv, symmetric tf.pad, nearest-neighbour resize, default layer naming) is
restated below from TensorFlow 1.x's documented behaviour, each op a few lines, and stays unverified against a real
TensorFlow build.

Graph-mode notions collapse: tensors are torch tensors, name scopes are no-ops, a variable is looked up by its full
TF name in `VARIABLES` (filled by the caller before the model runs; a missing name raises, `created` records the order in
which names were requested).
"""
import collections
import contextlib

import torch
import torch.nn.functional as F

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
AUTO_REUSE = "AUTO_REUSE"

COMPUTE_DTYPE = torch.float64    # what `tf.float32` means here: the vectors are generated in float64 end to end, so constants the
                                 # reference creates as float32 (tf.ones / tf.zeros / tf.reshape of a list) must not inject
                                 # float32 rounding (1/9 of the variance filter: 7e-8) into an otherwise float64 evaluation


def _f(dtype):
  return COMPUTE_DTYPE if dtype in (None, torch.float32) else dtype


VARIABLES = {}       # full TF variable name -> torch tensor (TF layouts: conv [kh,kw,cin,cout], transposed conv [kh,kw,cout,cin])
created = []         # names in first-request order


def reset(variables=None):
  VARIABLES.clear()
  if variables:
    VARIABLES.update(variables)
  del created[:]
  _scope_stack[:] = [""]
  _scope_counts.clear()


# ------------------------------------------------------------------------------------------------ scopes and variables
# tf.variable_scope as in tensorflow/python/ops/variable_scope.py (1.x): a default-named scope ("conv2d") is made unique
# against the per-graph counts of OPENED scope names under the current scope ("conv2d", "conv2d_1", ...); leaving a scope
# zeroes the counts of its sub-scopes (close_variable_subscopes), so re-entering 'reused_core_architecture' numbers its
# layers from "conv2d" again - which is what makes `reuse=True` find the variables of the first pass.
_scope_stack = [""]
_scope_counts = collections.defaultdict(int)


def _join(a, b):
  return a + "/" + b if a else b


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, reuse=None):
  cur = _scope_stack[-1]
  if name_or_scope is None:
    base = _join(cur, default_name)
    name = default_name
    if _scope_counts[base] > 0:
      idx = 1
      while _scope_counts[base + "_%d" % idx] > 0:
        idx += 1
      name = default_name + "_%d" % idx
  else:
    name = name_or_scope
  full = _join(cur, name)
  _scope_counts[full] += 1
  _scope_stack.append(full)
  try:
    yield full
  finally:
    _scope_stack.pop()
    for k in list(_scope_counts):
      if k.startswith(full + "/"):
        _scope_counts[k] = 0


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
  yield name


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
  full = _join(_scope_stack[-1], name)
  if full not in created:
    created.append(full)
  if full not in VARIABLES:
    raise KeyError("variable '%s' was requested by the reference but not provided" % full)
  v = VARIABLES[full]
  if shape is not None and tuple(int(s) for s in shape) != tuple(v.shape):
    raise ValueError("variable '%s': the reference asks for shape %s, provided %s" % (full, tuple(shape), tuple(v.shape)))
  return v


# ------------------------------------------------------------------------------------------------ element-wise
def _t(x, like=None):
  if isinstance(x, torch.Tensor):
    return x
  return torch.as_tensor(x, dtype=like.dtype if isinstance(like, torch.Tensor) and like.is_floating_point() else None)


def add(x, y, name=None): return _t(x, y) + _t(y, x)
def subtract(x, y, name=None): return _t(x, y) - _t(y, x)
def multiply(x, y, name=None): return _t(x, y) * _t(y, x)
def divide(x, y, name=None): return _t(x, y) / _t(y, x)
def scalar_mul(scalar, x, name=None): return x * scalar
def squared_difference(x, y, name=None): return (x - y) * (x - y)
def abs(x, name=None): return torch.abs(x)                      # noqa: A001
def sign(x, name=None): return torch.sign(x)
def square(x, name=None): return x * x
def _c(x): return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=COMPUTE_DTYPE)     # python scalars: compute dtype
def sqrt(x, name=None): return torch.sqrt(_c(x))
def log(x, name=None): return torch.log(x)
def log1p(x, name=None): return torch.log1p(x)
def expm1(x, name=None): return torch.expm1(x)
def exp(x, name=None): return torch.exp(x)
def sigmoid(x, name=None): return torch.sigmoid(x)
def minimum(x, y, name=None): return torch.minimum(_t(x, y), _t(y, x).to(_t(x, y).dtype))
def maximum(x, y, name=None): return torch.maximum(_t(x, y), _t(y, x).to(_t(x, y).dtype))
def negative(x, name=None): return -x
def sin(x, name=None): return torch.sin(_c(x))
def cos(x, name=None): return torch.cos(_c(x))
def matmul(a, b, name=None): return torch.matmul(a, b.to(a.dtype))
def equal(x, y, name=None): return _t(x) == _t(y)
def less(x, y, name=None): return _t(x, y) < _t(y, x)
def greater(x, y, name=None): return _t(x, y) > _t(y, x)
def where(condition, x=None, y=None, name=None): return torch.where(condition, x, y)
def add_n(inputs, name=None):
  total = 0.
  for v in inputs:
    total = total + v
  return total if isinstance(total, torch.Tensor) else torch.tensor(total, dtype=torch.float64)


def _axes(axis):
  return None if axis is None else (tuple(axis) if isinstance(axis, (list, tuple)) else int(axis))


def reduce_sum(x, axis=None, keepdims=False, name=None, keep_dims=None):
  keep = bool(keepdims or keep_dims)
  return torch.sum(x) if axis is None else torch.sum(x, dim=_axes(axis), keepdim=keep)


def reduce_mean(x, axis=None, keepdims=False, name=None, keep_dims=None):
  keep = bool(keepdims or keep_dims)
  return torch.mean(x) if axis is None else torch.mean(x, dim=_axes(axis), keepdim=keep)


# ------------------------------------------------------------------------------------------------ shapes and layout
def shape(x, name=None): return [int(s) for s in x.shape]
def constant(value, dtype=None, shape=None, name=None):             # noqa: A002
  t = torch.as_tensor(value, dtype=_f(dtype) if isinstance(value, float) or dtype is not None else None)
  return t.reshape(shape) if shape is not None else t
def ones(shape, dtype=float32, name=None): return torch.ones([int(s) for s in shape], dtype=_f(dtype))     # noqa: A002
def zeros(shape, dtype=float32, name=None): return torch.zeros([int(s) for s in shape], dtype=_f(dtype))   # noqa: A002
def reshape(tensor, shape, name=None):                               # noqa: A002
  if isinstance(tensor, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in tensor):
    tensor = torch.stack([_t(v).to(COMPUTE_DTYPE) for v in tensor])
  t = tensor if isinstance(tensor, torch.Tensor) else torch.tensor(tensor, dtype=COMPUTE_DTYPE)
  return t.reshape([int(s) for s in shape])
def transpose(a, perm=None, name=None): return a.permute(*perm) if perm is not None else a.t()
def concat(values, axis, name=None): return torch.cat(list(values), dim=int(axis))
def stack(values, axis=0, name=None): return torch.stack(list(values), dim=int(axis))
def tile(input, multiples, name=None): return input.repeat(*[int(m) for m in multiples])               # noqa: A002
def identity(x, name=None): return x


def split(value, num_or_size_splits, axis=0, num=None, name=None):
  if isinstance(num_or_size_splits, int):
    n = value.shape[axis]
    assert n % num_or_size_splits == 0, "tf.split: %d does not divide %d" % (num_or_size_splits, n)
    return list(torch.split(value, n // num_or_size_splits, dim=axis))
  return list(torch.split(value, [int(s) for s in num_or_size_splits], dim=axis))


def slice(input_, begin, size, name=None):                          # noqa: A001
  index = []
  for b, s in zip(begin, size):
    b, s = int(b), int(s)
    index.append(builtins_slice(b, None if s < 0 else b + s))
  return input_[tuple(index)]


import builtins as _builtins  # noqa: E402
builtins_slice = _builtins.slice


def pad(tensor, paddings, mode="CONSTANT", name=None, constant_values=0):
  """tf.pad: CONSTANT / REFLECT (mirror without the edge) / SYMMETRIC (mirror including the edge), case-insensitive."""
  mode = mode.upper()
  out = tensor
  for axis, (before, after) in enumerate(paddings):
    before, after = int(before), int(after)
    if before == 0 and after == 0:
      continue
    n = out.shape[axis]
    if mode == "CONSTANT":
      shp = list(out.shape)
      parts = []
      if before:
        shp[axis] = before
        parts.append(torch.full(shp, constant_values, dtype=out.dtype))
      parts.append(out)
      if after:
        shp[axis] = after
        parts.append(torch.full(shp, constant_values, dtype=out.dtype))
      out = torch.cat(parts, dim=axis)
      continue
    if mode == "SYMMETRIC":
      assert before <= n and after <= n
      idx = list(range(before - 1, -1, -1)) + list(range(n)) + list(range(n - 1, n - 1 - after, -1))
    elif mode == "REFLECT":
      assert before < n and after < n
      idx = list(range(before, 0, -1)) + list(range(n)) + list(range(n - 2, n - 2 - after, -1))
    else:
      raise ValueError("tf.pad mode '%s'" % mode)
    out = out.index_select(axis, torch.tensor(idx, dtype=torch.long))
  return out


def map_fn(fn, elems, dtype=None, name=None):
  return torch.stack([fn(elems[i]) for i in range(elems.shape[0])], dim=0)


def cond(pred, true_fn=None, false_fn=None, name=None):
  return true_fn() if bool(pred) else false_fn()


def case(pred_fn_pairs, default=None, exclusive=False, name=None):
  hits = [fn for pred, fn in pred_fn_pairs if bool(pred)]
  if exclusive and len(hits) > 1:
    raise ValueError("tf.case(exclusive=True): more than one predicate is true")
  return hits[0]() if hits else default()


# ------------------------------------------------------------------------------------------------ SAME padding (TF)
def _same_pad(size, k, s):
  """TensorFlow SAME: out = ceil(size / s); total = max((out - 1) s + k - size, 0); the smaller half goes in front."""
  out = -(-size // s)
  total = max((out - 1) * s + k - size, 0)
  return total // 2, total - total // 2


def _pair(v):
  return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _to_nchw(x, data_format):
  return x.permute(0, 3, 1, 2) if data_format in ("channels_last", "NHWC") else x


def _from_nchw(x, data_format):
  return x.permute(0, 2, 3, 1) if data_format in ("channels_last", "NHWC") else x


def _conv2d_nchw(x, kernel_hwio, strides, padding):
  kh, kw = int(kernel_hwio.shape[0]), int(kernel_hwio.shape[1])
  w = kernel_hwio.permute(3, 2, 0, 1).to(x.dtype)
  if padding.upper() == "SAME":
    pt, pb = _same_pad(x.shape[2], kh, strides[0])
    pl, pr = _same_pad(x.shape[3], kw, strides[1])
    x = F.pad(x, (pl, pr, pt, pb))
  else:
    assert padding.upper() == "VALID"
  return F.conv2d(x, w, stride=strides)


class _NN(object):
  relu = staticmethod(lambda x, name=None: torch.relu(x))
  sigmoid = staticmethod(sigmoid)

  @staticmethod
  def softmax(logits, axis=-1, name=None, dim=None):
    return torch.softmax(logits, dim=int(axis if dim is None else dim))

  @staticmethod
  def embedding_lookup(params, ids, name=None):
    return params[torch.as_tensor(ids, dtype=torch.long)]

  @staticmethod
  def conv2d(input, filter=None, strides=None, padding=None, data_format="NHWC", name=None, filters=None):   # noqa: A002
    k = filter if filter is not None else filters
    x = _to_nchw(input, data_format)
    st = (int(strides[1]), int(strides[2])) if data_format == "NHWC" else (int(strides[2]), int(strides[3]))
    return _from_nchw(_conv2d_nchw(x, k, st, padding), data_format)


nn = _NN()


# ------------------------------------------------------------------------------------------------ tf.layers
def _no_bn(*a, **k):
  raise NotImplementedError("tf.layers.batch_normalization: the reference's Architecture never enables it "
                            "(Architecture.py:506 passes use_batch_normalization=False)")


class _Layers(object):
  batch_normalization = staticmethod(_no_bn)

  @staticmethod
  def dropout(inputs, rate=0.5, training=False, name=None):
    if training and rate > 0.:
      raise NotImplementedError("tf.layers.dropout in training mode (Architecture.py:506 passes dropout_rate=0.)")
    return inputs

  @staticmethod
  def flatten(inputs, name=None):
    return inputs.reshape(inputs.shape[0], -1)

  @staticmethod
  def conv2d(inputs, filters, kernel_size, strides=(1, 1), padding="valid", data_format="channels_last", dilation_rate=(1, 1),
             activation=None, use_bias=True, name=None, reuse=None, **unused):
    kh, kw = _pair(kernel_size)
    x = _to_nchw(inputs, data_format)
    with variable_scope(name, default_name="conv2d"):
      kernel = get_variable("kernel", [kh, kw, int(x.shape[1]), int(filters)])
      bias = get_variable("bias", [int(filters)]) if use_bias else None
    y = _conv2d_nchw(x, kernel, _pair(strides), padding)
    if bias is not None:
      y = y + bias.to(y.dtype).reshape(1, -1, 1, 1)
    if activation is not None:
      y = activation(y)
    return _from_nchw(y, data_format)

  @staticmethod
  def conv2d_transpose(inputs, filters, kernel_size, strides=(1, 1), padding="valid", data_format="channels_last",
                       activation=None, use_bias=True, name=None, reuse=None, **unused):
    """Gradient of conv2d with respect to its input (tf.nn.conv2d_transpose): SAME gives out = in * stride; the forward
    convolution it transposes pads that output size with TF's SAME rule, so the full transposed result is cropped by the
    forward padding."""
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(strides)
    x = _to_nchw(inputs, data_format)
    with variable_scope(name, default_name="conv2d_transpose"):
      kernel = get_variable("kernel", [kh, kw, int(filters), int(x.shape[1])])
      bias = get_variable("bias", [int(filters)]) if use_bias else None
    full = F.conv_transpose2d(x, kernel.permute(3, 2, 0, 1).to(x.dtype), stride=(sh, sw))
    if padding.upper() == "SAME":
      oh, ow = x.shape[2] * sh, x.shape[3] * sw
      pt, _ = _same_pad(oh, kh, sh)
      pl, _ = _same_pad(ow, kw, sw)
      # the full result has (in - 1) s + k rows; rows the forward convolution never reads do not exist when k < s
      full = F.pad(full, (0, max(0, pl + ow - full.shape[3]), 0, max(0, pt + oh - full.shape[2])))
      y = full[:, :, pt:pt + oh, pl:pl + ow]
    else:
      y = full
    if bias is not None:
      y = y + bias.to(y.dtype).reshape(1, -1, 1, 1)
    if activation is not None:
      y = activation(y)
    return _from_nchw(y, data_format)

  @staticmethod
  def max_pooling2d(inputs, pool_size, strides, padding="valid", data_format="channels_last", name=None):
    kh, kw = _pair(pool_size)
    sh, sw = _pair(strides)
    x = _to_nchw(inputs, data_format)
    if padding.upper() == "SAME":          # padded positions never win
      pt, pb = _same_pad(x.shape[2], kh, sh)
      pl, pr = _same_pad(x.shape[3], kw, sw)
      x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
    return _from_nchw(F.max_pool2d(x, (kh, kw), stride=(sh, sw)), data_format)

  @staticmethod
  def average_pooling2d(inputs, pool_size, strides, padding="valid", data_format="channels_last", name=None):
    kh, kw = _pair(pool_size)
    sh, sw = _pair(strides)
    x = _to_nchw(inputs, data_format)
    if padding.upper() == "SAME":          # padded positions are excluded from the divisor
      pt, pb = _same_pad(x.shape[2], kh, sh)
      pl, pr = _same_pad(x.shape[3], kw, sw)
      ones_ = F.pad(torch.ones_like(x[:1, :1]), (pl, pr, pt, pb))
      x = F.pad(x, (pl, pr, pt, pb))
      total = F.avg_pool2d(x, (kh, kw), stride=(sh, sw)) * (kh * kw)
      count = F.avg_pool2d(ones_, (kh, kw), stride=(sh, sw)) * (kh * kw)
      return _from_nchw(total / count, data_format)
    return _from_nchw(F.avg_pool2d(x, (kh, kw), stride=(sh, sw)), data_format)


layers = _Layers()


# ------------------------------------------------------------------------------------------------ tf.image
class _ResizeMethod(object):
  BILINEAR, NEAREST_NEIGHBOR, BICUBIC, AREA = 0, 1, 2, 3


class _Image(object):
  ResizeMethod = _ResizeMethod

  @staticmethod
  def resize_images(images, size, method=0, align_corners=False):
    """NEAREST_NEIGHBOR, align_corners False: out[y, x] = in[floor(y * in_h / out_h), floor(x * in_w / out_w)] (NHWC)."""
    if method != _ResizeMethod.NEAREST_NEIGHBOR or align_corners:
      raise NotImplementedError("only the nearest-neighbour resize of MultiScalePrediction.py:27 is restated")
    batched = images.dim() == 4
    x = images if batched else images[None]
    oh, ow = int(size[0]), int(size[1])
    ih, iw = x.shape[1], x.shape[2]
    ys = torch.tensor([min(ih - 1, (y * ih) // oh) for y in range(oh)], dtype=torch.long)
    xs = torch.tensor([min(iw - 1, (c * iw) // ow) for c in range(ow)], dtype=torch.long)
    y = x.index_select(1, ys).index_select(2, xs)
    return y if batched else y[0]

  @staticmethod
  def flip_left_right(image):
    """Reverses the width axis ([h, w, c] or [n, h, w, c])."""
    return torch.flip(image, dims=[image.dim() - 2])

  @staticmethod
  def rot90(image, k=1, name=None):
    """k counter-clockwise quarter turns of the (height, width) plane."""
    return torch.rot90(image, k=int(k) % 4, dims=(image.dim() - 3, image.dim() - 2))

  @staticmethod
  def ssim_multiscale(*a, **k):
    raise NotImplementedError("tf.image.ssim_multiscale is not restated in the shim (Training.py:200)")


image = _Image()


# ------------------------------------------------------------------------------------------------ estimator / summary stubs
class _ModeKeys(object):
  TRAIN, EVAL, PREDICT = "train", "eval", "infer"


class _Estimator(object):
  ModeKeys = _ModeKeys
  EstimatorSpec = collections.namedtuple("EstimatorSpec", ["mode", "loss", "train_op", "eval_metric_ops", "predictions"])
  EstimatorSpec.__new__.__defaults__ = (None, None, None, None, None)


estimator = _Estimator()


class _Summary(object):
  scalar = staticmethod(lambda *a, **k: None)
  histogram = staticmethod(lambda *a, **k: None)
  image = staticmethod(lambda *a, **k: None)


summary = _Summary()


class _Metrics(object):
  mean = staticmethod(lambda values, *a, **k: (torch.mean(values) if isinstance(values, torch.Tensor) else values, None))


metrics = _Metrics()


# ------------------------------------------------------------------------------------------------ Training.main() plumbing
# Training.py builds its loss objects (FeatureTraining / CombinedFeatureTraining / CombinedImageFeatureTraining) inside main()
# and hands them to tf.estimator.Estimator as `params`.  The stand-in Estimator records (model_fn, params) and stops main()
# right there (SetupCaptured), so the reference's own construction code runs and the caller can invoke the reference's
# model_fn on tensors of its choice.
class SetupCaptured(Exception):
  def __init__(self, model_fn, params):
    Exception.__init__(self, "tf.estimator.Estimator constructed")
    self.model_fn, self.params = model_fn, params


class _Anything(object):
  """ConfigProto / RunConfig / OptimizerOptions: attribute bags main() writes into and never reads back."""

  def __init__(self, *a, **k):
    pass

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)
    value = _Anything()
    object.__setattr__(self, name, value)
    return value


ConfigProto = _Anything
OptimizerOptions = _Anything()
estimator.RunConfig = _Anything


def _estimator(model_fn=None, model_dir=None, config=None, params=None, **unused):
  raise SetupCaptured(model_fn, params)


estimator.Estimator = _estimator


# ------------------------------------------------------------------------------------------------ Prediction.main() plumbing
# Prediction.py writes its tiles into a temporary TFRecord file and reads them back through tf.data (Prediction.py:85-155, 316-372).
# Here the "file" is a list of {feature name: bytes} kept in memory (an empty file is created so that main()'s os.remove works),
# and `Estimator.predict` (ESTIMATOR_MODE = "predict") feeds every record, as a batch of one [tile, tile, 3] float32 example per
# feature, to the reference's model_fn and yields its prediction dictionary without the batch axis - what
# tf.estimator.Estimator.predict yields.  The tile grid, the crop / stitch arithmetic and the combination of the passes are
# the reference's own statements.
ESTIMATOR_MODE = "capture"
RECORDS = {}


def enable_eager_execution(*a, **k):
  return None


class _Compat(object):
  as_bytes = staticmethod(lambda b: bytes(b))


compat = _Compat()


class _Train(object):
  class BytesList(object):
    def __init__(self, value=None):
      self.value = list(value or [])

  class Feature(object):
    def __init__(self, bytes_list=None):
      self.bytes_list = bytes_list

  class Features(object):
    def __init__(self, feature=None):
      self.feature = dict(feature or {})

  class Example(object):
    def __init__(self, features=None):
      self.features = features

    def SerializeToString(self):
      return self


train = _Train()


class _TFRecordWriter(object):
  def __init__(self, path):
    import os
    self.key = os.path.abspath(path)
    RECORDS[self.key] = []
    open(path, "wb").close()

  def write(self, example):
    RECORDS[self.key].append({k: v.bytes_list.value[0] for k, v in example.features.feature.items()})

  def close(self):
    pass


class _PythonIO(object):
  TFRecordWriter = _TFRecordWriter


python_io = _PythonIO()


class _PredictingEstimator(object):
  def __init__(self, model_fn=None, model_dir=None, config=None, params=None, **unused):
    self.model_fn, self.params = model_fn, params

  def predict(self, input_fn=None, **unused):
    import numpy as np
    records = list(RECORDS.values())[-1]
    for record in records:
      features = {}
      for name, raw in record.items():
        flat = np.frombuffer(raw, dtype=np.float32)
        side = int(round((flat.size // 3) ** 0.5))
        assert side * side * 3 == flat.size, "Prediction.py reshapes every source to [tile, tile, 3] (Prediction.py:70)"
        features[name] = torch.as_tensor(flat.reshape(1, side, side, 3).copy(), dtype=COMPUTE_DTYPE)
      spec = self.model_fn(features, None, _ModeKeys.PREDICT, self.params)
      yield {k: v[0].detach().numpy() for k, v in spec.predictions.items()}


def _estimator(model_fn=None, model_dir=None, config=None, params=None, **unused):      # noqa: F811
  if ESTIMATOR_MODE == "predict":
    return _PredictingEstimator(model_fn, model_dir, config, params)
  raise SetupCaptured(model_fn, params)


estimator.Estimator = _estimator
