"""ORACLE (test infrastructure, not product code) - CPU restatement of the reference's model assembly:
Architecture.predict(features, mode) and everything it calls.

PARITY UNPINNED: TensorFlow 1.x cannot run here and the reference ships no tests or golden vectors, so this
file IS the definition of "reference results" for the repo.  It follows the reference class by class (each
class / method cites the file:line it restates, paths relative to /root/reference/TensorFlow) and uses the
TF operator semantics written down in SURVEY.md Appendix A through a backend module:

    oracle.np_ops    + float64 : the oracle proper
    oracle.torch_ops + float32 : the restated reference on CPU (bench.py cpu_baseline / --impl reference)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Variables follow TF-1.x auto-naming (tf.layers.conv2d -> "<scope>/conv2d", "conv2d_1", ... in creation order,
counters restart when a variable_scope is re-entered with reuse=True), so a weight dictionary is
interchangeable between this oracle and the CUDA product path.  Kernels use TF layouts
([kh,kw,cin,cout]; conv2d_transpose [kh,kw,cout,cin]); tf.layers defaults: glorot-uniform kernels, zero bias.
"""
import numpy as np

from . import np_ops


# ------------------------------------------------------------------------------------------------ variables
class VariableStore:
  """Stand-in for TF's variable store + tf.layers default naming.  `values` (name -> ndarray) supplies the
  weights; missing variables are created glorot-uniform (kernels, embedding) / zero (bias) from `seed`."""

  def __init__(self, values=None, seed=4321):
    self.values = dict(values) if values is not None else {}
    self.rng = np.random.default_rng(seed)
    self.created = []          # creation order, for tests
    self._scope = []
    self._counts = {}

  # tf.variable_scope(name, reuse=...): re-entering restarts the default-name counters of that scope
  def enter_scope(self, name):
    self._scope.append(name)
    prefix = "/".join(self._scope) + "/"
    for key in [k for k in self._counts if k.startswith(prefix)]:
      del self._counts[key]

  def exit_scope(self):
    self._scope.pop()

  def _layer_name(self, default_name):
    prefix = "/".join(self._scope) + "/" if self._scope else ""
    key = prefix + default_name
    count = self._counts.get(key, 0)
    self._counts[key] = count + 1
    return prefix + (default_name if count == 0 else "%s_%d" % (default_name, count))

  def _get(self, name, shape, fan_in, fan_out):
    if name not in self.values:
      if fan_in is None:
        self.values[name] = np.zeros(shape, dtype=np.float32)
      else:
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        self.values[name] = self.rng.uniform(-limit, limit, size=shape).astype(np.float32)
    value = self.values[name]
    assert tuple(value.shape) == tuple(shape), (name, value.shape, shape)
    if name not in self.created:
      self.created.append(name)
    return value

  def conv2d(self, ksize, cin, cout):
    layer = self._layer_name("conv2d")
    receptive = ksize * ksize
    kernel = self._get(layer + "/kernel", (ksize, ksize, cin, cout), receptive * cin, receptive * cout)
    bias = self._get(layer + "/bias", (cout,), None, None)
    return kernel, bias

  def conv2d_transpose(self, ksize, cin, cout):
    layer = self._layer_name("conv2d_transpose")
    receptive = ksize * ksize
    # keras glorot on shape [kh,kw,cout,cin]: fan_in = shape[-2]*receptive, fan_out = shape[-1]*receptive
    kernel = self._get(layer + "/kernel", (ksize, ksize, cout, cin), receptive * cout, receptive * cin)
    bias = self._get(layer + "/bias", (cout,), None, None)
    return kernel, bias

  def embedding(self, vocabulary, dimension):
    return self._get("embedding/feature_flags_embedding_matrix", (vocabulary, dimension), vocabulary, dimension)


# ------------------------------------------------------------------------------------------------ render passes / naming
def number_of_channels(render_pass_name):
  """RenderPasses.number_of_channels (RenderPasses.py:40-44)."""
  return 1 if render_pass_name in ("Alpha", "Depth") else 3


def source_feature_name(name, index):
  """Naming.source_feature_name (Naming.py:57-65) with samples_per_pixel=None."""
  return "source_image/%d/%s" % (index, name)


def feature_flags_name(name):
  """Naming.feature_flags_name (Naming.py:68-70)."""
  return "feature_flag/" + name


def feature_prediction_name(name):
  """Naming.feature_prediction_name (Naming.py:79-81)."""
  return "prediction/" + name


# ------------------------------------------------------------------------------------------------ feature bookkeeping
class FeatureStandardization:
  """Architecture.FeatureStandardization (Architecture.py:25-55)."""

  def __init__(self, ops, use_log1p, mean, variance):
    self.ops, self.use_log1p, self.mean, self.variance = ops, use_log1p, mean, variance

  def standardize(self, x):
    if self.use_log1p:
      x = self.ops.signed_log1p(x)
    if self.mean != 0.:
      x = x - self.mean
    if self.variance != 1.:
      x = x / float(np.sqrt(self.variance))
    return x

  def invert(self, x):
    if self.variance != 1.:
      x = x * float(np.sqrt(self.variance))
    if self.mean != 0.:
      x = x + self.mean
    if self.use_log1p:
      x = self.ops.signed_expm1(x)
    return x


class FeaturePrediction:
  """Architecture.FeaturePrediction (Architecture.py:82-165)."""

  def __init__(self, ops, kind, load_data, number_of_sources, preserve_source, is_target, standardization,
               invert_standardization, variance_json, channels, name):
    self.ops = ops
    self.kind = kind                      # 'Color' | 'Direct' | 'Indirect' | 'Auxiliary'
    self.load_data = load_data
    self.number_of_sources = number_of_sources
    self.preserve_source = preserve_source
    self.is_target = is_target
    self.standardization = standardization
    self.invert_standardization = invert_standardization
    self.variance_json = variance_json
    self.number_of_channels = channels
    self.name = name
    self.predictions = []

  @property
  def use_variance(self):
    return self.variance_json["use_variance"]

  def initialize_sources(self, dictionary, dtype):                       # Architecture.py:102-112
    self.source, self.variance, self.preserved_source = [], [], []
    self.predictions = []
    for index in range(self.number_of_sources):
      src = self.ops.asarray(dictionary[source_feature_name(self.name, index)], dtype)
      self.source.append(src)
      if self.preserve_source:
        self.preserved_source.append(src)

  def _variance(self, x):                                                # Architecture.py:69-74
    v = self.variance_json
    return self.ops.variance_feature(x, variance_mode=v["variance_mode"], relative_variance=v["relative_variance"],
                                     compress_to_one_channel=v["compress_to_one_channel"], epsilon=1e-4)

  def standardize(self):                                                 # Architecture.py:114-132
    before = self.variance_json["compute_before_standardization"]
    if self.use_variance and before:
      self.variance = [self._variance(s) for s in self.source]
    if self.standardization is not None:
      self.source = [self.standardization.standardize(s) for s in self.source]
    if self.use_variance and not before:
      self.variance = [self._variance(s) for s in self.source]

  def prediction_invert_standardization(self):                           # Architecture.py:134-138
    if self.standardization is not None:
      self.predictions = [self.standardization.invert(p) for p in self.predictions]

  def add_prediction(self, scale_index, prediction):                     # Architecture.py:140-145
    assert self.is_target
    while len(self.predictions) <= scale_index:
      self.predictions.append(None)
    self.predictions[scale_index] = prediction

  def add_prediction_to_dictionary(self, scale_index, dictionary):       # Architecture.py:147-165
    if not self.is_target:
      return
    prediction = self.predictions[scale_index]
    if not self.load_data:
      # generated data: the "prediction" is a slice of the (standardised) source so it cannot hurt training
      n, h, w, c = prediction.shape
      prediction = self.source[0][:n, :h, :w, :c]
    if prediction.shape[3] != self.number_of_channels:
      assert self.number_of_channels == 1
      prediction = prediction[..., :1]
    dictionary[feature_prediction_name(self.name)] = prediction


class FeaturePredictionTuple:
  """Architecture.FeaturePredictionTuple (Architecture.py:185-191)."""

  def __init__(self, feature_predictions, name):
    self.feature_predictions = feature_predictions
    self.name = name


# ------------------------------------------------------------------------------------------------ networks
class UNet:
  """UNet.predict (UNet.py:61-99); batch norm / dropout are dead code (Architecture.py:506)."""

  def __init__(self, ops, filters, convs_per_block, multiscale):
    self.ops, self.filters, self.n, self.multiscale = ops, filters, convs_per_block, multiscale

  def _block(self, store, x, cout):                                      # UNet.py:25-36
    for _ in range(self.n):
      kernel, bias = store.conv2d(3, x.shape[3], cout)
      x = self.ops.conv2d_same(x, kernel, bias, relu=True)
    return x

  def predict(self, store, x):
    ops, results, skips = self.ops, [], []
    steps = len(self.filters) - 1
    for i in range(steps):                                               # UNet.py:71-79
      x = self._block(store, x, self.filters[i])
      skips.append(x)
      x = ops.max_pool_same_s2(x, 3)                                     # UNet.py:38-52
    for i in range(steps):                                               # UNet.py:82-92
      index = steps - i
      x = self._block(store, x, self.filters[index])
      if self.multiscale:
        results.append(x)
      kernel, bias = store.conv2d_transpose(2, x.shape[3], self.filters[index - 1])
      x = ops.conv2d_transpose_same_s2(x, kernel, bias, relu=True)       # UNet.py:54-59
      x = ops.concat([skips[index - 1], x], axis=3)
    x = self._block(store, x, self.filters[0])                           # UNet.py:95-97
    results.append(x)
    return results


class Tiramisu:
  """Tiramisu.predict (Tiramisu.py:67-111)."""

  def __init__(self, ops, pre_filters, filters, convs_per_block, multiscale):
    self.ops, self.pre, self.filters, self.n, self.multiscale = ops, pre_filters, filters, convs_per_block, multiscale

  def _block(self, store, x, growth):                                    # Tiramisu.py:26-41
    for _ in range(self.n):
      kernel, bias = store.conv2d(3, x.shape[3], growth)
      layer = self.ops.conv2d_same(self.ops.relu(x), kernel, bias, relu=False)
      x = self.ops.concat([x, layer], axis=3)
    return x

  def _down(self, store, x):                                             # Tiramisu.py:43-58
    c = x.shape[3]
    kernel, bias = store.conv2d(1, c, c)
    x = self.ops.conv2d_same(self.ops.relu(x), kernel, bias, relu=False)
    return self.ops.max_pool_same_s2(x, 2)

  def predict(self, store, x):
    ops, results, skips = self.ops, [], []
    steps = len(self.filters) - 1
    kernel, bias = store.conv2d(3, x.shape[3], self.pre)                 # Tiramisu.py:76-79
    x = ops.conv2d_same(x, kernel, bias, relu=True)
    for i in range(steps):                                               # Tiramisu.py:82-90
      x = self._block(store, x, self.filters[i])
      skips.append(x)
      x = self._down(store, x)
    for i in range(steps):                                               # Tiramisu.py:93-104
      index = steps - i
      x = self._block(store, x, self.filters[index])
      if self.multiscale:
        results.append(x)
      kernel, bias = store.conv2d_transpose(3, x.shape[3], self.filters[index - 1])
      x = ops.conv2d_transpose_same_s2(x, kernel, bias, relu=True)       # Tiramisu.py:60-65
      x = ops.concat([skips[index - 1], x], axis=3)
    x = self._block(store, x, self.filters[0])                           # Tiramisu.py:107-109
    results.append(x)
    return results


def compose_scales(ops, store, small, large):
  """MultiScalePrediction.compose_scales + weight network (MultiScalePrediction.py:36-93)."""
  small_up = ops.resize_nearest_x2(small)
  x = ops.concat([small_up, large], axis=3)                              # :59
  kernel, bias = store.conv2d(1, x.shape[3], 24)                         # :62-66
  x = ops.conv2d_same(x, kernel, bias, relu=True)
  for _ in range(2):                                                     # :69-71, _residual_block :81-93
    residual = x
    for _ in range(2):
      kernel, bias = store.conv2d(3, 24, 24)
      residual = ops.conv2d_same(ops.relu(residual), kernel, bias, relu=False)
    x = x + residual
  kernel, bias = store.conv2d(1, 24, 1)                                  # :73-75 (ReLU) then sigmoid :77
  weights = ops.sigmoid(ops.conv2d_same(x, kernel, bias, relu=True))
  low = ops.resize_nearest_x2(ops.avg_pool_same(large, 2))               # :45-46
  return large - weights * low + weights * small_up                      # :48-52


# ------------------------------------------------------------------------------------------------ Architecture
class Architecture:
  """Architecture.__init__ / __prepare_feature_predictions / __prepare_architecture / predict
  (Architecture.py:341-617)."""

  def __init__(self, parsed_json, ops=np_ops, dtype=np.float64, weights=None, seed=4321):
    self.ops, self.dtype = ops, dtype
    self.store = VariableStore(weights, seed)
    self.model_directory = parsed_json["model_directory"]
    self.number_of_sources_per_target = parsed_json["number_of_sources_per_target"]
    arch = parsed_json["architecture"]
    self.tuple_type = arch["source_encoder"]["feature_prediction_tuple_type"]      # 'SINGLE' | 'COMBINED'
    self.flag_mode = arch["source_encoder"]["feature_flag_mode"]                  # NONE | ONE_HOT_ENCODING | EMBEDDING
    kp = arch["kernel_prediction"]
    self.use_kernel_prediction = kp["use_kernel_prediction"]
    self.kernel_size = kp["kernel_size"]
    self.use_standardized_source = kp["use_standardized_source_for_kernel_prediction"]
    self.preserve_source = not self.use_standardized_source
    ms = arch["multiscale_prediction"]
    self.use_multiscale = ms["use_multiscale_predictions"]
    self.invert_after_multiscale = ms["invert_standardization_after_multiscale_predictions"]
    core = arch["core_architecture"]
    self.core_name = core["name"]
    filters = core["number_of_filters_for_convolution_blocks"]
    n = core["number_of_convolutions_per_block"]
    if self.core_name == "U-Net":                                                  # Architecture.py:194-227
      self.core = UNet(ops, filters, n, self.use_multiscale)
    else:
      assert self.core_name == "Tiramisu"
      self.core = Tiramisu(ops, filters[0], filters, n, self.use_multiscale)
    self._prepare_feature_predictions(parsed_json["combined_features"], parsed_json["combined_features_handling"],
                                      parsed_json["auxiliary_features"])
    tuple_size = 1 if self.tuple_type == "SINGLE" else 3                           # Architecture.py:510-522
    if self.use_kernel_prediction:
      self.number_of_output_channels = self.number_of_sources_per_target * tuple_size * self.kernel_size ** 2
    else:
      self.number_of_output_channels = tuple_size * 3
    self.flag_names = sorted(t.name for t in self.feature_prediction_tuples)       # FeatureFlags.py:22

  def _prepare_feature_predictions(self, combined, handling, auxiliary):           # Architecture.py:367-473
    ops = self.ops
    self.auxiliary_features = []
    for name in sorted(auxiliary.keys()):
      f = auxiliary[name]
      st = f["standardization"]
      self.auxiliary_features.append(FeaturePrediction(
          ops, "Auxiliary", True, self.number_of_sources_per_target, self.preserve_source, False,
          FeatureStandardization(ops, st["use_log1p"], st["mean"], st["variance"]), False, f["feature_variance"],
          f["number_of_channels"], name))
    self.feature_predictions, self.feature_prediction_tuples = [], []
    for combined_name in sorted(combined.keys()):
      members = []
      for kind in ("Color", "Direct", "Indirect"):
        feature_name = combined[combined_name][kind]
        h = handling[kind]
        st = h["standardization"]
        channels = number_of_channels(feature_name)
        load_data = True
        if feature_name is None or feature_name == "":
          feature_name = combined_name + " " + kind
          load_data = False
        fp = None
        if load_data or self.tuple_type == "COMBINED":
          fp = FeaturePrediction(ops, kind, load_data, self.number_of_sources_per_target, self.preserve_source, True,
                                 FeatureStandardization(ops, st["use_log1p"], st["mean"], st["variance"]),
                                 h["invert_standardization"], h["feature_variance"], channels, feature_name)
          self.feature_predictions.append(fp)
        members.append(fp)
      if self.tuple_type == "COMBINED":
        self.feature_prediction_tuples.append(FeaturePredictionTuple(members, combined_name))
    if self.tuple_type == "SINGLE":
      for fp in self.feature_predictions:
        self.feature_prediction_tuples.append(FeaturePredictionTuple([fp], fp.name))

  # SourceEncoder.prepare_neural_network_input (SourceEncoder.py:29-79), channels_last throughout
  def _network_input(self, tup, features):
    ops = self.ops
    members = list(tup.feature_predictions) + list(self.auxiliary_features)
    parts = []
    for index in range(len(tup.feature_predictions[0].source)):
      for f in members:
        src = f.source[index]
        if src.shape[3] != 3:
          assert src.shape[3] == 1
          src = ops.concat([src, src, src], axis=3)
        parts.append(src)
        if f.use_variance:
          parts.append(f.variance[index])
    x = ops.concat(parts, axis=3)
    if self.flag_mode == "ONE_HOT_ENCODING":
      x = ops.concat([x, ops.asarray(features[feature_flags_name(tup.name)], self.dtype)], axis=3)
    elif self.flag_mode == "EMBEDDING":                                            # FeatureFlags.py:50-69
      vocabulary = len(self.flag_names)
      matrix = self.store.embedding(vocabulary, vocabulary // 2)
      row = ops.asarray(matrix[self.flag_names.index(tup.name)], self.dtype)
      n, h, w, _ = x.shape
      x = ops.concat([x, ops.tile_hw(row, n, h, w)], axis=3)
    return x

  def _postprocess(self, x):                                                       # Architecture.py:230-244
    k1, b1 = self.store.conv2d(1, x.shape[3], self.number_of_output_channels)
    x = self.ops.conv2d_same(x, k1, b1, relu=True)
    k2, b2 = self.store.conv2d(1, self.number_of_output_channels, self.number_of_output_channels)
    return self.ops.conv2d_same(x, k2, b2, relu=False)

  def _kernel_predictor(self, fp):                                                 # Architecture.py:260-289
    if not self.use_kernel_prediction:
      return
    ops = self.ops
    source = fp.source[0] if self.use_standardized_source else fp.preserved_source[0]
    if source.shape[3] != 3:
      assert source.shape[3] == 1
      source = ops.concat([source, source, source], axis=3)
    for scale_index in range(len(fp.predictions)):
      scaled = source if scale_index == 0 else ops.avg_pool_same(source, 2 ** scale_index)
      fp.add_prediction(scale_index, ops.kernel_prediction(scaled, fp.predictions[scale_index], self.kernel_size))

  def _multiscale_predictor(self, fp):                                             # Architecture.py:302-325
    if not self.invert_after_multiscale and fp.invert_standardization:
      fp.prediction_invert_standardization()
    if self.use_multiscale:
      for scale_index in range(len(fp.predictions) - 1, 0, -1):
        self.store.enter_scope("reused_compose_scales")
        composed = compose_scales(self.ops, self.store, fp.predictions[scale_index], fp.predictions[scale_index - 1])
        self.store.exit_scope()
        fp.add_prediction(scale_index - 1, composed)
    if self.invert_after_multiscale and fp.invert_standardization:
      fp.prediction_invert_standardization()

  def predict(self, features, mode=None):                                          # Architecture.py:537-617
    for f in self.feature_predictions + self.auxiliary_features:
      f.initialize_sources(features, self.dtype)
    for f in self.feature_predictions + self.auxiliary_features:
      f.standardize()
    for tup in self.feature_prediction_tuples:
      x = self._network_input(tup, features)
      self.store.enter_scope("reused_core_architecture")
      outputs = self.core.predict(self.store, x)
      outputs = [self._postprocess(o) for o in outputs]
      self.store.exit_scope()
      if self.use_multiscale:
        outputs = list(reversed(outputs))
      members = tup.feature_predictions
      for scale_index, out in enumerate(outputs):
        per = out.shape[3] // len(members)
        assert per * len(members) == out.shape[3]
        for index, fp in enumerate(members):                                       # tf.split, :581-587
          fp.add_prediction(scale_index, out[..., index * per:(index + 1) * per])
      for fp in members:
        self._kernel_predictor(fp)
      for fp in members:
        self._multiscale_predictor(fp)
    target = next(fp for fp in self.feature_predictions if fp.is_target)
    dictionaries = []
    for scale_index in range(len(target.predictions)):
      d = {}
      for fp in self.feature_predictions:
        fp.add_prediction_to_dictionary(scale_index, d)
      dictionaries.append(d)
    return dictionaries

  # convenience for tests / bench: numpy outputs
  def predict_numpy(self, features):
    return [{k: self.ops.to_numpy(v) for k, v in d.items()} for d in self.predict(features)]


def synthetic_source(fp, height, width, dtype=np.float32):
  """The constant passes fed for load_data == False features (Prediction.py:246-252, Training.py:531-537):
  ones for Color, 0.5 for Direct / Indirect."""
  value = 1.0 if fp.kind == "Color" else 0.5
  return np.full((height, width, fp.number_of_channels), value, dtype=dtype)
