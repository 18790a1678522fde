"""Model assembly - the drop-in for the reference's Architecture.py (the seam is
Architecture.predict(features, mode), TensorFlow/Architecture.py:537-617).

Same constructor arguments, same public attributes (feature_predictions, auxiliary_features,
feature_prediction_tuples, feature_flags, model_directory, number_of_sources_per_target,
feature_prediction_tuple_type, data_format, source_data_format), same dictionary keys in and out:

    features : {'source_image/<i>/<Pass>': float32 NHWC [N,H,W,3|1]}   (+ 'feature_flag/<tuple>' for ONE_HOT_ENCODING)
    returns  : [ {'prediction/<Pass>': float32 NHWC [N,H/2^s,W/2^s,C_pass]} for s in scales ]   (largest first)

but the arithmetic runs in libdd_b200.so (hand-written sm_100a CUDA behind a C ABI) instead of a TensorFlow
graph.  Differences that are deliberate:
  * tensors are torch CUDA tensors (numpy / CPU tensors are uploaded); the result tensors live on the GPU;
  * the T feature-prediction tuples (17 SINGLE / 8 COMBINED passes with shared weights, Architecture.py:561-571)
    are batched along N instead of being unrolled into T copies of the graph;
  * the device layout is always NHWC; `data_format` is accepted and stored for interface compatibility only
    (the reference transposes to NCHW for cuDNN, Conv2dUtilities.py:50-66 - pure HBM traffic here);
  * an optional "b200" block in the JSON selects the arithmetic: {"dtype": "float16" | "bfloat16" | "float32", ...}.
There is no CPU fallback: predict() raises if the CUDA library or a B200 is missing.
"""
import json
from enum import Enum

import numpy as np
import torch

from . import _lib
from .FeatureFlags import FeatureFlagMode, FeatureFlags
from .Naming import Naming
from .RenderPasses import RenderPasses
from .network import DeviceNetwork, NetworkSpec, V


class ModeKeys:
  """String constants of tf.estimator.ModeKeys."""
  TRAIN = "train"
  EVAL = "eval"
  PREDICT = "infer"


class FeatureStandardization:
  """Parameters of the per-pass standardisation (Architecture.py:25-55); the arithmetic is
  dd_standardize_variance / dd_invert_standardization."""

  def __init__(self, use_log1p, mean, variance, name):
    self.use_log1p, self.mean, self.variance, self.name = use_log1p, mean, variance, name

  def use_mean(self):
    return self.mean != 0.

  def use_variance(self):
    return self.variance != 1.

  def key(self):
    return (bool(self.use_log1p), float(self.mean), float(self.variance))

  def invert_params(self):
    return _lib.dd_invert_params(int(bool(self.use_log1p)), float(self.mean), float(self.variance))


class FeatureVariance:
  """Parameters of the local-variance input feature (Architecture.py:58-79)."""

  def __init__(self, use_variance, variance_mode, relative_variance, compute_before_standardization,
               compress_to_one_channel, name):
    assert variance_mode in ("uniform", "neighbor"), variance_mode
    self.use_variance = use_variance
    self.variance_mode = variance_mode
    self.relative_variance = relative_variance
    self.compute_before_standardization = compute_before_standardization
    self.compress_to_one_channel = compress_to_one_channel
    self.name = name

  def channels(self, source_channels):
    if not self.use_variance:
      return 0
    return 1 if self.compress_to_one_channel else source_channels


class FeaturePredictionType(Enum):
  COLOR = 1
  DIRECT = 2
  INDIRECT = 3
  AUXILIARY = 4


_TYPE_STRINGS = {FeaturePredictionType.COLOR: "Color", FeaturePredictionType.DIRECT: "Direct",
                 FeaturePredictionType.INDIRECT: "Indirect", FeaturePredictionType.AUXILIARY: "Auxiliary"}


class FeaturePrediction:
  """One render pass that is fed to (and, if is_target, predicted by) the network (Architecture.py:82-183)."""

  def __init__(self, feature_prediction_type, load_data, number_of_sources, preserve_source, is_target,
               feature_standardization, invert_standardization, feature_variance, number_of_channels, name):
    self.feature_prediction_type = feature_prediction_type
    self.load_data = load_data
    self.number_of_sources = number_of_sources
    self.preserve_source = preserve_source
    self.is_target = is_target
    self.feature_standardization = feature_standardization
    self.invert_standardization = invert_standardization
    self.feature_variance = feature_variance
    self.number_of_channels = number_of_channels
    self.name = name
    self.bank_index = None       # image-bank slot (set by Architecture)
    self.predictions = []        # per scale, filled by Architecture.predict

  def synthetic_source(self, batch, height, width, device=None):
    """Constant pass for load_data == False features: ones for Color, 0.5 for Direct / Indirect
    (Prediction.py:246-252, Training.py:531-537)."""
    assert not self.load_data and self.feature_prediction_type != FeaturePredictionType.AUXILIARY
    value = 1.0 if self.feature_prediction_type == FeaturePredictionType.COLOR else 0.5
    return torch.full((batch, height, width, self.number_of_channels), value, dtype=torch.float32, device=device)

  @staticmethod
  def feature_prediction_type_to_string(feature_prediction_type):
    return _TYPE_STRINGS.get(feature_prediction_type, "")


class FeaturePredictionTupleType(Enum):
  SINGLE = 1
  COMBINED = 2


class FeaturePredictionTuple:

  def __init__(self, feature_predictions, feature_prediction_tuple_type, name):
    self.feature_predictions = feature_predictions
    self.feature_prediction_tuple_type = feature_prediction_tuple_type
    self.name = name


class Architecture:

  def __init__(self, parsed_json, source_data_format="channels_last", data_format="channels_first", device=0,
               weights=None, seed=4321):
    if source_data_format != "channels_last":
      raise ValueError("sources are NHWC at both reference call sites (Training.py:963, Prediction.py:214)")
    self.source_data_format = source_data_format
    self.data_format = data_format
    self.model_directory = parsed_json["model_directory"]
    self.number_of_sources_per_target = parsed_json["number_of_sources_per_target"]
    if self.number_of_sources_per_target != 1:
      # the reference asserts channels == K^2 in KernelPrediction.py:15, so only 1 ever worked there either
      raise ValueError("number_of_sources_per_target must be 1 (ArchitectureExample.json:5)")
    architecture_json = parsed_json["architecture"]
    self.feature_prediction_tuple_type = FeaturePredictionTupleType[
        architecture_json["source_encoder"]["feature_prediction_tuple_type"]]
    kp = architecture_json["kernel_prediction"]
    self.use_kernel_prediction = kp["use_kernel_prediction"]
    self.kernel_size = kp["kernel_size"]
    self.use_standardized_source_for_kernel_prediction = kp["use_standardized_source_for_kernel_prediction"]
    self._preserve_source = not self.use_standardized_source_for_kernel_prediction
    ms = architecture_json["multiscale_prediction"]
    self.use_multiscale_predictions = ms["use_multiscale_predictions"]
    self.invert_standardization_after_multiscale_predictions = ms["invert_standardization_after_multiscale_predictions"]

    self._prepare_feature_predictions(parsed_json["combined_features"], parsed_json["combined_features_handling"],
                                      parsed_json["auxiliary_features"])

    # feature flags (Architecture.py:484-494): the object is only kept in EMBEDDING mode
    self.feature_flag_mode = FeatureFlagMode[architecture_json["source_encoder"]["feature_flag_mode"]]
    flags = FeatureFlags([t.name for t in self.feature_prediction_tuples], self.feature_flag_mode, "channels_last")
    self._flags = flags
    self.feature_flags = flags if self.feature_flag_mode == FeatureFlagMode.EMBEDDING else None

    core = architecture_json["core_architecture"]
    tuple_size = 1 if self.feature_prediction_tuple_type == FeaturePredictionTupleType.SINGLE else 3
    self.features_per_tuple = tuple_size
    if self.use_kernel_prediction:
      self.number_of_output_channels = self.number_of_sources_per_target * tuple_size * self.kernel_size ** 2
    else:
      self.number_of_output_channels = tuple_size * 3
    embedding_shape = None
    if self.feature_flag_mode == FeatureFlagMode.EMBEDDING:
      embedding_shape = (flags.vocabulary_size, flags.embedding_dimension)
    self.input_layouts = [self.input_layout(t) for t in self.feature_prediction_tuples]
    widths = {len(l) for l in self.input_layouts}
    assert len(widths) == 1, "all tuples must produce the same number of input channels"
    self.number_of_input_channels = widths.pop()
    # BN / dropout keys of the JSON are ignored exactly like the reference does (Architecture.py:506)
    self.spec = NetworkSpec(core["name"], core["number_of_filters_for_convolution_blocks"],
                            core["number_of_convolutions_per_block"], self.number_of_input_channels,
                            self.number_of_output_channels, self.use_multiscale_predictions, embedding_shape)
    options = dict(parsed_json.get("b200", {}))
    # "float16x2": split-fp16 tensor-core arithmetic (fp16 hi + lo pairs, 3 MMA passes): the mode that meets the 1e-4 parity
    # bound of the fp32 reference on tensor cores
    self.dtype = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32,
                  "float16x2": "float16x2"}[options.get("dtype", "float16")]
    self.logits_dtype = {"float16": torch.float16, "float32": torch.float32}[options.get("logits_dtype", "float32")]
    self.max_chunk_pixels = int(options.get("max_chunk_pixels", 16 * 1024 * 1024))
    self.weights = dict(weights) if weights is not None else self.spec.init_weights(seed)
    self.device_index = device
    self.ctx = None
    self.network = None

  # ---------------------------------------------------------------------------------------------- host logic
  @classmethod
  def from_json_file(cls, filename, **kwargs):
    with open(filename, "r") as f:
      return cls(json.load(f), **kwargs)

  def _prepare_feature_predictions(self, combined_features_json, combined_features_handling_json,
                                   auxiliary_features_json):
    """Architecture.__prepare_feature_predictions (Architecture.py:367-473)."""
    n_src, keep = self.number_of_sources_per_target, self._preserve_source

    def variance_of(j, name):
      return FeatureVariance(j["use_variance"], j["variance_mode"], j["relative_variance"],
                             j["compute_before_standardization"], j["compress_to_one_channel"], name)

    def standardization_of(j, name):
      return FeatureStandardization(j["use_log1p"], j["mean"], j["variance"], name)

    # auxiliaries: sorted by name so the channel order is reproducible (Architecture.py:369-392)
    self.auxiliary_features = []
    for name in sorted(auxiliary_features_json.keys()):
      j = auxiliary_features_json[name]
      self.auxiliary_features.append(FeaturePrediction(
          FeaturePredictionType.AUXILIARY, True, n_src, keep, False, standardization_of(j["standardization"], name),
          False, variance_of(j["feature_variance"], name), j["number_of_channels"], name))

    self.feature_predictions, self.feature_prediction_tuples = [], []
    combined = self.feature_prediction_tuple_type == FeaturePredictionTupleType.COMBINED
    kinds = (FeaturePredictionType.COLOR, FeaturePredictionType.DIRECT, FeaturePredictionType.INDIRECT)
    for combined_name in sorted(combined_features_json.keys()):
      members = []
      for kind in kinds:
        kind_name = _TYPE_STRINGS[kind]
        handling = combined_features_handling_json[kind_name]
        pass_name = combined_features_json[combined_name][kind_name]
        channels = RenderPasses.number_of_channels(pass_name)
        load_data = not (pass_name is None or pass_name == "")
        if not load_data:
          pass_name = combined_name + " " + kind_name      # synthetic constant pass (Architecture.py:439-441)
        member = None
        if load_data or combined:
          member = FeaturePrediction(kind, load_data, n_src, keep, True,
                                     standardization_of(handling["standardization"], pass_name),
                                     handling["invert_standardization"],
                                     variance_of(handling["feature_variance"], pass_name), channels, pass_name)
          self.feature_predictions.append(member)
        members.append(member)
      if combined:
        self.feature_prediction_tuples.append(
            FeaturePredictionTuple(members, self.feature_prediction_tuple_type, combined_name))
    if not combined:
      for fp in self.feature_predictions:
        self.feature_prediction_tuples.append(FeaturePredictionTuple([fp], self.feature_prediction_tuple_type, fp.name))
    # image-bank slots: targets first (tuple order == creation order), then auxiliaries
    for i, fp in enumerate(self.feature_predictions + self.auxiliary_features):
      fp.bank_index = i

  def input_layout(self, feature_prediction_tuple):
    """Channel-by-channel description of the network input of one tuple, in SourceEncoder order
    (SourceEncoder.py:36-74): [('source', fp, ch) | ('variance', fp, ch) | ('one_hot', j) | ('embedding', j)]."""
    layout = []
    for fp in list(feature_prediction_tuple.feature_predictions) + list(self.auxiliary_features):
      for ch in range(3):
        layout.append(("source", fp, ch))      # 1-channel passes are replicated to 3 (SourceEncoder.py:49-51)
      for ch in range(fp.feature_variance.channels(fp.number_of_channels)):
        layout.append(("variance", fp, ch))
    if self.feature_flag_mode == FeatureFlagMode.ONE_HOT_ENCODING:
      layout += [("one_hot", feature_prediction_tuple.name, j) for j in range(len(self._flags.feature_flag_names))]
    elif self.feature_flag_mode == FeatureFlagMode.EMBEDDING:
      layout += [("embedding", feature_prediction_tuple.name, j) for j in range(self._flags.embedding_dimension)]
    return layout

  def required_features(self):
    return self.auxiliary_features + self.feature_predictions

  def mac_per_pixel(self):
    return self.spec.mac_per_pixel(self.features_per_tuple)

  # ---------------------------------------------------------------------------------------------- device state
  def _ensure_device(self):
    if self.ctx is not None:
      return
    self.ctx = _lib.Context(self.device_index)      # raises when the library / GPU is missing: no CPU fallback
    self.network = DeviceNetwork(self.ctx, self.spec, self.weights, self.dtype, self.logits_dtype)
    self._sync_embedding()

  def _sync_embedding(self):
    if self.feature_flag_mode == FeatureFlagMode.EMBEDDING:
      m = torch.from_numpy(np.ascontiguousarray(self.weights["embedding/feature_flags_embedding_matrix"]))
      self._flags.embedding_matrix = m.to(self.ctx.device)

  def set_weights(self, weights):
    self.weights = dict(weights)
    if self.ctx is not None:
      self.network.load_weights(self.weights)
      self._sync_embedding()

  def _as_device(self, value):
    if isinstance(value, np.ndarray):
      value = torch.from_numpy(value)
    t = value.to(device=self.ctx.device, dtype=torch.float32, non_blocking=True)
    if t.dim() == 3:
      t = t.unsqueeze(0)
    return t.contiguous()

  @staticmethod
  def _std_params(fp):
    st, fv = fp.feature_standardization, fp.feature_variance
    return _lib.dd_standardize_params(
        int(bool(st.use_log1p)), float(st.mean), float(st.variance), int(bool(fv.use_variance)),
        0 if fv.variance_mode == "uniform" else 1, int(bool(fv.relative_variance)),
        int(bool(fv.compute_before_standardization)), int(bool(fv.compress_to_one_channel)), 1e-4)

  _ENTRY = np.dtype([("ptr", "<u8"), ("cstride", "<i4"), ("cidx", "<i4"), ("constant", "<f4"), ("pad", "<i4")])

  def _gather_table(self, std_bank, var_bank, var_width, features, n, c0p):
    """dd_gather_entry rows (one per tuple) that make dd_assemble_input reproduce SourceEncoder's concat."""
    img_std = std_bank.shape[1] * std_bank.shape[2] * 3 * 4 * n
    img_var = std_bank.shape[1] * std_bank.shape[2] * max(var_width, 1) * 4 * n
    table = np.zeros((len(self.feature_prediction_tuples), c0p), dtype=self._ENTRY)
    keep = []
    for t, layout in enumerate(self.input_layouts):
      for ch, entry in enumerate(layout):
        kind = entry[0]
        if kind == "source":
          fp = entry[1]
          table[t, ch] = (std_bank.data_ptr() + fp.bank_index * img_std, 3, entry[2], 0.0, 0)
        elif kind == "variance":
          fp = entry[1]
          table[t, ch] = (var_bank.data_ptr() + fp.bank_index * img_var, var_width, entry[2], 0.0, 0)
        elif kind == "embedding":
          m = self._flags.embedding_matrix
          row = self._flags.index(entry[1])
          table[t, ch] = (m.data_ptr() + (row * m.shape[1] + entry[2]) * 4, 0, 0, 0.0, 0)
        else:  # one_hot: caller-provided planes
          flags = self._as_device(features[Naming.feature_flags_name(entry[1])])
          if flags.shape[0] != n:
            flags = flags.expand(n, -1, -1, -1).contiguous()
          keep.append(flags)
          table[t, ch] = (flags.data_ptr(), flags.shape[3], entry[2], 0.0, 0)
    dev = torch.from_numpy(table.view(np.uint8).reshape(-1)).to(self.ctx.device)
    return dev, keep

  # ---------------------------------------------------------------------------------------------- predict
  def predict(self, features, mode=ModeKeys.PREDICT):
    """Architecture.predict (Architecture.py:537-617)."""
    self._ensure_device()
    ctx, net, dev = self.ctx, self.network, self.ctx.device
    targets, auxiliaries = self.feature_predictions, self.auxiliary_features
    every = targets + auxiliaries
    sources = [self._as_device(features[Naming.source_feature_name(fp.name, index=0)]) for fp in every]
    n, h, w = sources[0].shape[0], sources[0].shape[1], sources[0].shape[2]
    for fp, s in zip(every, sources):
      if s.shape[:3] != (n, h, w) or s.shape[3] not in (1, 3):
        raise ValueError("source '%s' has shape %s, expected [%d,%d,%d,1|3]" % (fp.name, tuple(s.shape), n, h, w))
    n_scales = (self.spec.steps + 1) if self.use_multiscale_predictions else 1
    if h % (1 << self.spec.steps) or w % (1 << self.spec.steps):
      raise ValueError("height and width must be divisible by %d" % (1 << self.spec.steps))

    # 1. standardise every pass once + its variance feature (Architecture.py:549-555)
    var_width = max([fp.feature_variance.channels(fp.number_of_channels) for fp in every] + [0])
    std_bank = net._buf("bank.std", (len(every) * n, h, w, 3), torch.float32)
    var_bank = net._buf("bank.var", (len(every) * n, h, w, max(var_width, 1)), torch.float32)
    raw_bank = net._buf("bank.raw", (len(targets) * n, h, w, 3), torch.float32) if self._preserve_source else None
    jobs = []                                    # every pass in ONE launch (dd_standardize_variance_batch)
    for fp, s in zip(every, sources):
      lo, hi = fp.bank_index * n, (fp.bank_index + 1) * n
      vc = fp.feature_variance.channels(fp.number_of_channels)
      jobs.append((_lib.desc(s), self._std_params(fp), _lib.desc(std_bank[lo:hi]),
                   _lib.desc(var_bank[lo:hi], vc, 0) if vc else None))
      if self._preserve_source and fp.is_target:
        identity = _lib.dd_standardize_params(0, 0.0, 1.0, 0, 0, 0, 0, 0, 1e-4)
        jobs.append((_lib.desc(s), identity, _lib.desc(raw_bank[lo:hi]), None))
    if len(jobs) * n <= 65535:
      job_table = net._buf("std.jobs", (len(jobs) * int(ctx.lib.dd_standardize_variance_job_bytes()),), torch.uint8)
      ctx.standardize_variance_batch(jobs, job_table)
    else:
      for job in jobs:
        ctx.standardize_variance(*job)

    # kernel-prediction sources per scale: avg-pool by 2^s of the full-resolution source (Architecture.py:280-283)
    nt = len(targets) * n
    kp_full = raw_bank if self._preserve_source else std_bank[:nt]
    kp_sources = [kp_full]
    for s in range(1, n_scales):
      pooled = net._buf("bank.kpsrc%d" % s, (nt, h >> s, w >> s, 3), torch.float32)
      if self.use_kernel_prediction:
        ctx.avgpool(_lib.desc(kp_full), 1 << s, _lib.desc(pooled))
      kp_sources.append(pooled)

    # 2. network input of every tuple (SourceEncoder.py:29-79)
    c0 = self.number_of_input_channels
    c0p = (c0 + 7) // 8 * 8
    table, keep_alive = self._gather_table(std_bank, var_bank, max(var_width, 1), features, n, c0p)

    finals = [torch.empty((nt, h >> s, w >> s, 3), dtype=torch.float32, device=dev) for s in range(n_scales)]
    tuples = self.feature_prediction_tuples
    ft = self.features_per_tuple
    per_chunk = max(1, min(len(tuples), self.max_chunk_pixels // max(1, n * h * w)))
    # balanced chunks (17 tuples at 1080p: 6 + 6 + 5 instead of 8 + 8 + 1): no launch is left with a single image
    per_chunk = -(-len(tuples) // -(-len(tuples) // per_chunk))
    entry_bytes = self._ENTRY.itemsize
    for t0 in range(0, len(tuples), per_chunk):
      t1 = min(len(tuples), t0 + per_chunk)
      bc = (t1 - t0) * n
      # 3. core architecture + 1x1 post-processing, all tuples of the chunk batched along N
      if net.split:
        x0 = net._buf("net.x0x2", (bc, h, w, 2 * c0p))
        ctx.assemble_input_split(table[t0 * c0p * entry_bytes:], t1 - t0, n, _lib.desc(x0, c0p, 0), _lib.desc(x0, c0p, c0p))
        core = net.forward_core(net._sv(x0, c0, 0))           # coarsest first
      else:
        x0 = net._buf("net.x0", (bc, h, w, c0p))
        ctx.assemble_input(table[t0 * c0p * entry_bytes:], t1 - t0, n, _lib.desc(x0))
        core = net.forward_core(V(x0, c0, 0))                 # coarsest first
      fuse = self.use_kernel_prediction and net.can_fuse_post_kp(self.kernel_size, ft)
      logits = None if fuse else net.post_process(core)       # largest first
      # 4. split per feature + kernel prediction per scale (Architecture.py:581-591)
      lo, hi = t0 * ft * n, t1 * ft * n
      stage = []
      for s in range(n_scales):
        last = (s == n_scales - 1)
        dst = finals[s][lo:hi] if last else net._buf("kp.out%d" % s, ((t1 - t0) * ft * n, h >> s, w >> s, 3),
                                                     torch.float32)
        if fuse:
          # 1x1 post-processing + kernel prediction in one kernel: the logits never reach HBM
          k = (len(core) - 1 - s) if self.use_multiscale_predictions else 0
          net.post_kernel_predict(k, core[k], _lib.desc(kp_sources[s][lo:hi]), self.kernel_size, ft, n, _lib.desc(dst))
        elif self.use_kernel_prediction:
          ctx.kernel_predict(_lib.desc(kp_sources[s][lo:hi]), logits[s].d, self.kernel_size, ft, n, _lib.desc(dst))
        else:
          self._split_direct(logits[s], dst, t1 - t0, n)
        stage.append(dst)
      # 5. multi-scale composition + inverse standardisation (Architecture.py:302-325)
      members = targets[t0 * ft:t1 * ft]
      if not self.invert_standardization_after_multiscale_predictions:
        for s in range(n_scales):
          self._invert(members, stage[s], n)
      for s in range(n_scales - 1, 0, -1):
        out = finals[s - 1][lo:hi]
        net.compose(V(stage[s]), V(stage[s - 1]), V(out))
        stage[s - 1] = out
      if self.invert_standardization_after_multiscale_predictions:
        for s in range(n_scales):
          self._invert(members, stage[s], n)
    del keep_alive

    # 6. prediction dictionaries (Architecture.py:602-617)
    dictionaries = []
    for s in range(n_scales):
      d = {}
      for fp in targets:
        lo, hi = fp.bank_index * n, (fp.bank_index + 1) * n
        if fp.load_data:
          prediction = finals[s][lo:hi]
        else:
          # generated pass: a crop of the (standardised) source, so it is harmless in training (:151-157)
          prediction = std_bank[lo:hi, :h >> s, :w >> s, :].clone()
        if fp.number_of_channels != 3:
          assert fp.number_of_channels == 1
          prediction = prediction[..., :1]
        d[Naming.feature_prediction_name(fp.name)] = prediction
      dictionaries.append(d)
    for fp in targets:
      fp.predictions = [d[Naming.feature_prediction_name(fp.name)] for d in dictionaries]
    return dictionaries

  def _invert(self, members, bank, n):
    """prediction_invert_standardization (Architecture.py:134-138) on a pass-major bank slice, one launch per
    run of consecutive passes with identical parameters."""
    i = 0
    while i < len(members):
      fp = members[i]
      j = i + 1
      while (j < len(members) and members[j].invert_standardization == fp.invert_standardization and
             members[j].feature_standardization.key() == fp.feature_standardization.key()):
        j += 1
      st = fp.feature_standardization
      if fp.invert_standardization and st is not None and st.key() != (False, 0.0, 1.0):
        view = _lib.desc(bank[i * n:j * n])
        self.ctx.invert_standardization(view, st.invert_params(), view)
      i = j

  def _split_direct(self, logits, dst, n_tuples, n):
    """No kernel prediction: the post-processed tensor is split into one 3-channel prediction per feature."""
    ft = self.features_per_tuple
    for t in range(n_tuples):
      for f in range(ft):
        src = V(logits.t[t * n:(t + 1) * n], 3, logits.coff + 3 * f)
        self.ctx.cast_copy(src.d, _lib.desc(dst[(t * ft + f) * n:(t * ft + f + 1) * n]))
