"""Feature flags: how the network is told which render pass it is denoising (interface of the reference's
FeatureFlags.py:12-81).  EMBEDDING keeps a trainable [V, V//2] matrix whose row (looked up by the position
of the tuple name in the SORTED name list) is broadcast over the image; ONE_HOT_ENCODING puts constant
one-hot planes into the source dictionary under Naming.feature_flags_name()."""
from enum import Enum

import torch

from .Naming import Naming


class FeatureFlagMode(Enum):
  NONE = 1
  FLAGS = 2
  ONE_HOT_ENCODING = 3
  EMBEDDING = 4


class FeatureFlags:

  def __init__(self, feature_flag_names, feature_flag_mode, data_format="channels_last"):
    self.feature_flag_names = sorted(feature_flag_names)
    self.feature_flag_mode = feature_flag_mode
    self.data_format = data_format
    if feature_flag_mode == FeatureFlagMode.EMBEDDING:
      self.vocabulary_size = len(self.feature_flag_names)
      self.embedding_dimension = len(self.feature_flag_names) // 2   # hard-coded in the reference, FeatureFlags.py:47
    self.embedding_matrix = None   # device fp32 [V, V//2]; owned by the Architecture's weights

  def index(self, feature_flag_name):
    return self.feature_flag_names.index(feature_flag_name)

  def feature_flags(self, feature_flag_name, height, width, data_format="channels_last"):
    """The embedding row of `feature_flag_name` tiled to [height, width, V//2] (FeatureFlags.py:50-69)."""
    assert self.feature_flag_mode == FeatureFlagMode.EMBEDDING
    if self.embedding_matrix is None:
      raise RuntimeError("embedding matrix not initialised (it lives in the Architecture's weights)")
    row = self.embedding_matrix[self.index(feature_flag_name)]
    return row.reshape(1, 1, -1).expand(height, width, -1)

  def add_to_source_dictionary(self, sources, height, width, device=None, batch=None):
    """ONE_HOT_ENCODING: sources['feature_flag/<name>'] = one-hot planes [height, width, V]
    ([batch, height, width, V] when batch is given) (FeatureFlags.py:71-81)."""
    if self.feature_flag_mode != FeatureFlagMode.ONE_HOT_ENCODING:
      return
    v = len(self.feature_flag_names)
    for i, name in enumerate(self.feature_flag_names):
      shape = (height, width, v) if batch is None else (batch, height, width, v)
      flags = torch.zeros(shape, dtype=torch.float32, device=device)
      flags[..., i] = 1.0
      sources[Naming.feature_flags_name(name)] = flags
