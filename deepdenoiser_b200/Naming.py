"""Dictionary-key and summary-name contract (interface of the reference's Naming.py:8-102)."""
from .RenderPasses import RenderPasses


def _suffix(name, masked=False, internal=False, scale_index=None):
  if masked:
    name += " Masked"
  if internal:
    name += " Internal"
  if scale_index is not None:
    name += "/" + str(2 ** scale_index)
  return name


class Naming:

  # ---- summary (TensorBoard-style) names, Naming.py:12-51
  @staticmethod
  def tensorboard_name(name):
    return name.lower().replace(" ", "_")

  @staticmethod
  def _tensorboard_statistics_name(name, statistics_name, masked=False, internal=False, scale_index=None):
    if RenderPasses.is_combined_feature_render_pass(name):
      name = "Combined " + name
    return Naming.tensorboard_name(_suffix(name + statistics_name, masked, internal, scale_index))

  # the reference only forwards `internal` for mean_name (Naming.py:17-19); the others drop it
  @staticmethod
  def difference_name(name, masked=False, internal=False, scale_index=None):
    return Naming._tensorboard_statistics_name(name, " Difference", masked=masked, scale_index=scale_index)

  @staticmethod
  def mean_name(name, masked=False, internal=False, scale_index=None):
    return Naming._tensorboard_statistics_name(name, " Mean", masked=masked, internal=internal, scale_index=scale_index)

  @staticmethod
  def variation_difference_name(name, masked=False, internal=False, scale_index=None):
    return Naming._tensorboard_statistics_name(name, " Variation Difference", masked=masked, scale_index=scale_index)

  @staticmethod
  def variation_mean_name(name, masked=False, internal=False, scale_index=None):
    return Naming._tensorboard_statistics_name(name, " Variation Mean", masked=masked, scale_index=scale_index)

  @staticmethod
  def ms_ssim_name(name, masked=False, internal=False):
    return Naming._tensorboard_statistics_name(name, " MS SSIM", masked=masked)

  # ---- feature dictionary keys, Naming.py:57-81
  @staticmethod
  def source_feature_name(name, samples_per_pixel=None, index=None, masked=False):
    parts = ["source_image"]
    if samples_per_pixel is not None:
      parts.append(str(samples_per_pixel))
    if index is not None:
      parts.append(str(index))
    parts.append(_suffix(name, masked))
    return "/".join(parts)

  @staticmethod
  def feature_flags_name(name):
    return "feature_flag/" + name

  @staticmethod
  def target_feature_name(name, masked=False):
    return "target_image/" + _suffix(name, masked)

  @staticmethod
  def feature_prediction_name(name):
    return "prediction/" + name
