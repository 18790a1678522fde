"""Render-pass vocabulary of the denoiser (mirrors the interface of the reference's RenderPasses.py:5-227:
same constant names, same static helpers, same RenderPassesUsage ordering) - table driven."""

_PASSES = (
    # (constant, pass name, channels, rgb colour data?, usage flag)
    ("COMBINED", "Combined", 3, True, None),
    ("COMBINED_DIFFUSE", "Diffuse", 3, True, None),
    ("COMBINED_GLOSSY", "Glossy", 3, True, None),
    ("COMBINED_SUBSURFACE", "Subsurface", 3, True, None),
    ("COMBINED_TRANSMISSION", "Transmission", 3, True, None),
    ("ALPHA", "Alpha", 1, False, "use_alpha"),
    ("DEPTH", "Depth", 1, False, "use_depth"),
    ("MIST", "Mist", 3, False, "use_mist"),
    ("NORMAL", "Normal", 3, False, "use_normal"),
    ("SCREEN_SPACE_NORMAL", "Screen Space Normal", 3, False, "use_screen_space_normal"),
    ("MOTION_VECTOR", "Motion Vector", 3, False, "use_motion_vector"),
    ("OBJECT_ID", "Object ID", 3, False, "use_object_id"),
    ("MATERIAL_ID", "Material ID", 3, False, "use_material_id"),
    ("UV", "UV", 3, False, "use_uv"),
    ("SHADOW", "Shadow", 3, True, "use_shadow"),
    ("AMBIENT_OCCLUSION", "Ambient Occlusion", 3, True, "use_ambient_occlusion"),
    ("EMISSION", "Emission", 3, True, "use_emission"),
    ("ENVIRONMENT", "Environment", 3, True, "use_environment"),
    ("DIFFUSE_COLOR", "Diffuse Color", 3, True, "use_diffuse_color"),
    ("DIFFUSE_DIRECT", "Diffuse Direct", 3, True, "use_diffuse_direct"),
    ("DIFFUSE_INDIRECT", "Diffuse Indirect", 3, True, "use_diffuse_indirect"),
    ("GLOSSY_COLOR", "Glossy Color", 3, True, "use_glossy_color"),
    ("GLOSSY_DIRECT", "Glossy Direct", 3, True, "use_glossy_direct"),
    ("GLOSSY_INDIRECT", "Glossy Indirect", 3, True, "use_glossy_indirect"),
    ("TRANSMISSION_COLOR", "Transmission Color", 3, True, "use_transmission_color"),
    ("TRANSMISSION_DIRECT", "Transmission Direct", 3, True, "use_transmission_direct"),
    ("TRANSMISSION_INDIRECT", "Transmission Indirect", 3, True, "use_transmission_indirect"),
    ("SUBSURFACE_COLOR", "Subsurface Color", 3, True, "use_subsurface_color"),
    ("SUBSURFACE_DIRECT", "Subsurface Direct", 3, True, "use_subsurface_direct"),
    ("SUBSURFACE_INDIRECT", "Subsurface Indirect", 3, True, "use_subsurface_indirect"),
    ("VOLUME_DIRECT", "Volume Direct", 3, True, "use_volume_direct"),
    ("VOLUME_INDIRECT", "Volume Indirect", 3, True, "use_volume_indirect"),
)
_CHANNELS = {name: channels for _, name, channels, _, _ in _PASSES}
_NOT_RGB = {name for _, name, _, rgb, _ in _PASSES if not rgb}
# passes whose "colour" pass is the pass itself (RenderPasses.py:92-103,112-123)
_SELF_COLORED = ("Alpha", "Emission", "Environment", "Ambient Occlusion", "Shadow")


def _apply_special_cases(name):
  for prefix in _SELF_COLORED:
    if name.startswith(prefix):
      return prefix
  return name


class RenderPasses:

  @staticmethod
  def number_of_channels(render_pass_name):
    """1 for Alpha / Depth, else 3 (RenderPasses.py:40-44); unknown names (incl. '' and None) are 3."""
    return 1 if render_pass_name in ("Alpha", "Depth") else 3

  @staticmethod
  def is_combined_feature_render_pass(render_pass_name):
    return render_pass_name in ("Diffuse", "Glossy", "Subsurface", "Transmission")

  @staticmethod
  def is_volume_render_pass(render_pass_name):
    return "Volume" in render_pass_name

  @staticmethod
  def is_direct_or_indirect_render_pass(render_pass_name):
    return render_pass_name.endswith((" Direct", " Indirect"))

  @staticmethod
  def is_color_render_pass(render_pass_name):
    return render_pass_name.endswith(" Color")

  @staticmethod
  def is_rgb_color_render_pass(render_pass_name):
    return render_pass_name not in _NOT_RGB

  @staticmethod
  def direct_or_indirect_to_color_render_pass(render_pass_name, replicate_reference_typo=True):
    """' Direct' -> ' Color'.  The reference replaces ' Inirect' (sic, RenderPasses.py:88) for indirect
    passes, i.e. returns them unchanged; that behaviour is kept unless replicate_reference_typo=False."""
    result = None
    if render_pass_name.endswith(" Direct"):
      result = render_pass_name[:-len(" Direct")] + " Color"
    elif render_pass_name.endswith(" Indirect"):
      result = render_pass_name if replicate_reference_typo else render_pass_name[:-len(" Indirect")] + " Color"
    return _apply_special_cases(result)

  @staticmethod
  def combined_to_color_render_pass(render_pass_name):
    return _apply_special_cases(render_pass_name + " Color")

  @staticmethod
  def combined_to_direct_render_pass(render_pass_name):
    return render_pass_name + " Direct"

  @staticmethod
  def combined_to_indirect_render_pass(render_pass_name):
    return render_pass_name + " Indirect"


for _const, _name, _, _, _ in _PASSES:
  setattr(RenderPasses, _const, _name)


class RenderPassesUsage:
  """Ordered subset of passes selected by use_* flags (RenderPassesUsage, RenderPasses.py:132-227)."""
  _FLAGS = tuple((flag, name) for _, name, _, _, flag in _PASSES if flag)

  def __init__(self, **flags):
    for flag, _ in self._FLAGS:
      setattr(self, flag, bool(flags.pop(flag, False)))
    if flags:
      raise TypeError("unknown render pass flags: %s" % sorted(flags))

  def render_passes(self):
    return [name for flag, name in self._FLAGS if getattr(self, flag)]
