"""ORACLE (test infrastructure, not product code) - the training loss of the reference restated on torch-CPU so
that torch.autograd provides the gradient oracle for the CUDA backward kernels.

Follows Training.py: model_fn loss assembly :611-660, BaseFeatureTraining.loss :210-243 (scale weights (1/4^s) / sum),
mean :126-129, FeatureTraining / CombinedFeatureTraining / CombinedImageFeatureTraining.initialize :374-495 and
LossDifference.difference (LossDifference.py:15-36), variation_mean :139-186 + :304-346 and masked_mean :131-137 with the
masks of FeatureTraining / CombinedFeatureTraining.initialize :374-392, :434-437, and ms_ssim :188-204 with
tf.image.ssim_multiscale restated from TensorFlow's image_ops_impl.py [external].  PARITY: loss and every parameter gradient
are pinned against the reference's own Training.main() + model_fn executed over oracle/tf_shim
(tests/golden/refshim_training_example.npz: equal to the last bit for the loss); MS-SSIM is not part of that pin (the shim
does not restate tf.image.ssim_multiscale) and TensorFlow's kernels stay unpinned (see oracle/np_ops.py).
"""
import torch

from . import torch_ops

LIGHTS = ("Diffuse", "Glossy", "Subsurface", "Transmission")
IMAGE_TERMS = ("Volume Direct", "Volume Indirect", "Emission", "Environment")


def multiscale_targets(labels, n_scales):
  """Training.py:611-623: scale s = average pooling by 2^s of every label."""
  out = [labels]
  for s in range(1, n_scales):
    out.append({k: torch_ops.avg_pool_same(v, 2 ** s) for k, v in labels.items()})
  return out


def variation_mean(predicted, target, kind):
  """Training.py:139-186: mean over the concatenated horizontal / vertical variation differences."""
  hor = torch_ops.loss_difference(predicted[:, :, 1:, :] - predicted[:, :, :-1, :], target[:, :, 1:, :] - target[:, :, :-1, :], kind)
  ver = torch_ops.loss_difference(predicted[:, 1:, :, :] - predicted[:, :-1, :, :], target[:, 1:, :, :] - target[:, :-1, :, :], kind)
  return torch.cat([hor.reshape(hor.shape[0], -1), ver.reshape(ver.shape[0], -1)], dim=1).mean()


def non_zero_mask(x):
  """Conv2dUtilities.non_zero_mask (Conv2dUtilities.py:69-74)."""
  return torch.sign(x.abs().sum(dim=3))


def masked_mean(predicted, target, mask_source, kind):
  """Training.py:131-137."""
  mask = non_zero_mask(mask_source)
  total = mask.sum()
  if float(total) <= 0:
    return torch.zeros((), dtype=predicted.dtype)
  return (torch_ops.loss_difference(predicted, target, kind) * mask / total).sum()


def _fspecial_gauss(size, sigma, dtype):
  """tf.image's _fspecial_gauss [external]: softmax-normalised 2-D Gaussian."""
  coords = torch.arange(size, dtype=dtype) - (size - 1) / 2.0
  g = coords ** 2 * (-0.5 / sigma ** 2)
  g = g.reshape(1, -1) + g.reshape(-1, 1)
  return torch.softmax(g.reshape(-1), dim=0).reshape(size, size)


def ssim_multiscale(img1, img2, max_val=1.0, power_factors=(0.0448, 0.2856, 0.3001), filter_size=11, filter_sigma=1.5,
                    k1=0.01, k2=0.03):
  """tf.image.ssim_multiscale [external] on NHWC tensors with even sizes at every level: returns [N]."""
  import torch.nn.functional as F
  x, y = img1.permute(0, 3, 1, 2), img2.permute(0, 3, 1, 2)
  ch = x.shape[1]
  kernel = _fspecial_gauss(filter_size, filter_sigma, x.dtype).reshape(1, 1, filter_size, filter_size).repeat(ch, 1, 1, 1)
  reducer = lambda t: F.conv2d(t, kernel, groups=ch)      # noqa: E731  depthwise, VALID
  c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
  mcs = []
  for k in range(len(power_factors)):
    if k > 0:
      assert x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
      x, y = F.avg_pool2d(x, 2), F.avg_pool2d(y, 2)
    mean0, mean1 = reducer(x), reducer(y)
    num0, den0 = mean0 * mean1 * 2.0, mean0 ** 2 + mean1 ** 2
    luminance = (num0 + c1) / (den0 + c1)
    num1, den1 = reducer(x * y) * 2.0, reducer(x ** 2 + y ** 2)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    ssim_per_channel = (luminance * cs).mean(dim=(2, 3))
    mcs.append(torch.relu(cs.mean(dim=(2, 3))))
  mcs.pop()
  stacked = torch.stack(mcs + [torch.relu(ssim_per_channel)], dim=-1)            # [N, C, levels]
  pw = torch.tensor(power_factors, dtype=x.dtype)
  return torch.prod(stacked ** pw, dim=-1).mean(dim=-1)


def ms_ssim_loss(predicted, target):
  """BaseFeatureTraining.ms_ssim (Training.py:188-204): 1 - mean(ssim_multiscale(..., max_val = 1, three power factors))."""
  return 1.0 - ssim_multiscale(predicted, target, 1.0).mean()


def feature_loss(predicted, target, kind, weight, use_multiscale_loss=True, variation_weight=0.0, masked_weight=0.0,
                 mask_source=None, ms_ssim_weight=0.0):
  """BaseFeatureTraining.loss (Training.py:210-243) without the MS-SSIM terms."""
  scales = len(target) if use_multiscale_loss else 1
  norm = 1.0 / sum(1.0 / 4.0 ** s for s in range(scales))
  result = 0.0
  for s in range(scales):
    factor = norm / 4.0 ** s
    if weight > 0:
      result = result + weight * factor * torch_ops.loss_difference(predicted[s], target[s], kind).mean()
    if variation_weight > 0:
      result = result + variation_weight * factor * variation_mean(predicted[s], target[s], kind)
    if masked_weight > 0:
      result = result + masked_weight * factor * masked_mean(predicted[s], target[s], mask_source[s], kind)
  if ms_ssim_weight > 0:
    result = result + ms_ssim_weight * ms_ssim_loss(predicted[0], target[0])
  return result


def mask_pass(name):
  """FeatureTraining.initialize (Training.py:380-388) incl. the reference's ' Inirect' typo (RenderPasses.py:88)."""
  if name.endswith(" Color") or name in ("Environment", "Emission", "Volume Direct", "Volume Indirect"):
    return name
  if name.endswith(" Direct"):
    return name.replace(" Direct", " Color")
  if name.endswith(" Indirect"):
    return name
  return None


def combined_tuples_of(model):
  """[(tuple name, [(member name, load_data, is Color) x 3])] of a COMBINED-tuple architecture (any object with
  .feature_prediction_tuples whose members have .name / .load_data / .kind or .feature_prediction_type), None for SINGLE tuples."""
  tuples = getattr(model, "feature_prediction_tuples", None) or []
  out = []
  for tup in tuples:
    members = list(tup.feature_predictions)
    if len(members) != 3:
      return None
    rows = []
    for index, fp in enumerate(members):
      rows.append((fp.name, bool(fp.load_data), index == 0))
    out.append((tup.name, rows))
  return out or None


def combined_to_color_pass(name):
  """RenderPasses.combined_to_color_render_pass (RenderPasses.py:106-122): the pass whose target masks a combined feature."""
  return name if name in ("Alpha", "Emission", "Environment", "Ambient Occlusion", "Shadow") else name + " Color"


def total_loss(predictions, labels, loaded_names, kind="SMAPE", feature_weight=1.0, combined_feature_weight=5.0,
               combined_image_weight=10.0, use_multiscale_loss=True, feature_variation_weight=0.0, feature_masked_weight=0.0,
               combined_feature_variation_weight=0.0, combined_feature_masked_weight=0.0, combined_image_variation_weight=0.0,
               feature_ms_ssim_weight=0.0, combined_feature_ms_ssim_weight=0.0, combined_image_ms_ssim_weight=0.0,
               combined_tuples=None):
  """predictions: list over scales of {'prediction/<Pass>': tensor}; labels: {'target_image/<Pass>': tensor}.

  combined_tuples (combined_tuples_of(model)): COMBINED tuple type.  Training.main() then builds a CombinedFeatureTraining for
  EVERY tuple (Training.py:1142-1171), also for those whose Direct / Indirect (or Color) members are generated: their
  'prediction' is the standardised generated source (Architecture.py:151-157), their target the raw constant of
  FeatureTrainingLoader.add_to_targets_dictionary (Training.py:538-549: 1 for Color, 0.5 for Direct / Indirect)."""
  targets = multiscale_targets(labels, len(predictions))
  p = lambda name: [d["prediction/" + name] for d in predictions]     # noqa: E731
  t = lambda name: [d["target_image/" + name] for d in targets]       # noqa: E731
  loss = 0.0
  if feature_weight > 0 or feature_variation_weight > 0 or feature_masked_weight > 0 or feature_ms_ssim_weight > 0:
    for name in loaded_names:
      mask = t(mask_pass(name)) if (feature_masked_weight > 0 and mask_pass(name) is not None) else None
      loss = loss + feature_loss(p(name), t(name), kind, feature_weight, use_multiscale_loss, feature_variation_weight,
                                 feature_masked_weight if mask is not None else 0.0, mask, feature_ms_ssim_weight)
  use_combined = (combined_feature_weight > 0 or combined_feature_variation_weight > 0 or combined_feature_masked_weight > 0 or
                  combined_feature_ms_ssim_weight > 0)
  lights = [l for l in LIGHTS if all((l + k) in loaded_names for k in (" Color", " Direct", " Indirect"))]
  comb_p, comb_t = {}, {}
  if combined_tuples is None:
    groups = [(l, [(l + " Color", True, True), (l + " Direct", True, False), (l + " Indirect", True, False)]) for l in lights]
  else:
    groups = list(combined_tuples)
  for name, members in groups:
    def member_target(member, s):
      m_name, loaded, is_color = member
      if loaded:
        return t(m_name)[s]
      like = predictions[s]["prediction/" + m_name]
      return torch.full_like(like, 1.0 if is_color else 0.5)
    n_scales = len(predictions)
    preds = [[p(m[0])[s] for m in members] for s in range(n_scales)]
    tgts = [[member_target(m, s) for m in members] for s in range(n_scales)]
    comb_p[name] = [c * (d + i) for c, d, i in preds]
    comb_t[name] = [c * (d + i) for c, d, i in tgts]
    if use_combined:
      color_pass = combined_to_color_pass(name)
      if combined_feature_masked_weight > 0:
        source = next((m for m in members if m[0] == color_pass), None)
        mask = [member_target(source, s) for s in range(n_scales)] if source is not None else t(color_pass)
      else:
        mask = None
      loss = loss + feature_loss(comb_p[name], comb_t[name], kind, combined_feature_weight, use_multiscale_loss,
                                 combined_feature_variation_weight, combined_feature_masked_weight, mask,
                                 combined_feature_ms_ssim_weight)
  terms = [x for x in IMAGE_TERMS if x in loaded_names]
  if ((combined_image_weight > 0 or combined_image_variation_weight > 0 or combined_image_ms_ssim_weight > 0) and
      len(lights) == 4 and len(terms) == 4):
    img_p = [sum(comb_p[l][s] for l in lights) + sum(p(x)[s] for x in terms) for s in range(len(predictions))]
    img_t = [sum(comb_t[l][s] for l in lights) + sum(t(x)[s] for x in terms) for s in range(len(predictions))]
    loss = loss + feature_loss(img_p, img_t, kind, combined_image_weight, use_multiscale_loss, combined_image_variation_weight,
                               ms_ssim_weight=combined_image_ms_ssim_weight)
  return loss
