"""ORACLE (test infrastructure, not product code) - the training loss of the reference restated on torch-CPU so
that torch.autograd provides the gradient oracle for the CUDA backward kernels.

Follows Training.py: model_fn loss assembly :611-660, BaseFeatureTraining.loss :210-243 (scale weights (1/4^s) / sum),
mean :126-129, FeatureTraining / CombinedFeatureTraining / CombinedImageFeatureTraining.initialize :374-495 and
LossDifference.difference (LossDifference.py:15-36).  Masked / variation / MS-SSIM terms have weight 0 in
TrainingExample.json:31-98 and are not restated.  PARITY UNPINNED (see oracle/np_ops.py).
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


def feature_loss(predicted, target, kind, weight, use_multiscale_loss=True):
  """BaseFeatureTraining.loss with only mean_weight > 0 (Training.py:210-243)."""
  scales = len(target) if use_multiscale_loss else 1
  norm = 1.0 / sum(1.0 / 4.0 ** s for s in range(scales))
  result = 0.0
  for s in range(scales):
    mean = torch_ops.loss_difference(predicted[s], target[s], kind).mean()
    result = result + weight * (norm / 4.0 ** s) * mean
  return result


def total_loss(predictions, labels, loaded_names, kind="SMAPE", feature_weight=1.0, combined_feature_weight=5.0,
               combined_image_weight=10.0, use_multiscale_loss=True):
  """predictions: list over scales of {'prediction/<Pass>': tensor}; labels: {'target_image/<Pass>': tensor}."""
  targets = multiscale_targets(labels, len(predictions))
  p = lambda name: [d["prediction/" + name] for d in predictions]     # noqa: E731
  t = lambda name: [d["target_image/" + name] for d in targets]       # noqa: E731
  loss = 0.0
  if feature_weight > 0:
    for name in loaded_names:
      loss = loss + feature_loss(p(name), t(name), kind, feature_weight, use_multiscale_loss)
  lights = [l for l in LIGHTS if all((l + k) in loaded_names for k in (" Color", " Direct", " Indirect"))]
  comb_p, comb_t = {}, {}
  for l in lights:
    comb_p[l] = [c * (d + i) for c, d, i in zip(p(l + " Color"), p(l + " Direct"), p(l + " Indirect"))]
    comb_t[l] = [c * (d + i) for c, d, i in zip(t(l + " Color"), t(l + " Direct"), t(l + " Indirect"))]
    if combined_feature_weight > 0:
      loss = loss + feature_loss(comb_p[l], comb_t[l], kind, combined_feature_weight, use_multiscale_loss)
  terms = [x for x in IMAGE_TERMS if x in loaded_names]
  if combined_image_weight > 0 and len(lights) == 4 and len(terms) == 4:
    img_p = [sum(comb_p[l][s] for l in lights) + sum(p(x)[s] for x in terms) for s in range(len(predictions))]
    img_t = [sum(comb_t[l][s] for l in lights) + sum(t(x)[s] for x in terms) for s in range(len(predictions))]
    loss = loss + feature_loss(img_p, img_t, kind, combined_image_weight, use_multiscale_loss)
  return loss
