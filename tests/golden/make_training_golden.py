"""Generates tests/golden/training_example.npz: loss and parameter gradients of the float64 torch oracle
(oracle/reference_model.py + oracle/reference_loss.py, torch autograd) for a small U-Net ([16, 24] filters, one conv per
block, K = 3, the passes / tuples / flags of ArchitectureExample.json) with the loss weights of TrainingExample.json plus
non-zero variation / masked-mean weights.

A regression pin of the restated reference; the same problem through the reference's own Training.main() / model_fn is
tests/golden/refshim_training_example.npz (make_reference_golden.py) - the losses agree to the last bit.
Re-run after an intentional oracle change:  python tests/golden/make_training_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
from deepdenoiser_b200 import synthetic  # noqa: E402
from oracle import reference_loss, reference_model, torch_ops  # noqa: E402

LOSS_ARGS = dict(kind="SMAPE", feature_weight=1.0, combined_feature_weight=5.0, combined_image_weight=10.0,
                 feature_variation_weight=0.25, combined_feature_variation_weight=0.5, combined_feature_masked_weight=1.0)


def problem():
  from deepdenoiser_b200.Architecture import Architecture
  j = cases._small(synthetic.example_architecture_json(), [16, 24], 1, 3)
  arch = Architecture(j, seed=4321)
  weights = synthetic.randomize_biases(arch.weights)
  arch.weights = weights
  n, h, w = 2, 16, 16
  features = synthetic.synthetic_features(arch, n, h, w, seed=1234)
  clean = synthetic.synthetic_features(arch, n, h, w, seed=4242)
  targets = {"target_image/" + fp.name: clean["source_image/0/" + fp.name] for fp in arch.feature_predictions if fp.load_data}
  return j, arch, weights, features, targets


COMBINED_LOSS_ARGS = dict(kind="ABSOLUTE", feature_weight=1.0, combined_feature_weight=5.0, combined_image_weight=10.0,
                          feature_variation_weight=0.25, combined_feature_variation_weight=0.5)


def problem_combined():
  """COMBINED tuples (Color, Direct, Indirect per pass; Alpha / Emission / Environment / Volume with generated members)."""
  from deepdenoiser_b200.Architecture import Architecture
  j = cases._small(synthetic.example_architecture_json(), [16, 24], 1, 3)
  j["architecture"]["source_encoder"]["feature_prediction_tuple_type"] = "COMBINED"
  arch = Architecture(j, seed=4321)
  weights = synthetic.randomize_biases(arch.weights)
  arch.weights = weights
  n, h, w = 1, 16, 16
  features = synthetic.synthetic_features(arch, n, h, w, seed=1234)
  clean = synthetic.synthetic_features(arch, n, h, w, seed=4242)
  targets = {"target_image/" + fp.name: clean["source_image/0/" + fp.name] for fp in arch.feature_predictions if fp.load_data}
  return j, arch, weights, features, targets


def oracle_loss_and_gradients(j, weights, features, targets, loss_args=None):
  params = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in weights.items()}
  model = reference_model.Architecture(j, ops=torch_ops, dtype=torch.float64, weights=params)
  preds = model.predict(features)
  labels = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in targets.items()}
  loaded = [fp.name for fp in model.feature_predictions if fp.load_data]
  loss = reference_loss.total_loss(preds, labels, loaded, combined_tuples=reference_loss.combined_tuples_of(model), **(loss_args or LOSS_ARGS))
  loss.backward()
  grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(weights[k], dtype=np.float64)) for k, v in params.items()}
  return float(loss.detach()), grads


def main():
  j, arch, weights, features, targets = problem()
  loss, grads = oracle_loss_and_gradients(j, weights, features, targets)
  payload = {"loss": np.array(loss)}
  for k, g in grads.items():
    payload["grad|" + k] = g.astype(np.float32)
  np.savez_compressed(os.path.join(HERE, "training_example.npz"), **payload)
  print("loss %.6f, %d gradient tensors" % (loss, len(grads)))


if __name__ == "__main__":
  main()
