"""Training path on the GPU (exact fp32 kernels through the C ABI) against torch-autograd of the oracle:
loss value, every parameter gradient, one TF-form Adam step, and that a few steps reduce the loss."""
import copy

import numpy as np
import pytest
import torch

import cases
from deepdenoiser_b200 import synthetic
from deepdenoiser_b200.Architecture import Architecture
from deepdenoiser_b200.training import Trainer, TrainingSettings
from oracle import reference_loss, reference_model, torch_ops

pytestmark = pytest.mark.gpu


def small_example(filters=(16, 24, 32), n_convs=2, k=3, tuple_type="SINGLE", invert_after=True):
  j = synthetic.example_architecture_json()
  core = j["architecture"]["core_architecture"]
  core["number_of_filters_for_convolution_blocks"] = list(filters)
  core["number_of_convolutions_per_block"] = n_convs
  j["architecture"]["kernel_prediction"]["kernel_size"] = k
  j["architecture"]["source_encoder"]["feature_prediction_tuple_type"] = tuple_type
  j["architecture"]["multiscale_prediction"]["invert_standardization_after_multiscale_predictions"] = invert_after
  j["b200"] = {"dtype": "float32"}
  return j


def make_problem(j, n=2, h=16, w=16, seed=21):
  host = Architecture(j)
  weights = synthetic.randomize_biases(host.weights, scale=0.05)
  features = synthetic.synthetic_features(host, n, h, w, seed=seed)
  clean = synthetic.synthetic_features(host, n, h, w, seed=seed + 1)     # "ground truth" renders
  targets = {"target_image/" + fp.name: clean["source_image/0/" + fp.name] for fp in host.feature_predictions if fp.load_data}
  return host, weights, features, targets


def oracle_loss_and_grads(j, weights, features, targets, **loss_args):
  params = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in weights.items()}
  model = reference_model.Architecture(j, ops=torch_ops, dtype=torch.float64, weights=params)
  preds = model.predict(features)
  preds = [{k: v for k, v in d.items()} for d in preds]
  labels = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in targets.items()}
  loaded = [fp.name for fp in model.feature_predictions if fp.load_data]
  loss = reference_loss.total_loss(preds, labels, loaded, **loss_args)
  loss.backward()
  grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(weights[k], dtype=np.float64)) for k, v in params.items()}
  return float(loss.detach()), grads, preds


def check_gradients(trainer, want, rtol=2e-3):
  got = trainer.gradients()
  worst, bad = 0.0, []
  for name, g in want.items():
    scale = max(1e-6, float(np.abs(g).max()))
    err = float(np.abs(got[name] - g).max()) / scale
    worst = max(worst, err)
    if err > rtol:
      bad.append("%s: relative gradient error %.3e (|g|max %.3e)" % (name, err, scale))
  assert not bad, "\n".join(bad)
  return worst


@pytest.mark.parametrize("kind", ["SMAPE", "SQUARED", "ABSOLUTE", "SMOOTH_ABSOLUTE"])
def test_loss_and_all_gradients_match_autograd_single_tuples(kind):
  j = small_example()
  host, weights, features, targets = make_problem(j)
  settings = TrainingSettings({"loss_difference": kind})
  trainer = Trainer(Architecture(j, weights=weights), settings)
  trainer.forward({k: torch.from_numpy(v) for k, v in features.items()})
  loss = float(trainer.loss_and_gradient({k: torch.from_numpy(v) for k, v in targets.items()}).item())
  trainer.backward()
  want_loss, want_grads, want_preds = oracle_loss_and_grads(j, weights, features, targets, kind=kind)
  # forward of the training path == oracle
  got_preds = trainer.predictions()
  for s in range(len(want_preds)):
    for k_, v in want_preds[s].items():
      assert np.abs(got_preds[s][k_].cpu().numpy() - v.detach().numpy()).max() <= 1e-4
  assert abs(loss - want_loss) <= 1e-4 * max(1.0, abs(want_loss)), (loss, want_loss)
  worst = check_gradients(trainer, want_grads)
  print(kind, "loss %.6f (oracle %.6f), worst relative gradient error %.2e" % (loss, want_loss, worst))


def test_gradients_combined_tuples_and_invert_before_compose():
  j = small_example(filters=(16, 16), n_convs=1, k=3, tuple_type="COMBINED", invert_after=False)
  host, weights, features, targets = make_problem(j, n=1, h=8, w=12)
  trainer = Trainer(Architecture(j, weights=weights), TrainingSettings())
  trainer.forward({k: torch.from_numpy(v) for k, v in features.items()})
  loss = float(trainer.loss_and_gradient({k: torch.from_numpy(v) for k, v in targets.items()}).item())
  trainer.backward()
  want_loss, want_grads, _ = oracle_loss_and_grads(j, weights, features, targets)
  assert abs(loss - want_loss) <= 1e-4 * max(1.0, abs(want_loss))
  check_gradients(trainer, want_grads)


def test_gradients_tiramisu_backbone():
  """Dense blocks (shared concat gradient), 1x1 transition + 2x2 max-pool, 3x3 stride-2 transposed convolution."""
  j = small_example(filters=(8, 16, 16), n_convs=2, k=3)
  j["architecture"]["core_architecture"]["name"] = "Tiramisu"
  host, weights, features, targets = make_problem(j, n=2, h=16, w=24)
  trainer = Trainer(Architecture(j, weights=weights), TrainingSettings())
  trainer.forward({k: torch.from_numpy(v) for k, v in features.items()})
  loss = float(trainer.loss_and_gradient({k: torch.from_numpy(v) for k, v in targets.items()}).item())
  trainer.backward()
  want_loss, want_grads, want_preds = oracle_loss_and_grads(j, weights, features, targets)
  got_preds = trainer.predictions()
  for s in range(len(want_preds)):
    for k_, v in want_preds[s].items():
      assert np.abs(got_preds[s][k_].cpu().numpy() - v.detach().numpy()).max() <= 1e-4
  assert abs(loss - want_loss) <= 1e-4 * max(1.0, abs(want_loss)), (loss, want_loss)
  worst = check_gradients(trainer, want_grads)
  print("Tiramisu: loss %.6f (oracle %.6f), worst relative gradient error %.2e" % (loss, want_loss, worst))


def test_adam_step_is_tf_form_and_training_reduces_the_loss():
  j = small_example(filters=(16, 24), n_convs=1, k=3)
  host, weights, features, targets = make_problem(j, n=2, h=16, w=16)
  trainer = Trainer(Architecture(j, weights=weights), TrainingSettings({"learning_rate": 1e-3}))
  f = {k: torch.from_numpy(v) for k, v in features.items()}
  t = {k: torch.from_numpy(v) for k, v in targets.items()}
  first = float(trainer.train_step(f, t).item())
  # one TF-form Adam step from zero moments: theta -= lr * sqrt(1-b2)/(1-b1) * (1-b1) g / (sqrt((1-b2) g^2) + eps)
  _, g, _ = oracle_loss_and_grads(j, weights, features, targets)
  name = "reused_core_architecture/conv2d/kernel"
  lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
  want = weights[name] - lr_t * (0.1 * g[name]) / (np.sqrt(0.001 * g[name] ** 2) + 1e-8)
  got = trainer.get_weights()[name]
  assert np.abs(got - want).max() <= 2e-5
  losses = [first]
  for _ in range(15):
    losses.append(float(trainer.train_step(f, t).item()))
  print("loss trajectory", ["%.4f" % l for l in losses])
  assert losses[-1] < losses[0] - 0.5 and all(b < a + 1e-3 for a, b in zip(losses, losses[1:]))   # steady descent
  # checkpoint round trip
  import os, tempfile
  path = os.path.join(tempfile.mkdtemp(), "ckpt.npz")
  trainer.save_checkpoint(path)
  other = Trainer(Architecture(j, weights=weights), TrainingSettings())
  other.load_checkpoint(path)
  assert other.step_count == trainer.step_count
  assert torch.equal(other.theta, trainer.theta) and torch.equal(other.adam_v, trainer.adam_v)
