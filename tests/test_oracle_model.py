"""Oracle model (oracle/reference_model.py) against the committed golden vectors, both backends, plus the
model-level properties the reference's math implies."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import np_ops, reference_model, torch_ops

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(name):
  z = np.load(os.path.join(GOLDEN, name + ".npz"))
  out = {}
  for k in z.files:
    if "|" in k:
      s, key = k.split("|", 1)
      out.setdefault(int(s), {})[key] = z[k]
  return [out[s] for s in sorted(out)], str(z["inputs_sha256"]), str(z["weights_sha256"])


@pytest.mark.parametrize("name", cases.GOLDEN_CASES)
def test_numpy_float64_oracle_reproduces_golden(name):
  from golden.make_golden import digest
  j, arch, weights, features = cases.build(name)
  gold, in_sha, w_sha = _golden(name)
  assert digest(features) == in_sha, "synthetic input generator changed: regenerate the golden vectors"
  assert digest(weights) == w_sha, "weight initialisation changed: regenerate the golden vectors"
  out = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
  assert len(out) == len(gold)
  for s in range(len(gold)):
    assert set(out[s]) == set(gold[s])
    for k in gold[s]:
      np.testing.assert_allclose(out[s][k], gold[s][k], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", cases.GOLDEN_CASES)
def test_torch_float32_restatement_matches_golden(name):
  j, arch, weights, features = cases.build(name)
  gold, _, _ = _golden(name)
  out = reference_model.Architecture(j, ops=torch_ops, dtype=torch.float32, weights=weights).predict_numpy(features)
  for s in range(len(gold)):
    for k in gold[s]:
      scale = max(1.0, float(np.abs(gold[s][k]).max()))
      assert np.abs(out[s][k] - gold[s][k]).max() <= 2e-5 * scale, (s, k)


def test_output_shapes_keys_and_channel_trimming():
  j, arch, weights, features = cases.build("example")
  out = reference_model.Architecture(j, weights=weights).predict_numpy(features)
  assert len(out) == 3                                   # largest first (Architecture.py:577-579)
  assert len(out[0]) == 17
  for s, d in enumerate(out):
    for k, v in d.items():
      assert k.startswith("prediction/")
      c = 1 if k == "prediction/Alpha" else 3            # Alpha is trimmed to one channel (:159-163)
      assert v.shape == (1, 16 >> s, 16 >> s, c), (k, v.shape)


def test_variable_names_follow_tf_creation_order():
  j, arch, weights, features = cases.build("example")
  oracle = reference_model.Architecture(j, weights={})     # let the oracle create its own variables
  oracle.predict(features)
  names = oracle.store.created
  assert names[0] == "embedding/feature_flags_embedding_matrix"
  assert names[1] == "reused_core_architecture/conv2d/kernel"
  assert "reused_core_architecture/conv2d_transpose_1/kernel" in names
  assert "reused_core_architecture/conv2d_25/bias" in names       # 20 backbone + 6 post-process convs
  assert "reused_core_architecture/conv2d_26/kernel" not in names
  assert names[-1] == "reused_compose_scales/conv2d_5/bias"
  assert oracle.store.values["embedding/feature_flags_embedding_matrix"].shape == (17, 8)
  # the product's static variable list is the same list (names, order, shapes)
  assert [(n, tuple(oracle.store.values[n].shape)) for n in names] == arch.spec.variable_shapes()


def test_compose_scales_properties():
  """compose(small, large) == large when up(small) == up(down(large)) for ANY weight net; weights lie in
  [0.5, 1) because of sigmoid(relu(.)) (MultiScalePrediction.py:48-52,73-77)."""
  rng = np.random.default_rng(3)
  large = rng.standard_normal((1, 8, 8, 3))
  small = np_ops.avg_pool_same(large, 2)
  store = reference_model.VariableStore(seed=11)
  store.enter_scope("reused_compose_scales")
  out = reference_model.compose_scales(np_ops, store, small, large)
  store.exit_scope()
  np.testing.assert_allclose(out, large, atol=1e-12)
  # a different small image changes only the low-frequency part: out - large == w * (up(small') - up(down(large)))
  small2 = small + 1.0
  store.enter_scope("reused_compose_scales")
  out2 = reference_model.compose_scales(np_ops, store, small2, large)
  store.exit_scope()
  w = (out2 - large) / 1.0
  assert np.all(w >= 0.5 - 1e-12) and np.all(w < 1.0)
  np.testing.assert_allclose(w[..., 0], w[..., 1])       # one weight per pixel, shared by the channels


def test_non_loaded_passes_return_source_crop():
  j, arch, weights, features = cases.build("combined_onehot")
  out = reference_model.Architecture(j, weights=weights).predict_numpy(features)
  # 'Volume Color' is generated (ones); standardised: log1p(1); the prediction is that constant at every scale
  for s, d in enumerate(out):
    np.testing.assert_allclose(d["prediction/Volume Color"], np.log1p(1.0))
    np.testing.assert_allclose(d["prediction/Alpha Direct"], np.log1p(0.5))
    assert d["prediction/Volume Color"].shape[1] == 16 >> s


def test_training_golden_is_reproduced_by_the_oracle():
  """tests/golden/training_example.npz (loss + every parameter gradient of the restated loss, incl. variation and masked-mean
  terms) is what the float64 torch oracle computes today."""
  import importlib.util
  import os
  here = os.path.dirname(os.path.abspath(__file__))
  spec = importlib.util.spec_from_file_location("make_training_golden", os.path.join(here, "golden", "make_training_golden.py"))
  gen = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(gen)
  j, arch, weights, features, targets = gen.problem()
  loss, grads = gen.oracle_loss_and_gradients(j, weights, features, targets)
  z = np.load(os.path.join(here, "golden", "training_example.npz"))
  assert abs(loss - float(z["loss"])) <= 1e-9 * max(1.0, abs(loss))
  for k, g in grads.items():
    want = z["grad|" + k].astype(np.float64)
    assert np.abs(g - want).max() <= 1e-6 * max(1e-6, np.abs(want).max()) + 1e-12, k
