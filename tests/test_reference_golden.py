"""The oracle against the reference's own code.

tests/golden/refshim_<case>.npz hold float64 outputs of the UNMODIFIED reference modules (/root/reference/TensorFlow/
Architecture.py and everything it imports) executed over oracle/tf_shim - a torch-backed stand-in for the TensorFlow 1.x
symbols they use - by tests/golden/make_reference_golden.py.  Here:
  * oracle/reference_model.py (numpy float64 and torch float32 backends) must reproduce them,
  * the variable names the reference requested, in its creation order, must be the oracle's and the product's,
  * the shim's restatements of TensorFlow kernels are checked against independent definitions (torch's own 'same' padding,
    numpy.pad, autograd of the forward convolution for conv2d_transpose, a literal loop for SAME pooling),
  * when the reference sources are present (this container, not the GPU box) the fixtures are regenerated and compared.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import cases
from oracle import np_ops, reference_model, torch_ops

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def _maker():
  spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GOLDEN, "make_reference_golden.py"))
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m


def _shim():
  """The shim under a private module name (so that `tensorflow` stays unclaimed in this process)."""
  spec = importlib.util.spec_from_file_location("dd_tf_shim", os.path.join(os.path.dirname(HERE), "oracle", "tf_shim", "tensorflow",
                                                                          "__init__.py"))
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m


def _fixture(name):
  z = np.load(os.path.join(GOLDEN, "refshim_" + name + ".npz"))
  out = {}
  for k in z.files:
    if "|" in k:
      s, key = k.split("|", 1)
      out.setdefault(int(s), {})[key] = z[k]
  return [out[s] for s in sorted(out)], z


@pytest.mark.parametrize("name", cases.GOLDEN_CASES)
def test_numpy_oracle_reproduces_the_reference_code(name):
  from golden.make_golden import digest
  j, arch, weights, features = cases.build(name)
  want, z = _fixture(name)
  assert digest(features) == str(z["inputs_sha256"]) and digest(weights) == str(z["weights_sha256"]), \
      "synthetic inputs / weights changed: regenerate with tests/golden/make_reference_golden.py"
  got = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
  assert len(got) == len(want)
  for s in range(len(want)):
    assert set(got[s]) == set(want[s])
    for k in want[s]:
      assert got[s][k].shape == want[s][k].shape, (s, k)
      scale = max(1.0, float(np.abs(want[s][k]).max()))
      assert np.abs(got[s][k] - want[s][k]).max() <= 1e-8 * scale, (s, k)       # float64 against float64 (measured: 1e-10, summation order)


@pytest.mark.parametrize("name", cases.GOLDEN_CASES)
def test_torch_float32_oracle_matches_the_reference_code(name):
  j, arch, weights, features = cases.build(name)
  want, _ = _fixture(name)
  got = reference_model.Architecture(j, ops=torch_ops, dtype=torch.float32, weights=weights).predict_numpy(features)
  for s in range(len(want)):
    for k in want[s]:
      scale = max(1.0, float(np.abs(want[s][k]).max()))
      assert np.abs(got[s][k] - want[s][k]).max() <= 2e-5 * scale, (s, k)


@pytest.mark.parametrize("name", cases.BASELINE_CASES)
def test_oracle_reproduces_the_reference_code_on_the_baseline_architectures(name):
  """BASELINE.json's own networks (U-Net [64,96,128]x4 K=5 with 32 input channels - the benchmarked one -, the Tiramisu K=21,
  the 9-channel cfg1) on one small tile; the fixtures are stored in float32."""
  j, arch, weights, features = cases.build(name)
  want, z = _fixture(name)
  got = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
  assert len(got) == len(want) == 3
  for s in range(len(want)):
    assert set(want[s]) <= set(got[s])
    for k in want[s]:
      scale = max(1.0, float(np.abs(want[s][k]).max()))
      assert np.abs(got[s][k] - want[s][k]).max() <= 2e-7 * scale, (s, k)       # float32 rounding of the stored fixture
  requested = str(z["variables_in_creation_order"]).split("\n")
  assert [n for n, _ in arch.spec.variable_shapes()] == requested


@pytest.mark.parametrize("name", ["example", "tiramisu"])
def test_variable_names_and_creation_order_are_the_references(name):
  j, arch, weights, features = cases.build(name)
  _, z = _fixture(name)
  requested = str(z["variables_in_creation_order"]).split("\n")
  oracle = reference_model.Architecture(j, weights={})
  oracle.predict(features)
  assert oracle.store.created == requested
  assert [n for n, _ in arch.spec.variable_shapes()] == requested


# ------------------------------------------------------------------------------------------------ building blocks
def _components():
  m = _maker()
  z = np.load(os.path.join(GOLDEN, "refshim_components.npz"))
  return m.component_inputs(), z


@pytest.mark.parametrize("k", [3, 5, 7, 21])
def test_kernel_prediction_oracle_matches_the_reference_code(k):
  inp, z = _components()
  got = np_ops.kernel_prediction(inp["kp%d|src" % k], inp["kp%d|logits" % k], k)
  assert np.abs(got - z["kp%d|out" % k]).max() <= 1e-12


def test_variance_loss_difference_and_log_transforms_match_the_reference_code():
  inp, z = _components()
  for mode in ("uniform", "neighbor"):
    for rel in (False, True):
      for one in (False, True):
        got = np_ops.variance_feature(inp["var|x"], mode, rel, one)
        want = z["var|%s|%d|%d" % (mode, rel, one)]
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max()), (mode, rel, one)
  for kind in ("DIFFERENCE", "ABSOLUTE", "SMOOTH_ABSOLUTE", "SQUARED", "SMAPE"):
    got = np_ops.loss_difference(inp["loss|p"], inp["loss|t"], kind)
    assert np.abs(got - z["loss|" + kind]).max() <= 1e-12, kind
    got = torch_ops.loss_difference(torch.from_numpy(inp["loss|p"]), torch.from_numpy(inp["loss|t"]), kind).numpy()
    assert np.abs(got - z["loss|" + kind]).max() <= 1e-12, kind
  assert np.abs(np_ops.signed_log1p(inp["util|x"]) - z["util|log1p"]).max() <= 1e-13
  assert np.abs(np_ops.signed_expm1(inp["util|x"]) - z["util|expm1"]).max() <= 1e-12 * np.abs(z["util|expm1"]).max()


def test_compose_scales_oracle_matches_the_reference_code():
  inp, z = _components()
  weights = {"reused_compose_scales/" + k.split("|", 1)[1]: v for k, v in inp.items() if k.startswith("compose|conv2d")}
  store = reference_model.VariableStore(weights)
  store.enter_scope("reused_compose_scales")
  got = reference_model.compose_scales(np_ops, store, inp["compose|small"], inp["compose|large"])
  store.exit_scope()
  assert np.abs(got - z["compose|out"]).max() <= 1e-12
  assert np.abs(np_ops.avg_pool_same(inp["compose|large"], 4) - z["compose|down4"]).max() <= 1e-13
  assert np.array_equal(np_ops.resize_nearest_x2(inp["compose|small"]), z["compose|up"])


def test_augmentation_oracle_and_host_matrix_match_the_reference_code():
  """oracle/np_augment.py and the product's host-side rotation matrix (deepdenoiser_b200/augmentation.py) against
  DataAugmentation.py itself (flip / rot90 incl. the screen-space-normal sign rules, the five RGB permutations,
  random_rotation_matrix, rotate_normal)."""
  from deepdenoiser_b200 import augmentation
  from oracle import np_augment
  m = _maker()
  z = np.load(os.path.join(GOLDEN, "refshim_components.npz"))
  image = m.det((5, 7, 3), 77)
  for name in ("Diffuse Color", "Screen Space Normal"):
    tag = name.replace(" ", "")
    for flip in (0, 1):
      assert np.array_equal(np_augment.flip_left_right(image, name, flip), z["augment|flip|%s|%d" % (tag, flip)]), (name, flip)
    for k in range(4):
      assert np.array_equal(np_augment.rotate_90(image, k, name), z["augment|rot|%s|%d" % (tag, k)]), (name, k)
  for permute in range(6):
    assert np.array_equal(np_augment.permute_rgb(image, permute), z["augment|perm|%d" % permute]), permute
  for i, vec in enumerate(m.AUGMENT_VECTORS):
    want = z["augment|matrix|%d" % i]
    got = augmentation.random_rotation_matrix(vec)
    assert np.abs(np.asarray(got, dtype=np.float64) - want).max() <= 1e-6, vec          # the product builds it in float32
    assert np.abs(np_augment.rotate_normal(image, want) - z["augment|normal|%d" % i]).max() <= 1e-13


def test_augmentation_pipeline_order_matches_the_references_feature_training_augmentation():
  """FeatureTrainingAugmentation driven as input_fn_tfrecords drives it (Training.py:551-605, 803-819) against
  oracle/np_augment.augment_example: flip -> rot90 -> RGB permutation (colour passes only) -> normal rotation ('Normal' only),
  sources and targets, 1- and 3-channel passes."""
  from deepdenoiser_b200.RenderPasses import RenderPasses
  from oracle import np_augment
  m = _maker()
  z = np.load(os.path.join(GOLDEN, "refshim_components.npz"))
  for tag, passes, flip, rot, perm, vec in m.AUGMENT_PIPELINES:
    matrix = z["augment|matrix|%d" % m.AUGMENT_VECTORS.index(vec)] if vec is not None else None
    for i, (name, c, is_target) in enumerate(passes):
      for kind, seed in (("source_image/0/", 300 + i), ("target_image/", 400 + i)):
        if kind == "target_image/" and not is_target:
          continue
        got = np_augment.augment_example(m.det((6, 6, c), seed), name, RenderPasses.is_rgb_color_render_pass(name), flip=flip, rot=rot,
                                         perm=perm, rotation=matrix)
        want = z["augment|pipeline|%s|%s%s" % (tag, kind, name)]
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-13, (tag, kind, name)


def test_oracle_loss_and_gradients_match_the_references_model_fn():
  """tests/golden/refshim_training_example.npz: the reference's Training.main() built its loss objects from the training JSON,
  its model_fn (Training.py:607-725) produced the loss, torch autograd through the shim the gradients."""
  spec = importlib.util.spec_from_file_location("make_training_golden", os.path.join(GOLDEN, "make_training_golden.py"))
  mtg = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mtg)
  j, arch, weights, features, targets = mtg.problem()
  loss, grads = mtg.oracle_loss_and_gradients(j, weights, features, targets)
  z = np.load(os.path.join(GOLDEN, "refshim_training_example.npz"))
  assert abs(loss - float(z["loss"])) <= 1e-12 * abs(float(z["loss"]))
  checked = 0
  for key in z.files:
    if key.startswith("grad|"):
      want = z[key]
      got = grads[key[5:]]
      assert np.abs(got - want).max() <= 1e-9 * max(1e-6, np.abs(want).max()), key
      checked += 1
  assert checked == len(weights)
  # the reference tracked a mean metric for every loaded pass, every combined light and the combined image at every scale
  metrics = str(z["metrics"]).split("\n")
  assert "combined_mean/1" in metrics and "combined_diffuse_mean/2" in metrics and "alpha_mean/1" in metrics


def test_oracle_loss_for_combined_tuples_matches_the_references_model_fn():
  """COMBINED tuples: Training.main() builds a combined-feature loss for EVERY tuple, also for Alpha / Emission / Environment /
  Volume whose generated members 'predict' their standardised source against the loader's constant targets
  (tests/golden/refshim_training_combined.npz; ABSOLUTE differences, variation terms)."""
  spec = importlib.util.spec_from_file_location("make_training_golden", os.path.join(GOLDEN, "make_training_golden.py"))
  mtg = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mtg)
  j, arch, weights, features, targets = mtg.problem_combined()
  loss, grads = mtg.oracle_loss_and_gradients(j, weights, features, targets, mtg.COMBINED_LOSS_ARGS)
  z = np.load(os.path.join(GOLDEN, "refshim_training_combined.npz"))
  assert abs(loss - float(z["loss"])) <= 1e-12 * abs(float(z["loss"])), (loss, float(z["loss"]))
  for key in z.files:
    if key.startswith("grad|"):
      want = z[key].astype(np.float64)
      assert np.abs(grads[key[5:]] - want).max() <= 1e-6 * max(1e-6, np.abs(want).max()), key     # float32-stored fixture


def test_tiled_prediction_matches_the_references_prediction_main():
  """tests/golden/refshim_prediction.npz: the reference's Prediction.main() (Prediction.py:188-519 - tile grid, per-tile
  prediction, crop / stitch, lighting = colour x (direct + indirect), combined image) on a 40 x 72 frame with 32-pixel tiles and
  4 pixels of overlap.  Here: the product's tile grid / stitch / combine (deepdenoiser_b200/prediction.py) around the ORACLE
  as the per-tile network."""
  from deepdenoiser_b200 import prediction
  m = _maker()
  j, arch, weights, frame, h, w, tile, overlap = m.prediction_problem()
  z = np.load(os.path.join(GOLDEN, "refshim_prediction.npz"))
  oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights)
  feats = {"source_image/0/" + name: img for name, img in frame.items()}
  got = prediction.predict_image(
      None, feats, h, w, tile, overlap, tiles_per_batch=1,
      predict_fn=lambda f: {k: torch.from_numpy(v) for k, v in oracle.predict_numpy({kk: vv.numpy() for kk, vv in f.items()})[0].items()})
  image, combined = prediction.combine_passes(got)
  assert sorted(z.files) == sorted(["Combined"] + [k[len("prediction/"):] for k in got])
  for key in z.files:
    mine = image if key == "Combined" else got["prediction/" + key]
    assert tuple(mine.shape) == z[key].shape, (key, tuple(mine.shape), z[key].shape)
    assert np.abs(mine.numpy() - z[key]).max() <= 2e-6 * max(1.0, float(np.abs(z[key]).max())), key    # float32-stored fixture
  assert z["Alpha"].shape == (h, w, 1) and z["Combined"].shape == (h, w, 3)
  tiles, _, _ = prediction.tile_grid(h, w, tile, overlap)
  assert len(tiles) == 6                                   # 2 x 3: first / interior / last column


# ------------------------------------------------------------------------------------------------ the shim's kernels
def test_shim_conv2d_same_is_torchs_same_padding_and_valid_is_unpadded():
  tf = _shim()
  g = torch.Generator().manual_seed(3)
  x = torch.randn(2, 9, 7, 5, generator=g, dtype=torch.float64)
  for k in (1, 3, 5):
    w = torch.randn(k, k, 5, 4, generator=g, dtype=torch.float64)
    tf.reset({"conv2d/kernel": w, "conv2d/bias": torch.zeros(4, dtype=torch.float64)})
    got = tf.layers.conv2d(x, 4, (k, k), padding="same", data_format="channels_last")
    want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding="same").permute(0, 2, 3, 1)
    assert torch.allclose(got, want, atol=1e-12)
    got = tf.nn.conv2d(x, filter=w, strides=[1, 1, 1, 1], padding="VALID", data_format="NHWC")
    assert got.shape == (2, 9 - k + 1, 7 - k + 1, 4)
  # NCHW entry point == NHWC entry point
  tf.reset({"conv2d/kernel": w, "conv2d/bias": torch.ones(4, dtype=torch.float64)})
  a = tf.layers.conv2d(x, 4, (5, 5), padding="same", activation=tf.nn.relu, data_format="channels_last")
  tf.reset({"conv2d/kernel": w, "conv2d/bias": torch.ones(4, dtype=torch.float64)})
  b = tf.layers.conv2d(x.permute(0, 3, 1, 2), 4, (5, 5), padding="same", activation=tf.nn.relu, data_format="channels_first")
  assert torch.allclose(a, b.permute(0, 2, 3, 1), atol=1e-12)


@pytest.mark.parametrize("k,s", [(2, 2), (3, 2)])
def test_shim_conv2d_transpose_is_the_gradient_of_the_same_padded_forward_convolution(k, s):
  """tf.layers.conv2d_transpose is DEFINED as conv2d's input gradient: for a forward convolution `y = conv(u, W)` (SAME,
  stride s, u of the upsampled size), conv2d_transpose(x, W) = d<y, x>/du."""
  tf = _shim()
  g = torch.Generator().manual_seed(5)
  cin, cout, h, w = 3, 4, 5, 6
  x = torch.randn(1, h, w, cin, generator=g, dtype=torch.float64)
  kernel = torch.randn(k, k, cout, cin, generator=g, dtype=torch.float64)       # TF layout of the transposed layer
  tf.reset({"conv2d_transpose/kernel": kernel, "conv2d_transpose/bias": torch.zeros(cout, dtype=torch.float64)})
  got = tf.layers.conv2d_transpose(x, cout, (k, k), strides=(s, s), padding="same", data_format="channels_last")
  assert got.shape == (1, h * s, w * s, cout)
  u = torch.zeros(1, h * s, w * s, cout, dtype=torch.float64, requires_grad=True)
  y = tf.nn.conv2d(u, filter=kernel, strides=[1, s, s, 1], padding="SAME", data_format="NHWC")   # [k,k,in=cout,out=cin]
  assert y.shape == x.shape
  (y * x).sum().backward()
  assert torch.allclose(got, u.grad, atol=1e-12)


def test_shim_same_pooling_pads_the_tail_and_excludes_padding():
  tf = _shim()
  g = torch.Generator().manual_seed(7)
  x = torch.randn(1, 6, 8, 2, generator=g, dtype=torch.float64)
  got = tf.layers.max_pooling2d(x, (3, 3), (2, 2), padding="same", data_format="channels_last")
  want = torch.empty(1, 3, 4, 2, dtype=torch.float64)
  for i in range(3):
    for j in range(4):          # SAME, k 3, s 2, even size: one padded row / column at the END (pad_before = 0)
      want[0, i, j] = x[0, 2 * i:min(6, 2 * i + 3), 2 * j:min(8, 2 * j + 3)].amax(dim=(0, 1))
  assert torch.equal(got, want)
  odd = torch.randn(1, 5, 5, 1, generator=g, dtype=torch.float64)
  got = tf.layers.average_pooling2d(odd, 2, 2, padding="same", data_format="channels_last")
  assert got.shape == (1, 3, 3, 1)
  assert torch.allclose(got[0, 2, 2, 0], odd[0, 4, 4, 0]) and torch.allclose(got[0, 0, 2, 0], odd[0, 0:2, 4, 0].mean())
  assert torch.allclose(got[0, 0, 0, 0], odd[0, 0:2, 0:2, 0].mean())


def test_same_padding_rule_is_the_one_hugging_face_uses_to_run_tensorflow_checkpoints():
  """An outside witness for TensorFlow's SAME rule: transformers' `apply_tf_padding` (MobileNet ports, validated against real
  TensorFlow checkpoints) pads exactly like the shim's `_same_pad` and the oracle's `np_ops.same_padding` for every
  (size, kernel, stride) the path uses and more; torch's 'nearest' interpolation is the shim's resize for any ratio."""
  mobilenet = pytest.importorskip("transformers.models.mobilenet_v2.modeling_mobilenet_v2")
  tf = _shim()
  for k in (1, 2, 3, 5, 7):
    for stride in (1, 2, 3):
      conv = torch.nn.Conv2d(1, 1, k, stride=stride)
      for size in range(1, 20):
        x = torch.zeros(1, 1, size, size + 3)
        padded = mobilenet.apply_tf_padding(x, conv)
        top, bottom = tf._same_pad(size, k, stride)
        left, right = tf._same_pad(size + 3, k, stride)
        assert tuple(padded.shape[-2:]) == (size + top + bottom, size + 3 + left + right), (k, stride, size)
        probe = torch.zeros(1, 1, size, size + 3)
        probe[0, 0, 0, 0] = 1.0
        where = torch.nonzero(mobilenet.apply_tf_padding(probe, conv)[0, 0])[0].tolist()
        assert where == [top, left], (k, stride, size)                     # the smaller half goes in front
        assert tuple(np_ops.same_padding(size, k, stride)[1:]) == (top, bottom), (k, stride, size)
  x = torch.arange(2 * 5 * 7 * 3, dtype=torch.float64).reshape(2, 5, 7, 3)
  for oh, ow in ((10, 14), (8, 9), (3, 4)):
    got = tf.image.resize_images(x, (oh, ow), method=tf.image.ResizeMethod.NEAREST_NEIGHBOR)
    want = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2), size=(oh, ow), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(got, want), (oh, ow)


def test_shim_pad_modes_are_numpys_and_resize_is_pixel_replication():
  tf = _shim()
  x = torch.arange(2 * 4 * 5 * 3, dtype=torch.float64).reshape(2, 4, 5, 3)
  for mode in ("symmetric", "REFLECT", "constant"):
    got = tf.pad(x, [[0, 0], [2, 2], [1, 3], [0, 0]], mode)
    want = np.pad(x.numpy(), [(0, 0), (2, 2), (1, 3), (0, 0)], mode=mode.lower())
    assert np.array_equal(got.numpy(), want), mode
  up = tf.image.resize_images(x, (8, 10), method=tf.image.ResizeMethod.NEAREST_NEIGHBOR)
  assert np.array_equal(up.numpy(), np.repeat(np.repeat(x.numpy(), 2, axis=1), 2, axis=2))


def test_shim_variable_scopes_number_default_names_like_tf1():
  tf = _shim()
  zeros = {n: torch.zeros(1, 1, 1, 1) for n in ("a/conv2d/kernel", "a/conv2d_1/kernel", "a/b/conv2d/kernel", "conv2d/kernel")}
  zeros.update({n.replace("kernel", "bias"): torch.zeros(1) for n in list(zeros)})
  tf.reset(zeros)
  x = torch.zeros(1, 2, 2, 1)
  for reuse in (False, True):                       # re-entering 'a' restarts the numbering: reuse finds the same variables
    with tf.variable_scope("a", reuse=reuse):
      tf.layers.conv2d(x, 1, (1, 1))
      tf.layers.conv2d(x, 1, (1, 1))
      with tf.variable_scope("b"):
        tf.layers.conv2d(x, 1, (1, 1))
  tf.layers.conv2d(x, 1, (1, 1))
  assert tf.created == ["a/conv2d/kernel", "a/conv2d/bias", "a/conv2d_1/kernel", "a/conv2d_1/bias", "a/b/conv2d/kernel",
                        "a/b/conv2d/bias", "conv2d/kernel", "conv2d/bias"]
  with pytest.raises(KeyError):
    with tf.variable_scope("a"):
      for _ in range(3):
        tf.layers.conv2d(x, 1, (1, 1))              # a/conv2d_2 was never provided


# ------------------------------------------------------------------------------------------------ live, where the reference is
@pytest.mark.skipif(not os.path.isdir("/root/reference/TensorFlow"), reason="reference sources are not on this machine")
def test_reference_prediction_main_over_the_shim_reproduces_the_committed_fixture():
  m = _maker()
  saved_path, saved_mods = list(sys.path), dict(sys.modules)
  try:
    tf, mods = m.load_reference()
    got = m.reference_prediction(tf)
  finally:
    sys.path[:] = saved_path
    for k in list(sys.modules):
      if k not in saved_mods:
        del sys.modules[k]
  z = np.load(os.path.join(GOLDEN, "refshim_prediction.npz"))
  assert sorted(z.files) == sorted(got)
  for k in got:
    assert np.array_equal(got[k].astype(np.float32), z[k]), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/TensorFlow"), reason="reference sources are not on this machine")
def test_reference_building_blocks_over_the_shim_reproduce_the_committed_fixture():
  m = _maker()
  saved_path, saved_mods = list(sys.path), dict(sys.modules)
  try:
    tf, mods = m.load_reference()
    got = m.reference_components(tf, mods)
    got.update({"augment|" + k: v for k, v in m.reference_augmentation(tf).items()})
  finally:
    sys.path[:] = saved_path
    for k in list(sys.modules):
      if k not in saved_mods:
        del sys.modules[k]
  z = np.load(os.path.join(GOLDEN, "refshim_components.npz"))
  assert sorted(z.files) == sorted(got)
  for k in got:
    assert np.abs(got[k] - z[k]).max() <= 1e-12 * max(1.0, float(np.abs(z[k]).max())), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/TensorFlow"), reason="reference sources are not on this machine")
@pytest.mark.parametrize("name", ["direct", "variants"])
def test_reference_code_over_the_shim_reproduces_the_committed_fixtures(name):
  m = _maker()
  saved_path, saved_mods = list(sys.path), dict(sys.modules)
  try:
    tf, mods = m.load_reference()
    j, arch, weights, features = cases.build(name)
    got, _ = m.run_reference(tf, mods, j, weights, features, "channels_first")
  finally:
    sys.path[:] = saved_path
    for k in list(sys.modules):
      if k not in saved_mods:
        del sys.modules[k]
  want, _ = _fixture(name)
  for s in range(len(want)):
    for k in want[s]:
      assert np.abs(got[s][k] - want[s][k]).max() <= 1e-12 * max(1.0, float(np.abs(want[s][k]).max())), (s, k)


@pytest.mark.skipif(not os.path.isdir("/root/reference/TensorFlow"), reason="reference sources are not on this machine")
def test_random_architectures_oracle_matches_the_reference_code_live():
  """cases.random_case: a seeded sweep over the architecture JSON - the reference's Architecture.predict over the shim against
  the oracle (float64), and the product's variable list against the variables the reference requested."""
  m = _maker()
  saved_path, saved_mods = list(sys.path), dict(sys.modules)
  try:
    tf, mods = m.load_reference()
    for trial in range(16):
      j, arch, weights, features = cases.random_case(trial)
      want, requested = m.run_reference(tf, mods, j, weights, features, "channels_first" if trial % 2 else "channels_last")
      got = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
      assert [n for n, _ in arch.spec.variable_shapes()] == requested, trial
      # the product's host-side bookkeeping (deepdenoiser_b200/Architecture.py) against the reference's objects
      book = m.describe_reference(mods, j)
      row = lambda fp: (fp.name, bool(fp.load_data), bool(fp.is_target), int(fp.number_of_channels), int(fp.number_of_sources),     # noqa: E731
                        bool(fp.preserve_source), bool(fp.invert_standardization) if fp.is_target else None)
      assert [row(fp) for fp in arch.feature_predictions] == book["feature_predictions"], trial
      assert [row(fp) for fp in arch.auxiliary_features] == book["auxiliary_features"], trial
      assert [(t.name, [fp.name for fp in t.feature_predictions]) for t in arch.feature_prediction_tuples] == book["tuples"], trial
      assert len(got) == len(want), trial
      for s in range(len(want)):
        assert set(got[s]) == set(want[s]), trial
        for k in want[s]:
          assert got[s][k].shape == want[s][k].shape, (trial, s, k)
          assert np.abs(got[s][k] - want[s][k]).max() <= 1e-9 * max(1.0, float(np.abs(want[s][k]).max())), (trial, s, k)
  finally:
    sys.path[:] = saved_path
    for k in list(sys.modules):
      if k not in saved_mods:
        del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir("/root/reference/TensorFlow"), reason="reference sources are not on this machine")
def test_source_index_tuples_draw_like_the_references_live():
  """Training.source_index_tuples (Training.py:879-913) under the same `random` seed: same tuples, same required indices."""
  import random
  from deepdenoiser_b200 import tfrecords
  m = _maker()
  saved_path, saved_mods = list(sys.path), dict(sys.modules)
  try:
    tf, mods = m.load_reference()
    sys.path.insert(0, m.REFERENCE)
    sys.path.insert(0, m.SHIM)
    import importlib
    training = importlib.import_module("Training")
    for per_example, tuples, per_target in ((4, 8, 1), (4, 10, 1), (3, 2, 1), (5, 7, 2), (2, 6, 2)):
      for seed in (0, 1, 2):
        random.seed(seed)
        want = training.source_index_tuples(per_example, tuples, per_target)
        random.seed(seed)
        got = tfrecords.source_index_tuples(per_example, tuples, per_target)
        assert (list(got[0]), list(got[1])) == (list(want[0]), list(want[1])), (per_example, tuples, per_target, seed)
    for bad in ((1, 4, 2), (4, 4, 3)):
      with pytest.raises(Exception):
        training.source_index_tuples(*bad)
      with pytest.raises(Exception):
        tfrecords.source_index_tuples(*bad)
  finally:
    sys.path[:] = saved_path
    for k in list(sys.modules):
      if k not in saved_mods:
        del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir("/root/reference/TensorFlow"), reason="reference sources are not on this machine")
def test_render_passes_and_naming_agree_with_the_references_live():
  """The pure-Python helpers of the drop-in boundary (RenderPasses.py, Naming.py - dictionary keys, pass classification, mask
  lookups incl. the ' Inirect' typo) against the reference's own functions, over every pass constant of both trees, every
  combined feature name and every flag combination."""
  import itertools
  from deepdenoiser_b200.Naming import Naming
  from deepdenoiser_b200.RenderPasses import RenderPasses
  m = _maker()
  saved_path, saved_mods = list(sys.path), dict(sys.modules)
  try:
    tf, mods = m.load_reference()
    ref_rp, ref_nm = mods["RenderPasses"].RenderPasses, mods["Naming"].Naming
    constants = {k: v for k, v in vars(ref_rp).items() if k.isupper() and isinstance(v, str)}
    assert constants, "no pass constants found in the reference"
    for key, value in constants.items():
      assert getattr(RenderPasses, key) == value, key
    names = sorted(set(constants.values()) | {"Alpha Direct", "Volume Color", "Foo", "Foo Direct", "Foo Indirect", "Foo Color", ""})
    for name in names:
      for fn in ("number_of_channels", "is_combined_feature_render_pass", "is_volume_render_pass", "is_direct_or_indirect_render_pass",
                 "is_color_render_pass", "is_rgb_color_render_pass", "combined_to_color_render_pass", "combined_to_direct_render_pass",
                 "combined_to_indirect_render_pass"):
        assert getattr(RenderPasses, fn)(name) == getattr(ref_rp, fn)(name), (fn, name)
      if ref_rp.is_direct_or_indirect_render_pass(name):
        assert RenderPasses.direct_or_indirect_to_color_render_pass(name) == ref_rp.direct_or_indirect_to_color_render_pass(name), name
      assert Naming.feature_prediction_name(name) == ref_nm.feature_prediction_name(name)
      assert Naming.feature_flags_name(name) == ref_nm.feature_flags_name(name)
      assert Naming.tensorboard_name(name) == ref_nm.tensorboard_name(name)
      for masked in (False, True):
        assert Naming.target_feature_name(name, masked=masked) == ref_nm.target_feature_name(name, masked=masked)
        for spp, index in itertools.product((None, 16), (None, 0, 3)):
          assert (Naming.source_feature_name(name, samples_per_pixel=spp, index=index, masked=masked) ==
                  ref_nm.source_feature_name(name, samples_per_pixel=spp, index=index, masked=masked)), (name, spp, index, masked)
        for internal, scale in itertools.product((False, True), (None, 0, 2)):
          for fn in ("difference_name", "mean_name", "variation_difference_name", "variation_mean_name"):
            assert (getattr(Naming, fn)(name, masked=masked, internal=internal, scale_index=scale) ==
                    getattr(ref_nm, fn)(name, masked=masked, internal=internal, scale_index=scale)), (fn, name, masked, internal, scale)
          assert Naming.ms_ssim_name(name, masked=masked, internal=internal) == ref_nm.ms_ssim_name(name, masked=masked, internal=internal)
  finally:
    sys.path[:] = saved_path
    for k in list(sys.modules):
      if k not in saved_mods:
        del sys.modules[k]
