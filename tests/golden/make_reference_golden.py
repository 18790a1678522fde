"""Generates tests/golden/refshim_<case>.npz: outputs of the REFERENCE'S OWN Python (the unmodified modules under
/root/reference/TensorFlow, imported from where they lie) for the seeded synthetic inputs / weights of tests/cases.py.

TensorFlow 1.x cannot be installed in this image, so `import tensorflow` inside the reference modules resolves to
oracle/tf_shim/tensorflow - a torch-backed eager stand-in for the ~100 TF symbols these paths use (its header states what
that leaves unverified: TF's kernels are restated, the reference's code is executed as written).  Both data formats of
the reference ('channels_last', its CPU mode, and 'channels_first', its GPU default with the NHWC <-> NCHW conversions of
Conv2dUtilities.convert_to_data_format) are run in float64 and must agree before anything is written.

The vectors pin oracle/reference_model.py (tests/test_reference_golden.py) and, through it, the CUDA path.
/root/reference does not exist on the GPU box: only the .npz files travel.

  python tests/golden/make_reference_golden.py            # writes the fixtures
  python tests/golden/make_reference_golden.py --check    # compares with the committed fixtures, writes nothing
"""
import importlib
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("DD_REFERENCE_DIR", "/root/reference/TensorFlow")
SHIM = os.path.join(ROOT, "oracle", "tf_shim")

REFERENCE_MODULES = ("RenderPasses", "Naming", "Conv2dUtilities", "Utilities", "FeatureFlags", "FeatureEngineering", "KernelPrediction",
                     "MultiScalePrediction", "SourceEncoder", "UNet", "Tiramisu", "LossDifference", "Architecture")


STORED_PASSES = ("prediction/Alpha", "prediction/Diffuse Color", "prediction/Glossy Direct", "prediction/Emission",
                 "prediction/Volume Indirect")


def load_reference():
  """The reference modules (by their own top-level names) with the shim as `tensorflow`; returns (tf shim, modules)."""
  if not os.path.isdir(REFERENCE):
    raise RuntimeError("reference sources not found at %s" % REFERENCE)
  for name in ("tensorflow",) + REFERENCE_MODULES:
    sys.modules.pop(name, None)
  sys.path.insert(0, REFERENCE)
  sys.path.insert(0, SHIM)
  try:
    tf = importlib.import_module("tensorflow")
    assert os.path.abspath(tf.__file__).startswith(SHIM), tf.__file__
    mods = {name: importlib.import_module(name) for name in REFERENCE_MODULES}
    for name, m in mods.items():
      assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE)), (name, m.__file__)
  finally:
    sys.path.remove(SHIM)
    sys.path.remove(REFERENCE)
  return tf, mods


def run_reference(tf, mods, j, weights, features, data_format, dtype=torch.float64):
  """Architecture.predict of the reference (Architecture.py:537-617) -> list over scales of {name: numpy array}."""
  tf.reset({k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in weights.items()})
  arch = mods["Architecture"].Architecture(j, source_data_format="channels_last", data_format=data_format)
  feats = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in features.items()}
  out = arch.predict(feats, tf.estimator.ModeKeys.PREDICT)
  return [{k: v.detach().numpy() for k, v in d.items()} for d in out], list(tf.created)


def describe_reference(mods, j):
  """Bookkeeping of the reference's Architecture object (Architecture.py:367-473): passes, tuples, auxiliaries."""
  arch = mods["Architecture"].Architecture(j, source_data_format="channels_last", data_format="channels_last")
  row = lambda fp: (fp.name, bool(fp.load_data), bool(fp.is_target), int(fp.number_of_channels), int(fp.number_of_sources),     # noqa: E731
                    bool(fp.preserve_source), bool(fp.invert_standardization) if fp.is_target else None)
  return {"feature_predictions": [row(fp) for fp in arch.feature_predictions],
          "auxiliary_features": [row(fp) for fp in arch.auxiliary_features],
          "tuples": [(t.name, [fp.name for fp in t.feature_predictions]) for t in arch.feature_prediction_tuples]}


def reference_training_setup(tf, j, training_json, data_format="channels_last"):
  """Runs the reference's Training.main() (Training.py:944-1232) on JSON files written to a scratch directory until it
  constructs its tf.estimator.Estimator; returns (Training module, model_fn, params) - the loss objects in `params` were
  built by the reference's own code from `training_json`."""
  import argparse
  import json
  import tempfile
  sys.modules.pop("Training", None)
  sys.modules.pop("DataAugmentation", None)
  sys.path.insert(0, REFERENCE)
  sys.path.insert(0, SHIM)
  try:
    training = importlib.import_module("Training")
    assert os.path.abspath(training.__file__).startswith(os.path.abspath(REFERENCE)), training.__file__
  finally:
    sys.path.remove(SHIM)
    sys.path.remove(REFERENCE)
  with tempfile.TemporaryDirectory() as scratch:
    records = os.path.join(scratch, "records")
    os.makedirs(records)
    tj = dict(training_json)
    tj["architecture"] = "architecture.json"
    tj["base_tfrecords_directory"] = records
    with open(os.path.join(scratch, "architecture.json"), "w") as f:
      json.dump(j, f)
    with open(os.path.join(scratch, "training.json"), "w", encoding="utf-8") as f:
      json.dump(tj, f)
    for mode in ("training", "validation"):
      with open(os.path.join(records, mode + ".json"), "w", encoding="utf-8") as f:
        json.dump({"source_samples_per_pixel_list": [16], "tiles_height_width": 16, "number_of_sources_per_example": 1}, f)
    args = argparse.Namespace(json_filename=os.path.join(scratch, "training.json"), threads=1, data_format=data_format,
                              validate=True, train_epochs=1, validation_interval=1)
    try:
      training.main(args)
    except tf.SetupCaptured as captured:
      return training, captured.model_fn, captured.params
  raise RuntimeError("Training.main() returned without constructing an Estimator")


def run_reference_training(tf, mods, j, training_json, weights, features, targets, dtype=torch.float64):
  """loss and d loss / d variable of the reference's model_fn (Training.py:607-725, EVAL mode: no optimizer)."""
  variables = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in weights.items()}
  tf.reset(variables)
  training, model_fn, params = reference_training_setup(tf, j, training_json)
  feats = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in features.items()}
  labels = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in targets.items()}
  spec = model_fn(feats, labels, tf.estimator.ModeKeys.EVAL, params)
  loss = spec.loss
  loss.backward()
  grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape))) for k, v in variables.items()}
  return float(loss.detach()), grads, sorted(spec.eval_metric_ops)


def training_json_for(loss_args):
  """TrainingExample.json (the repo's configs/ copy of the reference schema) with the loss weights of `loss_args`."""
  import json
  with open(os.path.join(ROOT, "configs", "TrainingExample.json")) as f:
    tj = json.load(f)
  tj["loss_difference"] = loss_args.get("kind", "SMAPE")
  tj["use_multiscale_loss"] = loss_args.get("use_multiscale_loss", True)
  groups = (("features_training_settings", "feature"), ("combined_features_training_settings", "combined_feature"),
            ("combined_image_training_settings", "combined_image"))
  defaults = {"feature_weight": 1.0, "combined_feature_weight": 5.0, "combined_image_weight": 10.0}
  for section, prefix in groups:
    lw = tj[section]["loss_weights"]
    lw["mean"] = loss_args.get(prefix + "_weight", defaults[prefix + "_weight"])
    lw["variation"] = loss_args.get(prefix + "_variation_weight", 0.0)
    lw["ms_ssim"] = 0.0
    if "loss_weights_masked" in tj[section]:
      tj[section]["loss_weights_masked"].update(mean=loss_args.get(prefix + "_masked_weight", 0.0), variation=0.0, ms_ssim=0.0)
  return tj


def main_training(tf, mods, check):
  spec = importlib.util.spec_from_file_location("make_training_golden", os.path.join(HERE, "make_training_golden.py"))
  mtg = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mtg)
  j, arch, weights, features, targets = mtg.problem()
  loss, grads, metrics = run_reference_training(tf, mods, j, training_json_for(mtg.LOSS_ARGS), weights, features, targets)
  payload = {"loss": np.array(loss), "metrics": np.array("\n".join(metrics))}
  for k, g in grads.items():
    payload["grad|" + k] = g.astype(np.float64)
  _write_training(os.path.join(HERE, "refshim_training_example.npz"), payload, check)
  # COMBINED tuples: the reference's loader also feeds constant targets for the generated members (Training.py:538-549)
  j, arch, weights, features, targets = mtg.problem_combined()
  full_targets = dict(targets)
  n, h, w = next(iter(targets.values())).shape[:3]
  for tup in arch.feature_prediction_tuples:
    for index, fp in enumerate(tup.feature_predictions):
      if not fp.load_data:
        full_targets["target_image/" + fp.name] = np.full((n, h, w, fp.number_of_channels), 1.0 if index == 0 else 0.5, np.float32)
  loss, grads, metrics = run_reference_training(tf, mods, j, training_json_for(mtg.COMBINED_LOSS_ARGS), weights, features,
                                                full_targets)
  payload = {"loss": np.array(loss), "metrics": np.array("\n".join(metrics))}
  for k, g in grads.items():
    payload["grad|" + k] = g.astype(np.float32)
  _write_training(os.path.join(HERE, "refshim_training_combined.npz"), payload, check)


def _write_training(path, payload, check):
  if check:
    z = np.load(path)
    assert float(z["loss"]) == float(payload["loss"])
    print(os.path.basename(path), "matches the committed fixture")
  else:
    np.savez_compressed(path, **payload)
    print("%s: loss %.9f, %d gradient tensors" % (os.path.basename(path), float(payload["loss"]),
                                                  sum(k.startswith("grad|") for k in payload)))


def det(shape, seed, scale=1.0):
  """Deterministic pseudo-random float64 values in (-scale, scale) from a closed form - the component inputs are NOT stored in
  the fixture, the test regenerates them (an ulp of difference in sin() moves an input by 1e-12, far below the tolerance)."""
  n = int(np.prod(shape))
  v = np.sin(np.arange(n, dtype=np.float64) * 12.9898 + seed * 78.233) * 43758.5453
  return ((v - np.floor(v)) * 2.0 - 1.0).reshape(shape) * scale


COMPOSE_SHAPES = [(1, 1, 6, 24), (3, 3, 24, 24), (3, 3, 24, 24), (3, 3, 24, 24), (3, 3, 24, 24), (1, 1, 24, 1)]


def component_inputs():
  """{key: numpy float64} inputs of reference_components()."""
  inp = {}
  for k in (3, 5, 7, 21):          # symmetric tf.pad needs pad <= size (K 21: pad 10)
    inp["kp%d|src" % k], inp["kp%d|logits" % k] = det((2, 12, 13, 3), k), det((2, 12, 13, k * k), 100 + k, 3.0)
  inp["var|x"] = det((2, 8, 10, 3), 7)
  inp["loss|p"], inp["loss|t"] = det((2, 6, 7, 3), 8, 2.5), det((2, 6, 7, 3), 9)
  inp["util|x"] = det((64,), 10, 4.0)
  for i, shp in enumerate(COMPOSE_SHAPES):
    name = "conv2d" if i == 0 else "conv2d_%d" % i
    inp["compose|" + name + "/kernel"], inp["compose|" + name + "/bias"] = det(shp, 20 + i, 0.4), det((shp[3],), 40 + i, 0.2)
  inp["compose|small"], inp["compose|large"] = det((2, 4, 6, 3), 60), det((2, 8, 12, 3), 61)
  return inp


def reference_components(tf, mods, dtype=torch.float64):
  """Direct calls of the reference's building blocks at sizes the end-to-end cases do not reach (K = 7 / 21 kernel prediction,
  every LossDifference kind, every variance variant, compose_scales on its own): {key: numpy array} of OUTPUTS."""
  inp = {k: torch.as_tensor(v, dtype=dtype) for k, v in component_inputs().items()}
  out = {}
  # KernelPrediction.kernel_prediction (KernelPrediction.py:11-63)
  kp = mods["KernelPrediction"].KernelPrediction.kernel_prediction
  for k in (3, 5, 7, 21):
    src, logits = inp["kp%d|src" % k], inp["kp%d|logits" % k]
    last = kp(src, logits, k, data_format="channels_last")
    first = kp(src.permute(0, 3, 1, 2), logits.permute(0, 3, 1, 2), k, data_format="channels_first").permute(0, 2, 3, 1)
    assert torch.allclose(last, first, atol=1e-13)
    out["kp%d|out" % k] = last.numpy()
  # FeatureEngineering.variance (FeatureEngineering.py:57-70), all eight parametrisations
  for mode in ("uniform", "neighbor"):
    for rel in (False, True):
      for one in (False, True):
        v = mods["FeatureEngineering"].FeatureEngineering.variance(inp["var|x"], variance_mode=mode, relative_variance=rel,
                                                                   compress_to_one_channel=one, data_format="channels_last")
        out["var|%s|%d|%d" % (mode, rel, one)] = v.numpy()
  # LossDifference.difference (LossDifference.py:15-36), every kind
  enum = mods["LossDifference"].LossDifferenceEnum
  for kind in ("DIFFERENCE", "ABSOLUTE", "SMOOTH_ABSOLUTE", "SQUARED", "SMAPE"):
    out["loss|" + kind] = mods["LossDifference"].LossDifference.difference(inp["loss|p"], inp["loss|t"], enum[kind]).numpy()
  # Utilities.signed_log1p / signed_expm1 (Utilities.py:3-7)
  out["util|log1p"] = mods["Utilities"].signed_log1p(inp["util|x"]).numpy()
  out["util|expm1"] = mods["Utilities"].signed_expm1(inp["util|x"]).numpy()
  # MultiScalePrediction.scale_down / scale_up / compose_scales (MultiScalePrediction.py:11-93)
  tf.reset({k.split("|", 1)[1]: v for k, v in inp.items() if k.startswith("compose|conv2d")})
  msp = mods["MultiScalePrediction"].MultiScalePrediction
  out["compose|out"] = msp.compose_scales(inp["compose|small"], inp["compose|large"], data_format="channels_last").numpy()
  out["compose|down4"] = msp.scale_down(inp["compose|large"], heigh_width_scale_factor=4, data_format="channels_last").numpy()
  out["compose|up"] = msp.scale_up(inp["compose|small"], data_format="channels_last").numpy()
  return out


def prediction_problem():
  """A small all-passes network (ArchitectureExample.json with [8, 8] filters, one convolution per block, K = 3, without the
  'Screen Space Normal' auxiliary: Prediction.py matches files by SUBSTRING, 'Normal' would also match that file) and a
  40 x 72 frame cut into 32-pixel tiles with 4 pixels of overlap: 2 x 3 tiles, first / interior / last in the width."""
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.dirname(HERE))
  import cases
  from deepdenoiser_b200 import synthetic
  from deepdenoiser_b200.Architecture import Architecture
  j = cases._small(synthetic.example_architecture_json(), [8, 8], 1, 3)
  j["auxiliary_features"] = {k: v for k, v in j["auxiliary_features"].items() if k != "Screen Space Normal"}
  arch = Architecture(j, seed=777)
  weights = synthetic.randomize_biases(arch.weights)
  arch.weights = weights
  h, w = 40, 72
  features = synthetic.synthetic_features(arch, 1, h, w, seed=99)
  frame = {}
  for fp in arch.required_features():
    if fp.load_data:
      img = features["source_image/0/" + fp.name][0]
      frame[fp.name] = np.repeat(img, 3, axis=2) if img.shape[2] == 1 else img        # an EXR always decodes to 3 channels
  return j, arch, weights, frame, h, w, 32, 4


def reference_prediction(tf):
  """Prediction.main() of the reference (Prediction.py:188-519) on a scratch directory: returns {file stem: array} of the
  .npy files it saved (every pass + 'Combined')."""
  import argparse
  import json
  import tempfile
  j, arch, weights, frame, h, w, tile, overlap = prediction_problem()
  for name in ("Prediction", "OpenEXRDirectory", "cv2"):
    sys.modules.pop(name, None)
  sys.path.insert(0, REFERENCE)
  sys.path.insert(0, SHIM)
  try:
    prediction = importlib.import_module("Prediction")
    assert os.path.abspath(prediction.__file__).startswith(os.path.abspath(REFERENCE)), prediction.__file__
  finally:
    sys.path.remove(SHIM)
    sys.path.remove(REFERENCE)
  tf.reset({k: torch.as_tensor(np.asarray(v), dtype=torch.float64) for k, v in weights.items()})
  tf.ESTIMATOR_MODE = "predict"
  cwd = os.getcwd()
  out = {}
  try:
    with tempfile.TemporaryDirectory() as scratch:
      frame_dir = os.path.join(scratch, "frame")
      os.makedirs(frame_dir)
      for name, img in frame.items():
        with open(os.path.join(frame_dir, name + ".exr"), "wb") as f:
          np.save(f, np.ascontiguousarray(img[..., ::-1]).astype(np.float32))           # BGR, as OpenCV decodes
      with open(os.path.join(scratch, "architecture.json"), "w") as f:
        json.dump(j, f)
      os.chdir(scratch)                                                                 # main() writes ./tmp.tfrecords
      prediction.main(argparse.Namespace(json_filename=os.path.join(scratch, "architecture.json"), input=frame_dir, tile_size=tile,
                                         tile_overlap_size=overlap, threads=1, data_format="channels_last"))
      for file in sorted(os.listdir(frame_dir)):
        if file.endswith(".npy"):
          out[file[:-4]] = np.load(os.path.join(frame_dir, file))
  finally:
    os.chdir(cwd)
    tf.ESTIMATOR_MODE = "capture"
  return out


AUGMENT_VECTORS = ((0.0, 0.0, 0.0), (0.25, 0.5, 0.75), (0.9, 0.1, 0.3))
# (tag, [(pass, channels, is target)], flip, rot90 k, RGB permutation, rotation vector); the reference refuses to flip 'Normal'
AUGMENT_PIPELINES = (
    ("rotate", [("Diffuse Color", 3, True), ("Normal", 3, False), ("Screen Space Normal", 3, False), ("Depth", 1, False),
                ("Alpha", 1, True)], None, 3, 4, (0.25, 0.5, 0.75)),
    ("flip", [("Glossy Direct", 3, True), ("Screen Space Normal", 3, False), ("Depth", 1, False)], 1, 1, 2, None))


def reference_augmentation(tf):
  """DataAugmentation.py on one [h, w, 3] example (the reference augments per example, Training.py:803-815): outputs only."""
  sys.path.insert(0, REFERENCE)
  sys.path.insert(0, SHIM)
  try:
    da = importlib.import_module("DataAugmentation").DataAugmentation
  finally:
    sys.path.remove(SHIM)
    sys.path.remove(REFERENCE)
  image = torch.as_tensor(det((5, 7, 3), 77), dtype=torch.float64)
  out = {}
  for name in ("Diffuse Color", "Screen Space Normal"):
    tag = name.replace(" ", "")
    for flip in (0, 1):
      out["flip|%s|%d" % (tag, flip)] = da.flip_left_right(image, name, flip).numpy()
    for k in range(4):
      out["rot|%s|%d" % (tag, k)] = da.rotate_90(image, k, name).numpy()
      first = da.rotate_90(image.permute(2, 0, 1), k, name, data_format="channels_first").permute(1, 2, 0)
      assert torch.equal(first, torch.as_tensor(out["rot|%s|%d" % (tag, k)]))
  for permute in range(6):
    out["perm|%d" % permute] = da.permute_rgb(image, permute).numpy()
  for i, vec in enumerate(AUGMENT_VECTORS):
    r = da.random_rotation_matrix([torch.tensor(v, dtype=torch.float64) for v in vec])
    out["matrix|%d" % i] = r.numpy()
    out["normal|%d" % i] = da.rotate_normal(image, r).numpy()
  # the whole per-example pipeline in the reference's order (FeatureTrainingAugmentation, Training.py:551-605, driven as
  # input_fn_tfrecords does, :803-819): flip -> rot90 -> RGB permutation -> normal rotation
  sys.path.insert(0, REFERENCE)
  sys.path.insert(0, SHIM)
  try:
    training = importlib.import_module("Training")
  finally:
    sys.path.remove(SHIM)
    sys.path.remove(REFERENCE)
  for tag, passes, flip, rot, perm, vec in AUGMENT_PIPELINES:
    sources = {"source_image/0/" + n: torch.as_tensor(det((6, 6, c), 300 + i), dtype=torch.float64) for i, (n, c, _) in enumerate(passes)}
    targets = {"target_image/" + n: torch.as_tensor(det((6, 6, c), 400 + i), dtype=torch.float64) for i, (n, c, t) in enumerate(passes) if t}
    matrix = da.random_rotation_matrix([torch.tensor(v, dtype=torch.float64) for v in vec]) if vec is not None else None
    for n, c, is_target in passes:
      fta = training.FeatureTrainingAugmentation(1, is_target, c, n)
      fta.intialize_from_dictionaries(sources, targets)
      if flip is not None:
        fta.flip_left_right(flip, "channels_last")
      fta.rotate_90(rot, "channels_last")
      fta.permute_rgb(perm, "channels_last")
      if matrix is not None:
        fta.rotate_normal(matrix, "channels_last")
      fta.add_to_sources_dictionary(sources)
      fta.add_to_targets_dictionary(targets)
    for k, v in list(sources.items()) + list(targets.items()):
      out["pipeline|%s|%s" % (tag, k)] = v.numpy()
  return out


def main():
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.dirname(HERE))
  import cases
  from golden.make_golden import digest
  check = "--check" in sys.argv
  tf, mods = load_reference()
  for name in cases.GOLDEN_CASES + cases.BASELINE_CASES:
    j, arch, weights, features = cases.build(name)
    last, names_last = run_reference(tf, mods, j, weights, features, "channels_last")
    first, names_first = run_reference(tf, mods, j, weights, features, "channels_first")
    assert names_last == names_first
    assert sorted(names_last) == sorted(weights), "the reference requested a different variable set than the case provides"
    worst = 0.0
    for a, b in zip(last, first):
      assert set(a) == set(b)
      for k in a:
        worst = max(worst, float(np.abs(a[k] - b[k]).max() / max(1.0, np.abs(a[k]).max())))
    assert worst < 1e-12, "channels_last and channels_first runs of the reference disagree: %g" % worst
    payload = {"inputs_sha256": np.array(digest(features)), "weights_sha256": np.array(digest(weights)),
               "variables_in_creation_order": np.array("\n".join(names_last))}
    baseline = name in cases.BASELINE_CASES       # the BASELINE.json architectures: float32 storage, and a subset of the 17 passes
    keep = None if name != "tiramisu32_small" else STORED_PASSES
    for s, d in enumerate(last):
      for k, v in d.items():
        if keep is None or k in keep:
          payload["%d|%s" % (s, k)] = v.astype(np.float32 if baseline else np.float64)
    path = os.path.join(HERE, "refshim_" + name + ".npz")
    if check:
      z = np.load(path)
      for k in payload:
        if "|" in k:
          assert np.array_equal(z[k], payload[k]), (name, k)
      print(name, "matches the committed fixture")
    else:
      np.savez_compressed(path, **payload)
      print(name, len(last), "scales", len(last[0]), "passes, NHWC vs NCHW max rel diff %.1e" % worst, "->", os.path.basename(path))
  main_training(tf, mods, check)
  pred = reference_prediction(tf)
  path = os.path.join(HERE, "refshim_prediction.npz")
  if check:
    z = np.load(path)
    for k in pred:
      assert np.array_equal(z[k], pred[k].astype(np.float32)), k
    print("tiled prediction matches the committed fixture")
  else:
    np.savez_compressed(path, **{k: v.astype(np.float32) for k, v in pred.items()})
    print("tiled prediction:", len(pred), "images", next(iter(pred.values())).shape, "->", os.path.basename(path))
  comp = reference_components(tf, mods)
  comp.update({"augment|" + k: v for k, v in reference_augmentation(tf).items()})
  path = os.path.join(HERE, "refshim_components.npz")
  if check:
    z = np.load(path)
    for k in comp:
      assert np.array_equal(z[k], comp[k]), k
    print("components match the committed fixture")
  else:
    np.savez_compressed(path, **comp)
    print("components:", len(comp), "arrays ->", os.path.basename(path))


if __name__ == "__main__":
  main()
