"""Seeded synthetic render-pass stacks and the architecture JSONs of the BASELINE configurations
(SURVEY.md section 8(d)).  No dataset or checkpoint ships with the reference, so every test, the bench and
smoke() draw their inputs from here.  Pure numpy; used identically by the CUDA path and by the oracle."""
import copy

import numpy as np

_HANDLING = {
    "feature_variance": {"use_variance": True, "variance_mode": "uniform", "relative_variance": True,
                         "compute_before_standardization": False, "compress_to_one_channel": True},
    "standardization": {"use_log1p": True, "mean": 0.0, "variance": 1.0},
    "invert_standardization": True,
}


def _auxiliary(channels=3, use_log1p=False):
  return {"number_of_channels": channels,
          "feature_variance": dict(_HANDLING["feature_variance"]),
          "standardization": {"use_log1p": use_log1p, "mean": 0.0, "variance": 1.0}}


def example_architecture_json():
  """Same content as the reference's ArchitectureExample.json (without the *_description strings)."""
  lights = {}
  for name in ("Diffuse", "Glossy", "Subsurface", "Transmission"):
    lights[name] = {"Color": name + " Color", "Direct": name + " Direct", "Indirect": name + " Indirect"}
  lights["Volume"] = {"Color": "", "Direct": "Volume Direct", "Indirect": "Volume Indirect"}
  for name in ("Emission", "Environment", "Alpha"):
    lights[name] = {"Color": name, "Direct": "", "Indirect": ""}
  return {
      "model_directory": "../Models/Example",
      "number_of_sources_per_target": 1,
      "architecture": {
          "source_encoder": {"feature_prediction_tuple_type": "SINGLE", "feature_flag_mode": "EMBEDDING"},
          "core_architecture": {"name": "U-Net", "number_of_filters_for_convolution_blocks": [64, 96, 128],
                                "number_of_convolutions_per_block": 4, "use_batch_normalization": False,
                                "dropout_rate": 0.0},
          "kernel_prediction": {"use_kernel_prediction": True, "kernel_size": 5,
                                "use_standardized_source_for_kernel_prediction": True},
          "multiscale_prediction": {"use_multiscale_predictions": True,
                                    "invert_standardization_after_multiscale_predictions": True},
      },
      "combined_features": lights,
      "combined_features_handling": {k: copy.deepcopy(_HANDLING) for k in ("Color", "Direct", "Indirect")},
      "auxiliary_features": {"Normal": _auxiliary(3)},
  }


def baseline_architecture_json(name="unet32"):
  """Architecture JSONs of the BASELINE.json configurations.

    unet32      cfg2/cfg4/cfg5: U-Net [64,96,128]x4, K=5, SINGLE tuples (17 passes), EMBEDDING flags (8 dims) and five
                auxiliaries -> (1 + 5) * 4 + 8 = 32 input channels
    tiramisu32  cfg3: Tiramisu [64,96,128]x4, K=21, otherwise as unet32
    rgb9        cfg1: RGB + Normal + albedo, 9 source channels, no variance / flags, one SINGLE tuple
  """
  j = example_architecture_json()
  if name in ("unet32", "tiramisu32"):
    j["auxiliary_features"] = {"Normal": _auxiliary(3), "Depth": _auxiliary(1, use_log1p=True),
                               "Shadow": _auxiliary(3), "Ambient Occlusion": _auxiliary(3),
                               "Screen Space Normal": _auxiliary(3)}
    if name == "tiramisu32":
      j["architecture"]["core_architecture"]["name"] = "Tiramisu"
      j["architecture"]["kernel_prediction"]["kernel_size"] = 21
    j["model_directory"] = "../Models/" + name
    return j
  if name == "rgb9":
    # one SINGLE tuple (the noisy RGB) + Normal + albedo as auxiliaries = 3 x 3 channels, nothing else
    j["combined_features"] = {"Diffuse": {"Color": "", "Direct": "Diffuse Direct", "Indirect": ""}}
    j["architecture"]["source_encoder"] = {"feature_prediction_tuple_type": "SINGLE", "feature_flag_mode": "NONE"}
    for kind in ("Color", "Direct", "Indirect"):
      j["combined_features_handling"][kind]["feature_variance"]["use_variance"] = False
    aux = {}
    for aux_name in ("Normal", "Diffuse Color"):
      aux[aux_name] = _auxiliary(3)
      aux[aux_name]["feature_variance"]["use_variance"] = False
    j["auxiliary_features"] = aux
    j["model_directory"] = "../Models/rgb9"
    return j
  raise ValueError(name)


def _smooth_field(rng, n, h, w, c, waves=4):
  """Sum of low-frequency sinusoids in [0, ~1]."""
  yy, xx = np.meshgrid(np.arange(h, dtype=np.float64) / max(h, 1), np.arange(w, dtype=np.float64) / max(w, 1),
                       indexing="ij")
  out = np.zeros((n, h, w, c))
  for _ in range(waves):
    fy, fx = rng.uniform(0.5, 4.0, size=2)
    phase = rng.uniform(0, 2 * np.pi, size=(n, 1, 1, c))
    amp = rng.uniform(0.2, 1.0, size=(n, 1, 1, c))
    out += amp * (0.5 + 0.5 * np.sin(2 * np.pi * (fy * yy + fx * xx))[None, :, :, None] * np.cos(phase) +
                  0.0 * phase)
  return out / waves


def render_pass(name, channels, n, h, w, rng, samples_per_pixel=16):
  """One synthetic pass [n,h,w,channels] float32: colour passes uniform, light passes = smooth
  HDR field x Gamma(spp) Monte-Carlo noise with 1% fireflies and ~20% exactly-zero blocks, normals unit length."""
  if name.endswith("Normal"):
    v = _smooth_field(rng, n, h, w, 3) - 0.5 + 0.05 * rng.standard_normal((n, h, w, 3))
    v /= np.maximum(np.linalg.norm(v, axis=-1, keepdims=True), 1e-6)
    return v.astype(np.float32)[..., :channels]
  if name.endswith(" Color") or name in ("Alpha",):
    return rng.uniform(0.0, 1.0, size=(n, h, w, channels)).astype(np.float32)
  if name in ("Depth", "Shadow", "Ambient Occlusion", "Mist"):
    scale = 20.0 if name == "Depth" else 1.0
    return (scale * _smooth_field(rng, n, h, w, channels)).astype(np.float32)
  clean = _smooth_field(rng, n, h, w, channels) * np.exp(rng.normal(0.0, 0.5, size=(n, 1, 1, channels)))
  noisy = clean * rng.gamma(samples_per_pixel, 1.0 / samples_per_pixel, size=clean.shape)
  fireflies = rng.uniform(size=(n, h, w, 1)) < 0.01
  noisy = np.where(fireflies, noisy * 50.0, noisy)
  by, bx = max(h // 8, 1), max(w // 8, 1)
  block_mask = rng.uniform(size=(n, -(-h // by), -(-w // bx), 1)) < 0.2
  mask = np.repeat(np.repeat(block_mask, by, axis=1), bx, axis=2)[:, :h, :w]
  return np.where(mask, 0.0, noisy).astype(np.float32)


def synthetic_features(architecture, n, h, w, seed=1234):
  """Source dictionary for `architecture` (anything exposing feature_predictions / auxiliary_features with
  .name, .number_of_channels, .load_data and a Color/Direct/Indirect kind): numpy float32 [n,h,w,C] per key
  'source_image/0/<Pass>'.  Non-loaded passes get the constants the reference feeds (1.0 Color, 0.5 else)."""
  rng = np.random.default_rng(seed)
  features = {}
  for fp in list(architecture.feature_predictions) + list(architecture.auxiliary_features):
    key = "source_image/0/" + fp.name
    if fp.load_data:
      features[key] = render_pass(fp.name, fp.number_of_channels, n, h, w, rng)
    else:
      kind = getattr(fp, "kind", None) or fp.feature_prediction_type.name.capitalize()
      value = 1.0 if kind == "Color" else 0.5
      features[key] = np.full((n, h, w, fp.number_of_channels), value, dtype=np.float32)
  return features


def randomize_biases(weights, seed=99, scale=0.05):
  """tf.layers initialises biases to zero; tests perturb them so the bias path is exercised."""
  rng = np.random.default_rng(seed)
  out = dict(weights)
  for k, v in weights.items():
    if k.endswith("/bias"):
      out[k] = (scale * rng.standard_normal(v.shape)).astype(np.float32)
  return out
