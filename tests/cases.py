"""Small architecture configurations shared by the golden-vector generator, the CPU oracle tests and the GPU
parity tests.  Each case: (architecture JSON, batch, height, width)."""
import copy

from deepdenoiser_b200 import synthetic


def _small(j, filters, n, k):
  j = copy.deepcopy(j)
  core = j["architecture"]["core_architecture"]
  core["number_of_filters_for_convolution_blocks"] = filters
  core["number_of_convolutions_per_block"] = n
  j["architecture"]["kernel_prediction"]["kernel_size"] = k
  return j


def case(name):
  ex = synthetic.example_architecture_json()
  if name == "example":                       # ArchitectureExample.json as shipped: U-Net SINGLE EMBEDDING K5
    return ex, 1, 16, 16
  if name == "combined_onehot":               # COMBINED tuples (3 features / pass), one-hot flags, 2 scales
    j = _small(ex, [16, 24], 2, 3)
    j["architecture"]["source_encoder"] = {"feature_prediction_tuple_type": "COMBINED",
                                           "feature_flag_mode": "ONE_HOT_ENCODING"}
    return j, 2, 16, 24
  if name == "tiramisu":                      # Tiramisu dense blocks, 3x3 transposed convs, 2x2 max-pool
    j = _small(ex, [16, 24, 32], 2, 5)
    j["architecture"]["core_architecture"]["name"] = "Tiramisu"
    j["architecture"]["source_encoder"]["feature_flag_mode"] = "NONE"
    j["combined_features"] = {k: ex["combined_features"][k] for k in ("Diffuse", "Alpha")}
    return j, 1, 24, 16
  if name == "variants":                      # raw-source KP, invert before compose, neighbor / uncompressed variance,
    j = _small(ex, [16, 16, 24], 1, 3)        # non-trivial mean / variance standardisation
    j["architecture"]["kernel_prediction"]["use_standardized_source_for_kernel_prediction"] = False
    j["architecture"]["multiscale_prediction"]["invert_standardization_after_multiscale_predictions"] = False
    j["combined_features"] = {k: ex["combined_features"][k] for k in ("Glossy", "Volume", "Emission")}
    h = j["combined_features_handling"]
    h["Direct"]["feature_variance"].update(variance_mode="neighbor", relative_variance=False,
                                           compress_to_one_channel=False)
    h["Direct"]["standardization"].update(mean=0.25, variance=2.0)
    h["Indirect"]["feature_variance"].update(compute_before_standardization=True)
    h["Color"]["invert_standardization"] = False
    # keep every tuple at the same input width: Color / Indirect (3 + 1), Direct (3 + 3) differ -> use one width
    h["Color"]["feature_variance"].update(compress_to_one_channel=False)
    h["Indirect"]["feature_variance"].update(compress_to_one_channel=False)
    return j, 1, 16, 16
  if name == "direct":                        # no kernel prediction, no multi-scale: 3-channel direct prediction
    j = _small(ex, [16, 24], 2, 5)
    j["architecture"]["kernel_prediction"]["use_kernel_prediction"] = False
    j["architecture"]["multiscale_prediction"]["use_multiscale_predictions"] = False
    j["combined_features"] = {k: ex["combined_features"][k] for k in ("Diffuse", "Environment")}
    return j, 2, 8, 12
  if name == "rgb9":                          # BASELINE cfg1: 9 source channels, one 64x64 tile
    return synthetic.baseline_architecture_json("rgb9"), 1, 64, 64
  if name == "unet32_small":                  # BASELINE cfg2 / cfg4 / cfg5 architecture (the benchmarked network) on one small tile
    return synthetic.baseline_architecture_json("unet32"), 1, 16, 16
  if name == "tiramisu32_small":              # BASELINE cfg3 architecture: K = 21 needs >= 10 pixels at the coarsest scale
    return synthetic.baseline_architecture_json("tiramisu32"), 1, 40, 40
  raise KeyError(name)


BASELINE_CASES = ("rgb9", "unet32_small", "tiramisu32_small")     # reference-code fixtures only (tests/golden/refshim_*.npz)
GOLDEN_CASES = ("example", "combined_onehot", "tiramisu", "variants", "direct")


def build(name, seed=4321):
  """(json, product-side Architecture (host logic only), weights, features) for a case."""
  from deepdenoiser_b200.Architecture import Architecture
  j, n, h, w = case(name)
  arch = Architecture(j, seed=seed)
  weights = synthetic.randomize_biases(arch.weights)
  arch.weights = weights
  features = synthetic.synthetic_features(arch, n, h, w, seed=1234)
  if arch.feature_flag_mode.name == "ONE_HOT_ENCODING":
    import numpy as np
    names = sorted(t.name for t in arch.feature_prediction_tuples)
    for i, nm in enumerate(names):
      planes = np.zeros((n, h, w, len(names)), dtype=np.float32)
      planes[..., i] = 1.0
      features["feature_flag/" + nm] = planes
  return j, arch, weights, features


def random_case(trial):
  """Seeded random point of the architecture JSON space the reference accepts (backbone, filters, convolutions per block,
  K 3 / 5 / 7, SINGLE / COMBINED tuples, NONE / ONE_HOT / EMBEDDING flags, kernel prediction on / off with standardised or raw
  source, multi-scale on / off, inversion before / after, pass and auxiliary subsets, variance mode / relative / before /
  compressed, log1p / mean / variance) -> (json, product-side Architecture, weights, features)."""
  import random
  import numpy as np
  from deepdenoiser_b200.Architecture import Architecture
  rnd = random.Random(1000 + trial)
  ex = synthetic.example_architecture_json()
  j = _small(ex, rnd.choice([[8, 8], [8, 16, 8], [16, 8, 8, 8]]), rnd.choice([1, 2, 3]), rnd.choice([3, 5, 7]))
  a = j["architecture"]
  a["core_architecture"]["name"] = rnd.choice(["U-Net", "Tiramisu"])
  a["source_encoder"] = {"feature_prediction_tuple_type": rnd.choice(["SINGLE", "COMBINED"]),
                         "feature_flag_mode": rnd.choice(["NONE", "ONE_HOT_ENCODING", "EMBEDDING"])}
  a["kernel_prediction"]["use_kernel_prediction"] = rnd.random() < 0.8
  a["kernel_prediction"]["use_standardized_source_for_kernel_prediction"] = rnd.random() < 0.5
  a["multiscale_prediction"]["use_multiscale_predictions"] = rnd.random() < 0.8
  a["multiscale_prediction"]["invert_standardization_after_multiscale_predictions"] = rnd.random() < 0.5
  names = rnd.sample(sorted(ex["combined_features"]), rnd.choice([2, 3, 8]))
  j["combined_features"] = {k: ex["combined_features"][k] for k in names}
  aux = sorted(ex["auxiliary_features"])             # the reference needs at least one (Architecture.py:404 reads a leaked name)
  j["auxiliary_features"] = {k: copy.deepcopy(ex["auxiliary_features"][k]) for k in rnd.sample(aux, rnd.randint(1, len(aux)))}
  one = rnd.random() < 0.5 or "Alpha" in names       # every tuple must feed the same number of input channels
  for kind in ("Color", "Direct", "Indirect"):
    hnd = j["combined_features_handling"][kind]
    hnd["feature_variance"].update(use_variance=True, variance_mode=rnd.choice(["uniform", "neighbor"]),
                                   relative_variance=rnd.random() < 0.5, compute_before_standardization=rnd.random() < 0.5,
                                   compress_to_one_channel=one)
    hnd["standardization"].update(use_log1p=rnd.random() < 0.7, mean=rnd.choice([0.0, 0.3]), variance=rnd.choice([1.0, 1.7]))
    hnd["invert_standardization"] = rnd.random() < 0.8
  arch = Architecture(j, seed=100 + trial)
  weights = synthetic.randomize_biases(arch.weights)
  arch.weights = weights
  steps = len(a["core_architecture"]["number_of_filters_for_convolution_blocks"]) - 1
  h = w = (8 << steps) if a["kernel_prediction"]["kernel_size"] < 7 else (16 << steps)
  features = synthetic.synthetic_features(arch, 1, h, w, seed=trial)
  if arch.feature_flag_mode.name == "ONE_HOT_ENCODING":
    tuple_names = sorted(t.name for t in arch.feature_prediction_tuples)
    for i, nm in enumerate(tuple_names):
      planes = np.zeros((1, h, w, len(tuple_names)), dtype=np.float32)
      planes[..., i] = 1.0
      features["feature_flag/" + nm] = planes
  return j, arch, weights, features
