"""End-to-end parity of Architecture.predict on the GPU (through the C ABI) against the oracle and the
committed golden vectors.

Tolerances (per-pixel L1, i.e. |out - oracle| per element, relative to max(1, |oracle|_max) of the pass):
  float32 mode   : <= 1e-4 - the north-star bound; the exact (SIMT) path accumulates in fp32 like the reference
  float16x2 mode : <= 1e-4 - the same bound on TENSOR CORES (fp16 hi + lo pairs, three tcgen05 passes per layer)
  float16 mode   : 2x the measured error of each case (FP16_BOUNDS) - fp16 storage of ~22 stacked conv layers (fp32
                   accumulate, fp32 logits / softmax / filter apply); see DESIGN.md "precision"
"""
import os

import numpy as np
import pytest
import torch

import cases
from deepdenoiser_b200 import synthetic
from deepdenoiser_b200.Architecture import Architecture, ModeKeys
from oracle import np_ops, reference_model

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_case(name, dtype):
  j, host_arch, weights, features = cases.build(name)
  j = dict(j)
  j["b200"] = {"dtype": dtype}
  arch = Architecture(j, weights=weights)
  out = arch.predict({k: torch.from_numpy(v) for k, v in features.items()}, ModeKeys.PREDICT)
  torch.cuda.synchronize()
  oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
  return arch, out, oracle


def errors(out, oracle):
  worst_max, worst_mean = 0.0, 0.0
  assert len(out) == len(oracle)
  for s in range(len(oracle)):
    assert set(out[s]) == set(oracle[s])
    for k, want in oracle[s].items():
      got = out[s][k].float().cpu().numpy().astype(np.float64)
      assert got.shape == want.shape, (k, got.shape, want.shape)
      assert np.isfinite(got).all(), k
      scale = max(1.0, float(np.abs(want).max()))
      worst_max = max(worst_max, float(np.abs(got - want).max()) / scale)
      worst_mean = max(worst_mean, float(np.abs(got - want).mean()) / scale)
  return worst_max, worst_mean


@pytest.mark.parametrize("name", cases.GOLDEN_CASES + ("rgb9",))
def test_predict_float32_matches_oracle_1e4(name):
  arch, out, oracle = run_case(name, "float32")
  mx, mean = errors(out, oracle)
  print(name, "fp32 max %.2e mean %.2e" % (mx, mean))
  assert mx <= 1e-4


# fp16 storage / fp32 accumulate against the float64 oracle.  Bounds = 2x what was measured on B200 in round 2 (max, mean of
# |out - oracle| / max(1, |oracle|max)): example 4.2e-5 / 4.0e-6, combined_onehot 2.2e-3 / 7.4e-5, tiramisu 3.7e-3 / 7.3e-5,
# variants 2.4e-3 / 2.6e-5, rgb9 1.2e-5 / 7.9e-7.  (The small random-weight nets amplify a logit perturbation far more than
# the benchmark network does: 3.1e-4 / 1.6e-6 on the full 1080p frame, tests/test_gpu_bench_shapes.py and bench.py.)
FP16_BOUNDS = {"example": (1e-4, 1e-5), "combined_onehot": (5e-3, 1.5e-4), "tiramisu": (8e-3, 1.5e-4),
               "variants": (5e-3, 6e-5), "rgb9": (3e-5, 2e-6)}


@pytest.mark.parametrize("name", ["example", "combined_onehot", "tiramisu", "variants", "rgb9"])
def test_predict_float16_tensor_core_path(name):
  arch, out, oracle = run_case(name, "float16")
  mx, mean = errors(out, oracle)
  print(name, "fp16 max %.2e mean %.2e" % (mx, mean))
  assert mx <= FP16_BOUNDS[name][0] and mean <= FP16_BOUNDS[name][1]


@pytest.mark.parametrize("name", ["example", "combined_onehot", "variants", "rgb9"])
def test_predict_float16x2_tensor_core_path_meets_1e4(name):
  """The high-accuracy TENSOR-CORE mode (fp16 hi + lo pairs, three tcgen05 passes per layer, fp32 accumulate) meets the
  reference's 1e-4 bound on every U-Net case - the same bound the exact SIMT path is held to."""
  arch, out, oracle = run_case(name, "float16x2")
  mx, mean = errors(out, oracle)
  print(name, "fp16x2 max %.2e mean %.2e" % (mx, mean))
  assert mx <= 1e-4


@pytest.mark.parametrize("name", cases.GOLDEN_CASES)
def test_predict_float32_matches_committed_golden(name):
  arch, out, _ = run_case(name, "float32")
  z = np.load(os.path.join(GOLDEN, name + ".npz"))
  checked = 0
  for key in z.files:
    if "|" not in key:
      continue
    s, k = key.split("|", 1)
    want = z[key].astype(np.float64)
    got = out[int(s)][k].float().cpu().numpy()
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), key
    checked += 1
  assert checked > 0


@pytest.mark.parametrize("name", cases.GOLDEN_CASES + cases.BASELINE_CASES)
@pytest.mark.parametrize("dtype,bound", [("float32", 1e-4), ("float16x2", 1e-4)])
def test_predict_matches_the_reference_code_fixtures(name, dtype, bound):
  """The CUDA path against tests/golden/refshim_<case>.npz: outputs of the reference's OWN Python modules (float64) executed
  over oracle/tf_shim (tests/golden/make_reference_golden.py) - the exact SIMT path and the high-accuracy tensor-core mode,
  north-star tolerance."""
  if dtype == "float16x2" and name not in ("example", "combined_onehot", "variants", "rgb9", "unet32_small"):
    pytest.skip("float16x2 is built for the U-Net with kernel prediction")
  arch, out, _ = run_case(name, dtype)
  z = np.load(os.path.join(GOLDEN, "refshim_" + name + ".npz"))
  checked = 0
  for key in z.files:
    if "|" not in key:
      continue
    s, k = key.split("|", 1)
    want = z[key]
    got = out[int(s)][k].float().cpu().numpy()
    assert got.shape == want.shape, key
    assert np.abs(got - want).max() <= bound * max(1.0, np.abs(want).max()), key
    checked += 1
  assert checked > 0


@pytest.mark.parametrize("trial", range(16))
def test_random_architectures_exact_path_matches_the_oracle(trial):
  """cases.random_case: seeded random points of the architecture JSON space (the same sweep tests/test_reference_golden.py runs
  through the reference's own code): the exact CUDA path against the float64 oracle, north-star tolerance."""
  j, host_arch, weights, features = cases.random_case(trial)
  j = dict(j)
  j["b200"] = {"dtype": "float32"}
  arch = Architecture(j, weights=weights)
  out = arch.predict({k: torch.from_numpy(v) for k, v in features.items()}, ModeKeys.PREDICT)
  torch.cuda.synchronize()
  oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
  if max(float(np.abs(v).max()) for d in oracle for v in d.values()) > 1e30:
    pytest.skip("signed_expm1 of a raw-source prediction leaves the fp32 range (the reference's float32 graph overflows too)")
  mx, mean = errors(out, oracle)
  assert mx <= 1e-4, (trial, mx)


@pytest.mark.parametrize("trial", range(0, 24, 2))
@pytest.mark.parametrize("dtype,bound", [("float16", 3e-2), ("bfloat16", 2.5e-1)])
def test_random_architectures_run_on_the_tensor_core_path(trial, dtype, bound):
  """The same sweep through the 16-bit tensor-core modes: every configuration must run (odd channel counts, 2- to 4-level
  backbones, K 3 / 5 / 7 fused or stand-alone heads, with and without kernel prediction / multi-scale) and stay finite; the
  bound only catches structural errors (a wrong channel mapping is O(1)) - the tiny random-weight nets push 16-bit rounding
  through signed_expm1 (measured: float16 <= 1.6e-2, bfloat16 <= 1.9e-1 where the outputs stay below 1e4)."""
  j, host_arch, weights, features = cases.random_case(trial)
  j = dict(j)
  j["b200"] = {"dtype": dtype}
  oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
  arch = Architecture(j, weights=weights)
  out = arch.predict({k: torch.from_numpy(v) for k, v in features.items()}, ModeKeys.PREDICT)
  torch.cuda.synchronize()
  if max(float(np.abs(v).max()) for d in oracle for v in d.values()) > 1e4:
    return                                  # ran; exp-amplified outputs are not a meaningful yardstick for 16-bit rounding
  mx, mean = errors(out, oracle)
  assert mx <= bound, (trial, mx)


def test_tuple_chunking_does_not_change_results():
  j, host_arch, weights, features = cases.build("example")
  outs = []
  for chunk in (1 << 30, 16 * 16 * 3):            # all 17 tuples at once vs 3 tuples per chunk
    jj = dict(j)
    jj["b200"] = {"dtype": "float32", "max_chunk_pixels": chunk}
    arch = Architecture(jj, weights=weights)
    outs.append(arch.predict({k: torch.from_numpy(v) for k, v in features.items()}))
  for s in range(3):
    for k in outs[0][s]:
      assert torch.equal(outs[0][s][k], outs[1][s][k]), (s, k)


def test_identity_network_reproduces_source():
  """Known answer through the whole path: zero weights => equal logits => box-filtered standardised source at every
  scale; composition of box filters of pooled images; checked against the oracle at tight tolerance."""
  j, host_arch, weights, features = cases.build("example")
  zero = {k: np.zeros_like(v) for k, v in weights.items()}
  jj = dict(j)
  jj["b200"] = {"dtype": "float16"}               # with zero weights the fp16 path is exact too
  arch = Architecture(jj, weights=zero)
  out = arch.predict({k: torch.from_numpy(v) for k, v in features.items()})
  oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=zero).predict_numpy(features)
  mx, _ = errors(out, oracle)
  assert mx <= 2e-5


def test_larger_image_and_batch_float16():
  """Tile shapes of the real workloads (width > 128 => several strips, rows % 4 != 0, batch > 1)."""
  j = synthetic.baseline_architecture_json("unet32")
  host = Architecture(j)
  weights = synthetic.randomize_biases(host.weights)
  features = synthetic.synthetic_features(host, 2, 36, 264, seed=5)
  jj = dict(j)
  jj["b200"] = {"dtype": "float16"}
  out = Architecture(jj, weights=weights).predict({k: torch.from_numpy(v) for k, v in features.items()})
  oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights).predict_numpy(features)
  mx, mean = errors(out, oracle)
  print("unet32 36x264 fp16 max %.2e mean %.2e" % (mx, mean))
  assert mx <= 3e-4 and mean <= 6e-6          # measured 1.24e-4 / 2.6e-6


@pytest.mark.parametrize("name", ["example", "combined_onehot"])
def test_fused_output_head_matches_layer_by_layer_path(name):
  """The fused 1x1 + 1x1 + kernel-prediction kernel against the unfused launches on the same fp16 network: the only
  difference is fp32 vs fp32 logits accumulated in another order (<= 2e-3 of the output scale)."""
  j, host_arch, weights, features = cases.build(name)
  jj = dict(j)
  jj["b200"] = {"dtype": "float16"}
  feats = {k: torch.from_numpy(v) for k, v in features.items()}
  fused = Architecture(jj, weights=weights)
  fused._ensure_device()
  assert fused.network.can_fuse_post_kp(fused.kernel_size, fused.features_per_tuple)
  a = fused.predict(feats)
  plain = Architecture(jj, weights=weights)
  plain._ensure_device()
  plain.network.fused_post_kp = False
  b = plain.predict(feats)
  for s in range(len(a)):
    for k in a[s]:
      x, y = a[s][k].float().cpu().numpy(), b[s][k].float().cpu().numpy()
      assert np.abs(x - y).max() <= 2e-3 * max(1.0, np.abs(y).max()), (s, k)
