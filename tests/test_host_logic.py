"""Host-side logic of the product (no GPU): feature layout, tuples, variable lists, work counts, naming, and that
the C-ABI library loads and exports every symbol include/dd_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
from deepdenoiser_b200 import _lib, synthetic
from deepdenoiser_b200.Architecture import Architecture, FeaturePredictionTupleType, FeaturePredictionType
from deepdenoiser_b200.FeatureFlags import FeatureFlagMode
from deepdenoiser_b200.Naming import Naming
from deepdenoiser_b200.RenderPasses import RenderPasses, RenderPassesUsage
from oracle import reference_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, "include", "dd_b200.h")).read()
  declared = set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", header))
  declared -= {"dd_ctx"}
  assert len(declared) >= 20
  assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
  lib = _lib.load_library()          # raises if the .so is not built
  for name in declared:
    assert hasattr(lib, name), name
  assert lib.dd_abi_version() == 1


def test_no_cpu_fallback_without_device():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(_lib.DDError):
    _lib.Context(0)
  lib = _lib.load_library()
  handle = ctypes.c_void_p()
  assert lib.dd_ctx_create(0, ctypes.byref(handle)) == -4      # DD_ERR_NO_DEVICE
  assert b"no CPU fallback" in lib.dd_last_error()
  arch = Architecture(synthetic.example_architecture_json())
  with pytest.raises(_lib.DDError):
    arch.predict(synthetic.synthetic_features(arch, 1, 16, 16))


def test_example_json_feature_layout():
  a = Architecture(synthetic.example_architecture_json())
  assert a.feature_prediction_tuple_type == FeaturePredictionTupleType.SINGLE
  names = [fp.name for fp in a.feature_predictions]
  # sorted combined names x [Color, Direct, Indirect], empty names skipped (Architecture.py:420-473)
  assert names == ["Alpha", "Diffuse Color", "Diffuse Direct", "Diffuse Indirect", "Emission", "Environment",
                   "Glossy Color", "Glossy Direct", "Glossy Indirect", "Subsurface Color", "Subsurface Direct",
                   "Subsurface Indirect", "Transmission Color", "Transmission Direct", "Transmission Indirect",
                   "Volume Direct", "Volume Indirect"]
  assert len(a.feature_prediction_tuples) == 17 and all(fp.load_data for fp in a.feature_predictions)
  assert [f.name for f in a.auxiliary_features] == ["Normal"]
  assert a.feature_predictions[0].number_of_channels == 1
  assert a.feature_flags is not None and a.feature_flags.embedding_dimension == 8
  assert a.number_of_input_channels == 16 and a.number_of_output_channels == 25
  kinds = [e[0] for e in a.input_layouts[0]]
  assert kinds == ["source"] * 3 + ["variance"] + ["source"] * 3 + ["variance"] + ["embedding"] * 8
  assert a.input_layouts[3][0][1].name == "Diffuse Indirect" and a.input_layouts[3][4][1].name == "Normal"


def test_combined_tuples_and_synthetic_members():
  j, n, h, w = cases.case("combined_onehot")
  a = Architecture(j)
  assert len(a.feature_prediction_tuples) == 8 and len(a.feature_predictions) == 24
  assert a.feature_flags is None and a.feature_flag_mode == FeatureFlagMode.ONE_HOT_ENCODING
  vol = a.feature_prediction_tuples[-1]
  assert vol.name == "Volume" and [fp.load_data for fp in vol.feature_predictions] == [False, True, True]
  assert vol.feature_predictions[0].name == "Volume Color"
  assert vol.feature_predictions[0].feature_prediction_type == FeaturePredictionType.COLOR
  assert a.number_of_input_channels == 4 * 4 + 8 and a.number_of_output_channels == 3 * 9
  assert float(vol.feature_predictions[0].synthetic_source(1, 2, 2)[0, 0, 0, 0]) == 1.0
  assert float(vol.feature_predictions[1].feature_standardization.use_log1p) == 1.0


def test_baseline_configs_match_survey_work_counts():
  a = Architecture(synthetic.baseline_architecture_json("unet32"))
  assert a.number_of_input_channels == 32
  assert int(a.mac_per_pixel()) == 566182                  # SURVEY Appendix B.1 / BASELINE.md section 3
  assert a.spec.parameter_count() == 1691218
  t = Architecture(synthetic.baseline_architecture_json("tiramisu32"))
  assert int(t.mac_per_pixel()) == 4157694
  assert t.spec.scale_channels == [1216, 1184, 640]        # Appendix B.2
  assert Architecture(synthetic.baseline_architecture_json("rgb9")).number_of_input_channels == 9
  assert int(Architecture(synthetic.example_architecture_json()).mac_per_pixel()) == 556966


@pytest.mark.parametrize("name", cases.GOLDEN_CASES + ("rgb9",))
def test_layout_and_variables_agree_with_oracle(name):
  j, arch, weights, features = cases.build(name)
  oracle = reference_model.Architecture(j, weights=weights)
  assert [t.name for t in oracle.feature_prediction_tuples] == [t.name for t in arch.feature_prediction_tuples]
  assert [f.name for f in oracle.feature_predictions] == [f.name for f in arch.feature_predictions]
  assert oracle.number_of_output_channels == arch.number_of_output_channels
  for f in oracle.feature_predictions + oracle.auxiliary_features:
    f.initialize_sources(features, np.float64)
    f.standardize()
  for tup, layout in zip(oracle.feature_prediction_tuples, arch.input_layouts):
    x = oracle._network_input(tup, features)
    assert x.shape[3] == len(layout) == arch.number_of_input_channels
    # spot-check the channel order: every 'source' channel equals the oracle's standardised source
    for ch, entry in enumerate(layout):
      if entry[0] == "source":
        fp = next(f for f in oracle.feature_predictions + oracle.auxiliary_features if f.name == entry[1].name)
        src = fp.source[0]
        np.testing.assert_array_equal(x[..., ch], src[..., entry[2] if src.shape[3] == 3 else 0])
  oracle2 = reference_model.Architecture(j, weights=weights)
  oracle2.predict(features)
  assert [(n, tuple(oracle2.store.values[n].shape)) for n in oracle2.store.created] == arch.spec.variable_shapes()


def test_render_passes_and_naming_contract():
  assert RenderPasses.number_of_channels("Alpha") == 1 and RenderPasses.number_of_channels("Depth") == 1
  assert RenderPasses.number_of_channels("Normal") == 3 and RenderPasses.number_of_channels("") == 3
  assert RenderPasses.is_combined_feature_render_pass("Glossy") and not RenderPasses.is_combined_feature_render_pass("Volume")
  assert RenderPasses.direct_or_indirect_to_color_render_pass("Diffuse Direct") == "Diffuse Color"
  assert RenderPasses.direct_or_indirect_to_color_render_pass("Diffuse Indirect") == "Diffuse Indirect"   # reference typo kept
  assert RenderPasses.combined_to_color_render_pass("Emission") == "Emission"
  assert not RenderPasses.is_rgb_color_render_pass("Screen Space Normal") and RenderPasses.is_rgb_color_render_pass("Shadow")
  usage = RenderPassesUsage(use_volume_indirect=True, use_alpha=True, use_glossy_color=True)
  assert usage.render_passes() == ["Alpha", "Glossy Color", "Volume Indirect"]
  assert Naming.source_feature_name("Normal", index=0) == "source_image/0/Normal"
  assert Naming.source_feature_name("Normal", samples_per_pixel=16, index=1, masked=True) == "source_image/16/1/Normal Masked"
  assert Naming.target_feature_name("Alpha") == "target_image/Alpha"
  assert Naming.feature_prediction_name("Diffuse Color") == "prediction/Diffuse Color"
  assert Naming.feature_flags_name("Diffuse") == "feature_flag/Diffuse"
  assert Naming.mean_name("Diffuse", masked=True, scale_index=2) == "combined_diffuse_mean_masked/4"
  assert Naming.difference_name("Volume Direct") == "volume_direct_difference"


# ------------------------------------------------------------------------------------------------ data-parallel exchange (gloo)
def _dp_worker(rank, world, port, out_dir):
  import os
  import numpy as np
  import torch
  import torch.distributed as dist
  from deepdenoiser_b200.training import data_parallel_reduce
  os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  # a toy "mean over the tiles of this rank" loss: L_r = mean_i 0.5 * (theta . x_i)^2 over the rank's 3 tiles
  rng = np.random.default_rng(7)
  theta = torch.from_numpy(rng.standard_normal(5).astype(np.float32))
  tiles = torch.from_numpy(rng.standard_normal((world * 3, 5)).astype(np.float32))
  mine = tiles[rank * 3:(rank + 1) * 3]
  proj = mine @ theta
  loss = (0.5 * proj ** 2).mean().reshape(1).clone()
  grad = ((proj[:, None] * mine).mean(dim=0)).clone()
  scale = data_parallel_reduce(grad, loss, world)
  np.savez(os.path.join(out_dir, "rank%d.npz" % rank), grad=(grad * scale).numpy(), loss=loss.numpy())
  dist.destroy_process_group()


def test_data_parallel_reduce_equals_the_global_batch_gradient(tmp_path):
  import socket
  import numpy as np
  import torch
  import torch.multiprocessing as mp
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  world = 2
  mp.spawn(_dp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  rng = np.random.default_rng(7)
  theta = rng.standard_normal(5).astype(np.float32)
  tiles = rng.standard_normal((world * 3, 5)).astype(np.float32)
  proj = tiles @ theta
  want_loss, want_grad = float((0.5 * proj ** 2).mean()), (proj[:, None] * tiles).mean(axis=0)
  for rank in range(world):
    z = np.load(str(tmp_path / ("rank%d.npz" % rank)))
    assert np.allclose(z["grad"], want_grad, atol=1e-6) and abs(float(z["loss"][0]) - want_loss) < 1e-6


def test_training_settings_parse_the_reference_json():
  """TrainingSettings reads the loss blocks of TrainingExample.json (:31-98) and refuses only what the reference itself
  cannot run."""
  import json
  import os
  import pytest
  from deepdenoiser_b200.training import TrainingSettings
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  s = TrainingSettings(json.load(open(os.path.join(root, "configs", "TrainingExample.json"))))
  assert (s.feature_weight, s.combined_feature_weight, s.combined_image_weight) == (1.0, 5.0, 10.0)
  assert s.loss_difference == "SMAPE" and s.use_multiscale_loss and s.learning_rate == 1e-3 and s.batch_size == 8
  assert s.feature_variation_weight == 0.0 and s.feature_ms_ssim_weight == 0.0 and s.combined_feature_masked_weight == 0.0
  s = TrainingSettings({"features_training_settings": {"loss_weights": {"mean": 2.0, "variation": 0.5, "ms_ssim": 0.25},
                                                       "loss_weights_masked": {"mean": 0.75}}})
  assert (s.feature_weight, s.feature_variation_weight, s.feature_ms_ssim_weight, s.feature_masked_weight) == (2.0, 0.5, 0.25, 0.75)
  with pytest.raises(NotImplementedError):
    TrainingSettings({"features_training_settings": {"loss_weights_masked": {"ms_ssim": 0.1}}})
  with pytest.raises(NotImplementedError):
    TrainingSettings({"combined_features_training_settings": {"loss_weights_masked": {"variation": 0.1}}})
  with pytest.raises(NotImplementedError):
    TrainingSettings({"combined_image_training_settings": {"loss_weights_masked": {"mean": 0.1}}})
