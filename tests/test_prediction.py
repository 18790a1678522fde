"""Prediction driver: tile grid / stitch / combine logic (CPU), EXR round trip, sharded gather with gloo, and the
end-to-end CLI on a synthetic EXR directory (GPU) against the oracle run tile by tile like the reference does."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
from deepdenoiser_b200 import prediction, synthetic
from deepdenoiser_b200.Architecture import Architecture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("h,w,tile,overlap", [(1080, 1920, 128, 14), (64, 64, 128, 14), (100, 77, 64, 7), (128, 128, 128, 14),
                                              (129, 300, 128, 14), (16, 16, 128, 14), (270, 481, 128, 0)])
def test_tile_grid_partitions_the_image(h, w, tile, overlap):
  tiles, ts, ov = prediction.tile_grid(h, w, tile, overlap)
  cover = np.zeros((h, w), dtype=np.int32)
  for t in tiles:
    assert 0 <= t.y and t.y + ts <= h and 0 <= t.x and t.x + ts <= w
    cy0, cy1, cx0, cx1 = t.crop
    assert 0 <= cy0 < cy1 <= ts and 0 <= cx0 < cx1 <= ts
    dy0, dy1, dx0, dx1 = t.dest
    assert (dy1 - dy0, dx1 - dx0) == (cy1 - cy0, cx1 - cx0)
    cover[dy0:dy1, dx0:dx1] += 1
  assert (cover == 1).all()                       # every pixel is produced by exactly one tile
  if (h, w, tile, overlap) == (1080, 1920, 128, 14):
    assert len(tiles) == 11 * 19                  # 209 tiles at 1080p
  # interior tiles keep the centre: overlap pixels dropped on shared sides (Prediction.py:396-427)
  if len(tiles) > 9 and min(h, w) >= tile:
    inner = [t for t in tiles if 0 < t.y < h - ts and 0 < t.x < w - ts]
    assert all(t.crop == (ov, ts - ov, ov, ts - ov) for t in inner)


def test_small_image_shrinks_tile_and_rejects_tiny():
  tiles, ts, ov = prediction.tile_grid(64, 80, 128, 14)
  assert ts == 64 and ov == 7
  with pytest.raises(ValueError):
    prediction.tile_grid(15, 200)


def test_stitch_of_tiles_is_identity_for_identity_network():
  rng = np.random.default_rng(0)
  image = torch.from_numpy(rng.standard_normal((150, 201, 3)).astype(np.float32))
  out = prediction.predict_image(None, {"source_image/0/X": image}, 150, 201, 64, 7, tiles_per_batch=5,
                                 predict_fn=lambda f: {"prediction/X": f["source_image/0/X"] * 2})
  assert torch.equal(out["prediction/X"], image * 2)
  full = prediction.predict_image(None, {"source_image/0/X": image}, 150, 201, full_frame=True,
                                  predict_fn=lambda f: {"prediction/X": f["source_image/0/X"] * 2})
  assert torch.equal(full["prediction/X"], image * 2)


def test_combine_passes_formula():
  rng = np.random.default_rng(1)
  p = {}
  for light in ("Diffuse", "Glossy", "Subsurface", "Transmission"):
    for kind in ("Color", "Direct", "Indirect"):
      p["prediction/%s %s" % (light, kind)] = torch.from_numpy(rng.uniform(size=(4, 5, 3)))
  for name in ("Volume Direct", "Volume Indirect", "Environment", "Emission"):
    p["prediction/" + name] = torch.from_numpy(rng.uniform(size=(4, 5, 3)))
  p["prediction/Alpha"] = torch.ones(4, 5, 1)
  image, combined = prediction.combine_passes(p)
  want = sum(p["prediction/%s Color" % l] * (p["prediction/%s Direct" % l] + p["prediction/%s Indirect" % l])
             for l in ("Diffuse", "Glossy", "Subsurface", "Transmission"))
  want = want + p["prediction/Volume Direct"] + p["prediction/Volume Indirect"] + p["prediction/Environment"] + p["prediction/Emission"]
  torch.testing.assert_close(image, want)
  assert set(combined) == {"Diffuse", "Glossy", "Subsurface", "Transmission"}


def _write_exr_dir(directory, arch, h, w, seed=3):
  feats = synthetic.synthetic_features(arch, 1, h, w, seed=seed)
  for fp in arch.required_features():
    if fp.load_data:
      img = feats["source_image/0/" + fp.name][0]
      if img.shape[2] == 1:
        img = np.repeat(img, 3, axis=2)
      prediction.save_exr(os.path.join(directory, "scene_16_0001_0_%s_0001.exr" % fp.name), img)
  return feats


def test_exr_roundtrip_and_pass_matching(tmp_path):
  arch = Architecture(synthetic.baseline_architecture_json("unet32"))
  feats = _write_exr_dir(str(tmp_path), arch, 24, 40)
  loaded, h, w = prediction.load_features(arch, str(tmp_path))
  assert (h, w) == (24, 40)
  for fp in arch.required_features():
    got = loaded["source_image/0/" + fp.name]
    want = feats["source_image/0/" + fp.name][0]
    assert got.shape == (24, 40, 3)               # Prediction loads every pass as 3 channels (Prediction.py:69-70)
    np.testing.assert_array_equal(got[..., :want.shape[2]], want)    # float32 EXR is lossless
  # 'Normal' must not pick up 'Screen Space Normal'
  files = prediction.exr_files(str(tmp_path))
  assert prediction.find_pass_file(files, "Normal").endswith("_Normal_0001.exr")
  assert prediction.find_pass_file(files, "Screen Space Normal").endswith("_Screen Space Normal_0001.exr")
  with pytest.raises(IOError):
    prediction.load_features(Architecture(synthetic.example_architecture_json()), _empty(tmp_path))


def _empty(tmp_path):
  d = tmp_path / "empty"
  d.mkdir()
  return str(d)


def _gloo_worker(rank, world, port, out_dir):
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  rng = np.random.default_rng(0)
  image = torch.from_numpy(rng.standard_normal((90, 130, 3)).astype(np.float32))
  calls = []

  def fake_predict(f):
    calls.append(f["source_image/0/X"].shape[0])
    return {"prediction/X": f["source_image/0/X"] + 1}

  out = prediction.predict_image(None, {"source_image/0/X": image}, 90, 130, 32, 4, tiles_per_batch=4, rank=rank,
                                 world_size=world, predict_fn=fake_predict)
  n_tiles = len(prediction.tile_grid(90, 130, 32, 4)[0])
  assert sum(calls) == len(range(rank, n_tiles, world))          # every rank did only its share
  if rank == 0:
    assert torch.equal(out["prediction/X"], image + 1)
    open(os.path.join(out_dir, "ok"), "w").write("1")
  else:
    assert out is None
  dist.barrier()
  dist.destroy_process_group()


def test_tiles_shard_over_two_ranks_with_gloo(tmp_path):
  import torch.multiprocessing as mp
  port = 29500 + (os.getpid() % 2000)
  mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  assert os.path.exists(os.path.join(str(tmp_path), "ok"))


@pytest.mark.gpu
def test_prediction_cli_matches_tiled_oracle(tmp_path):
  """BASELINE cfg1: 9-channel source stack (RGB + Normal + albedo), 64x64 image, through the CLI; compared with the
  oracle driven the way the reference drives TensorFlow (same tiles, batch 1)."""
  from oracle import np_ops, reference_model
  j = synthetic.baseline_architecture_json("rgb9")
  j["b200"] = {"dtype": "float32"}
  arch_host = Architecture(j)
  weights = synthetic.randomize_biases(arch_host.weights)
  np.savez(str(tmp_path / "weights.npz"), **weights)
  import json
  json.dump(j, open(str(tmp_path / "arch.json"), "w"))
  d = tmp_path / "frame"
  d.mkdir()
  h, w = 80, 96
  _write_exr_dir(str(d), arch_host, h, w)
  rc = subprocess.run([sys.executable, os.path.join(ROOT, "Prediction.py"), str(tmp_path / "arch.json"), "--input", str(d),
                       "--tile_size", "64", "--tile_overlap_size", "8", "--weights", str(tmp_path / "weights.npz")],
                      capture_output=True, text=True, cwd=ROOT)
  assert rc.returncode == 0, rc.stderr[-2000:]
  got = np.load(str(d / "Diffuse Direct.npy"))
  assert got.shape == (h, w, 3) and got.dtype == np.float32
  assert not os.path.exists(str(d / "Combined.npy"))   # no complete lighting triple in this 9-channel configuration
  # oracle, tile by tile
  feats, hh, ww = prediction.load_features(arch_host, str(d))
  oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights)
  want = prediction.predict_image(
      None, feats, hh, ww, 64, 8, tiles_per_batch=1,
      predict_fn=lambda f: {k: torch.from_numpy(v) for k, v in oracle.predict_numpy({kk: vv.numpy() for kk, vv in f.items()})[0].items()})
  err = np.abs(got - want["prediction/Diffuse Direct"].numpy()).max()
  assert err <= 1e-4, err


@pytest.mark.gpu
def test_tiled_prediction_on_the_device_matches_the_references_prediction_main():
  """The product path (frame resident on the device, tiles cut / pasted by dd_tiles_gather / dd_tiles_scatter, exact fp32
  network, lighting combination in libdd_b200) against tests/golden/refshim_prediction.npz - the output files of the
  reference's own Prediction.main() executed over oracle/tf_shim (tests/golden/make_reference_golden.py)."""
  import importlib.util
  golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
  spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(golden, "make_reference_golden.py"))
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  j, arch_host, weights, frame, h, w, tile, overlap = m.prediction_problem()
  j = dict(j)
  j["b200"] = {"dtype": "float32"}
  arch = Architecture(j, weights=weights)
  feats = {"source_image/0/" + name: img for name, img in frame.items()}
  got = prediction.predict_image(arch, feats, h, w, tile, overlap)
  image, combined = prediction.combine_passes(got, arch.ctx)
  z = np.load(os.path.join(golden, "refshim_prediction.npz"))
  for key in z.files:
    mine = image if key == "Combined" else got["prediction/" + key]
    assert tuple(mine.shape) == z[key].shape, key
    assert np.abs(mine.float().cpu().numpy() - z[key]).max() <= 1e-4 * max(1.0, float(np.abs(z[key]).max())), key


def test_tile_grid_partitions_any_image_exactly():
  """Property (ragged sizes): for any image at least 16 pixels wide the kept crops of the reference's tile grid
  (Prediction.py:259-310, 396-427) cover every pixel exactly once and every tile lies inside the image."""
  from hypothesis import given, settings, strategies as st

  @settings(max_examples=120, deadline=None)
  @given(h=st.integers(16, 400), w=st.integers(16, 400), tile=st.sampled_from([32, 64, 128]), overlap=st.integers(2, 14))
  def check(h, w, tile, overlap):
    if 2 * overlap >= tile // 2:
      return
    tiles, size, ov = prediction.tile_grid(h, w, tile, overlap)
    cover = np.zeros((h, w), dtype=np.int32)
    for t in tiles:
      assert 0 <= t.y and t.y + size <= h and 0 <= t.x and t.x + size <= w
      dy0, dy1, dx0, dx1 = t.dest
      assert (dy1 - dy0, dx1 - dx0) == (t.crop[1] - t.crop[0], t.crop[3] - t.crop[2])
      cover[dy0:dy1, dx0:dx1] += 1
    assert cover.min() == 1 and cover.max() == 1

  check()


def test_latest_checkpoint_is_found_and_stripped_of_optimizer_state(tmp_path):
  """Prediction restores the newest Training.py checkpoint of the model directory (the Estimator's behaviour)."""
  model = tmp_path / "Models" / "Example"
  model.mkdir(parents=True)
  assert prediction.latest_checkpoint("Models/Example", str(tmp_path)) is None
  for step in (5, 500, 20):
    np.savez(str(model / ("ckpt-%d.npz" % step)), step=np.array(step), **{"a/kernel": np.full((2, 2), float(step), np.float32),
             "adam_m/a/kernel": np.zeros((2, 2), np.float32), "adam_v/a/kernel": np.zeros((2, 2), np.float32)})
  path = prediction.latest_checkpoint("Models/Example", str(tmp_path))
  assert os.path.basename(path) == "ckpt-500.npz"
  weights = prediction.load_checkpoint_weights(path)
  assert list(weights) == ["a/kernel"] and float(weights["a/kernel"][0, 0]) == 500.0
