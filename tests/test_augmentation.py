"""Data augmentation: the numpy oracle (oracle/np_augment.py, restating DataAugmentation.py) against properties of the
reference's math on the CPU, and the device kernel (dd_augment_tiles) against the oracle on the GPU - bit exact, the kernel
only moves / negates values except for the normal rotation (fp32 fma order, <= 1e-6)."""
import numpy as np
import pytest

from deepdenoiser_b200 import augmentation
from oracle import np_augment


def test_random_rotation_matrix_is_a_rotation():
  rng = np.random.default_rng(0)
  for _ in range(20):
    r = augmentation.random_rotation_matrix(rng.random(3)).astype(np.float64)
    assert np.abs(r @ r.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(r) - 1.0) < 1e-5
  # known answer: random_vector = 0 -> theta = phi = z = 0 -> V = (0, 0, sqrt 2): result = diag(-1, -1, 1)
  assert np.allclose(augmentation.random_rotation_matrix([0, 0, 0]), np.diag([-1.0, -1.0, 1.0]))


def test_oracle_known_answers():
  img = np.arange(2 * 2 * 3, dtype=np.float32).reshape(2, 2, 3)
  # rot90 counter-clockwise: the right column becomes the top row
  assert np.array_equal(np_augment.rotate_90(img, 1, "Diffuse Color")[0, :, 0], img[:, 1, 0])
  assert np.array_equal(np_augment.rotate_90(np_augment.rotate_90(img, 1, "x"), 3, "x"), img)
  assert np.array_equal(np_augment.flip_left_right(img, "x", 1)[:, 0], img[:, 1])
  assert np.array_equal(np_augment.permute_rgb(img, 3)[..., 0], img[..., 1])          # [1, 2, 0]
  # screen-space normals: four quarter turns give the identity, two give (-x, -y)
  n = np.random.default_rng(1).standard_normal((4, 4, 3)).astype(np.float32)
  out = n
  for _ in range(4):
    out = np_augment.rotate_90(out, 1, "Screen Space Normal")
  assert np.array_equal(out, n)
  two = np_augment.rotate_90(np_augment.rotate_90(n, 1, "Screen Space Normal"), 1, "Screen Space Normal")
  assert np.array_equal(two, np_augment.rotate_90(n, 2, "Screen Space Normal"))


def test_usage_from_training_json_and_pass_kinds():
  usage = augmentation.DataAugmentationUsage.from_json({"data_augmentation": {"use_rotate_90": True, "use_rgb_permutation": True,
                                                                              "use_normal_rotation": True}})
  assert usage.use_rotate_90 and not usage.use_flip_left_right
  assert augmentation.pass_kind("Diffuse Color", 3, usage) == augmentation.AUG_COLOR
  assert augmentation.pass_kind("Normal", 3, usage) == augmentation.AUG_NORMAL
  assert augmentation.pass_kind("Screen Space Normal", 3, usage) == augmentation.AUG_SCREEN_SPACE_NORMAL
  assert augmentation.pass_kind("Depth", 1, usage) == augmentation.AUG_PLAIN
  d = augmentation.draw(usage, 5, np.random.default_rng(2))
  assert d["flip"] is None and d["rot"].shape == (5,) and d["perm"].max() <= 5 and d["rotation"].shape == (5, 3, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("flip_on", [False, True])
def test_device_augmentation_matches_oracle(ctx, flip_on):
  import torch
  usage = augmentation.DataAugmentationUsage(True, flip_on, True, not flip_on)
  rng = np.random.default_rng(5)
  e, s = 6, 12
  names = {"Diffuse Color": 3, "Screen Space Normal": 3, "Depth": 1, "Diffuse Direct": 3}
  if not flip_on:
    names["Normal"] = 3                       # the reference refuses to flip world-space normals
  sources = {"source_image/0/" + k: rng.standard_normal((e, s, s, c)).astype(np.float32) for k, c in names.items()}
  targets = {"target_image/Diffuse Color": rng.standard_normal((e, s, s, 3)).astype(np.float32)}
  draws = augmentation.draw(usage, e, rng)
  draws["rot"] = np.array([0, 1, 2, 3, 1, 2], dtype=np.int32)
  got_s, got_t = augmentation.DeviceAugmenter(ctx, usage)(sources, targets, draws)
  torch.cuda.synchronize()
  for group, got in ((sources, got_s), (targets, got_t)):
    for key, value in group.items():
      name = key.split("/")[-1]
      for i in range(e):
        want = np_augment.augment_example(
            value[i], name, is_color=augmentation.pass_kind(name, value.shape[3], usage) == augmentation.AUG_COLOR,
            flip=int(draws["flip"][i]) if draws["flip"] is not None else None, rot=int(draws["rot"][i]),
            perm=int(draws["perm"][i]), rotation=draws["rotation"][i] if draws["rotation"] is not None else None)
        g = got[key][i].cpu().numpy()
        if name == "Normal":
          assert np.abs(g - want).max() <= 1e-5, key
        else:
          assert np.array_equal(g, want), (key, i)
