"""On-device data augmentation of a training batch: the reference's DataAugmentation.py applied by
Training.input_fn_tfrecords.data_augmentation (Training.py:794-821) - per example one draw of flip / rot90 / RGB permutation
/ normal rotation shared by every pass of that example - as one gather kernel per pass (dd_augment_tiles)."""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from .Naming import Naming
from .RenderPasses import RenderPasses

AUG_PLAIN, AUG_COLOR, AUG_SCREEN_SPACE_NORMAL, AUG_NORMAL = 0, 1, 2, 3


class DataAugmentationUsage:
  """DataAugmentation.py:203-209 / TrainingExample.json:17-23."""

  def __init__(self, use_rotate_90, use_flip_left_right, use_rgb_permutation, use_normal_rotation):
    self.use_rotate_90 = use_rotate_90
    self.use_flip_left_right = use_flip_left_right
    self.use_rgb_permutation = use_rgb_permutation
    self.use_normal_rotation = use_normal_rotation

  @staticmethod
  def from_json(training_json):
    d = training_json.get("data_augmentation", {})
    return DataAugmentationUsage(bool(d.get("use_rotate_90", False)), bool(d.get("use_flip_left_right", False)),
                                 bool(d.get("use_rgb_permutation", False)), bool(d.get("use_normal_rotation", False)))


def random_rotation_matrix(random_vector):
  """DataAugmentation.random_rotation_matrix (DataAugmentation.py:127-182; Arvo, Graphics Gems III): 3 uniforms -> [3,3]."""
  x0, x1, x2 = (float(v) for v in random_vector)
  theta, phi, z = x0 * 2.0 * math.pi, x1 * 2.0 * math.pi, x2 * 2.0
  r = math.sqrt(z)
  vx, vy, vz = math.sin(phi) * r, math.cos(phi) * r, math.sqrt(2.0 - z)
  st, ct = math.sin(theta), math.cos(theta)
  sx, sy = vx * ct - vy * st, vx * st + vy * ct
  return np.array([[vx * sx - ct, vx * sy - st, vx * vz],
                   [vy * sx + st, vy * sy - ct, vy * vz],
                   [vz * sx, vz * sy, 1.0 - z]], dtype=np.float32)


def pass_kind(name, channels, usage):
  """Which per-pass fix-up applies (FeatureTrainingAugmentation, Training.py:556-605)."""
  if channels != 3:
    return AUG_PLAIN
  if name == RenderPasses.SCREEN_SPACE_NORMAL:
    return AUG_SCREEN_SPACE_NORMAL
  if name == RenderPasses.NORMAL:
    return AUG_NORMAL if usage.use_normal_rotation else AUG_PLAIN
  if usage.use_rgb_permutation and RenderPasses.is_rgb_color_render_pass(name):
    return AUG_COLOR
  return AUG_PLAIN


def draw(usage, examples, rng):
  """The random draws of data_augmentation (Training.py:796-801), one set per example.  rng: numpy Generator."""
  return {"flip": rng.integers(0, 2, size=examples).astype(np.int32) if usage.use_flip_left_right else None,
          "rot": rng.integers(0, 4, size=examples).astype(np.int32) if usage.use_rotate_90 else None,
          "perm": rng.integers(0, 6, size=examples).astype(np.int32) if usage.use_rgb_permutation else None,
          "rotation": (np.stack([random_rotation_matrix(rng.random(3)) for _ in range(examples)]).astype(np.float32)
                       if usage.use_normal_rotation else None)}


def _pass_name(key):
  return key.split("/")[-1]


class DeviceAugmenter:
  """Applies one batch's draws to every source / target tensor on the GPU."""

  def __init__(self, ctx, usage):
    self.ctx, self.usage = ctx, usage

  def __call__(self, sources, targets, draws):
    ctx, dev = self.ctx, self.ctx.device
    up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    flip, rot, perm, rotation = (up(draws[k]) for k in ("flip", "rot", "perm", "rotation"))
    ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())   # noqa: E731
    out = []
    for group in (sources, targets):
      res = {}
      for key, value in group.items():
        x = value if torch.is_tensor(value) else torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32))
        x = x.to(dev, torch.float32).contiguous()
        if key.startswith("feature_flag/"):
          res[key] = x
          continue
        name = _pass_name(key)
        if self.usage.use_flip_left_right and name == RenderPasses.NORMAL:
          raise Exception("Flipping for normals is not supported.")          # DataAugmentation.py:21-22
        kind = pass_kind(name, x.shape[3], self.usage)
        y = torch.empty_like(x)
        ctx.call("dd_augment_tiles", ctypes.byref(_lib.desc(x)), kind, ptr(flip), ptr(rot), ptr(perm),
                 ptr(rotation) if kind == AUG_NORMAL else None, ctypes.byref(_lib.desc(y)))
        res[key] = y
      out.append(res)
    return out[0], out[1]
