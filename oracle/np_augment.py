"""ORACLE (test infrastructure only - never imported by the product path): numpy restatement of the reference's
DataAugmentation.py applied per example in the order of Training.py:803-815.  tf.image.flip_left_right == np.fliplr,
tf.image.rot90(k) == np.rot90(k) (counter-clockwise) [TF documentation; external]."""
import numpy as np

PERMUTATIONS = {1: [0, 2, 1], 2: [1, 0, 2], 3: [1, 2, 0], 4: [2, 0, 1], 5: [2, 1, 0]}    # DataAugmentation.py:117-123


def flip_left_right(image, name, flip):
  """DataAugmentation.py:10-28 (+ _flip_screen_space_normals :31-45)."""
  if flip > 0:
    image = image[:, ::-1, :]
    if name == "Screen Space Normal":
      image = image * np.array([-1.0, 1.0, 1.0], dtype=image.dtype)
  return image


def rotate_90(image, k, name):
  """DataAugmentation.py:47-104."""
  image = np.rot90(image, k=k, axes=(0, 1))
  if name == "Screen Space Normal":
    x, y, z = image[..., 0], image[..., 1], image[..., 2]
    if k == 1:
      x, y = -y, x
    elif k == 2:
      x, y = -x, -y
    elif k == 3:
      x, y = y, -x
    image = np.stack([x, y, z], axis=-1)
  return image


def permute_rgb(image, permute):
  """DataAugmentation.py:106-125."""
  if permute in PERMUTATIONS:
    image = image[..., PERMUTATIONS[permute]]
  return image


def rotate_normal(image, rotation_matrix):
  """DataAugmentation.py:184-200: reshape [h*w,3] @ R."""
  h, w, _ = image.shape
  return (image.reshape(h * w, 3) @ rotation_matrix).reshape(h, w, 3)


def augment_example(image, name, is_color, flip=None, rot=None, perm=None, rotation=None):
  """One pass of one example through FeatureTrainingAugmentation (Training.py:556-605)."""
  if flip is not None:
    image = flip_left_right(image, name, flip)
  if rot is not None:
    image = rotate_90(image, rot, name)
  if perm is not None and is_color and image.shape[-1] == 3:
    image = permute_rgb(image, perm)
  if rotation is not None and name == "Normal":
    image = rotate_normal(image, rotation)
  return np.ascontiguousarray(image)
