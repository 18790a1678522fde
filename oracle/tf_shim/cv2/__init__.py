"""TEST INFRASTRUCTURE: stand-in for the three OpenCV symbols OpenEXRDirectory._load_exr touches (OpenEXRDirectory.py:126-148),
so that the reference's own loader runs on files tests/golden/make_reference_golden.py writes: an ".exr" there is a numpy .npy
payload holding the BGR float32 image (EXR decoding itself is outside the hot path)."""
import io

import numpy as np

IMREAD_UNCHANGED = -1
COLOR_BGR2RGB = 4
COLOR_RGB2BGR = 4


class _Image(np.ndarray):
  """ndarray with the `tostring` alias numpy 2 removed (Prediction.py:333 calls it on every tile)."""

  def tostring(self, order="C"):
    return self.tobytes(order)


def imdecode(buffer, flags):
  return np.load(io.BytesIO(np.asarray(buffer, dtype=np.uint8).tobytes())).view(_Image)


def cvtColor(image, code):
  return np.ascontiguousarray(image[..., ::-1]).view(_Image)
