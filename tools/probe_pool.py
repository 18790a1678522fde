"""GPU timing probe of the 3x3 / stride-2 max pooling at the U-Net shapes (GB/s of unique bytes: read once + write once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
for n, h, w, c in ((6, 1080, 1920, 64), (6, 540, 960, 96)):
  x = torch.randn(n, h, w, c, device="cuda").half()
  y = torch.empty(n, h // 2, w // 2, c, device="cuda", dtype=torch.float16)
  for _ in range(2):
    ctx.maxpool_s2(_lib.desc(x), 3, _lib.desc(y))
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    ctx.maxpool_s2(_lib.desc(x), 3, _lib.desc(y))
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 10
  gb = (x.numel() + y.numel()) * 2 / 1e9
  print("maxpool 3x3/s2 %dx%dx%dx%d: %.3f ms, %.0f GB/s" % (n, h, w, c, ms, gb / ms * 1e3), flush=True)
