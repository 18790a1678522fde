"""GPU timing probe: fused output head (dd_post_kp_fwd) vs the unfused conv1x1 x2 + kernel prediction launches."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)


def timeit(fn, iters=10):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters


def case(b, h, w, c, k=5, f=1):
  o = f * k * k
  x = torch.randn(b, h, w, c, device="cuda").abs().half()
  src = torch.randn(b * f, h, w, 3, device="cuda")
  out = torch.empty_like(src)
  w1, b1 = np.random.randn(1, 1, c, o).astype(np.float32) * 0.1, np.zeros(o, np.float32)
  w2, b2 = np.random.randn(1, 1, o, o).astype(np.float32) * 0.1, np.zeros(o, np.float32)
  blob = torch.from_numpy(_lib.pack_post_kp_weights(w1, b1, w2, b2, k, f)).cuda()
  t_fused = timeit(lambda: ctx.post_kp(_lib.desc(x), blob, _lib.desc(src), k, f, 1, _lib.desc(out)))
  p1 = ctx.pack_conv_weights(torch.from_numpy(w1), torch.float16)
  p2 = ctx.pack_conv_weights(torch.from_numpy(w2), torch.float16)
  bias = torch.zeros(64, device="cuda")
  o8 = (o + 7) // 8 * 8
  mid = torch.empty(b, h, w, o8, device="cuda", dtype=torch.float16)
  lg = torch.empty(b, h, w, o8, device="cuda")
  t1 = timeit(lambda: ctx.conv2d(_lib.desc(x), p1, bias, 1, _lib.desc(mid, o, 0), relu=True))
  t2 = timeit(lambda: ctx.conv2d(_lib.desc(mid, o, 0), p2, bias, 1, _lib.desc(lg, o, 0)))
  t3 = timeit(lambda: ctx.kernel_predict(_lib.desc(src), _lib.desc(lg, o, 0), k, f, 1, _lib.desc(out)))
  gb = b * h * w * (c * 2 + 24 * f) / 1e9
  print("B%d %dx%d C%d K%d F%d: fused %.3f ms (%.0f GB/s algorithmic) | unfused %.3f + %.3f + %.3f = %.3f ms" %
        (b, h, w, c, k, f, t_fused, gb / t_fused * 1e3, t1, t2, t3, t1 + t2 + t3), flush=True)


if __name__ == "__main__":
  case(8, 1080, 1920, 64)
  case(8, 540, 960, 96)
  case(8, 270, 480, 128)
  case(1, 1080, 1920, 64)
  case(2, 1080, 1920, 64, 5, 3)
