"""GPU timing probe for the tcgen05 row-pipeline convolution on the U-Net / Tiramisu layer shapes.
Writes gpurun_out/probe_conv.json (TFLOP/s per shape; correctness lives in tests/test_gpu_ops.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
dev = ctx.device
out = []


def time_case(n, h, w, cin, cout, ks, iters=20, f32_out=False):
  cs = (cin + 7) // 8 * 8                  # the tensor-core path wants 16-byte aligned pixels: 25 channels live in a 32-wide buffer
  x = (torch.randn(n, h, w, cs, device=dev) * 0.5).half()
  wt = torch.randn(ks, ks, cin, cout) * 0.05
  wp = ctx.pack_conv_weights(wt, torch.float16)
  bias = torch.zeros((cout + 15) // 16 * 16, device=dev)
  y = torch.empty(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float32 if f32_out else torch.float16, device=dev)
  xd, yd = _lib.desc(x, cin, 0), _lib.desc(y, cout, 0)
  for _ in range(3):
    ctx.conv2d(xd, wp, bias, ks, yd, relu=True)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    ctx.conv2d(xd, wp, bias, ks, yd, relu=True)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / iters
  flops = 2.0 * n * h * w * cin * cout * ks * ks
  gb = n * h * w * (cin * 2 + cout * (4 if f32_out else 2)) / 1e9
  rec = dict(shape=[n, h, w, cin, cout, ks], ms=ms, tflops=flops / ms / 1e9, gbps=gb / ms * 1e3)
  out.append(rec)
  print(rec, flush=True)


if __name__ == "__main__":
  os.makedirs("gpurun_out", exist_ok=True)
  for shape in [(1, 1080, 1920, 64, 64, 3), (1, 1080, 1920, 32, 64, 3), (1, 1080, 1920, 128, 64, 3),
                (1, 540, 960, 64, 96, 3), (1, 540, 960, 96, 96, 3), (1, 540, 960, 192, 96, 3),
                (1, 270, 480, 96, 128, 3), (1, 270, 480, 128, 128, 3), (8, 1080, 1920, 64, 64, 3),
                (8, 540, 960, 96, 96, 3), (8, 270, 480, 128, 128, 3), (1, 1080, 1920, 24, 24, 3),
                (1, 1080, 1920, 64, 25, 1), (1, 1080, 1920, 25, 25, 1), (1, 540, 960, 96, 25, 1)]:
    time_case(*shape)
  time_case(1, 1080, 1920, 25, 25, 1, f32_out=True)
  json.dump(out, open("gpurun_out/probe_conv.json", "w"), indent=1)
