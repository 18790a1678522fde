"""Sweep of the weight-streaming configuration of conv_rows_kernel (rows per weight pass G, weight stages) on the layers whose
weights do not fit in shared memory."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402
ctx = _lib.Context(0)
dev = ctx.device

def time_case(n, h, w, cin, cout, ks=3, iters=10):
  x = (torch.randn(n, h, w, cin, device=dev) * 0.5).half()
  wp = ctx.pack_conv_weights(torch.randn(ks, ks, cin, cout) * 0.05, torch.float16)
  bias = torch.zeros((cout + 15) // 16 * 16, device=dev)
  y = torch.empty(n, h, w, cout, dtype=torch.float16, device=dev)
  xd, yd = _lib.desc(x), _lib.desc(y)
  res = []
  for stages in (0, 2, 3, 4):
    for g in (0, 1, 2, 3):
      ctx.set_option("conv_b_stages", stages); ctx.set_option("conv_rows", g)
      try:
        for _ in range(2): ctx.conv2d(xd, wp, bias, ks, yd, relu=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ctx.conv2d(xd, wp, bias, ks, yd, relu=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res.append("s%d g%d: %.0f" % (stages, g, 2.0 * n * h * w * cin * cout * ks * ks / ms / 1e9))
      except Exception as e:
        res.append("s%d g%d: -" % (stages, g))
  ctx.set_option("conv_b_stages", 0); ctx.set_option("conv_rows", 0)
  print("%dx%dx%d %d->%d TFLOP/s | " % (n, h, w, cin, cout) + "  ".join(res), flush=True)

for shape in [(8, 540, 960, 96, 96), (8, 540, 960, 192, 96), (8, 540, 960, 64, 96), (8, 270, 480, 128, 128), (8, 270, 480, 96, 128)]:
  time_case(*shape)
