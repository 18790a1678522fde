"""Ablation timing of conv_rows_kernel: which stage of the row pipeline bounds a layer?  Each switch removes one stage
(results are wrong, only the time matters): see ConvRowsParams::dbg.  Prints ms, TFLOP/s-equivalent and SM cycles per input row."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
dev = ctx.device
MODES = [(0, "baseline"), (1, "no convert/stage/store"), (2, "no epilogue work at all"), (4, "no UMMA"), (6, "no UMMA, no epilogue"),
         (8, "A loads hit L2"), (9, "A loads hit L2, no store"), (16, "1 k-step per (chunk, shift)"), (18, "1 k-step, no epilogue")]


def run(n, h, w, cin, cout, ks, iters=10):
  x = (torch.randn(n, h, w, (cin + 7) // 8 * 8, device=dev) * 0.5).half()
  wt = torch.randn(ks, ks, cin, cout) * 0.05
  wp = ctx.pack_conv_weights(wt, torch.float16)
  bias = torch.zeros((cout + 15) // 16 * 16, device=dev)
  y = torch.empty(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float16, device=dev)
  xd, yd = _lib.desc(x, cin, 0), _lib.desc(y, cout, 0)
  rows_per_cta = n * ((w + 127) // 128) * h / 148.0
  print("conv %dx%dx%d %d->%d k%d (%.0f rows per CTA)" % (n, h, w, cin, cout, ks, rows_per_cta), flush=True)
  for mode, name in MODES:
    ctx.set_option("conv_dbg", mode)
    for _ in range(2):
      ctx.conv2d(xd, wp, bias, ks, yd, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
      ctx.conv2d(xd, wp, bias, ks, yd, relu=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * n * h * w * cin * cout * ks * ks / ms / 1e9
    print("  dbg %2d %-32s %7.3f ms  %7.1f TFLOP/s-eq  %6.2f us/row" % (mode, name, ms, tf, ms * 1e3 / rows_per_cta), flush=True)
  ctx.set_option("conv_dbg", 0)


if __name__ == "__main__":
  run(6, 1080, 1920, 64, 64, 3)
  run(6, 1080, 1920, 32, 64, 3)
  run(6, 1080, 1920, 128, 64, 3)
  run(6, 540, 960, 96, 96, 3)
  run(6, 540, 960, 192, 96, 3)
  run(6, 270, 480, 128, 128, 3)
