"""GPU probe for the tcgen05 weight-gradient kernel: a correctness spot check + TFLOP/s on the U-Net layer shapes.
Writes gpurun_out/probe_wgrad.json."""
import ctypes
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
_b = ctypes.byref
out = []


def check(variant, ks, cin, cout, n, h, w):
  x = torch.randn(n, h, w, cin, device="cuda").half()
  dz = torch.randn(n, h, w, cout, device="cuda").half()
  dw = torch.zeros(ks, ks, cin, cout, device="cuda")
  try:
    ctx.call("dd_conv2d_wgrad_tc", _b(_lib.desc(x)), _b(_lib.desc(dz)), ks, 0, ctypes.c_void_p(dw.data_ptr()), ctypes.c_float(1.0))
    torch.cuda.synchronize()
  except Exception as e:  # noqa: BLE001
    print("variant", variant, "FAILED", e, flush=True)
    return
  wt = torch.zeros(cout, cin, ks, ks, device="cuda", requires_grad=True)
  F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=ks // 2).backward(dz.float().permute(0, 3, 1, 2))
  want = wt.grad.permute(2, 3, 1, 0)
  err = float((dw - want).abs().max()) / float(want.abs().max())
  per_tap = [(float((dw[r, s] - want[r, s]).abs().max()) / float(want.abs().max())) for r in range(ks) for s in range(ks)]
  print("variant %d ks %d %d->%d @%dx%dx%d: rel err %.3e  per tap %s" % (variant, ks, cin, cout, n, h, w, err,
                                                                        ["%.1e" % e for e in per_tap]), flush=True)
  out.append(dict(kind="check", variant=variant, shape=[ks, cin, cout, n, h, w], err=err))


def time_case(ks, cin, cout, n, h, w, iters=10):
  x = torch.randn(n, h, w, cin, device="cuda").half()
  dz = torch.randn(n, h, w, cout, device="cuda").half()
  dw = torch.zeros(ks, ks, cin, cout, device="cuda")
  args = (_b(_lib.desc(x)), _b(_lib.desc(dz)), ks, 0, ctypes.c_void_p(dw.data_ptr()), ctypes.c_float(1.0))
  for _ in range(2):
    ctx.call("dd_conv2d_wgrad_tc", *args)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    ctx.call("dd_conv2d_wgrad_tc", *args)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / iters
  rec = dict(kind="time", shape=[ks, cin, cout, n, h, w], ms=ms, tflops=2.0 * n * h * w * cin * cout * ks * ks / ms / 1e9)
  out.append(rec)
  print(rec, flush=True)


if __name__ == "__main__":
  os.makedirs("gpurun_out", exist_ok=True)
  check(0, 3, 64, 64, 1, 9, 128)
  check(0, 1, 64, 64, 1, 9, 128)
  check(0, 3, 128, 96, 2, 33, 200)
  for shape in [(3, 64, 64, 16, 256, 256), (3, 32, 64, 16, 256, 256), (3, 128, 64, 16, 256, 256), (3, 96, 96, 16, 128, 128),
                (3, 192, 96, 16, 128, 128), (3, 128, 128, 16, 64, 64), (3, 64, 64, 8, 1080, 1920), (1, 64, 32, 16, 256, 256)]:
    time_case(*shape)
  json.dump(out, open("gpurun_out/probe_wgrad.json", "w"), indent=1)
