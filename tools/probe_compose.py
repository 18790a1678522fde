"""Timing probe of the fused compose kernel (dd_compose_scales_fwd): ms, pixels/s, useful TFLOP/s (20,904 MAC/px)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
dev = ctx.device
rng = np.random.default_rng(0)
_blob, _floats, _code = _lib.pack_compose_weights(
    rng.standard_normal((6, 24)).astype(np.float32) * 0.3, np.zeros(24, np.float32),
    [rng.standard_normal((3, 3, 24, 24)).astype(np.float32) * 0.08 for _ in range(4)], [np.zeros(24, np.float32)] * 4,
    rng.standard_normal(24).astype(np.float32) * 0.3, np.zeros(1, np.float32))
blob = (torch.from_numpy(_blob).to(dev), _floats, _code)
out = []
for (n, h, w) in [(8, 1080, 1920), (8, 540, 960), (1, 1080, 1920)]:
  small = torch.randn(n, h // 2, w // 2, 3, device=dev)
  large = torch.randn(n, h, w, 3, device=dev)
  dst = torch.empty(n, h, w, 3, device=dev)
  sd, ld, od = _lib.desc(small), _lib.desc(large), _lib.desc(dst)
  for _ in range(3):
    ctx.compose_scales(sd, ld, blob, None, od)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    ctx.compose_scales(sd, ld, blob, None, od)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 10
  px = n * h * w
  rec = dict(shape=[n, h, w], ms=ms, gpx_per_s=px / ms / 1e6, useful_tflops=px * 20904 * 2 / ms / 1e9)
  out.append(rec)
  print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_compose.json", "w"), indent=1)
