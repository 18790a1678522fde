"""HBM roofline probe of the kernel-prediction apply (dd_kernel_predict_fwd): achieved GB/s on ALGORITHMIC bytes
(K*K*sizeof(logit) + 12 B source + 12 B output per pixel) against the measured copy bandwidth."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
dev = ctx.device
peak = 6551.0
if os.path.exists("MEASURED_PEAKS.json"):
  peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
out = []
cases = [(5, torch.float32, 8, 1080, 1920), (5, torch.float32, 8, 540, 960), (5, torch.float16, 8, 1080, 1920),
                            (21, torch.float16, 2, 1080, 1920), (21, torch.float32, 1, 1080, 1920)]
if os.environ.get("KP_CASES"):
  cases = [cases[int(i)] for i in os.environ["KP_CASES"].split(",")]
for (k, dtype, n, h, w) in cases:
  k2 = k * k
  cs = (k2 + 7) // 8 * 8
  logits = torch.randn(n, h, w, cs, device=dev).to(dtype)
  src = torch.randn(n, h, w, 3, device=dev)
  dst = torch.empty(n, h, w, 3, device=dev)
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
  ld, sd, od = _lib.desc(logits, k2, 0), _lib.desc(src), _lib.desc(dst)
  for _ in range(3):
    ctx.kernel_predict(sd, ld, k, 1, 1, od)
  times = []
  for _ in range(10):
    ctx.l2_flush(flush)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.kernel_predict(sd, ld, k, 1, 1, od)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
  ms = sorted(times)[len(times) // 2]
  bytes_alg = n * h * w * (k2 * logits.element_size() + 24)
  rec = dict(K=k, logits=str(dtype), shape=[n, h, w], ms=ms, gbps=bytes_alg / ms / 1e6, frac_of_hbm_peak=bytes_alg / ms / 1e6 / peak)
  out.append(rec)
  print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_kp.json", "w"), indent=1)
