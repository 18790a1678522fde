"""Per-tile pipeline timeline of the tcgen05 conv kernel (CTA 0): where do the cycles of a tile go?"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib
ctx = _lib.Context(0)
dev = ctx.device

def run(n, h, w, cin, cout, ks, iters=3, f32_out=False, residual=False, relu_copy=False):
  x = (torch.randn(n, h, w, cin, device=dev) * 0.5).half()
  wt = torch.randn(ks, ks, cin, cout) * 0.05
  wp = ctx.pack_conv_weights(wt, torch.float16)
  bias = torch.zeros((cout + 15) // 16 * 16, device=dev)
  y = torch.empty(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float32 if f32_out else torch.float16, device=dev)
  res = _lib.desc(torch.zeros(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float16, device=dev), cout, 0) if residual else None
  yr = _lib.desc(torch.zeros(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float16, device=dev), cout, 0) if relu_copy else None
  xd, yd = _lib.desc(x), _lib.desc(y, cout, 0)
  trace = torch.zeros(64 * 8 + 256, dtype=torch.int64, device=dev)
  for _ in range(iters):
    ctx.conv2d(xd, wp, bias, ks, yd, relu=not residual, residual=res, y_relu=yr)
  ctx.set_trace_buffer(trace)
  ctx.conv2d(xd, wp, bias, ks, yd, relu=not residual, residual=res, y_relu=yr)
  torch.cuda.synchronize()
  ctx.set_trace_buffer(None)
  full = trace.cpu()
  t = full[:512].view(64, 8)
  print('  streamed path, cycles the issuing thread waited per group (weights | input rows + accumulators):',
        ' '.join('%d|%d' % (int(full[512 + i]), int(full[576 + i])) for i in range(12)))
  base = int(t[0, 0])
  print("conv %dx%dx%d %d->%d k%d: rows = tile iteration; cycles relative to start" % (n, h, w, cin, cout, ks))
  print("  it | grp:enter     plan   a_full   issued | epi:enter  t_full    done | mma_busy epi_busy  period | waits_done probe_a probe_e")
  prev = None
  for i in range(12):
    raw7 = int(t[i, 7])
    pa, pe = (raw7 >> 62) & 1, (raw7 >> 61) & 1
    t7 = (raw7 & ((1 << 61) - 1)) - base
    r = [int(v) - base for v in t[i]]
    period = (r[3] - prev) if prev is not None else 0
    prev = r[3]
    print("  %2d | %9d %8d %8d %8d | %9d %8d %8d | %8d %8d %8d | %9d %d %d" % (i, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[3] - r[2], r[6] - r[5], period, t7, pa, pe))

import sys as _s
which = _s.argv[1] if len(_s.argv) > 1 else "big"
if which == "small":
  run(2, 1080, 1920, 24, 24, 3)
  run(2, 1080, 1920, 24, 24, 3, residual=True, relu_copy=True)
  run(2, 1080, 1920, 32, 25, 1, f32_out=True)
  run(2, 1080, 1920, 64, 25, 1)
else:
  run(1, 1080, 1920, 64, 64, 3)
if which == "mid":
  run(8, 540, 960, 96, 96, 3)
  run(8, 540, 960, 192, 96, 3)
  run(8, 270, 480, 128, 128, 3)
