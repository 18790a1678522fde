"""Per-tile pipeline timeline of the tcgen05 conv kernel (CTA 0): where do the cycles of a tile go?"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import _lib
ctx = _lib.Context(0)
dev = ctx.device

def run(n, h, w, cin, cout, ks, iters=3, f32_out=False, residual=False, relu_copy=False, dbg=0):
  x = (torch.randn(n, h, w, cin, device=dev) * 0.5).half()
  wt = torch.randn(ks, ks, cin, cout) * 0.05
  wp = ctx.pack_conv_weights(wt, torch.float16)
  bias = torch.zeros((cout + 15) // 16 * 16, device=dev)
  y = torch.empty(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float32 if f32_out else torch.float16, device=dev)
  res = _lib.desc(torch.zeros(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float16, device=dev), cout, 0) if residual else None
  yr = _lib.desc(torch.zeros(n, h, w, (cout + 7) // 8 * 8, dtype=torch.float16, device=dev), cout, 0) if relu_copy else None
  xd, yd = _lib.desc(x), _lib.desc(y, cout, 0)
  trace = torch.zeros(2048, dtype=torch.int64, device=dev)
  for _ in range(iters):
    ctx.conv2d(xd, wp, bias, ks, yd, relu=not residual, residual=res, y_relu=yr)
  ctx.set_trace_buffer(trace)
  ctx.set_option('conv_dbg', dbg)
  ctx.conv2d(xd, wp, bias, ks, yd, relu=not residual, residual=res, y_relu=yr)
  torch.cuda.synchronize()
  ctx.set_trace_buffer(None)
  ctx.set_option('conv_dbg', 0)
  if dbg: print('ablation switches', dbg)
  full = trace.cpu()
  t = full[:512].view(64, 8)
  base = int(t[0, 0])
  if dbg & 256:
    print("  latency probe (cycles): 32 dependent IMAD %d | 8 dependent param loads %d | 8 dependent LDS %d | clock-clock %d %d" %
          tuple(int(full[1024 + i]) for i in range(5)))
  if int(full[576]) > 0:
    print("  resident path, issuing thread, cycles per section: awaits | fence | issue (+ next plan) | commit a_empty | commit acc_full + loop")
    for i in range(12):
      r = [int(v) for v in t[i]]
      si, sc = int(full[512 + i]), int(full[576 + i])
      print("  %2d | %5d %5d %5d %5d %5d" % (i, r[1] - r[0], r[2] - r[1], si - r[2], sc - si, r[3] - sc))
    print("  single-chunk rows, cycles relative to the start: producer saw a_empty | TMA issued || watcher saw a_full | saw acc_empty || issuer enters row | awaits done")
    b0 = int(t[0, 0])
    for i in range(16):
      print("  %2d | %7d %7d || %7d %7d || %7d %7d" % (i, int(full[896 + i]) - b0, int(full[960 + i]) - b0, int(full[768 + i]) - b0,
                                                     int(full[832 + i]) - b0, int(t[i, 0]) - b0, int(t[i, 1]) - b0))
  else:
    print('  streamed path, cycles the issuing thread waited per group (weights | input rows + accumulators):',
          ' '.join('%d|%d' % (int(full[640 + i]), int(full[704 + i])) for i in range(12)))
  print("conv %dx%dx%d %d->%d k%d: rows = tile iteration; cycles relative to start" % (n, h, w, cin, cout, ks))
  print("  it | grp:enter     plan   a_full   issued | epi:enter  t_full    done | mma_busy epi_busy  period ")
  prev = None
  for i in range(12):
    r = [int(v) - base for v in t[i]]
    period = (r[3] - prev) if prev is not None else 0
    prev = r[3]
    print("  %2d | %9d %8d %8d %8d | %9d %8d %8d | %8d %8d %8d" % (i, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[3] - r[2], r[6] - r[5], period))

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
if which == "first":
  run(6, 1080, 1920, 32, 64, 3)
  run(6, 1080, 1920, 64, 64, 3)
  run(6, 1080, 1920, 64, 64, 3, dbg=6)
  run(6, 1080, 1920, 64, 64, 3, dbg=4)
  run(6, 1080, 1920, 64, 64, 3, dbg=2)
  run(6, 1080, 1920, 64, 64, 3, dbg=256)
  run(6, 1080, 1920, 64, 64, 3, dbg=256 + 6)
