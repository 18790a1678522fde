"""Per-row pipeline timeline of compose_rows_kernel (CTA 0): clock64 stamps of every role for rows 20..31 of the CTA's range,
relative to the first stamp.  Roles: epilogue WG of layer 1..4 (enter, accumulator ready, block re-zeroed + released, output
slot free, row published), head (enter, computed, slot free, published), UMMA issuer of layer 1..4 (enter, operand row ready,
accumulator blocks free, issued + committed)."""
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
n, h, w = 8, 1080, 1920
small = torch.randn(n, h // 2, w // 2, 3, device=dev)
large = torch.randn(n, h, w, 3, device=dev)
dst = torch.empty(n, h, w, 3, device=dev)
sd, ld, od = _lib.desc(small), _lib.desc(large), _lib.desc(dst)
for _ in range(3):
  ctx.compose_scales(sd, ld, blob, None, od)
trace = torch.zeros(9 * 64 * 8, dtype=torch.int64, device=dev)
ctx.set_trace_buffer(trace)
ctx.compose_scales(sd, ld, blob, None, od)
torch.cuda.synchronize()
ctx.set_trace_buffer(None)
t = trace.cpu().view(9, 64, 8).numpy()
base = int(t[4, 0, 0])
names = ["epi L1", "epi L2", "epi L3", "epi L4", "head  ", "mma L1", "mma L2", "mma L3", "mma L4"]
cols = {0: 5, 1: 5, 2: 5, 3: 3, 4: 4, 5: 4, 6: 4, 7: 4, 8: 4}
for role in range(9):
  print(names[role], "(cycles since the head's first stamp; last column = period of the final stamp)")
  prev = None
  for g in range(20, 32):
    r = [int(v) - base for v in t[role, g, :cols[role]]]
    last = r[-1]
    print("   row %2d: " % g + " ".join("%8d" % v for v in r) + "   | +" + " +".join("%d" % (r[i + 1] - r[i]) for i in range(len(r) - 1)) +
          ("   period %d" % (last - prev) if prev is not None else ""))
    prev = last
