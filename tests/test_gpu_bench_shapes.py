"""GPU parity AT THE BENCHMARK SHAPES (BASELINE configs[1]: 8 tuple passes of a 1920x1080 frame batched along N, 148
persistent CTAs, RowsWalker segments crossing image / strip boundaries, resident and streamed weights).

conv_rows_kernel is compared with oracle.torch_ops.conv2d_same / conv2d_transpose_same_s2 on the SAME fp16-rounded
operands on sampled row bands of the full tensor (bands at the image borders, around the boundaries of the CTAs' row
ranges, and at seeded random positions; always the full width, so every 128-pixel strip boundary is covered).  One full
1x1080x1920 single-tuple Architecture.predict is compared with the float32 restated reference.

Tolerances: fp16 outputs 2e-3 of the output scale (one fp16 rounding of the result + fp32 accumulation order; operands are
identical); the predict test states the measured bound next to the assertion."""
import numpy as np
import pytest
import torch

from deepdenoiser_b200 import _lib, synthetic
from deepdenoiser_b200.Architecture import Architecture
from oracle import reference_model, torch_ops

pytestmark = pytest.mark.gpu


def pad_bias(b):
  out = torch.zeros((b.shape[0] + 15) // 16 * 16)
  out[:b.shape[0]] = b
  return out.cuda()


def bands_for(n, h, w, sm_count, seed):
  """[(image, first row, one past last row)] bands of <= 6 rows to check."""
  rng = np.random.default_rng(seed)
  strips = (w + 127) // 128
  total = n * strips * h
  rows_per_cta = -(-total // min(sm_count, total))
  picks = {(0, 0), (n - 1, h - 6), (n // 2, h // 2)}
  # boundaries of the CTAs' contiguous row ranges, in the kernel's linear (image, strip, row) order
  for k in rng.choice(np.arange(1, max(2, total // rows_per_cta)), size=min(6, max(1, total // rows_per_cta - 1)), replace=False):
    lin = int(k) * rows_per_cta
    col, y = divmod(lin, h)
    picks.add((col // strips, max(0, min(h - 6, y - 3))))
  for _ in range(3):
    picks.add((int(rng.integers(n)), int(rng.integers(0, h - 6))))
  return [(i, y, min(h, y + 6)) for i, y in sorted(picks)]


def check_bands(x, y, bands, conv_slab, scale_rows=1, tol=2e-3, what=""):
  """conv_slab(slab [1,rows,W,C] float32 CPU) -> output rows of the slab (rows * scale_rows); rows whose receptive field
  leaves the slab (but not the image) are skipped."""
  h = x.shape[1]
  worst = 0.0
  for i, y0, y1 in bands:
    a, b = max(0, y0 - 1), min(h, y1 + 1)
    slab = x[i:i + 1, a:b].float().cpu()
    want = conv_slab(slab)
    lo = y0 - a
    want = want[:, lo * scale_rows:(lo + (y1 - y0)) * scale_rows]
    got = y[i:i + 1, y0 * scale_rows:y1 * scale_rows].float().cpu()
    scale = max(1.0, float(want.abs().max()))
    err = float((got - want).abs().max()) / scale
    worst = max(worst, err)
    assert err <= tol, "%s: image %d rows [%d,%d): max err %.3e > %.1e" % (what, i, y0, y1, err, tol)
  return worst


BENCH_CONVS = [
    # ks, cin, cout, n, h, w, cstride_in, coff_in    (the 3x3 layers of the U-Net of configs[1], 8 tuple passes per chunk)
    (3, 64, 64, 8, 1080, 1920, 64, 0),
    (3, 32, 64, 8, 1080, 1920, 32, 0),
    (3, 128, 64, 8, 1080, 1920, 128, 0),      # first conv after the skip concat
    (3, 96, 96, 8, 540, 960, 96, 0),
    (3, 192, 96, 8, 540, 960, 192, 0),
    (3, 64, 96, 8, 540, 960, 64, 0),
    (3, 128, 128, 8, 270, 480, 128, 0),
    (3, 96, 128, 8, 270, 480, 96, 0),
    (3, 64, 64, 1, 1080, 1920, 64, 0),        # the 17th tuple pass runs alone
    (3, 96, 96, 1, 540, 960, 96, 0),
    (1, 64, 25, 8, 1080, 1920, 64, 0),        # post-process 1x1 at full resolution
]


@pytest.mark.parametrize("ks,cin,cout,n,h,w,cs,coff", BENCH_CONVS)
def test_conv_rows_at_benchmark_shape(ctx, ks, cin, cout, n, h, w, cs, coff):
  g = torch.Generator(device="cuda").manual_seed(ks * 100000 + cin * 100 + cout + n)
  x = (torch.randn(n, h, w, cs, device="cuda", generator=g) * 0.5).half()
  k = torch.randn(ks, ks, cin, cout, generator=torch.Generator().manual_seed(cin + cout)) / float(np.sqrt(ks * ks * cin))
  b = torch.randn(cout, generator=torch.Generator().manual_seed(7)) * 0.1
  wp = ctx.pack_conv_weights(k, torch.float16)
  c8 = (cout + 7) // 8 * 8
  y = torch.empty(n, h, w, c8, dtype=torch.float16, device="cuda")
  ctx.conv2d(_lib.desc(x, cin, coff), wp, pad_bias(b), ks, _lib.desc(y, cout, 0), relu=True)
  torch.cuda.synchronize()
  k16 = k.half().float()

  def slab_conv(slab):
    return torch_ops.conv2d_same(slab[..., coff:coff + cin], k16, b, relu=True)

  worst = check_bands(x, y[..., :cout], bands_for(n, h, w, ctx.sm_count(), cin + cout), slab_conv, tol=2e-3,
                      what="conv %dx%d %d->%d @%dx%dx%d" % (ks, ks, cin, cout, n, h, w))
  print("conv %dx%d %d->%d @%dx%dx%d: worst band error %.2e" % (ks, ks, cin, cout, n, h, w, worst))


@pytest.mark.parametrize("cin,cout,n,h,w", [(96, 64, 8, 540, 960), (128, 96, 8, 270, 480), (96, 64, 1, 540, 960)])
def test_transpose2x2_at_benchmark_shape(ctx, cin, cout, n, h, w):
  g = torch.Generator(device="cuda").manual_seed(cin + cout + n)
  x = (torch.randn(n, h, w, cin, device="cuda", generator=g) * 0.5).half()
  k = torch.randn(2, 2, cout, cin, generator=torch.Generator().manual_seed(3)) / float(np.sqrt(cin))
  b = torch.randn(cout, generator=torch.Generator().manual_seed(7)) * 0.1
  wp = ctx.pack_conv_weights(k, torch.float16, transposed=True)
  # written into the second half of a skip-concat buffer, like UNet.py:91-92
  y = torch.full((n, 2 * h, 2 * w, 2 * cout), float("nan"), dtype=torch.float16, device="cuda")
  ctx.conv2d_transpose2x2(_lib.desc(x), wp, pad_bias(b), _lib.desc(y, cout, cout), relu=True)
  torch.cuda.synchronize()
  k16 = k.half().float()
  # a transposed 2x2 stride-2 convolution has no halo: every slab row is valid
  worst = 0.0
  for i, y0, y1 in bands_for(n, h, w, ctx.sm_count(), cin):
    want = torch_ops.conv2d_transpose_same_s2(x[i:i + 1, y0:y1].float().cpu(), k16, b, relu=True)
    got = y[i:i + 1, 2 * y0:2 * y1, :, cout:].float().cpu()
    err = float((got - want).abs().max()) / max(1.0, float(want.abs().max()))
    worst = max(worst, err)
    assert err <= 2e-3, (i, y0, err)
  assert bool(torch.isnan(y[0, :4, :, :cout].float()).all()), "wrote outside the channel window"
  print("T2x2 %d->%d @%dx%dx%d: worst band error %.2e" % (cin, cout, n, h, w, worst))


def single_tuple_json():
  """The benchmark network (U-Net [64,96,128]x4, K=5, 3 scales, 5 auxiliaries with variance) with ONE feature-prediction
  tuple, so the float32 CPU oracle of a full 1080p frame takes seconds."""
  j = synthetic.baseline_architecture_json("unet32")
  j["combined_features"] = {"Diffuse": {"Color": "", "Direct": "Diffuse Direct", "Indirect": ""}}
  j["architecture"]["source_encoder"]["feature_flag_mode"] = "NONE"
  return j


# Measured on B200 (round 2): fp16 storage max 2.3e-4 / mean 9e-7 of the output scale at 1080p - the asserted bounds are
# ~2x that; float16x2 (the split-fp16 tensor-core mode) must meet the north-star 1e-4 outright.  float32 (exact path) is not
# run at this size: it is the SIMT kernel, pinned on the small cases.
@pytest.mark.parametrize("dtype,tol_max,tol_mean", [("float16", 6e-4, 1e-5), ("float16x2", 1e-4, 5e-6)])
def test_predict_full_1080p_frame_single_tuple_vs_oracle(dtype, tol_max, tol_mean):
  j = single_tuple_json()
  host = Architecture(j)
  weights = synthetic.randomize_biases(host.weights)
  feats = synthetic.synthetic_features(host, 1, 1080, 1920, seed=1234)
  jj = dict(j)
  jj["b200"] = {"dtype": dtype}
  arch = Architecture(jj, weights=weights)
  out = arch.predict({k: torch.from_numpy(v) for k, v in feats.items()})
  torch.cuda.synchronize()
  with torch.no_grad():
    want = reference_model.Architecture(j, ops=torch_ops, dtype=torch.float32, weights=weights).predict(feats)
  assert len(out) == len(want) == 3
  for s in range(3):
    for key, w in want[s].items():
      w = w.numpy()
      got = out[s][key].float().cpu().numpy()
      assert got.shape == w.shape
      scale = max(1.0, float(np.abs(w).max()))
      err = np.abs(got - w) / scale
      print("%s scale %d %s: mean %.2e max %.2e" % (dtype, s, key, err.mean(), err.max()))
      assert float(err.max()) <= tol_max and float(err.mean()) <= tol_mean, (s, key, float(err.max()), float(err.mean()))
