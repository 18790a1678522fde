"""Training on the B200 path - what the reference does with tf.estimator + autodiff (Training.py:607-703, 853-877):
forward with saved activations, the loss of BaseFeatureTraining.loss / LossDifference.difference, the backward of
every op as explicit libdd_b200 kernels, TF-form Adam on one flat fp32 parameter buffer and (multi-GPU) one
all-reduce of the flat gradient buffer per step.

Two arithmetic modes:
  * precision="float32": the EXACT path (CUDA-core convolutions), U-Net and Tiramisu backbones; every gradient is
    parity-tested against torch-autograd of the oracle to 1e-6 (tests/test_gpu_training.py).
  * precision="float16" | "bfloat16": the TENSOR-CORE path: 16-bit activations and activation gradients, fp32 master weights /
    gradients / Adam state, fp32 image-level arithmetic (kernel-prediction apply, composition blend, loss).  Forward and
    input gradients run on conv_rows_kernel (the input gradient of a 3x3 conv is the conv of dz with flipped, channel-
    swapped weights), weight gradients on wgrad_rows_kernel (tcgen05, MN-major), the 2x2 transposed conv's backward as
    1x1 GEMMs on a space-to-depth view.  A static loss scale keeps the fp16 gradients in range; Adam divides it out.
Loss weights of TrainingExample.json (mean weights; variation / MS-SSIM / masked weights must be 0).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .Architecture import Architecture, FeaturePredictionTupleType, ModeKeys
from .FeatureFlags import FeatureFlagMode
from .Naming import Naming
from .network import V
from .RenderPasses import RenderPasses

LOSS_KINDS = {"DIFFERENCE": 0, "ABSOLUTE": 1, "SMOOTH_ABSOLUTE": 2, "SQUARED": 3, "SMAPE": 4}
_LIGHTS = ("Diffuse", "Glossy", "Subsurface", "Transmission")
_IMAGE_TERMS = ("Volume Direct", "Volume Indirect", "Emission", "Environment")
_b = ctypes.byref


def _fp(t):
  return ctypes.c_void_p(t.data_ptr())


def data_parallel_reduce(grad, loss, world_size, comm=None):
  """The ONE exchange step of data-parallel training: every rank holds the gradient of the mean loss over ITS
  tiles in one flat fp32 buffer; the buffers (and the scalar loss, for logging) are summed over the ranks in place and the
  returned factor 1/world turns the sum into the gradient of the mean over the GLOBAL batch (Training.py:128 takes a mean).
  Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
  if world_size <= 1:
    return 1.0
  if comm is not None:                         # libdd_b200's own NCCL communicator (dd_comm_allreduce_sum_f32)
    comm.all_reduce_sum(grad)
    comm.all_reduce_sum(loss)
  else:
    import torch.distributed as dist
    dist.all_reduce(grad)                     # one flat bucket: 1.7 M floats for the U-Net (latency bound)
    dist.all_reduce(loss)
  loss /= world_size
  return 1.0 / world_size


class TrainingSettings:
  """The part of TrainingExample.json the loss needs (Training.py:969-989, 1009-1203)."""

  def __init__(self, parsed_json=None):
    j = parsed_json or {}
    self.learning_rate = float(j.get("learning_rate", 1e-3))
    self.batch_size = int(j.get("batch_size", 8))
    self.loss_difference = j.get("loss_difference", "SMAPE")
    self.use_multiscale_loss = bool(j.get("use_multiscale_loss", True))

    def weights(block, default_mean):
      """(mean, variation, masked mean, ms_ssim) weights of one *_training_settings block (TrainingExample.json:31-98)."""
      b = j.get(block, {})
      lw, lm = b.get("loss_weights", {}), b.get("loss_weights_masked", {})
      if float(lm.get("ms_ssim", 0.0)) != 0.0:
        raise NotImplementedError("%s: masked MS-SSIM raises 'Not implemented' in the reference too (Training.py:206-207)" % block)
      if float(lm.get("variation", 0.0)) != 0.0:
        raise NotImplementedError("%s: the reference's masked variation loss multiplies a [N, h(w-1)+(h-1)w] tensor with an "
                                  "[N,h,w] mask (Training.py:146-149) and cannot run; not built" % block)
      return (float(lw.get("mean", default_mean)), float(lw.get("variation", 0.0)), float(lm.get("mean", 0.0)),
              float(lw.get("ms_ssim", 0.0)))

    (self.feature_weight, self.feature_variation_weight, self.feature_masked_weight,
     self.feature_ms_ssim_weight) = weights("features_training_settings", 1.0)
    (self.combined_feature_weight, self.combined_feature_variation_weight, self.combined_feature_masked_weight,
     self.combined_feature_ms_ssim_weight) = weights("combined_features_training_settings", 5.0)
    (self.combined_image_weight, self.combined_image_variation_weight, masked,
     self.combined_image_ms_ssim_weight) = weights("combined_image_training_settings", 10.0)
    if masked != 0.0:
      raise NotImplementedError("combined_image_training_settings: the combined image has no mask in the reference "
                                "(CombinedImageFeatureTraining.initialize, Training.py:475-495)")


class Trainer:
  """Owns the flat fp32 parameters / gradients / Adam state of an Architecture and runs training steps."""

  def __init__(self, architecture, settings=None, precision="float32", loss_scale=None):
    assert isinstance(architecture, Architecture)
    assert precision in ("float32", "float16", "bfloat16"), precision
    self.arch = architecture
    self.settings = settings or TrainingSettings()
    self.mixed = precision in ("float16", "bfloat16")
    self.act_dtype = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}[precision]
    self.pack_flag = _lib.DD_PACK_BF16 if precision == "bfloat16" else 0
    self.loss_scale = loss_scale                  # None: chosen per batch (N*H*W/8) in mixed mode, 1 in exact mode
    architecture.dtype = torch.float32            # the Architecture's own (inference) network is not used for training
    architecture.logits_dtype = torch.float32
    architecture._ensure_device()
    self.ctx = architecture.ctx
    self.dev = self.ctx.device
    self.spec = architecture.spec
    # flat parameter buffer in TF creation order
    self.offsets, n = {}, 0
    for name, shape in self.spec.variable_shapes():
      size = int(np.prod(shape))
      self.offsets[name] = (n, shape)
      n += (size + 3) // 4 * 4                    # keep every variable 16-byte aligned
    self.count = n
    if self.mixed:
      if any(f % 8 for f in self.spec.filters):
        raise _lib.DDError("16-bit training needs filter counts that are multiples of 8")
    # + slack: the tensor-core conv reads biases in groups of 16 floats
    self.theta = torch.zeros(n + 64, dtype=torch.float32, device=self.dev)
    self.grad = torch.zeros_like(self.theta)
    self.adam_m = torch.zeros_like(self.theta)
    self.adam_v = torch.zeros_like(self.theta)
    self.step_count = 0
    # overflow guard of the fp16 path (dd_adam_step_guarded): [skipped, flag, applied, lr_t bits]; lives on the device so the
    # step never synchronises with the host.  `loss_scale_factor` multiplies the per-batch default scale and is lowered by
    # update_loss_scale() after skipped steps.
    self.guard = torch.zeros(4, dtype=torch.int32, device=self.dev)
    self.loss_scale_factor = 1.0
    self._skipped_seen, self._clean_since = 0, 0
    self.set_weights(architecture.weights)
    self._buffers = {}
    self.loss_value = torch.zeros(1, dtype=torch.float32, device=self.dev)

  # ------------------------------------------------------------------------------------------ parameters
  def param(self, name):
    off, shape = self.offsets[name]
    return self.theta[off:off + int(np.prod(shape))].view(shape)

  def param_grad(self, name):
    off, shape = self.offsets[name]
    return self.grad[off:off + int(np.prod(shape))].view(shape)

  def set_weights(self, weights):
    for name, (off, shape) in self.offsets.items():
      w = torch.from_numpy(np.ascontiguousarray(weights[name], dtype=np.float32)).reshape(-1)
      self.theta[off:off + w.numel()].copy_(w)
    self._repack()

  def get_weights(self):
    return {name: self.param(name).detach().cpu().numpy().copy() for name in self.offsets}

  def gradients(self):
    """Gradients of the last backward() with the loss scale divided out."""
    inv = 1.0 / getattr(self, "_scale_used", 1.0)
    return {name: (self.param_grad(name).detach() * inv).cpu().numpy().copy() for name in self.offsets}

  def _repack(self):
    """fp32 master weights -> forward / input-gradient convolution layouts (device side) + host copies of the few
    weights that travel in launch parameters (compose head / tail)."""
    ctx = self.ctx
    if self.mixed:
      return self._repack_mixed()
    self.fwd, self.bwd, self.bias = {}, {}, {}
    for var in self.spec.conv_variables():
      w = self.param(var.kernel_name)
      self.bias[var.name] = self.param(var.bias_name)
      if var in self.spec.compose and var.ksize == 1:
        continue
      key = var.name
      if key not in getattr(self, "_packed_store", {}):
        if not hasattr(self, "_packed_store"):
          self._packed_store = {}
        self._packed_store[key] = (torch.empty(w.numel(), dtype=torch.float32, device=self.dev),
                                   torch.empty(w.numel(), dtype=torch.float32, device=self.dev))
      f, b = self._packed_store[key]
      if var.transposed and var.ksize == 3:
        # forward = 4 output phases, each an ordinary 3x3 'same' kernel holding a subset of the taps (network.py
        # _pack_transpose3x3); SIMT layout [tap][cout][cin] == TF conv2d_transpose [kh,kw,cout,cin] per tap
        ctx.call("dd_conv2d_repack_f32", _fp(w), 3, var.cin, var.cout, 1, None, _fp(b))
        k = w.view(3, 3, var.cout, var.cin)
        phases = []
        for py in range(2):
          for px in range(2):
            ph = torch.zeros_like(k)
            for dy in (0, -1):
              for dx in (0, -1):
                r, c = py - 2 * dy, px - 2 * dx
                if r <= 2 and c <= 2:
                  ph[dy + 1, dx + 1] = k[r, c]
            phases.append(ph.contiguous())
        self.fwd[key], self.bwd[key] = phases, b
        continue
      ctx.call("dd_conv2d_repack_f32", _fp(w), var.ksize, var.cin, var.cout, int(var.transposed), _fp(f), _fp(b))
      self.fwd[key], self.bwd[key] = f, b
    if self.spec.compose:
      head, tail = self.spec.compose[0], self.spec.compose[-1]
      self.host_small = {k: self.param(k).detach().cpu().contiguous() for k in
                         (head.kernel_name, head.bias_name, tail.kernel_name, tail.bias_name)}
    if self.arch.feature_flag_mode == FeatureFlagMode.EMBEDDING:
      self.arch._flags.embedding_matrix = self.param("embedding/feature_flags_embedding_matrix")

  def _repack_mixed(self):
    """fp32 master weights -> fp16 layouts of conv_rows_kernel, on the device (dd_conv2d_pack_weights_dev): forward, the
    input-gradient convolution (flipped taps, swapped channels) and, for the 2x2 transposed convs, the 1x1 GEMM of their
    backward on the space-to-depth view (TF [2,2,cout,cin] read as a [4*cout, cin] 1x1 kernel)."""
    ctx, lib = self.ctx, self.ctx.lib
    if not hasattr(self, "_packed_store"):
      self._packed_store = {}
    self.fwd, self.bwd, self.bias = {}, {}, {}
    for var in self.spec.conv_variables():
      w = self.param(var.kernel_name)
      self.bias[var.name] = self.param(var.bias_name)
      if var in self.spec.compose and var.ksize == 1:
        continue
      if var.transposed and var.ksize == 3:
        self._repack_transpose3x3(var, w)
        continue
      store = self._packed_store.get(var.name)
      if store is None:
        if var.transposed:
          assert var.ksize == 2
          nf = lib.dd_conv2d_packed_bytes(2, var.cin, var.cout, _lib.DD_F16, 1)
          nb = lib.dd_conv2d_packed_bytes(1, 4 * var.cout, var.cin, _lib.DD_F16, 0)
        else:
          nf = lib.dd_conv2d_packed_bytes(var.ksize, var.cin, var.cout, _lib.DD_F16, 0)
          nb = lib.dd_conv2d_packed_bytes(var.ksize, var.cout, var.cin, _lib.DD_F16, 0)
        store = (torch.zeros(nf, dtype=torch.uint8, device=self.dev), torch.zeros(nb, dtype=torch.uint8, device=self.dev))
        self._packed_store[var.name] = store
      f, b = store
      if var.transposed:
        ctx.call("dd_conv2d_pack_weights_dev", _fp(w), 2, var.cin, var.cout, 2 | self.pack_flag, _fp(f))
        ctx.call("dd_conv2d_pack_weights_dev", _fp(w), 1, 4 * var.cout, var.cin, 0 | self.pack_flag, _fp(b))
      else:
        ctx.call("dd_conv2d_pack_weights_dev", _fp(w), var.ksize, var.cin, var.cout, 0 | self.pack_flag, _fp(f))
        ctx.call("dd_conv2d_pack_weights_dev", _fp(w), var.ksize, var.cin, var.cout, 1 | self.pack_flag, _fp(b))
      self.fwd[var.name], self.bwd[var.name] = f, b
    if self.spec.compose:
      head, tail = self.spec.compose[0], self.spec.compose[-1]
      self.host_small = {k: self.param(k).detach().cpu().contiguous() for k in
                         (head.kernel_name, head.bias_name, tail.kernel_name, tail.bias_name)}
    if self.arch.feature_flag_mode == FeatureFlagMode.EMBEDDING:
      self.arch._flags.embedding_matrix = self.param("embedding/feature_flags_embedding_matrix")

  def _repack_transpose3x3(self, var, w):
    """3x3 stride-2 'SAME' transposed conv (Tiramisu.py:62-64), tensor-core mode.  w: TF [3,3,cout,cin] master weights.
    forward  = 4 output phases, each a 3x3 'same' conv holding a tap subset (network.py::_pack_transpose3x3);
    backward = ONE 3x3 'same' conv on the space-to-depth view Z[i,j,sp*cout+o] = dz[2i+ay,2j+ax,o] (sp = 2ay+ax):
               dx[i,j,c] = sum_{r,s,o} dz[2i+r,2j+s,o] W[r,s,o,c] = sum Z[i+r//2, j+s//2, (r%2,s%2), o] W[r,s,o,c], i.e. conv
               taps (1+r//2, 1+s//2) with input-channel block sp = 2(r%2)+(s%2); the other taps / blocks are zero."""
    ctx, lib = self.ctx, self.ctx.lib
    cin, cout = var.cin, var.cout
    k = w.view(3, 3, cout, cin)
    store = self._packed_store.get(var.name)
    if store is None:
      nf = lib.dd_conv2d_packed_bytes(3, cin, cout, _lib.DD_F16, 0)
      nb = lib.dd_conv2d_packed_bytes(3, 4 * cout, cin, _lib.DD_F16, 0)
      store = ([torch.zeros(nf, dtype=torch.uint8, device=self.dev) for _ in range(4)],
               torch.zeros(nb, dtype=torch.uint8, device=self.dev))
      self._packed_store[var.name] = store
    phases, bwd = store
    for py in range(2):
      for px in range(2):
        ph = torch.zeros(3, 3, cin, cout, dtype=torch.float32, device=self.dev)
        for dy in (0, -1):
          for dx in (0, -1):
            r, c = py - 2 * dy, px - 2 * dx
            if r <= 2 and c <= 2:
              ph[dy + 1, dx + 1] = k[r, c].t()
        ctx.call("dd_conv2d_pack_weights_dev", _fp(ph), 3, cin, cout, 0 | self.pack_flag, _fp(phases[py * 2 + px]))
    wc = torch.zeros(3, 3, 4 * cout, cin, dtype=torch.float32, device=self.dev)
    for r in range(3):
      for c in range(3):
        sp = 2 * (r % 2) + (c % 2)
        wc[1 + r // 2, 1 + c // 2, sp * cout:(sp + 1) * cout] = k[r, c]
    ctx.call("dd_conv2d_pack_weights_dev", _fp(wc), 3, 4 * cout, cin, 0 | self.pack_flag, _fp(bwd))
    self.fwd[var.name], self.bwd[var.name] = phases, bwd

  # ------------------------------------------------------------------------------------------ helpers
  def _buf(self, key, shape, zero=False, dtype=torch.float32):
    k = (key, tuple(shape), dtype)
    t = self._buffers.get(k)
    if t is None:
      t = torch.empty(shape, dtype=dtype, device=self.dev)
      self._buffers[k] = t
    if zero:
      t.zero_()
    return t

  def _act(self, key, nhw, c, zero=False):
    """Network activation / activation-gradient tensor [n,h,w,c] in the arithmetic mode's storage type; the channel
    stride is padded to 16 bytes in fp16 mode (TMA)."""
    cs = (c + 7) // 8 * 8 if self.mixed else c
    return V(self._buf(key, tuple(nhw) + (cs,), zero=zero, dtype=self.act_dtype), c)

  def _conv(self, var, x, y, relu=False, residual=None, y_relu=None):
    self.ctx.conv2d(x.d, self.fwd[var.name], self.bias[var.name], var.ksize, y.d, relu=relu,
                    residual=residual.d if residual is not None else None, y_relu=y_relu.d if y_relu is not None else None)

  def _conv_bwd(self, var, x, dz, dx=None, bias_done=False):
    """dW, db (accumulated) and optionally dx of y = conv(x, W) + b given dz = dL/dy.  bias_done: db was already
    accumulated by the _relu_bwd that produced dz (tensor-core mode fuses the two passes)."""
    ctx = self.ctx
    if self.mixed:
      ctx.call("dd_conv2d_wgrad_tc", _b(x.d), _b(dz.d), var.ksize, 0, _fp(self.param_grad(var.kernel_name)), ctypes.c_float(1.0))
      if not bias_done:
        ctx.call("dd_relu_bwd_bias", _b(dz.d), None, None, _fp(self.param_grad(var.bias_name)), ctypes.c_float(1.0))
    else:
      ctx.call("dd_conv2d_wgrad", _b(x.d), _b(dz.d), var.ksize, 0, _fp(self.param_grad(var.kernel_name)),
               _fp(self.param_grad(var.bias_name)))
    if dx is not None:
      ctx.conv2d(dz.d, self.bwd[var.name], None, var.ksize, dx.d, relu=False)

  def _relu_bwd(self, dy, y, dz, bias_of=None):
    """dz = dy * [y > 0]; tensor-core mode also accumulates the bias gradient of layer `bias_of` in the same pass
    (returns True when it did)."""
    if self.mixed:
      db = _fp(self.param_grad(bias_of.bias_name)) if bias_of is not None else None
      self.ctx.call("dd_relu_bwd_bias", _b(dy.d), _b(y.d), _b(dz.d), db, ctypes.c_float(1.0))
      return bias_of is not None
    self.ctx.call("dd_relu_bwd", _b(dy.d), _b(y.d), _b(dz.d))
    return False

  # ------------------------------------------------------------------------------------------ U-Net forward / backward
  def _unet_forward(self, x0):
    spec, f, steps, ctx = self.spec, self.spec.filters, self.spec.steps, self.ctx
    b, h, w = x0.t.shape[0], x0.t.shape[1], x0.t.shape[2]
    dims = [(h >> i, w >> i) for i in range(steps + 1)]
    tape = {"x0": x0, "blocks": [], "pools": [], "ups": [], "cats": []}

    def block(key, layers, x, out):
      acts = [x]
      for i, var in enumerate(layers):
        dst = out if i == len(layers) - 1 else self._act("%s.a%d" % (key, i), (b,) + tuple(x.t.shape[1:3]), var.cout)
        self._conv(var, acts[-1], dst, relu=True)
        acts.append(dst)
      tape["blocks"].append((key, layers, acts))
      return out

    x = x0
    for i in range(steps):
      hh, ww = dims[i]
      cat = self._act("cat%d" % i, (b, hh, ww), 2 * f[i]).t
      tape["cats"].append(cat)
      skip = V(cat, f[i], 0)
      block("d%d" % i, spec.down[i], x, skip)
      pooled = self._act("pool%d" % i, (b, dims[i + 1][0], dims[i + 1][1]), f[i])
      if self.mixed and f[i] % 8 == 0:
        # the forward pass records the first maximum of every window (one byte per window and channel): pure-gather backward
        idx = self._buf("pool%d.index" % i, (b, dims[i + 1][0], dims[i + 1][1], f[i]), dtype=torch.uint8)
        ctx.call("dd_maxpool_s2_fwd_index", _b(skip.d), 3, _b(pooled.d), _fp(idx))
      else:
        idx = None
        ctx.maxpool_s2(skip.d, 3, pooled.d)
      tape["pools"].append((skip, pooled, idx))
      x = pooled
    results = []
    for i in range(steps):
      index = steps - i
      hh, ww = dims[index]
      out = self._act("out%d" % index, (b, hh, ww), f[index])
      block("u%d" % index, spec.up[i], x, out)
      if spec.use_multiscale:
        results.append(out)
      var = spec.upsample[i]
      up = V(tape["cats"][index - 1], f[index - 1], f[index - 1])
      ctx.conv2d_transpose2x2(out.d, self.fwd[var.name], self.bias[var.name], up.d, relu=True)
      tape["ups"].append((var, out, up))
      x = V(tape["cats"][index - 1])
    out = self._act("out0", (b, h, w), f[0])
    block("l", spec.last, x, out)
    results.append(out)
    tape["results"] = results
    self._post_forward(tape)
    return tape

  def _post_forward(self, tape):
    """AdjustNumberOfChannels (Architecture.py:230-244): conv1x1 + ReLU, conv1x1 on every core output (coarsest first)."""
    spec = self.spec
    tape["post"] = []
    logits = []
    for k, (r, (a, bvar)) in enumerate(zip(tape["results"], spec.post)):
      bb, hh, ww = r.t.shape[0], r.t.shape[1], r.t.shape[2]
      mid = self._act("post.mid%d" % k, (bb, hh, ww), spec.output_channels)
      self._conv(a, r, mid, relu=True)
      o8 = (spec.output_channels + 7) // 8 * 8 if self.mixed else spec.output_channels
      out_l = V(self._buf("post.out%d" % k, (bb, hh, ww, o8)), spec.output_channels)     # logits stay fp32
      self._conv(bvar, mid, out_l, relu=False)
      tape["post"].append((a, bvar, r, mid, out_l))
      logits.append(out_l)
    tape["logits_coarse_first"] = logits

  def _post_backward(self, tape, dlogits_coarse_first):
    """Returns {id(core output buffer): gradient V} and accumulates the 1x1 weights' gradients."""
    dres = {}
    for k, ((a, bvar, r, mid, out_l), dl) in enumerate(zip(tape["post"], dlogits_coarse_first)):
      nhw = tuple(mid.t.shape[:3])
      dmid = self._act("post.dmid%d" % k, nhw, mid.c)
      self._conv_bwd(bvar, mid, dl, dmid)
      dz = self._act("post.dz%d" % k, nhw, mid.c)
      done = self._relu_bwd(dmid, mid, dz, bias_of=a)
      dr = self._act("post.dr%d" % k, nhw, a.cin)
      self._conv_bwd(a, r, dz, dr, bias_done=done)
      dres[id(r.t)] = dr
    return dres

  def _block_bwd(self, key, layers, acts, dout):
    """Backward through n x [conv + ReLU]; dout = dL/d(acts[-1]); returns dL/d(acts[0]) (fresh buffer)."""
    if self.mixed:
      return self._block_bwd_fused(key, layers, acts, dout)
    dy = dout
    for i in reversed(range(len(layers))):
      var, x, y = layers[i], acts[i], acts[i + 1]
      dz = self._act("%s.dz%d" % (key, i), tuple(y.t.shape[:3]), var.cout)
      done = self._relu_bwd(dy, y, dz, bias_of=var)
      dx = self._act("%s.dx%d" % (key, i), tuple(x.t.shape[:3]), var.cin)
      self._conv_bwd(var, x, dz, dx, bias_done=done)
      dy = dx
    return dy

  def _block_bwd_fused(self, key, layers, acts, dout):
    """Tensor-core mode: the input-gradient conv of layer i applies the ReLU mask of layer i-1 in its epilogue
    (DD_CONV_RESIDUAL_MASK), so dz_{i-1} is produced directly and only a read-only pass remains for the bias gradient."""
    ctx = self.ctx
    last = len(layers) - 1
    dz = self._act("%s.dz%d" % (key, last), tuple(acts[-1].t.shape[:3]), layers[last].cout)
    self._relu_bwd(dout, acts[-1], dz, bias_of=layers[last])
    for i in reversed(range(len(layers))):
      var, x = layers[i], acts[i]
      ctx.call("dd_conv2d_wgrad_tc", _b(x.d), _b(dz.d), var.ksize, 0, _fp(self.param_grad(var.kernel_name)), ctypes.c_float(1.0))
      if i > 0:
        dz_prev = self._act("%s.dz%d" % (key, i - 1), tuple(x.t.shape[:3]), var.cin)
        # input gradient + ReLU mask of layer i-1 + its bias gradient (column sums of dz_prev) in one launch
        ctx.conv2d(dz.d, self.bwd[var.name], None, var.ksize, dz_prev.d, residual=x.d, residual_is_mask=True,
                   colsum=self.param_grad(layers[i - 1].bias_name))
        dz = dz_prev
      else:
        dx = self._act("%s.dx0" % key, tuple(x.t.shape[:3]), var.cin)
        ctx.conv2d(dz.d, self.bwd[var.name], None, var.ksize, dx.d)
        return dx

  def _unet_backward(self, tape, dlogits_coarse_first):
    """Backward of _unet_forward.  Returns dL/dx0 (V over [B,H,W,C0])."""
    spec, f, steps, ctx = self.spec, self.spec.filters, self.spec.steps, self.ctx
    blocks = {key: (layers, acts) for key, layers, acts in tape["blocks"]}
    # 1x1 post-processing -> gradient of every core output
    dres = self._post_backward(tape, dlogits_coarse_first)
    # decoder: last block, then (transposed conv, block) pairs from fine to coarse
    layers, acts = blocks["l"]
    dcat = {0: self._block_bwd("l", layers, acts, dres[id(tape["results"][-1].t)])}    # dL/d cat_0, both halves
    dpool = None
    ups = {steps - i: tape["ups"][i] for i in range(steps)}
    for index in range(1, steps + 1):
      level = index - 1
      var, out, up = ups[index]
      dup = V(dcat[level].t, f[level], f[level])                     # [f:] half of the concat gradient
      dout = self._act("up%d.dout" % index, tuple(out.t.shape[:3]), out.c)
      if self.mixed:
        # ReLU mask + space-to-depth: the stride-2 2x2 transposed conv's backward becomes 1x1 GEMMs on the coarse grid
        s2d = self._act("up%d.s2d" % index, tuple(out.t.shape[:3]), 4 * f[level])
        ctx.call("dd_space_to_depth2_mask", _b(dup.d), _b(up.d), _b(s2d.d))
        db4 = self._buf("up%d.db4" % index, (4, f[level]), zero=True)
        ctx.call("dd_relu_bwd_bias", _b(s2d.d), None, None, _fp(db4), ctypes.c_float(1.0))
        self.param_grad(var.bias_name).add_(db4.sum(dim=0))
        ctx.call("dd_conv2d_wgrad_tc", _b(out.d), _b(s2d.d), 1, 1, _fp(self.param_grad(var.kernel_name)), ctypes.c_float(1.0))
        ctx.conv2d(s2d.d, self.bwd[var.name], None, 1, dout.d, relu=False)
      else:
        dz = V(self._buf("up%d.dz" % index, tuple(up.t.shape[:3]) + (f[level],)))
        self._relu_bwd(dup, up, dz)
        ctx.call("dd_conv2d_wgrad", _b(out.d), _b(dz.d), 2, 1, _fp(self.param_grad(var.kernel_name)),
                 _fp(self.param_grad(var.bias_name)))
        ctx.call("dd_conv2d_transpose2x2_dgrad", _b(dz.d), _fp(self.bwd[var.name]), _b(dout.d))
      extra = dres.get(id(out.t))                                    # multi-scale output taken from this block
      if extra is not None:
        ctx.call("dd_axpy", ctypes.c_float(1.0), _b(extra.d), _b(dout.d))
      layers, acts = blocks["u%d" % index]
      dx = self._block_bwd("u%d" % index, layers, acts, dout)
      if index == steps:
        dpool = dx                                                   # the deepest block consumed pool_{steps-1}
      else:
        dcat[index] = dx                                             # block u_index consumed cat_index
    # encoder: pooling + down blocks from coarse to fine
    for i in reversed(range(steps)):
      skip, pooled, idx = tape["pools"][i]
      dskip = V(dcat[i].t, f[i], 0)                                  # [:f] half, written by the concat consumer
      if idx is not None:
        # gather by the recorded index: adds into the 16-bit concat gradient in place (no atomics, no fp32 scratch tensor)
        ctx.call("dd_maxpool_s2_bwd_index", _fp(idx), _b(dpool.d), 3, _b(dskip.d))
      elif self.mixed:
        ctx.call("dd_maxpool_s2_bwd_acc", _b(skip.d), _b(pooled.d), _b(dpool.d), 3, _b(dskip.d))
      else:
        ctx.call("dd_maxpool_s2_bwd", _b(skip.d), _b(pooled.d), _b(dpool.d), 3, _b(dskip.d))
      layers, acts = blocks["d%d" % i]
      dpool = self._block_bwd("d%d" % i, layers, acts, dskip)
    return dpool                                                     # dL/dx0

  # ------------------------------------------------------------------------------------------ Tiramisu forward / backward
  # Same buffer plan as the inference path (network.py::_forward_tiramisu): per level one raw and one ReLU'd buffer whose
  # channel windows ARE the concatenations (Tiramisu.py:40,104).  The backward keeps one gradient buffer per level with the
  # same channel layout; a consumer that read relu(x) adds  dact * [raw > 0]  into it (dd_relu_bwd_acc), so by the time a
  # layer's output window is used as dz every later reader has contributed.
  def _tiramisu_forward(self, x0):
    spec, f, steps, ctx = self.spec, self.spec.filters, self.spec.steps, self.ctx
    b, h, w = x0.t.shape[0], x0.t.shape[1], x0.t.shape[2]
    dims = [(h >> i, w >> i) for i in range(steps + 1)]
    n = spec.convs_per_block
    totals = {steps: spec.skip_channels[steps - 1] + n * f[steps]}
    for level in range(steps):
      totals[level] = spec.skip_channels[level] + f[level] + n * f[level]
    raws = {l: self._buf("tira.raw%d" % l, (b,) + dims[l] + (totals[l],), dtype=self.act_dtype) for l in totals}
    acts = {l: self._buf("tira.act%d" % l, (b,) + dims[l] + (totals[l],), dtype=self.act_dtype) for l in totals}
    tape = {"x0": x0, "raws": raws, "acts": acts, "dims": dims, "totals": totals, "trans": [], "blocks": []}

    def dense(layers, level, c):
      for var in layers:
        self._conv(var, V(acts[level], c, 0), V(raws[level], var.cout, c), relu=False, y_relu=V(acts[level], var.cout, c))
        c += var.cout
      return c

    self._conv(spec.pre, x0, V(raws[0], f[0], 0), relu=True, y_relu=V(acts[0], f[0], 0))
    c = f[0]
    for i in range(steps):
      c0 = c
      c = dense(spec.down[i], i, c)
      tape["blocks"].append(("down", i, i, spec.down[i], c0))
      z = self._buf("tira.z%d" % i, (b,) + dims[i] + (c,), dtype=self.act_dtype)
      za = self._buf("tira.za%d" % i, (b,) + dims[i] + (c,), dtype=self.act_dtype)
      self._conv(spec.transition[i], V(acts[i], c, 0), V(z), relu=False, y_relu=V(za))
      ctx.maxpool_s2(V(z).d, 2, V(raws[i + 1], c, 0).d)
      ctx.maxpool_s2(V(za).d, 2, V(acts[i + 1], c, 0).d)
      tape["trans"].append((i, c, z))
    results, ups = [], []
    for i in range(steps):
      index = steps - i
      c0 = c
      c = dense(spec.up[i], index, c)
      tape["blocks"].append(("up", i, index, spec.up[i], c0))
      if spec.use_multiscale:
        results.append(V(raws[index], c, 0))
      var, level = spec.upsample[i], index - 1
      cs = spec.skip_channels[level]
      ctx.conv2d_transpose3x3(V(raws[index], c, 0).d, self.fwd[var.name], self.bias[var.name], V(raws[level], var.cout, cs).d,
                              V(acts[level], var.cout, cs).d, relu=True)
      ups.append((var, index, c, level, cs))
      c = cs + var.cout
    c0 = c
    c = dense(spec.last, 0, c)
    tape["blocks"].append(("last", 0, 0, spec.last, c0))
    results.append(V(raws[0], c, 0))
    tape["results"], tape["ups"] = results, ups
    self._post_forward(tape)
    return tape

  def _tiramisu_backward(self, tape, dlogits_coarse_first):
    spec, f, steps, ctx = self.spec, self.spec.filters, self.spec.steps, self.ctx
    raws, acts, dims, totals = tape["raws"], tape["acts"], tape["dims"], tape["totals"]
    b = tape["x0"].t.shape[0]
    adt = self.act_dtype
    G = {l: self._buf("tira.g%d" % l, (b,) + dims[l] + (totals[l],), zero=True, dtype=adt) for l in totals}
    dres = self._post_backward(tape, dlogits_coarse_first)
    for r in tape["results"]:
      level = [l for l in raws if raws[l] is r.t][0]
      ctx.call("dd_axpy", ctypes.c_float(1.0), _b(dres[id(r.t)].d), _b(V(G[level], r.c, 0).d))

    def dense_bwd(layers, level, c0):
      c = c0 + sum(v.cout for v in layers)
      for var in reversed(layers):
        c -= var.cout
        dz = V(G[level], var.cout, c)                               # complete: every later reader already added
        dact = V(self._buf("tira.dact%d" % level, (b,) + dims[level] + (totals[level],), dtype=adt), c, 0)
        self._conv_bwd(var, V(acts[level], c, 0), dz, dact)
        ctx.call("dd_relu_bwd_acc", _b(dact.d), _b(V(raws[level], c, 0).d), _b(V(G[level], c, 0).d))

    blocks = {(kind, i): (level, layers, c0) for kind, i, level, layers, c0 in tape["blocks"]}
    level, layers, c0 = blocks[("last", 0)]
    dense_bwd(layers, level, c0)
    for i in reversed(range(steps)):
      var, index, c, level, cs = tape["ups"][i]
      # y = relu(conv2d_transpose(raw_index[0:c])) written to raw_level[cs:cs+f]
      x = V(raws[index], c, 0)
      dx = V(self._buf("tira.updx%d" % index, (b,) + dims[index] + (c,), dtype=adt))
      if self.mixed:
        # ReLU mask + space-to-depth of the fine gradient, then 3x3 'same' convs on the coarse grid (_repack_transpose3x3)
        s2d = self._act("tira.s2d%d" % index, (b,) + dims[index], 4 * var.cout)
        ctx.call("dd_space_to_depth2_mask", _b(V(G[level], var.cout, cs).d), _b(V(raws[level], var.cout, cs).d), _b(s2d.d))
        db4 = self._buf("tira.db4_%d" % index, (4, var.cout), zero=True)
        ctx.call("dd_relu_bwd_bias", _b(s2d.d), None, None, _fp(db4), ctypes.c_float(1.0))
        self.param_grad(var.bias_name).add_(db4.sum(dim=0))
        # dWc[tr,ts,(sp,o),c] = sum Z[i+tr-1, j+ts-1, (sp,o)] x[i,j,c]; W[r,s,o,c] sits at tr = 1+r//2, ts = 1+s//2, sp = 2(r%2)+(s%2)
        dwc = self._buf("tira.dwc%d" % index, (3, 3, 4 * var.cout, c), zero=True)
        ctx.call("dd_conv2d_wgrad_tc", _b(s2d.d), _b(x.d), 3, 0, _fp(dwc), ctypes.c_float(1.0))
        dw = self.param_grad(var.kernel_name)                      # [3,3,cout,cin]
        for r in range(3):
          for cc in range(3):
            sp = 2 * (r % 2) + (cc % 2)
            dw[r, cc].add_(dwc[1 + r // 2, 1 + cc // 2, sp * var.cout:(sp + 1) * var.cout])
        ctx.conv2d(s2d.d, self.bwd[var.name], None, 3, dx.d, relu=False)
      else:
        dz = V(self._buf("tira.updz%d" % level, (b,) + dims[level] + (var.cout,)))
        self._relu_bwd(V(G[level], var.cout, cs), V(raws[level], var.cout, cs), dz)
        ctx.call("dd_conv2d_wgrad", _b(x.d), _b(dz.d), 3, 1, _fp(self.param_grad(var.kernel_name)),
                 _fp(self.param_grad(var.bias_name)))
        ctx.call("dd_conv2d_transpose3x3_dgrad", _b(dz.d), _fp(self.bwd[var.name]), _b(dx.d))
      ctx.call("dd_axpy", ctypes.c_float(1.0), _b(dx.d), _b(V(G[index], c, 0).d))
      lvl, layers, c0 = blocks[("up", i)]
      dense_bwd(layers, lvl, c0)
    for i in reversed(range(steps)):
      _, c, z = tape["trans"][i]
      # raw_{i+1}[0:c] = maxpool2(z), z = conv1x1(act_i[0:c])
      dzp = V(self._buf("tira.dzp%d" % i, (b,) + dims[i] + (c,), zero=True))   # dd_maxpool_s2_bwd accumulates (fp32 atomics)
      ctx.call("dd_maxpool_s2_bwd", _b(V(z).d), _b(V(raws[i + 1], c, 0).d), _b(V(G[i + 1], c, 0).d), 2, _b(dzp.d))
      if self.mixed:
        dzp16 = V(self._buf("tira.dzp16_%d" % i, (b,) + dims[i] + (c,), dtype=adt))
        ctx.cast_copy(dzp.d, dzp16.d)
        dzp = dzp16
      var = spec.transition[i]
      dact = V(self._buf("tira.dact%d" % i, (b,) + dims[i] + (totals[i],), dtype=adt), c, 0)
      self._conv_bwd(var, V(acts[i], c, 0), dzp, dact)
      ctx.call("dd_relu_bwd_acc", _b(dact.d), _b(V(raws[i], c, 0).d), _b(V(G[i], c, 0).d))
      lvl, layers, c0 = blocks[("down", i)]
      dense_bwd(layers, lvl, c0)
    # pre-processing conv: raw_0[0:f0] = relu(conv(x0))
    dz = self._act("tira.predz", (b,) + dims[0], f[0])
    done = self._relu_bwd(V(G[0], f[0], 0), V(raws[0], f[0], 0), dz, bias_of=spec.pre)
    dx0 = self._act("tira.dx0", tuple(tape["x0"].t.shape[:3]), tape["x0"].c)
    self._conv_bwd(spec.pre, tape["x0"], dz, dx0, bias_done=done)
    return dx0

  # ------------------------------------------------------------------------------------------ compose net
  def _compose_forward(self, key, small, large, out):
    """MultiScalePrediction.compose_scales with every intermediate kept for the backward pass."""
    spec, ctx = self.spec, self.ctx
    head, c1, c2, c3, c4, tail = spec.compose
    i, h, w = large.t.shape[0], large.t.shape[1], large.t.shape[2]
    t = {n: self._act("%s.%s" % (key, n), (i, h, w), 24) for n in ("x0", "a1", "x1", "a2", "a3", "x2")}
    hs = self.host_small
    ctx.compose_head(small.d, large.d, hs[head.kernel_name].reshape(6, 24), hs[head.bias_name], 24, t["x0"].d)
    self._conv(c1, t["x0"], t["a1"], relu=True)                              # a1 = relu(r1)
    self._conv(c2, t["a1"], t["x1"], residual=t["x0"], y_relu=t["a2"])       # x1 = x0 + r2, a2 = relu(x1)
    self._conv(c3, t["a2"], t["a3"], relu=True)                              # a3 = relu(r3)
    self._conv(c4, t["a3"], t["x2"], residual=t["x1"])                       # x2 = x1 + r4
    ctx.compose_tail(t["x2"].d, hs[tail.kernel_name].reshape(24), hs[tail.bias_name], 24, small.d, large.d, None, out.d)
    t.update(small=small, large=large, key=key)
    return t

  def _compose_backward(self, t, dout, dsmall, dlarge):
    """dsmall / dlarge (fp32 image banks) are accumulated; compose weights' gradients are accumulated."""
    spec, ctx = self.spec, self.ctx
    head, c1, c2, c3, c4, tail = spec.compose
    key = t["key"]
    shape = tuple(t["x0"].t.shape)
    hs = self.host_small
    g = lambda n: self._act("%s.d%s" % (key, n), shape[:3], 24)   # noqa: E731
    dx2 = g("x2")
    ctx.call("dd_compose_tail_bwd", _b(t["x2"].d), _fp(hs[tail.kernel_name]), _fp(hs[tail.bias_name]), 24, _b(t["small"].d),
             _b(t["large"].d), _b(dout.d), _b(dx2.d), _b(dsmall.d), _b(dlarge.d), _fp(self.param_grad(tail.kernel_name)),
             _fp(self.param_grad(tail.bias_name)))
    # block 2: x2 = x1 + conv4(a3), a3 = relu(conv3(a2)), a2 = relu(x1)
    da3 = g("a3")
    self._conv_bwd(c4, t["a3"], dx2, da3)
    dz3 = g("z3")
    done = self._relu_bwd(da3, t["a3"], dz3, bias_of=c3)
    da2 = g("a2")
    self._conv_bwd(c3, t["a2"], dz3, da2, bias_done=done)
    dx1 = g("x1")
    self._relu_bwd(da2, t["a2"], dx1)                                         # through relu(x1)
    ctx.call("dd_axpy", ctypes.c_float(1.0), _b(dx2.d), _b(dx1.d))            # + identity path
    # block 1: x1 = x0 + conv2(a1), a1 = relu(conv1(x0)) (x0 >= 0 is already a ReLU output)
    da1 = g("a1")
    self._conv_bwd(c2, t["a1"], dx1, da1)
    dz1 = g("z1")
    done = self._relu_bwd(da1, t["a1"], dz1, bias_of=c1)
    dx0 = g("x0")
    self._conv_bwd(c1, t["x0"], dz1, dx0, bias_done=done)
    ctx.call("dd_axpy", ctypes.c_float(1.0), _b(dx1.d), _b(dx0.d))
    ctx.call("dd_compose_head_bwd", _b(t["small"].d), _b(t["large"].d), _fp(hs[head.kernel_name]), 24, _b(t["x0"].d), _b(dx0.d),
             _b(dsmall.d), _b(dlarge.d), _fp(self.param_grad(head.kernel_name)), _fp(self.param_grad(head.bias_name)))

  # ------------------------------------------------------------------------------------------ forward (all tuples, one chunk)
  def forward(self, features):
    """Architecture.predict for training: same arithmetic as the inference path, one chunk, everything kept."""
    arch, ctx = self.arch, self.ctx
    targets, every = arch.feature_predictions, arch.feature_predictions + arch.auxiliary_features
    sources = [arch._as_device(features[Naming.source_feature_name(fp.name, index=0)]) for fp in every]
    n, h, w = sources[0].shape[0], sources[0].shape[1], sources[0].shape[2]
    n_scales = (self.spec.steps + 1) if arch.use_multiscale_predictions else 1
    if h % (1 << self.spec.steps) or w % (1 << self.spec.steps):
      raise ValueError("height and width must be divisible by %d" % (1 << self.spec.steps))
    var_width = max([fp.feature_variance.channels(fp.number_of_channels) for fp in every] + [0])
    std_bank = self._buf("bank.std", (len(every) * n, h, w, 3))
    var_bank = self._buf("bank.var", (len(every) * n, h, w, max(var_width, 1)))
    raw_bank = self._buf("bank.raw", (len(targets) * n, h, w, 3)) if arch._preserve_source else None
    for fp, s in zip(every, sources):
      lo, hi = fp.bank_index * n, (fp.bank_index + 1) * n
      vc = fp.feature_variance.channels(fp.number_of_channels)
      ctx.standardize_variance(_lib.desc(s), arch._std_params(fp), _lib.desc(std_bank[lo:hi]),
                               _lib.desc(var_bank[lo:hi], vc, 0) if vc else None)
      if arch._preserve_source and fp.is_target:
        ctx.standardize_variance(_lib.desc(s), _lib.dd_standardize_params(0, 0.0, 1.0, 0, 0, 0, 0, 0, 1e-4),
                                 _lib.desc(raw_bank[lo:hi]), None)
    nt = len(targets) * n
    kp_full = raw_bank if arch._preserve_source else std_bank[:nt]
    kp_sources = [kp_full]
    for s in range(1, n_scales):
      pooled = self._buf("bank.kpsrc%d" % s, (nt, h >> s, w >> s, 3))
      if arch.use_kernel_prediction:
        ctx.avgpool(_lib.desc(kp_full), 1 << s, _lib.desc(pooled))
      kp_sources.append(pooled)
    c0 = arch.number_of_input_channels
    table, keep = arch._gather_table(std_bank, var_bank, max(var_width, 1), features, n, c0)
    tuples, ft = arch.feature_prediction_tuples, arch.features_per_tuple
    x0v = self._act("net.x0", (len(tuples) * n, h, w), c0)
    x0 = x0v.t
    ctx.assemble_input(table, len(tuples), n, x0v.d)
    tape = self._unet_forward(x0v) if self.spec.core_name == "U-Net" else self._tiramisu_forward(x0v)
    logits = list(tape["logits_coarse_first"])
    if arch.use_multiscale_predictions:
      logits.reverse()                                   # largest first (Architecture.py:577-579)
    kp = []
    for s in range(n_scales):
      dst = self._buf("kp.out%d" % s, (nt, h >> s, w >> s, 3))
      if arch.use_kernel_prediction:
        ctx.kernel_predict(_lib.desc(kp_sources[s]), logits[s].d, arch.kernel_size, ft, n, _lib.desc(dst))
      else:
        # direct prediction (Architecture.py:519-521): the post-processed tensor is split into 3 channels per feature
        arch._split_direct(logits[s], dst, len(tuples), n)
      kp.append(dst)
    # multi-scale composition / inverse standardisation; `pre` = values fed to the inversion (needed by its backward)
    stage = list(kp)
    pre_inv = [None] * n_scales
    inv_first = not arch.invert_standardization_after_multiscale_predictions
    if inv_first:
      for s in range(n_scales):
        pre_inv[s] = stage[s]
        inv = self._buf("inv.out%d" % s, tuple(stage[s].shape))
        inv.copy_(stage[s])
        arch._invert(targets, inv, n)
        stage[s] = inv
    compose_tapes = [None] * n_scales
    composed_in = list(stage)                            # `large` operands (before composition)
    for s in range(n_scales - 1, 0, -1):
      out = self._buf("cmp.out%d" % (s - 1), tuple(stage[s - 1].shape))
      compose_tapes[s - 1] = self._compose_forward("cmp%d" % (s - 1), V(stage[s]), V(stage[s - 1]), V(out))
      stage[s - 1] = out
    finals = []
    for s in range(n_scales):
      if inv_first:
        finals.append(stage[s])
      else:
        pre_inv[s] = stage[s]
        fin = self._buf("final%d" % s, tuple(stage[s].shape))
        fin.copy_(stage[s])
        arch._invert(targets, fin, n)
        finals.append(fin)
    del keep
    self._state = dict(tape=tape, logits=logits, kp=kp, kp_sources=kp_sources, compose_tapes=compose_tapes, pre_inv=pre_inv,
                       finals=finals, n=n, h=h, w=w, n_scales=n_scales, x0=x0, std_bank=std_bank, inv_first=inv_first)
    return finals

  def predictions(self):
    """Prediction dictionaries of the last forward() (same structure as Architecture.predict)."""
    st, arch = self._state, self.arch
    out = []
    for s in range(st["n_scales"]):
      d = {}
      for fp in arch.feature_predictions:
        lo, hi = fp.bank_index * st["n"], (fp.bank_index + 1) * st["n"]
        p = st["finals"][s][lo:hi] if fp.load_data else st["std_bank"][lo:hi, :st["h"] >> s, :st["w"] >> s, :]
        d[Naming.feature_prediction_name(fp.name)] = p[..., :1] if fp.number_of_channels == 1 else p
      out.append(d)
    return out

  # ------------------------------------------------------------------------------------------ loss + its gradient
  def loss_and_gradient(self, targets_dict):
    """model_fn's loss (Training.py:611-660) on the last forward(); fills dL/d(final predictions) and returns the
    device scalar.  targets_dict: {'target_image/<Pass>': [N,H,W,C]}."""
    arch, ctx, st, cfg = self.arch, self.ctx, self._state, self.settings
    n, h, w, n_scales = st["n"], st["h"], st["w"], st["n_scales"]
    kind = LOSS_KINDS[cfg.loss_difference]
    # static loss scale of the fp16 path: the per-pixel gradient of a mean over N*H*W pixels would underflow fp16
    S = float(self.loss_scale) if self.loss_scale is not None else (n * h * w / 8.0 if self.act_dtype == torch.float16 else 1.0)
    if self.act_dtype == torch.float16:
      S *= self.loss_scale_factor
    self._scale_used = S
    loss_scales = n_scales if cfg.use_multiscale_loss else 1
    norm = 1.0 / sum(1.0 / 4.0 ** s for s in range(loss_scales))
    self.loss_value.zero_()
    dfin = [self._buf("dfinal%d" % s, tuple(st["finals"][s].shape), zero=True) for s in range(n_scales)]
    loaded = [fp for fp in arch.feature_predictions if fp.is_target and fp.load_data]
    by_name = {fp.name: fp for fp in arch.feature_predictions}
    # targets at every loss scale (Training.py:616-623): avg-pool 2^s of the labels
    tgt = {}
    for fp in loaded:
      full = arch._as_device(targets_dict[Naming.target_feature_name(fp.name)])
      tgt[fp.name] = [full]
      for s in range(1, loss_scales):
        t = self._buf("tgt.%s.%d" % (fp.name, s), (n, h >> s, w >> s, full.shape[3]))
        ctx.avgpool(_lib.desc(full), 1 << s, _lib.desc(t))
        tgt[fp.name].append(t)

    def pred_view(bank, fp, channels=None):
      lo, hi = fp.bank_index * n, (fp.bank_index + 1) * n
      return _lib.desc(bank[lo:hi], channels or fp.number_of_channels, 0)

    def mask_pass(name):
      """FeatureTraining.initialize (Training.py:374-392): the pass whose target defines the mask of feature `name`."""
      if RenderPasses.is_color_render_pass(name) or name in (RenderPasses.ENVIRONMENT, RenderPasses.EMISSION,
                                                               RenderPasses.VOLUME_DIRECT, RenderPasses.VOLUME_INDIRECT):
        return name
      if RenderPasses.is_direct_or_indirect_render_pass(name):
        return RenderPasses.direct_or_indirect_to_color_render_pass(name)
      return None

    def ms_ssim_term(pred_d, tgt_d, grad_d, weight, what):
      """weight * (1 - mean(tf.image.ssim_multiscale(pred, target, 1, (0.0448, 0.2856, 0.3001)))) on the full-resolution
      prediction (BaseFeatureTraining.ms_ssim / loss, Training.py:188-204, 229-230) and its gradient, accumulated into grad_d."""
      powers = (0.0448, 0.2856, 0.3001)
      c = pred_d.c
      if c != 3:
        raise NotImplementedError("MS-SSIM of '%s': %d-channel passes take the reference's channels_first detour "
                                  "(Training.py:197-199); only RGB passes are built" % (what, c))
      if (h % 4) or (w % 4) or (h >> 2) < 11 or (w >> 2) < 11:
        raise ValueError("MS-SSIM needs tiles divisible by 4 with at least 44 pixels per side (11x11 filter at the third scale)")
      xs, ys = [pred_d], [tgt_d]
      for k in (1, 2):
        xk = self._buf("msssim.x.%s.%d" % (what, k), (n, h >> k, w >> k, c))
        yk = self._buf("msssim.y.%s.%d" % (what, k), (n, h >> k, w >> k, c))
        ctx.avgpool(xs[-1], 2, _lib.desc(xk))
        ctx.avgpool(ys[-1], 2, _lib.desc(yk))
        xs.append(_lib.desc(xk)); ys.append(_lib.desc(yk))
      stats, sums = [], []
      for k in range(3):
        hk, wk = (h >> k) - 10, (w >> k) - 10
        st_k = _lib.desc(self._buf("msssim.stats.%s.%d" % (what, k), (n, hk, wk, 4 * c)))
        sm_k = self._buf("msssim.sums.%s.%d" % (what, k), (n, c, 2), zero=True)
        ctx.call("dd_ssim_stats", _b(xs[k]), _b(ys[k]), _b(st_k))
        ctx.call("dd_ssim_reduce", _b(st_k), c, ctypes.c_float(1.0), _fp(sm_k))
        stats.append(st_k); sums.append(sm_k / float(hk * wk))
      # per (image, channel, level) algebra on [n, c, 3] device tensors
      v = torch.stack([sums[0][..., 0], sums[1][..., 0], sums[2][..., 1]], dim=-1)
      m = torch.relu(v)
      pw = torch.tensor(powers, dtype=torch.float32, device=self.dev)
      ms = torch.prod(torch.pow(m, pw), dim=-1)
      self.loss_value += S * weight * (1.0 - ms.mean())
      dms = -S * weight / float(n * c)
      dv = torch.where(m > 0, dms * pw * ms.unsqueeze(-1) / m.clamp_min(1e-30), torch.zeros_like(m))     # [n, c, 3]
      zero = torch.zeros_like(dv[..., 0])
      coefs = [torch.stack([dv[..., 0], zero], dim=-1).contiguous(), torch.stack([dv[..., 1], zero], dim=-1).contiguous(),
               torch.stack([zero, dv[..., 2]], dim=-1).contiguous()]
      d2 = self._buf("msssim.d2.%s" % what, (n, h >> 2, w >> 2, c), zero=True)
      d1 = self._buf("msssim.d1.%s" % what, (n, h >> 1, w >> 1, c), zero=True)
      ctx.call("dd_ssim_bwd", _b(xs[2]), _b(ys[2]), _b(stats[2]), _fp(coefs[2]), ctypes.c_float(1.0), _b(_lib.desc(d2)))
      ctx.call("dd_ssim_bwd", _b(xs[1]), _b(ys[1]), _b(stats[1]), _fp(coefs[1]), ctypes.c_float(1.0), _b(_lib.desc(d1)))
      ctx.call("dd_avgpool2_adjoint", _b(_lib.desc(d2)), _b(_lib.desc(d1)))
      ctx.call("dd_ssim_bwd", _b(xs[0]), _b(ys[0]), _b(stats[0]), _fp(coefs[0]), ctypes.c_float(1.0), _b(grad_d))
      ctx.call("dd_avgpool2_adjoint", _b(_lib.desc(d1)), _b(grad_d))

    def extra_terms(pred_d, tgt_d, grad_d, s, variation_weight, masked_weight, mask_name, what):
      """variation_mean / masked_mean of one feature at scale s (BaseFeatureTraining.loss, Training.py:226-240); the gradient is
      accumulated into grad_d."""
      hs, ws = h >> s, w >> s
      factor_s = norm / 4.0 ** s
      if variation_weight > 0:
        count = float(n * (hs * (ws - 1) + (hs - 1) * ws))
        ctx.call("dd_loss_variation_fwd_bwd", _b(pred_d), _b(tgt_d), kind, ctypes.c_float(S * variation_weight * factor_s / count),
                 ctypes.c_float(1e-2), _fp(self.loss_value), _b(grad_d))
      if masked_weight > 0:
        if mask_name is None or mask_name not in tgt:
          raise Exception("Masking is not supported for '%s': no corresponding colour target (Training.py:102-113, 380-392)" % what)
        mask_d = _lib.desc(tgt[mask_name][s])
        msum = self._buf("mask.sum.%s.%d" % (what, s), (1,), zero=True)
        ctx.call("dd_mask_sum", _b(mask_d), _fp(msum))
        ctx.call("dd_loss_masked_fwd_bwd", _b(pred_d), _b(tgt_d), _b(mask_d), _fp(msum), kind,
                 ctypes.c_float(S * masked_weight * factor_s), ctypes.c_float(1e-2), _fp(self.loss_value), _b(grad_d))

    for s in range(loss_scales):
      px = float(n * (h >> s) * (w >> s))
      factor = norm / 4.0 ** s
      # single-feature losses (FeatureTraining, weight from features_training_settings)
      if cfg.feature_weight > 0:
        for fp in loaded:
          ctx.call("dd_loss_fwd_bwd", _b(pred_view(st["finals"][s], fp)), _b(_lib.desc(tgt[fp.name][s])), kind,
                   ctypes.c_float(S * cfg.feature_weight * factor / px), ctypes.c_float(1e-2), _fp(self.loss_value),
                   _b(pred_view(dfin[s], fp)), 1)
      if cfg.feature_variation_weight > 0 or cfg.feature_masked_weight > 0:
        for fp in loaded:
          extra_terms(pred_view(st["finals"][s], fp), _lib.desc(tgt[fp.name][s]), pred_view(dfin[s], fp), s,
                      cfg.feature_variation_weight, cfg.feature_masked_weight, mask_pass(fp.name), fp.name)
      if cfg.feature_ms_ssim_weight > 0 and s == 0:
        for fp in loaded:
          ms_ssim_term(pred_view(st["finals"][0], fp), _lib.desc(tgt[fp.name][0]), pred_view(dfin[0], fp),
                       cfg.feature_ms_ssim_weight, fp.name)
      # combined lighting passes color * (direct + indirect) and the combined image (sum of everything)
      lights = [l for l in _LIGHTS if all((l + k) in by_name and by_name[l + k].load_data for k in (" Color", " Direct", " Indirect"))]
      image_terms = [t for t in _IMAGE_TERMS if t in by_name and by_name[t].load_data]
      use_image = ((cfg.combined_image_weight > 0 or cfg.combined_image_variation_weight > 0 or
                    cfg.combined_image_ms_ssim_weight > 0) and len(lights) == 4 and len(image_terms) == 4)
      wants_image = (cfg.combined_image_weight > 0 or cfg.combined_image_variation_weight > 0 or
                     cfg.combined_image_ms_ssim_weight > 0)
      if wants_image and not use_image and not getattr(self, "_warned_image", False):
        # the reference would fail with a KeyError here (CombinedImageFeatureTraining.initialize, Training.py:475-495 reads all
        # 16 passes); architectures that load fewer passes train without the combined-image term - said once, not silently
        import warnings
        warnings.warn("combined_image_training_settings has non-zero weights but the architecture does not load the 4 lighting "
                      "triples + Emission / Environment / Volume passes: the combined-image loss term is omitted")
        self._warned_image = True
      use_combined = (cfg.combined_feature_weight > 0 or cfg.combined_feature_variation_weight > 0 or
                      cfg.combined_feature_masked_weight > 0 or cfg.combined_feature_ms_ssim_weight > 0)
      shape = (n, h >> s, w >> s, 3)
      g_img = None
      if use_image:
        img_p = self._buf("img.p%d" % s, shape, zero=True)
        img_t = self._buf("img.t%d" % s, shape, zero=True)
        g_img = self._buf("img.g%d" % s, shape)
      comb = {}
      if lights and (use_combined or use_image):
        for l in lights:
          c, d, i = (by_name[l + k] for k in (" Color", " Direct", " Indirect"))
          cp = self._buf("cmb.p.%s.%d" % (l, s), shape)
          ct = self._buf("cmb.t.%s.%d" % (l, s), shape)
          ctx.call("dd_muladd_fwd", _b(pred_view(st["finals"][s], c)), _b(pred_view(st["finals"][s], d)),
                   _b(pred_view(st["finals"][s], i)), _b(_lib.desc(cp)))
          ctx.call("dd_muladd_fwd", _b(_lib.desc(tgt[c.name][s])), _b(_lib.desc(tgt[d.name][s])), _b(_lib.desc(tgt[i.name][s])),
                   _b(_lib.desc(ct)))
          comb[l] = (cp, ct, self._buf("cmb.g.%s.%d" % (l, s), shape, zero=True))
          if use_image:
            ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(cp)), _b(_lib.desc(img_p)))
            ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(ct)), _b(_lib.desc(img_t)))
      if use_image:
        for t in image_terms:
          fp = by_name[t]
          ctx.call("dd_axpy", ctypes.c_float(1.0), _b(pred_view(st["finals"][s], fp)), _b(_lib.desc(img_p)))
          ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(tgt[t][s])), _b(_lib.desc(img_t)))
        ctx.call("dd_loss_fwd_bwd", _b(_lib.desc(img_p)), _b(_lib.desc(img_t)), kind,
                 ctypes.c_float(S * cfg.combined_image_weight * factor / px), ctypes.c_float(1e-2), _fp(self.loss_value),
                 _b(_lib.desc(g_img)), 0)
        extra_terms(_lib.desc(img_p), _lib.desc(img_t), _lib.desc(g_img), s, cfg.combined_image_variation_weight, 0.0, None,
                    "Combined")
        if cfg.combined_image_ms_ssim_weight > 0 and s == 0:
          ms_ssim_term(_lib.desc(img_p), _lib.desc(img_t), _lib.desc(g_img), cfg.combined_image_ms_ssim_weight, "Combined")
        for t in image_terms:
          ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(g_img)), _b(pred_view(dfin[s], by_name[t])))
      for l, (cp, ct, gc) in comb.items():
        if cfg.combined_feature_weight > 0:
          ctx.call("dd_loss_fwd_bwd", _b(_lib.desc(cp)), _b(_lib.desc(ct)), kind,
                   ctypes.c_float(S * cfg.combined_feature_weight * factor / px), ctypes.c_float(1e-2), _fp(self.loss_value),
                   _b(_lib.desc(gc)), 1)
        extra_terms(_lib.desc(cp), _lib.desc(ct), _lib.desc(gc), s, cfg.combined_feature_variation_weight,
                    cfg.combined_feature_masked_weight, RenderPasses.combined_to_color_render_pass(l), l)
        if cfg.combined_feature_ms_ssim_weight > 0 and s == 0:
          ms_ssim_term(_lib.desc(cp), _lib.desc(ct), _lib.desc(gc), cfg.combined_feature_ms_ssim_weight, l)
        if use_image:
          ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(g_img)), _b(_lib.desc(gc)))
        c, d, i = (by_name[l + k] for k in (" Color", " Direct", " Indirect"))
        inc = self._buf("cmb.inc%d" % s, shape)
        ctx.call("dd_muladd_bwd", _b(pred_view(st["finals"][s], c)), _b(pred_view(st["finals"][s], d)),
                 _b(pred_view(st["finals"][s], i)), _b(_lib.desc(gc)), _b(pred_view(dfin[s], c)), _b(_lib.desc(inc)))
        ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(inc)), _b(pred_view(dfin[s], d)))
        ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(inc)), _b(pred_view(dfin[s], i)))
      # COMBINED tuples: Training.main() builds a CombinedFeatureTraining for EVERY tuple (Training.py:1142-1171), also for the
      # ones with generated members (Alpha, Emission, Environment: generated Direct / Indirect; Volume: generated Color).  A
      # generated member's "prediction" is the top-left crop of its standardised source (Architecture.py:151-157), its target
      # the loader's constant (Training.py:538-549: 1 for Color, 0.5 for Direct / Indirect); no gradient flows into it.
      if use_combined and arch.feature_prediction_tuple_type == FeaturePredictionTupleType.COMBINED:
        hs, ws = h >> s, w >> s
        for tup in arch.feature_prediction_tuples:
          members = list(tup.feature_predictions)
          if len(members) != 3 or all(m.load_data for m in members) or not any(m.load_data for m in members):
            continue                                     # fully loaded tuples are the lights above
          tag = "%s.%d" % (tup.name, s)
          P, T, G = [], [], []                           # 3-channel predictions / targets / gradient accumulators per member
          for index, m in enumerate(members):
            pb = self._buf("ctup.p%d.%s" % (index, tag), shape)
            tb = self._buf("ctup.t%d.%s" % (index, tag), shape)
            if m.load_data:
              pv, tv = pred_view(st["finals"][s], m), _lib.desc(tgt[m.name][s])
              if m.number_of_channels == 3:
                ctx.cast_copy(pv, _lib.desc(pb)); ctx.cast_copy(tv, _lib.desc(tb))
              else:                                      # Alpha: one channel broadcast against the 3-channel generated members
                for ch in range(3):
                  ctx.cast_copy(pv, _lib.desc(pb, 1, ch)); ctx.cast_copy(tv, _lib.desc(tb, 1, ch))
            else:
              lo = m.bank_index * n
              pb.copy_(st["std_bank"][lo:lo + n, :hs, :ws, :])
              ctx.call("dd_fill", ctypes.c_float(1.0 if index == 0 else 0.5), _b(_lib.desc(tb)))
            P.append(pb); T.append(tb)
            G.append(self._buf("ctup.g%d.%s" % (index, tag), shape, zero=True))
          cp = self._buf("ctup.cp.%s" % tag, shape)
          ct = self._buf("ctup.ct.%s" % tag, shape)
          gc = self._buf("ctup.gc.%s" % tag, shape, zero=True)
          ctx.call("dd_muladd_fwd", _b(_lib.desc(P[0])), _b(_lib.desc(P[1])), _b(_lib.desc(P[2])), _b(_lib.desc(cp)))
          ctx.call("dd_muladd_fwd", _b(_lib.desc(T[0])), _b(_lib.desc(T[1])), _b(_lib.desc(T[2])), _b(_lib.desc(ct)))
          if cfg.combined_feature_weight > 0:
            ctx.call("dd_loss_fwd_bwd", _b(_lib.desc(cp)), _b(_lib.desc(ct)), kind,
                     ctypes.c_float(S * cfg.combined_feature_weight * factor / px), ctypes.c_float(1e-2), _fp(self.loss_value),
                     _b(_lib.desc(gc)), 1)
          if cfg.combined_feature_variation_weight > 0 or cfg.combined_feature_masked_weight > 0:
            # mask = non-zero mask of the tuple's colour target (RenderPasses.combined_to_color_render_pass)
            color_pass = tup.name if tup.name in (RenderPasses.ALPHA, RenderPasses.EMISSION, RenderPasses.ENVIRONMENT) else tup.name + " Color"
            mask_t = next((T[i] for i, m in enumerate(members) if m.name == color_pass), None)
            if cfg.combined_feature_variation_weight > 0:
              count = float(n * (hs * (ws - 1) + (hs - 1) * ws))
              ctx.call("dd_loss_variation_fwd_bwd", _b(_lib.desc(cp)), _b(_lib.desc(ct)), kind,
                       ctypes.c_float(S * cfg.combined_feature_variation_weight * factor / count), ctypes.c_float(1e-2),
                       _fp(self.loss_value), _b(_lib.desc(gc)))
            if cfg.combined_feature_masked_weight > 0:
              if mask_t is None:
                raise Exception("Masking is not supported for '%s': no corresponding colour target" % tup.name)
              msum = self._buf("mask.sum.ctup.%s" % tag, (1,), zero=True)
              ctx.call("dd_mask_sum", _b(_lib.desc(mask_t)), _fp(msum))
              ctx.call("dd_loss_masked_fwd_bwd", _b(_lib.desc(cp)), _b(_lib.desc(ct)), _b(_lib.desc(mask_t)), _fp(msum), kind,
                       ctypes.c_float(S * cfg.combined_feature_masked_weight * factor), ctypes.c_float(1e-2), _fp(self.loss_value),
                       _b(_lib.desc(gc)))
          if cfg.combined_feature_ms_ssim_weight > 0 and s == 0:
            ms_ssim_term(_lib.desc(cp), _lib.desc(ct), _lib.desc(gc), cfg.combined_feature_ms_ssim_weight, tup.name)
          inc = self._buf("ctup.inc.%s" % tag, shape)
          ctx.call("dd_muladd_bwd", _b(_lib.desc(P[0])), _b(_lib.desc(P[1])), _b(_lib.desc(P[2])), _b(_lib.desc(gc)),
                   _b(_lib.desc(G[0])), _b(_lib.desc(inc)))
          for index, m in enumerate(members):
            if not m.load_data:
              continue
            g_m = G[0] if index == 0 else inc
            if m.number_of_channels == 3:
              ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(g_m)), _b(pred_view(dfin[s], m)))
            else:
              for ch in range(3):
                ctx.call("dd_axpy", ctypes.c_float(1.0), _b(_lib.desc(g_m, 1, ch)), _b(pred_view(dfin[s], m)))
    self._dfinal = dfin
    if S != 1.0:
      self.loss_value.div_(S)
    return self.loss_value

  # ------------------------------------------------------------------------------------------ backward of everything
  def backward(self, accumulate=False):
    """Fills self.grad with the gradient of the last loss_and_gradient(); `accumulate` adds to what is there (every
    parameter-gradient kernel accumulates), which is how train_step runs micro-batches."""
    arch, ctx, st = self.arch, self.ctx, self._state
    n, n_scales = st["n"], st["n_scales"]
    targets = arch.feature_predictions
    if not accumulate:
      self.grad.zero_()
    g = list(self._dfinal)                       # dL/d finals

    def invert_bwd(dy, x, s, tag):
      """through prediction_invert_standardization, one launch per run of passes with identical parameters"""
      out = self._buf("ginv.%s%d" % (tag, s), tuple(dy.shape))
      out.copy_(dy)
      i = 0
      while i < len(targets):
        fp = targets[i]
        j = i + 1
        while (j < len(targets) and targets[j].invert_standardization == fp.invert_standardization and
               targets[j].feature_standardization.key() == fp.feature_standardization.key()):
          j += 1
        stdz = fp.feature_standardization
        if fp.invert_standardization and stdz.key() != (False, 0.0, 1.0):
          ctx.call("dd_invert_standardization_bwd", _b(_lib.desc(dy[i * n:j * n])), _b(_lib.desc(x[i * n:j * n])),
                   _b(stdz.invert_params()), _b(_lib.desc(out[i * n:j * n])))
        i = j
      return out

    if not st["inv_first"]:
      g = [invert_bwd(g[s], st["pre_inv"][s], s, "a") for s in range(n_scales)]
    # composition chain: composed_{s} = compose(small = composed_{s+1}, large = stage_in_s)
    dlarge = [None] * n_scales
    for s in range(n_scales - 1):
      t = st["compose_tapes"][s]
      dl = self._buf("dlarge%d" % s, tuple(g[s].shape), zero=True)
      gs_next = self._buf("gsmall%d" % (s + 1), tuple(g[s + 1].shape))
      gs_next.copy_(g[s + 1])
      self._compose_backward(t, V(g[s]), V(gs_next), V(dl))
      g[s + 1] = gs_next
      dlarge[s] = dl
    dlarge[n_scales - 1] = g[n_scales - 1]
    if st["inv_first"]:
      dlarge = [invert_bwd(dlarge[s], st["pre_inv"][s], s, "b") for s in range(n_scales)]
    # kernel prediction -> logits gradients (largest first), then the network
    dlogits = []
    for s in range(n_scales):
      dl = self._act("dlogits%d" % s, tuple(st["logits"][s].t.shape[:3]), st["logits"][s].c)
      if arch.use_kernel_prediction:
        ctx.call("dd_kernel_predict_bwd", _b(_lib.desc(st["kp_sources"][s])), _b(st["logits"][s].d), _b(_lib.desc(dlarge[s])),
                 arch.kernel_size, arch.features_per_tuple, n, _b(dl.d))
      else:
        # adjoint of the per-feature split: feature f of tuple t owns channels [3f, 3f+3) of the tuple's images
        ft = arch.features_per_tuple
        for t in range(len(arch.feature_prediction_tuples)):
          for f in range(ft):
            g_tf = dlarge[s][(t * ft + f) * n:(t * ft + f + 1) * n]
            ctx.cast_copy(_lib.desc(g_tf), V(dl.t[t * n:(t + 1) * n], 3, dl.coff + 3 * f).d)
      dlogits.append(dl)
    if arch.use_multiscale_predictions:
      dlogits.reverse()                           # coarsest first, the order of the core outputs
    if self.spec.core_name == "U-Net":
      dx0 = self._unet_backward(st["tape"], dlogits)
    else:
      dx0 = self._tiramisu_backward(st["tape"], dlogits)
    # embedding rows: sum of the input gradient over the pixels of each tuple's images
    if arch.feature_flag_mode == FeatureFlagMode.EMBEDDING:
      dim = arch._flags.embedding_dimension
      first = arch.number_of_input_channels - dim
      tuples = arch.feature_prediction_tuples
      sums = self._buf("emb.sums", (len(tuples), dim), zero=True)
      ctx.call("dd_channel_sum", _b(_lib.desc(dx0.t, dim, first)), len(tuples), _fp(sums))
      rows = torch.tensor([arch._flags.index(t.name) for t in tuples], device=self.dev)
      self.param_grad("embedding/feature_flags_embedding_matrix").index_add_(0, rows, sums)
    return self.grad

  # ------------------------------------------------------------------------------------------ one optimizer step
  def train_step(self, features, targets_dict, world_size=1, comm=None, micro_batch=None):
    """forward + loss + backward + (all-reduce) + Adam.  Returns the loss as a device scalar.

    `micro_batch`: tiles per forward/backward pass.  The batch (dim 0 of every [N,H,W,C] entry) is cut into N/micro_batch
    equal micro-batches whose gradients accumulate in the flat fp32 buffer before the ONE all-reduce and the ONE optimizer
    step, so the result is the step of the whole batch (the loss is a mean, Training.py:128) with the activation memory of
    a micro-batch: batch 128 x 256^2 x 32 ch (TrainingExample.json:15 scaled to BASELINE configs[4]) fits one GPU."""
    n = next(iter(targets_dict.values())).shape[0]
    if micro_batch is None or micro_batch >= n:
      self.forward(features)
      loss = self.loss_and_gradient(targets_dict)
      self.backward()
      parts = 1
    else:
      if n % micro_batch:
        raise ValueError("batch of %d tiles is not a multiple of the micro-batch %d" % (n, micro_batch))
      parts = n // micro_batch
      total = torch.zeros(1, dtype=torch.float32, device=self.dev)

      def cut(d, lo):
        return {k: (v[lo:lo + micro_batch] if v.shape[0] == n else v) for k, v in d.items()}

      for m in range(parts):
        self.forward(cut(features, m * micro_batch))
        total += self.loss_and_gradient(cut(targets_dict, m * micro_batch))
        self.backward(accumulate=m > 0)
      loss = total / parts
    scale = data_parallel_reduce(self.grad, loss, world_size, comm)
    self.apply_gradients(scale / parts)
    return loss

  def apply_gradients(self, scale=1.0):
    """TF-form Adam on the flat buffers (the loss scale of the fp16 path is divided out here) + weight repack.  The fp16
    path uses the guarded step: a gradient holding inf / NaN skips the update on the device (no host sync)."""
    scale /= getattr(self, "_scale_used", 1.0)
    self.step_count += 1
    if self.act_dtype == torch.float16:
      self.ctx.call("dd_adam_step_guarded", _fp(self.theta), _fp(self.grad), _fp(self.adam_m), _fp(self.adam_v),
                    ctypes.c_size_t(self.count), ctypes.c_float(self.settings.learning_rate), ctypes.c_float(0.9),
                    ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(scale), _fp(self.guard))
    else:
      self.ctx.call("dd_adam_step", _fp(self.theta), _fp(self.grad), _fp(self.adam_m), _fp(self.adam_v),
                    ctypes.c_size_t(self.count), ctypes.c_float(self.settings.learning_rate), ctypes.c_float(0.9),
                    ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_int64(self.step_count), ctypes.c_float(scale))
    self._repack()

  def update_loss_scale(self, growth_interval=500):
    """Dynamic part of the fp16 loss scale; call it wherever the host synchronises anyway (logging / checkpoints).  Reads
    the number of optimizer steps the guard skipped: any new one halves the scale factor, `growth_interval` clean steps
    double it again (never above the static default).  Returns (skipped steps in total, current factor)."""
    if self.act_dtype != torch.float16:
      return 0, 1.0
    skipped = int(self.guard[0].item())
    if skipped > self._skipped_seen:
      self.loss_scale_factor *= 0.5 ** (skipped - self._skipped_seen)
      self._skipped_seen, self._clean_since = skipped, self.step_count
    elif self.step_count - self._clean_since >= growth_interval and self.loss_scale_factor < 1.0:
      self.loss_scale_factor = min(1.0, self.loss_scale_factor * 2.0)
      self._clean_since = self.step_count
    return skipped, self.loss_scale_factor

  def applied_steps(self):
    """Optimizer steps actually applied (the guarded fp16 step may skip some)."""
    return int(self.guard[2].item()) if self.act_dtype == torch.float16 else self.step_count

  # ------------------------------------------------------------------------------------------ checkpoints
  def save_checkpoint(self, path, keep=5):
    """ckpt-<step>.npz written atomically (temp file + os.replace, so an interrupted write never leaves a truncated
    'latest' checkpoint); refuses non-finite weights; keeps the newest `keep` ckpt-*.npz of the directory."""
    if not bool(torch.isfinite(self.theta[:self.count]).all().item()):
      raise _lib.DDError("refusing to checkpoint non-finite weights (step %d)" % self.step_count)
    state = {"step": np.array(self.applied_steps())}
    for name in self.offsets:
      off, shape = self.offsets[name]
      size = int(np.prod(shape))
      state[name] = self.theta[off:off + size].view(shape).cpu().numpy()
      state["adam_m/" + name] = self.adam_m[off:off + size].view(shape).cpu().numpy()
      state["adam_v/" + name] = self.adam_v[off:off + size].view(shape).cpu().numpy()
    path = str(path)
    if not path.endswith(".npz"):
      path += ".npz"
    tmp = path + ".tmp.%d" % os.getpid()
    with open(tmp, "wb") as f:
      np.savez(f, **state)
      f.flush()
      os.fsync(f.fileno())
    os.replace(tmp, path)
    directory, base = os.path.split(path)
    if keep and base.startswith("ckpt-"):
      found = []
      for name in os.listdir(directory or "."):
        if name.startswith("ckpt-") and name.endswith(".npz"):
          try:
            found.append((int(name[5:-4]), name))
          except ValueError:
            pass
      for _, name in sorted(found)[:-keep]:
        os.remove(os.path.join(directory or ".", name))

  def load_checkpoint(self, path):
    z = np.load(path)
    self.step_count = int(z["step"])
    self.guard.zero_()
    self.guard[2] = self.step_count
    for name, (off, shape) in self.offsets.items():
      size = int(np.prod(shape))
      self.theta[off:off + size].copy_(torch.from_numpy(z[name]).reshape(-1))
      self.adam_m[off:off + size].copy_(torch.from_numpy(z["adam_m/" + name]).reshape(-1))
      self.adam_v[off:off + size].copy_(torch.from_numpy(z["adam_v/" + name]).reshape(-1))
    self._repack()
