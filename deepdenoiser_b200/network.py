"""The convolutional part of the denoiser as a sequence of libdd_b200 calls.

What the reference builds out of tf.layers.* calls (UNet.py:61-99, Tiramisu.py:67-111,
Architecture.AdjustNumberOfChannels :230-244, MultiScalePrediction._compose_scales_neural_network :57-93) is
expressed here as

  * a *variable list* in TensorFlow's creation order with TF-1.x auto names ("reused_core_architecture/conv2d_7/kernel",
    ...) and TF layouts, so a weight dictionary is interchangeable with the oracle (and a converted TF checkpoint), and
  * an *executor* that walks the same structure and issues C-ABI calls on NHWC buffers.  Concatenations never
    copy: producers write straight into a channel window of the wider buffer (dd_tensor.coff / cstride).

There is no CPU fallback: everything below requires a Context (a CUDA device and the built shared library).
"""
import numpy as np
import torch

from . import _lib


def _round_up(v, m):
  return (v + m - 1) // m * m


class ConvVariable:
  """One tf.layers.conv2d / conv2d_transpose layer: kernel + bias (Appendix A.1 / A.5 layouts)."""

  def __init__(self, name, ksize, cin, cout, transposed=False):
    self.name, self.ksize, self.cin, self.cout, self.transposed = name, ksize, cin, cout, transposed

  @property
  def kernel_name(self):
    return self.name + "/kernel"

  @property
  def bias_name(self):
    return self.name + "/bias"

  @property
  def kernel_shape(self):
    if self.transposed:
      return (self.ksize, self.ksize, self.cout, self.cin)
    return (self.ksize, self.ksize, self.cin, self.cout)

  def glorot_limit(self):
    # keras glorot_uniform on the kernel shape: fan_in = shape[-2] * receptive, fan_out = shape[-1] * receptive
    shape = self.kernel_shape
    receptive = shape[0] * shape[1]
    return float(np.sqrt(6.0 / (receptive * (shape[2] + shape[3]))))


class _Namer:
  """tf.layers default naming inside one variable scope: conv2d, conv2d_1, ... / conv2d_transpose, ..."""

  def __init__(self, scope):
    self.scope, self.counts = scope, {}

  def __call__(self, default_name):
    i = self.counts.get(default_name, 0)
    self.counts[default_name] = i + 1
    return "%s/%s" % (self.scope, default_name if i == 0 else "%s_%d" % (default_name, i))


class NetworkSpec:
  """Static description (no device state) of the network of one architecture JSON."""

  def __init__(self, core_name, filters, convs_per_block, input_channels, output_channels, use_multiscale,
               embedding_shape=None):
    assert core_name in ("U-Net", "Tiramisu"), core_name
    self.core_name = core_name
    self.filters = list(filters)
    self.convs_per_block = int(convs_per_block)
    self.input_channels = int(input_channels)
    self.output_channels = int(output_channels)
    self.use_multiscale = bool(use_multiscale)
    self.embedding_shape = embedding_shape
    self.steps = len(self.filters) - 1
    name = _Namer("reused_core_architecture")
    self.core = []          # in creation order
    self.scale_channels = []  # channels of the core outputs, coarsest first (order of `results` in the reference)
    if core_name == "U-Net":
      self._build_unet(name)
    else:
      self._build_tiramisu(name)
    if not self.use_multiscale:
      self.scale_channels = self.scale_channels[-1:]
    self.post = []          # [(conv1, conv2)] per scale, coarsest first (Architecture.py:573-575)
    for c in self.scale_channels:
      self.post.append((ConvVariable(name("conv2d"), 1, c, self.output_channels),
                        ConvVariable(name("conv2d"), 1, self.output_channels, self.output_channels)))
    cname = _Namer("reused_compose_scales")
    self.compose = []
    if self.use_multiscale and self.steps > 0:
      self.compose = ([ConvVariable(cname("conv2d"), 1, 6, 24)] +
                      [ConvVariable(cname("conv2d"), 3, 24, 24) for _ in range(4)] +
                      [ConvVariable(cname("conv2d"), 1, 24, 1)])

  # -- structure ---------------------------------------------------------------------------------------
  def _block(self, name, cin, cout):
    layers = []
    for _ in range(self.convs_per_block):
      layers.append(ConvVariable(name("conv2d"), 3, cin, cout))
      cin = cout
    return layers

  def _build_unet(self, name):
    f, steps = self.filters, self.steps
    self.down, self.up, self.upsample = [], [], []
    cin = self.input_channels
    for i in range(steps):
      blk = self._block(name, cin, f[i])
      self.down.append(blk)
      self.core += blk
      cin = f[i]
    for i in range(steps):
      index = steps - i
      blk = self._block(name, cin, f[index])
      self.up.append(blk)
      self.core += blk
      self.scale_channels.append(f[index])
      t = ConvVariable(name("conv2d_transpose"), 2, f[index], f[index - 1], transposed=True)
      self.upsample.append(t)
      self.core.append(t)
      cin = 2 * f[index - 1]
    self.last = self._block(name, cin, f[0])
    self.core += self.last
    self.scale_channels.append(f[0])

  def _dense_block(self, name, cin, growth):
    layers = []
    for _ in range(self.convs_per_block):
      layers.append(ConvVariable(name("conv2d"), 3, cin, growth))
      cin += growth
    return layers, cin

  def _build_tiramisu(self, name):
    f, steps = self.filters, self.steps
    self.pre = ConvVariable(name("conv2d"), 3, self.input_channels, f[0])
    self.core.append(self.pre)
    self.down, self.transition, self.up, self.upsample = [], [], [], []
    self.skip_channels = []
    c = f[0]
    for i in range(steps):
      blk, c = self._dense_block(name, c, f[i])
      self.down.append(blk)
      self.core += blk
      self.skip_channels.append(c)
      t = ConvVariable(name("conv2d"), 1, c, c)
      self.transition.append(t)
      self.core.append(t)
    for i in range(steps):
      index = steps - i
      blk, c = self._dense_block(name, c, f[index])
      self.up.append(blk)
      self.core += blk
      self.scale_channels.append(c)
      t = ConvVariable(name("conv2d_transpose"), 3, c, f[index - 1], transposed=True)
      self.upsample.append(t)
      self.core.append(t)
      c = self.skip_channels[index - 1] + f[index - 1]
    self.last, c = self._dense_block(name, c, f[0])
    self.core += self.last
    self.scale_channels.append(c)

  # -- variables ---------------------------------------------------------------------------------------
  def conv_variables(self):
    out = list(self.core)
    for a, b in self.post:
      out += [a, b]
    return out + list(self.compose)

  def variable_shapes(self):
    """[(name, shape)] in TF creation order (embedding first: it is created by the first SourceEncoder call)."""
    out = []
    if self.embedding_shape is not None:
      out.append(("embedding/feature_flags_embedding_matrix", tuple(self.embedding_shape)))
    for v in self.conv_variables():
      out.append((v.kernel_name, v.kernel_shape))
      out.append((v.bias_name, (v.cout,)))
    return out

  def init_weights(self, seed=4321):
    """tf.layers defaults: glorot-uniform kernels / embedding, zero biases; numpy Generator(seed) stream in
    creation order."""
    rng = np.random.default_rng(seed)
    weights = {}
    if self.embedding_shape is not None:
      v, d = self.embedding_shape
      limit = np.sqrt(6.0 / (v + d))
      weights["embedding/feature_flags_embedding_matrix"] = rng.uniform(-limit, limit, size=(v, d)).astype(np.float32)
    for var in self.conv_variables():
      limit = var.glorot_limit()
      weights[var.kernel_name] = rng.uniform(-limit, limit, size=var.kernel_shape).astype(np.float32)
      weights[var.bias_name] = np.zeros((var.cout,), dtype=np.float32)
    return weights

  def parameter_count(self):
    return int(sum(int(np.prod(s)) for _, s in self.variable_shapes()))

  def mac_per_pixel(self, features_per_tuple=1):
    """Multiply-accumulates per full-resolution pixel of one tuple pass."""
    total = 0.0
    res = {}
    # spatial scale of every conv variable
    if self.core_name == "U-Net":
      for i, blk in enumerate(self.down):
        for v in blk:
          res[v.name] = i
      for i, blk in enumerate(self.up):
        for v in blk:
          res[v.name] = self.steps - i
        res[self.upsample[i].name] = self.steps - i
      for v in self.last:
        res[v.name] = 0
    else:
      res[self.pre.name] = 0
      for i, blk in enumerate(self.down):
        for v in blk:
          res[v.name] = i
        res[self.transition[i].name] = i
      for i, blk in enumerate(self.up):
        for v in blk:
          res[v.name] = self.steps - i
        res[self.upsample[i].name] = self.steps - i
      for v in self.last:
        res[v.name] = 0
    for v in self.core:
      total += v.ksize * v.ksize * v.cin * v.cout / 4.0 ** res[v.name]
    nscales = len(self.post)
    for k, (a, b) in enumerate(self.post):
      s = nscales - 1 - k
      total += (a.cin * a.cout + b.cin * b.cout) / 4.0 ** s
    if self.compose:
      per = sum(v.ksize * v.ksize * v.cin * v.cout for v in self.compose)
      total += features_per_tuple * per * sum(1.0 / 4.0 ** s for s in range(nscales - 1))
    return total


# ------------------------------------------------------------------------------------------------ device side
class V:
  """Channel window [coff, coff+c) of a contiguous NHWC torch tensor (the Python face of dd_tensor)."""
  __slots__ = ("t", "c", "coff", "d", "lo")

  def __init__(self, t, c=None, coff=0):
    self.t = t
    self.coff = coff
    self.c = t.shape[3] - coff if c is None else c
    self.d = _lib.desc(t, self.c, coff)
    self.lo = None        # float16x2 mode: the view of the low halves (same channels, second half of the buffer)

  def window(self, c, coff):
    return V(self.t, c, self.coff + coff)


class DeviceNetwork:
  """Weights of a NetworkSpec packed on one GPU + the forward executor."""

  def __init__(self, ctx, spec, weights, dtype=torch.float16, logits_dtype=torch.float32):
    assert dtype in (torch.float16, torch.bfloat16, torch.float32, "float16x2")
    # "float16x2": the high-accuracy tensor-core mode - every activation / weight is an fp16 (hi, lo) pair, three MMA passes
    # per layer (dd_conv2d_fwd_split); buffers hold [hi channels | lo channels]
    self.split = (dtype == "float16x2")
    self.pack_dtype = dtype
    if self.split:
      dtype = torch.float16
      if spec.core_name != "U-Net":
        raise _lib.DDError("the float16x2 mode is built for the U-Net core (the benchmarked network) only")
    self.ctx, self.spec, self.dtype = ctx, spec, dtype
    # bfloat16 storage runs the same tensor-core kernels (kind::f16 with bf16 operands), the fused compose kernel included;
    # the fused output-head kernel is fp16 mma.sync code, so bf16 uses the layer-by-layer launches there
    self.logits_dtype = logits_dtype if (dtype == torch.float16 and not self.split) else torch.float32
    self.fused_compose = True   # tests flip this to compare against the layer-by-layer path
    self.fused_post_kp = True   # likewise: 1x1 post-processing + kernel-prediction apply in one kernel
    self.align = 8 if dtype in (torch.float16, torch.bfloat16) else 1
    if dtype in (torch.float16, torch.bfloat16):
      for f in spec.filters:
        if f % 8:
          raise _lib.DDError("the 16-bit paths need filter counts that are multiples of 8 (got %s)" % spec.filters)
    self._buffers = {}
    self.load_weights(weights)

  def load_weights(self, weights):
    """(Re)packs a weight dictionary (TF names / layouts) for the device."""
    ctx, dev = self.ctx, self.ctx.device
    self.packed, self.bias, self.host = {}, {}, {}
    for var in self.spec.conv_variables():
      k = np.asarray(weights[var.kernel_name], dtype=np.float32)
      b = np.asarray(weights[var.bias_name], dtype=np.float32)
      assert k.shape == var.kernel_shape, (var.name, k.shape, var.kernel_shape)
      self.host[var.name] = (k, b)
      if var in self.spec.compose and var.ksize == 1:
        continue  # compose head / tail weights travel in the launch parameters
      if var.transposed and var.ksize == 3:
        # one packed 3x3 weight set per output phase is built lazily by _transpose3x3
        self.packed[var.name] = self._pack_transpose3x3(k)
      else:
        self.packed[var.name] = ctx.pack_conv_weights(torch.from_numpy(k), self.pack_dtype, transposed=var.transposed)
      bias = torch.zeros(_round_up(var.cout, 16), dtype=torch.float32)
      bias[:var.cout] = torch.from_numpy(b)
      self.bias[var.name] = bias.to(dev)
    self.post_kp_packed = {}    # (scale index in spec.post, K, features) -> device blob, built on first use
    self.compose_packed = None
    if self.spec.compose and self.dtype in (torch.float16, torch.bfloat16):
      head, c1, c2, c3, c4, tail = self.spec.compose
      blob, floats, code = _lib.pack_compose_weights(
          self.host[head.name][0], self.host[head.name][1], [self.host[c.name][0] for c in (c1, c2, c3, c4)],
          [self.host[c.name][1] for c in (c1, c2, c3, c4)], self.host[tail.name][0], self.host[tail.name][1],
          dtype=_lib.DD_F16 if self.dtype == torch.float16 else _lib.DD_BF16)
      self.compose_packed = (torch.from_numpy(blob).to(dev), floats, code)

  # conv2d_transpose 3x3 stride 2 'same' (Tiramisu.py:62-64) = 4 output phases, each a stride-1
  # convolution of the input with a subset of the taps: out[2y+py, 2x+px] = sum_{dy,dx in {0,-1}}
  # x[y+dy, x+dx] . W[py-2dy, px-2dx] (taps with index > 2 do not exist).  Packed as ordinary 3x3 'same'
  # kernels (zero taps elsewhere) so the same implicit-GEMM kernel runs them with a pixel-shuffle epilogue.
  def _pack_transpose3x3(self, k):
    kh, kw, cout, cin = k.shape
    phases = []
    for py in range(2):
      for px in range(2):
        w = np.zeros((3, 3, cin, cout), dtype=np.float32)
        for dy in (0, -1):
          r = py - 2 * dy
          if r > 2:
            continue
          for dx in (0, -1):
            s = px - 2 * dx
            if s > 2:
              continue
            w[dy + 1, dx + 1] = k[r, s].T
        phases.append(self.ctx.pack_conv_weights(torch.from_numpy(w), self.dtype))
    return phases

  # -- buffers -----------------------------------------------------------------------------------------
  def _buf(self, key, shape, dtype=None):
    dtype = dtype or self.dtype
    k = (key, tuple(shape), dtype)
    t = self._buffers.get(k)
    if t is None:
      t = torch.empty(shape, dtype=dtype, device=self.ctx.device)
      self._buffers[k] = t
    return t

  def release_buffers(self):
    self._buffers.clear()

  # -- primitive wrappers ------------------------------------------------------------------------------
  def _conv(self, var, x, y, relu=False, residual=None, y_relu=None):
    assert x.c == var.cin and y.c == var.cout, (var.name, x.c, var.cin, y.c, var.cout)
    self.ctx.conv2d(x.d, self.packed[var.name], self.bias[var.name], var.ksize, y.d, relu=relu,
                    residual=residual.d if residual is not None else None,
                    y_relu=y_relu.d if y_relu is not None else None)

  def _block(self, key, layers, x, out):
    """n x [conv3x3 + ReLU] (UNet.py:25-36); the last conv writes into `out` (a window of a concat buffer)."""
    b, h, w = x.t.shape[0], x.t.shape[1], x.t.shape[2]
    cur = x
    for i, var in enumerate(layers):
      if i == len(layers) - 1:
        dst = out
      else:
        dst = V(self._buf("%s.pp%d" % (key, i % 2), (b, h, w, _round_up(var.cout, self.align))), var.cout)
      self._conv(var, cur, dst, relu=True)
      cur = dst
    return out

  # -- U-Net on fp16 (hi, lo) pairs (float16x2 mode) -------------------------------------------------------------
  def _sv(self, t, c=None, coff=0):
    """Split view: channels [coff, coff+c) of the hi half and of the lo half of a [.., 2*Ct] buffer."""
    ct = t.shape[3] // 2
    v = V(t, ct - coff if c is None else c, coff)
    v.lo = V(t, v.c, ct + coff)
    return v

  def _sbuf(self, key, shape):
    return self._buf("x2." + key, tuple(shape[:3]) + (2 * shape[3],))

  def _conv_split(self, var, x, y, relu=False):
    assert x.c == var.cin and y.c == var.cout and x.lo is not None
    self.ctx.conv2d_split(x.d, x.lo.d, self.packed[var.name], self.bias[var.name], var.ksize, y.d,
                          y.lo.d if y.lo is not None else None, relu=relu)

  def _block_split(self, key, layers, x, out):
    b, h, w = x.t.shape[0], x.t.shape[1], x.t.shape[2]
    cur = x
    for i, var in enumerate(layers):
      dst = out if i == len(layers) - 1 else self._sv(self._sbuf("%s.pp%d" % (key, i % 2), (b, h, w, _round_up(var.cout, 8))), var.cout)
      self._conv_split(var, cur, dst, relu=True)
      cur = dst
    return out

  def _forward_unet_split(self, x0):
    spec, f, steps, ctx = self.spec, self.spec.filters, self.spec.steps, self.ctx
    b, h, w = x0.t.shape[0], x0.t.shape[1], x0.t.shape[2]
    dims = [(h, w)]
    for i in range(steps):
      hh, ww = dims[-1]
      if hh % 2 or ww % 2:
        raise _lib.DDError("height/width must be divisible by 2^%d (got %dx%d)" % (steps, h, w))
      dims.append((hh // 2, ww // 2))
    cats, x = [], x0
    for i in range(steps):
      hh, ww = dims[i]
      cat = self._sbuf("unet.cat%d" % i, (b, hh, ww, 2 * f[i]))
      cats.append(cat)
      skip = self._sv(cat, f[i], 0)
      self._block_split("unet.d%d" % i, spec.down[i], x, skip)
      pooled = self._sv(self._sbuf("unet.pool%d" % i, (b, dims[i + 1][0], dims[i + 1][1], f[i])))
      ctx.maxpool_s2_split(skip.d, skip.lo.d, 3, pooled.d, pooled.lo.d)
      x = pooled
    results = []
    for i in range(steps):
      index = steps - i
      hh, ww = dims[index]
      out = self._sv(self._sbuf("unet.out%d" % index, (b, hh, ww, f[index])))
      self._block_split("unet.u%d" % index, spec.up[i], x, out)
      if spec.use_multiscale:
        results.append(out)
      cat = cats[index - 1]
      var = spec.upsample[i]
      up = self._sv(cat, f[index - 1], f[index - 1])
      ctx.conv2d_transpose2x2_split(out.d, out.lo.d, self.packed[var.name], self.bias[var.name], up.d, up.lo.d, relu=True)
      x = self._sv(cat)
    out = self._sv(self._sbuf("unet.out0", (b, h, w, f[0])))
    self._block_split("unet.l", spec.last, x, out)
    results.append(out)
    return results

  # -- U-Net ---------------------------------------------------------------------------------------------
  def _forward_unet(self, x0):
    spec, f, steps, ctx = self.spec, self.spec.filters, self.spec.steps, self.ctx
    b, h, w = x0.t.shape[0], x0.t.shape[1], x0.t.shape[2]
    dims = [(h, w)]
    for i in range(steps):
      hh, ww = dims[-1]
      if hh % 2 or ww % 2:
        raise _lib.DDError("height/width must be divisible by 2^%d (got %dx%d): the skip concat of UNet.py:91-92 "
                           "requires it" % (steps, h, w))
      dims.append((hh // 2, ww // 2))
    cats = []
    x = x0
    for i in range(steps):
      hh, ww = dims[i]
      cat = self._buf("unet.cat%d" % i, (b, hh, ww, 2 * f[i]))
      cats.append(cat)
      skip = V(cat, f[i], 0)
      self._block("unet.d%d" % i, spec.down[i], x, skip)
      pooled = V(self._buf("unet.pool%d" % i, (b, dims[i + 1][0], dims[i + 1][1], f[i])))
      ctx.maxpool_s2(skip.d, 3, pooled.d)
      x = pooled
    results = []
    for i in range(steps):
      index = steps - i
      hh, ww = dims[index]
      out = V(self._buf("unet.out%d" % index, (b, hh, ww, f[index])))
      self._block("unet.u%d" % index, spec.up[i], x, out)
      if spec.use_multiscale:
        results.append(out)
      cat = cats[index - 1]
      var = spec.upsample[i]
      ctx.conv2d_transpose2x2(out.d, self.packed[var.name], self.bias[var.name], V(cat, f[index - 1], f[index - 1]).d,
                              relu=True)
      x = V(cat)
    out = V(self._buf("unet.out0", (b, h, w, f[0])))
    self._block("unet.l", spec.last, x, out)
    results.append(out)
    return results

  # -- Tiramisu ------------------------------------------------------------------------------------------
  # Every tensor of the dense path exists twice: raw (consumed by 1x1 post-process, conv2d_transpose and as
  # the concatenated block output) and ReLU'd (the input of every dense conv, Tiramisu.py:34), so the
  # pre-activation never has to be applied on load.  Both live in channel windows of block-wide buffers.
  def _dense_block(self, layers, raw, act, c0):
    c = c0
    for var in layers:
      self._conv(var, V(act, c, 0), V(raw, var.cout, c), relu=False, y_relu=V(act, var.cout, c))
      c += var.cout
    return c

  def _forward_tiramisu(self, x0):
    spec, f, steps, ctx = self.spec, self.spec.filters, self.spec.steps, self.ctx
    b, h, w = x0.t.shape[0], x0.t.shape[1], x0.t.shape[2]
    dims = [(h, w)]
    for i in range(steps):
      hh, ww = dims[-1]
      if hh % 2 or ww % 2:
        raise _lib.DDError("height/width must be divisible by 2^%d (got %dx%d)" % (steps, h, w))
      dims.append((hh // 2, ww // 2))
    n = spec.convs_per_block
    # channel plan of the up-path buffers: [skip | upsampled | dense growth]
    up_total = {}
    for i in range(steps):
      index = steps - i
      up_total[index - 1] = spec.skip_channels[index - 1] + f[index - 1] + n * f[index - 1]
    raws, acts = {}, {}
    for level in range(steps):
      hh, ww = dims[level]
      raws[level] = self._buf("tira.raw%d" % level, (b, hh, ww, up_total[level]))
      acts[level] = self._buf("tira.act%d" % level, (b, hh, ww, up_total[level]))
    hb, wb = dims[steps]
    c_bottom_in = spec.skip_channels[steps - 1]
    bottom_total = c_bottom_in + n * f[steps]
    raws[steps] = self._buf("tira.raw%d" % steps, (b, hb, wb, bottom_total))
    acts[steps] = self._buf("tira.act%d" % steps, (b, hb, wb, bottom_total))
    # pre-processing conv (ReLU activation => raw == act)
    self._conv(spec.pre, x0, V(raws[0], f[0], 0), relu=True, y_relu=V(acts[0], f[0], 0))
    c = f[0]
    for i in range(steps):
      c = self._dense_block(spec.down[i], raws[i], acts[i], c)
      assert c == spec.skip_channels[i]
      # transition down: ReLU -> conv1x1 -> maxpool 2x2 (Tiramisu.py:43-58); pool(relu(z)) == relu(pool(z))
      hh, ww = dims[i]
      z = self._buf("tira.z%d" % i, (b, hh, ww, c))
      za = self._buf("tira.za%d" % i, (b, hh, ww, c))
      self._conv(spec.transition[i], V(acts[i], c, 0), V(z), relu=False, y_relu=V(za))
      ctx.maxpool_s2(V(z).d, 2, V(raws[i + 1], c, 0).d)
      ctx.maxpool_s2(V(za).d, 2, V(acts[i + 1], c, 0).d)
    results = []
    for i in range(steps):
      index = steps - i
      c = self._dense_block(spec.up[i], raws[index], acts[index], c)
      if spec.use_multiscale:
        results.append(V(raws[index], c, 0))
      var = spec.upsample[i]
      level = index - 1
      cs = spec.skip_channels[level]
      self._transpose3x3(var, V(raws[index], c, 0), V(raws[level], var.cout, cs), V(acts[level], var.cout, cs))
      c = cs + var.cout
    c = self._dense_block(spec.last, raws[0], acts[0], c)
    results.append(V(raws[0], c, 0))
    return results

  def _transpose3x3(self, var, x, y, y_act):
    self.ctx.conv2d_transpose3x3(x.d, self.packed[var.name], self.bias[var.name], y.d, y_act.d, relu=True)

  # -- public ------------------------------------------------------------------------------------------
  def forward_core(self, x0):
    """Core architecture only: the multi-scale outputs, COARSEST first (the order of spec.post)."""
    spec = self.spec
    assert x0.c == spec.input_channels, (x0.c, spec.input_channels)
    if self.split:
      return self._forward_unet_split(x0)
    return self._forward_unet(x0) if spec.core_name == "U-Net" else self._forward_tiramisu(x0)

  def can_fuse_post_kp(self, ksize, features):
    return (self.fused_post_kp and self.dtype == torch.float16 and not self.split and
            bool(self.ctx.lib.dd_post_kp_supported(int(ksize), int(features))) and
            self.spec.output_channels == features * ksize * ksize)

  def post_kernel_predict(self, k, r, src, ksize, features, images_per_tuple, out):
    """AdjustNumberOfChannels of core output `r` (index k of spec.post) + kernel prediction on `src`, fused."""
    key = (k, ksize, features)
    blob = self.post_kp_packed.get(key)
    if blob is None:
      a, bvar = self.spec.post[k]
      blob = torch.from_numpy(_lib.pack_post_kp_weights(self.host[a.name][0], self.host[a.name][1], self.host[bvar.name][0],
                                                        self.host[bvar.name][1], ksize, features)).to(self.ctx.device)
      self.post_kp_packed[key] = blob
    self.ctx.post_kp(r.d, blob, src, ksize, features, images_per_tuple, out)

  def forward(self, x0):
    """x0: V over [B,H,W,C0] (dtype of the network).  Returns the post-processed outputs (logits), LARGEST
    scale first (Architecture.py:577-579), as V over [B,h,w,O] tensors of `logits_dtype`."""
    spec = self.spec
    results = self.forward_core(x0)
    return self.post_process(results)

  def post_process(self, results):
    spec = self.spec
    outs = []
    o = spec.output_channels
    for k, (r, (a, bvar)) in enumerate(zip(results, spec.post)):
      b, h, w = r.t.shape[0], r.t.shape[1], r.t.shape[2]
      out = V(self._buf("post.out%d" % k, (b, h, w, _round_up(o, 8)), self.logits_dtype), o)
      if self.split:
        mid = self._sv(self._sbuf("post.mid%d" % k, (b, h, w, _round_up(o, 8))), o)
        self._conv_split(a, r, mid, relu=True)
        self._conv_split(bvar, mid, out, relu=False)        # fp32 logits: no low half
      else:
        mid = V(self._buf("post.mid%d" % k, (b, h, w, _round_up(o, self.align))), o)
        self._conv(a, r, mid, relu=True)
        self._conv(bvar, mid, out, relu=False)
      outs.append(out)
    if spec.use_multiscale:
      outs.reverse()
    return outs

  def compose(self, small, large, out, inv=None):
    """MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93) on fp32 image banks
    small [I,h/2,w/2,3], large [I,h,w,3] -> out [I,h,w,3]; `inv` fuses the inverse standardisation."""
    spec, ctx = self.spec, self.ctx
    if self.compose_packed is not None and self.fused_compose:
      # tensor-core mode: the whole weight net + blend is one tcgen05 launch, intermediates never leave shared memory / TMEM
      ctx.compose_scales(small.d, large.d, self.compose_packed, inv, out.d)
      return
    i, h, w = large.t.shape[0], large.t.shape[1], large.t.shape[2]
    head, c1, c2, c3, c4, tail = spec.compose
    ta = V(self._buf("compose.a", (i, h, w, 24)))
    tb = V(self._buf("compose.b", (i, h, w, 24)))
    tc = V(self._buf("compose.c", (i, h, w, 24)))
    td = V(self._buf("compose.d", (i, h, w, 24)))
    hw_, hb_ = self.host[head.name]
    ctx.compose_head(small.d, large.d, torch.from_numpy(np.ascontiguousarray(hw_.reshape(6, 24))),
                     torch.from_numpy(hb_), 24, ta.d)
    # residual block 1: x1 = x0 + conv(relu(conv(relu(x0)))); x0 = ta >= 0 already (ReLU'd head)
    self._conv(c1, ta, tb, relu=True)                            # tb = relu(conv(x0))
    self._conv(c2, tb, tc, relu=False, residual=ta, y_relu=td)   # tc = x1, td = relu(x1)
    # residual block 2: x2 = x1 + conv(relu(conv(relu(x1))))
    self._conv(c3, td, tb, relu=True)                            # tb = relu(conv(relu(x1)))
    self._conv(c4, tb, ta, relu=False, residual=tc)              # ta = x2
    tb = ta
    tw, tbias = self.host[tail.name]
    ctx.compose_tail(tb.d, torch.from_numpy(np.ascontiguousarray(tw.reshape(24))), torch.from_numpy(tbias), 24,
                     small.d, large.d, inv, out.d)
