"""ctypes binding of libdd_b200.so (include/dd_b200.h).

This is the *only* way the Python host reaches the GPU kernels: plain pointers and sizes, PyTorch tensors
are nothing more than the device-memory container (``tensor.data_ptr()``).  There is no CPU fallback:
if the shared library is missing, or a call fails, an exception is raised.
"""
import ctypes
import os

import torch

DD_F32, DD_F16, DD_BF16 = 0, 1, 2
DD_F16X2 = 3   # weight-packing code of the split-fp16 mode
DD_PACK_BF16 = 16
DD_CONV_RELU, DD_CONV_RELU_COPY, DD_CONV_RESIDUAL_MASK = 1, 2, 4

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdd_b200.so")


class DDError(RuntimeError):
  pass


class dd_tensor(ctypes.Structure):
  _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("n", ctypes.c_int32), ("h", ctypes.c_int32),
              ("w", ctypes.c_int32), ("c", ctypes.c_int32), ("cstride", ctypes.c_int32), ("coff", ctypes.c_int32)]


class dd_standardize_params(ctypes.Structure):
  _fields_ = [("use_log1p", ctypes.c_int32), ("mean", ctypes.c_float), ("variance", ctypes.c_float),
              ("use_variance", ctypes.c_int32), ("variance_mode", ctypes.c_int32),
              ("relative_variance", ctypes.c_int32), ("compute_before_standardization", ctypes.c_int32),
              ("compress_to_one_channel", ctypes.c_int32), ("epsilon", ctypes.c_float)]


class dd_invert_params(ctypes.Structure):
  _fields_ = [("use_log1p", ctypes.c_int32), ("mean", ctypes.c_float), ("variance", ctypes.c_float)]


class dd_gather_entry(ctypes.Structure):
  _fields_ = [("ptr", ctypes.c_void_p), ("cstride", ctypes.c_int32), ("cidx", ctypes.c_int32),
              ("constant", ctypes.c_float), ("pad_", ctypes.c_int32)]


_P = ctypes.POINTER
_vp, _i, _u32, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t
_T = _P(dd_tensor)

# name -> (restype, argtypes); mirrors include/dd_b200.h declaration by declaration
SIGNATURES = {
    "dd_abi_version": (_i, []),
    "dd_last_error": (ctypes.c_char_p, []),
    "dd_ctx_create": (_i, [_i, _P(_vp)]),
    "dd_ctx_destroy": (_i, [_vp]),
    "dd_ctx_sm_count": (_i, [_vp]),
    "dd_ctx_set_option": (_i, [_vp, ctypes.c_char_p, _i]),
    "dd_ctx_launch_count": (ctypes.c_int64, [_vp]),
    "dd_ctx_set_trace_buffer": (_i, [_vp, _vp]),
    "dd_conv2d_packed_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dd_conv2d_pack_weights": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "dd_conv2d_fwd": (_i, [_vp, _T, _vp, _vp, _i, _u32, _T, _T, _T, _vp]),
    "dd_conv2d_transpose2x2_fwd": (_i, [_vp, _T, _vp, _vp, _u32, _T, _vp]),
    "dd_conv2d_fwd_split": (_i, [_vp, _T, _T, _vp, _vp, _i, _u32, _T, _T, _vp]),
    "dd_conv2d_fwd_colsum": (_i, [_vp, _T, _vp, _vp, _i, _u32, _T, _T, _vp, _vp]),
    "dd_conv2d_transpose2x2_fwd_split": (_i, [_vp, _T, _T, _vp, _vp, _u32, _T, _T, _vp]),
    "dd_maxpool_s2_fwd_split": (_i, [_vp, _T, _T, _i, _T, _T, _vp]),
    "dd_assemble_input_split": (_i, [_vp, _vp, _i, _i, _T, _T, _vp]),
    "dd_conv2d_transpose3x3_fwd": (_i, [_vp, _T, _P(_vp), _vp, _u32, _T, _T, _vp]),
    "dd_maxpool_s2_fwd": (_i, [_vp, _T, _i, _T, _vp]),
    "dd_avgpool_fwd": (_i, [_vp, _T, _i, _T, _vp]),
    "dd_standardize_variance": (_i, [_vp, _T, _P(dd_standardize_params), _T, _T, _vp]),
    "dd_standardize_variance_batch": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "dd_standardize_variance_job_bytes": (ctypes.c_size_t, []),
    "dd_assemble_input": (_i, [_vp, _vp, _i, _i, _T, _vp]),
    "dd_kernel_predict_fwd": (_i, [_vp, _T, _T, _i, _i, _i, _T, _vp]),
    "dd_compose_head_fwd": (_i, [_vp, _T, _T, _vp, _vp, _i, _T, _vp]),
    "dd_compose_tail_fwd": (_i, [_vp, _T, _vp, _vp, _i, _T, _T, _P(dd_invert_params), _T, _vp]),
    "dd_compose_weights_bytes": (_sz, []),
    "dd_compose_params_floats": (_sz, []),
    "dd_compose_pack_weights": (_i, [_P(_vp), _P(_vp), _i, _vp]),
    "dd_compose_scales_fwd": (_i, [_vp, _T, _T, _vp, _vp, _i, _P(dd_invert_params), _T, _vp]),
    "dd_invert_standardization": (_i, [_vp, _T, _P(dd_invert_params), _T, _vp]),
    "dd_relu_bwd": (_i, [_vp, _T, _T, _T, _vp]),
    "dd_relu_bwd_acc": (_i, [_vp, _T, _T, _T, _vp]),
    "dd_muladd_fwd": (_i, [_vp, _T, _T, _T, _T, _vp]),
    "dd_muladd_bwd": (_i, [_vp, _T, _T, _T, _T, _T, _T, _vp]),
    "dd_axpy": (_i, [_vp, ctypes.c_float, _T, _T, _vp]),
    "dd_fill": (_i, [_vp, ctypes.c_float, _T, _vp]),
    "dd_invert_standardization_bwd": (_i, [_vp, _T, _T, _P(dd_invert_params), _T, _vp]),
    "dd_loss_fwd_bwd": (_i, [_vp, _T, _T, _i, ctypes.c_float, ctypes.c_float, _vp, _T, _i, _vp]),
    "dd_loss_variation_fwd_bwd": (_i, [_vp, _T, _T, _i, ctypes.c_float, ctypes.c_float, _vp, _T, _vp]),
    "dd_mask_sum": (_i, [_vp, _T, _vp, _vp]),
    "dd_loss_masked_fwd_bwd": (_i, [_vp, _T, _T, _T, _vp, _i, ctypes.c_float, ctypes.c_float, _vp, _T, _vp]),
    "dd_ssim_stats": (_i, [_vp, _T, _T, _T, _vp]),
    "dd_ssim_reduce": (_i, [_vp, _T, _i, ctypes.c_float, _vp, _vp]),
    "dd_ssim_bwd": (_i, [_vp, _T, _T, _T, _vp, ctypes.c_float, _T, _vp]),
    "dd_avgpool2_adjoint": (_i, [_vp, _T, _T, _vp]),
    "dd_conv2d_wgrad": (_i, [_vp, _T, _T, _i, _i, _vp, _vp, _vp]),
    "dd_conv2d_wgrad_tc": (_i, [_vp, _T, _T, _i, _i, _vp, ctypes.c_float, _vp]),
    "dd_conv2d_pack_weights_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "dd_space_to_depth2_mask": (_i, [_vp, _T, _T, _T, _vp]),
    "dd_relu_bwd_bias": (_i, [_vp, _T, _T, _T, _vp, ctypes.c_float, _vp]),
    "dd_conv2d_transpose2x2_dgrad": (_i, [_vp, _T, _vp, _T, _vp]),
    "dd_conv2d_transpose3x3_dgrad": (_i, [_vp, _T, _vp, _T, _vp]),
    "dd_conv2d_repack_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dd_maxpool_s2_bwd": (_i, [_vp, _T, _T, _T, _i, _T, _vp]),
    "dd_maxpool_s2_bwd_acc": (_i, [_vp, _T, _T, _T, _i, _T, _vp]),
    "dd_maxpool_s2_fwd_index": (_i, [_vp, _T, _i, _T, _vp, _vp]),
    "dd_maxpool_s2_bwd_index": (_i, [_vp, _vp, _T, _i, _T, _vp]),
    "dd_kernel_predict_bwd": (_i, [_vp, _T, _T, _T, _i, _i, _i, _T, _vp]),
    "dd_compose_tail_bwd": (_i, [_vp, _T, _vp, _vp, _i, _T, _T, _T, _T, _T, _T, _vp, _vp, _vp]),
    "dd_compose_head_bwd": (_i, [_vp, _T, _T, _vp, _i, _T, _T, _T, _T, _vp, _vp, _vp]),
    "dd_channel_sum": (_i, [_vp, _T, _i, _vp, _vp]),
    "dd_adam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                          ctypes.c_int64, ctypes.c_float, _vp]),
    "dd_adam_step_guarded": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                  ctypes.c_float, _vp, _vp]),
    "dd_crc32c": (ctypes.c_uint32, [ctypes.c_char_p, _sz]),
    "dd_augment_tiles": (_i, [_vp, _T, _i, _vp, _vp, _vp, _vp, _T, _vp]),
    "dd_tiles_gather": (_i, [_vp, _T, _vp, _T, _vp]),
    "dd_tiles_scatter": (_i, [_vp, _T, _vp, _T, _vp]),
    "dd_post_kp_supported": (_i, [_i, _i]),
    "dd_post_kp_weights_bytes": (_sz, [_i, _i, _i]),
    "dd_post_kp_pack_weights": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dd_post_kp_fwd": (_i, [_vp, _T, _vp, _T, _i, _i, _i, _T, _vp]),
    "dd_comm_unique_id": (_i, [_vp]),
    "dd_comm_init": (_i, [_vp, _vp, _i, _i, _P(_vp)]),
    "dd_comm_allreduce_sum_f32": (_i, [_vp, _vp, _sz, _vp]),
    "dd_comm_destroy": (_i, [_vp]),
    "dd_cast_copy": (_i, [_vp, _T, _T, _vp]),
    "dd_l2_flush": (_i, [_vp, _vp, _sz, _vp]),
}

_lib = None


def load_library(path=None):
  """Loads libdd_b200.so and installs the prototypes.  Raises DDError when it is not built."""
  global _lib
  if _lib is not None and path is None:
    return _lib
  path = path or _LIB_PATH
  if not os.path.exists(path):
    raise DDError("%s not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                  "(there is no CPU fallback)" % path)
  lib = ctypes.CDLL(path)
  for name, (restype, argtypes) in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
  if lib.dd_abi_version() != 1:
    raise DDError("libdd_b200 ABI version mismatch")
  _lib = lib
  return lib


def _dtype_code(t):
  if t.dtype == torch.float32:
    return DD_F32
  if t.dtype == torch.float16:
    return DD_F16
  if t.dtype == torch.bfloat16:
    return DD_BF16
  raise DDError("unsupported tensor dtype %s" % t.dtype)


def desc(t, c=None, coff=0):
  """dd_tensor view of channels [coff, coff+c) of a contiguous NHWC torch tensor."""
  assert t.dim() == 4 and t.is_contiguous(), "NHWC contiguous tensor expected"
  n, h, w, cs = t.shape
  if c is None:
    c = cs - coff
  assert 0 <= coff and coff + c <= cs
  d = dd_tensor(t.data_ptr(), _dtype_code(t), n, h, w, c, cs, coff)
  d._keep = t   # the descriptor keeps its storage alive (kernels run asynchronously)
  return d


class Context:
  """Owns a dd_ctx for one CUDA device and wraps every entry point with error checking."""

  def __init__(self, device=0):
    self.lib = load_library()
    if not torch.cuda.is_available():
      raise DDError("no CUDA device: deepdenoiser_b200 has no CPU fallback")
    self.device = torch.device("cuda", device)
    handle = ctypes.c_void_p()
    self._check(self.lib.dd_ctx_create(device, ctypes.byref(handle)))
    self.handle = handle

  def __del__(self):
    try:
      if getattr(self, "handle", None):
        self.lib.dd_ctx_destroy(self.handle)
        self.handle = None
    except Exception:
      pass

  def _check(self, rc):
    if rc != 0:
      raise DDError("libdd_b200 error %d: %s" % (rc, self.lib.dd_last_error().decode()))

  def _stream(self=None):
    """The caller's current stream ON THIS CONTEXT'S DEVICE (not on whatever device is current)."""
    device = self.device if isinstance(self, Context) else None
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)

  # -- context
  def sm_count(self):
    return self.lib.dd_ctx_sm_count(self.handle)

  def launch_count(self):
    return int(self.lib.dd_ctx_launch_count(self.handle))

  def set_option(self, name, value):
    self._check(self.lib.dd_ctx_set_option(self.handle, name.encode(), int(value)))

  def set_trace_buffer(self, t):
    self._check(self.lib.dd_ctx_set_trace_buffer(self.handle, t.data_ptr() if t is not None else None))

  # -- conv
  def pack_conv_weights(self, w, dtype, transposed=False):
    """w: CPU float32 tensor in TF layout [kh,kw,cin,cout] ([kh,kw,cout,cin] if transposed).
    Returns (packed device uint8 tensor)."""
    w = w.detach().to(torch.float32).contiguous().cpu()
    ks = w.shape[0]
    cin, cout = (w.shape[3], w.shape[2]) if transposed else (w.shape[2], w.shape[3])
    code = DD_F16X2 if dtype == "float16x2" else (DD_F16 if dtype == torch.float16 else (DD_BF16 if dtype == torch.bfloat16 else DD_F32))
    nbytes = self.lib.dd_conv2d_packed_bytes(ks, cin, cout, code, int(transposed))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
    self._check(self.lib.dd_conv2d_pack_weights(self.handle, w.data_ptr(), ks, cin, cout, code, int(transposed),
                                                packed.data_ptr(), self._stream()))
    return packed

  def conv2d(self, x, w_packed, bias, ksize, y, relu=False, residual=None, y_relu=None, residual_is_mask=False, colsum=None):
    flags = ((DD_CONV_RELU if relu else 0) | (DD_CONV_RELU_COPY if y_relu is not None else 0) |
             (DD_CONV_RESIDUAL_MASK if residual_is_mask else 0))
    if colsum is not None:       # fp32 device tensor: column sums of y are accumulated into it (fused BiasAddGrad)
      assert y_relu is None
      self._check(self.lib.dd_conv2d_fwd_colsum(
          self.handle, ctypes.byref(x), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None, ksize, flags,
          ctypes.byref(residual) if residual is not None else None, ctypes.byref(y), colsum.data_ptr(), self._stream()))
      return
    self._check(self.lib.dd_conv2d_fwd(
        self.handle, ctypes.byref(x), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None, ksize,
        flags, ctypes.byref(residual) if residual is not None else None, ctypes.byref(y),
        ctypes.byref(y_relu) if y_relu is not None else None, self._stream()))

  def conv2d_split(self, x_hi, x_lo, w_packed, bias, ksize, y_hi, y_lo, relu=False):
    """float16x2 mode (dd_conv2d_fwd_split): fp16 pairs in, an fp16 pair or one fp32 tensor (y_lo None) out."""
    self._check(self.lib.dd_conv2d_fwd_split(
        self.handle, ctypes.byref(x_hi), ctypes.byref(x_lo), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
        ksize, DD_CONV_RELU if relu else 0, ctypes.byref(y_hi), ctypes.byref(y_lo) if y_lo is not None else None, self._stream()))

  def conv2d_transpose2x2_split(self, x_hi, x_lo, w_packed, bias, y_hi, y_lo, relu=False):
    self._check(self.lib.dd_conv2d_transpose2x2_fwd_split(
        self.handle, ctypes.byref(x_hi), ctypes.byref(x_lo), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
        DD_CONV_RELU if relu else 0, ctypes.byref(y_hi), ctypes.byref(y_lo), self._stream()))

  def maxpool_s2_split(self, x_hi, x_lo, ksize, y_hi, y_lo):
    self._check(self.lib.dd_maxpool_s2_fwd_split(self.handle, ctypes.byref(x_hi), ctypes.byref(x_lo), ksize, ctypes.byref(y_hi),
                                                 ctypes.byref(y_lo), self._stream()))

  def assemble_input_split(self, table_dev, tuples, n, out_hi, out_lo):
    self._check(self.lib.dd_assemble_input_split(self.handle, table_dev.data_ptr(), tuples, n, ctypes.byref(out_hi),
                                                 ctypes.byref(out_lo), self._stream()))

  def conv2d_transpose2x2(self, x, w_packed, bias, y, relu=False):
    self._check(self.lib.dd_conv2d_transpose2x2_fwd(
        self.handle, ctypes.byref(x), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
        DD_CONV_RELU if relu else 0, ctypes.byref(y), self._stream()))

  def conv2d_transpose3x3(self, x, w_phases, bias, y, y_relu=None, relu=False):
    arr = (ctypes.c_void_p * 4)(*[w.data_ptr() for w in w_phases])
    flags = (DD_CONV_RELU if relu else 0) | (DD_CONV_RELU_COPY if y_relu is not None else 0)
    self._check(self.lib.dd_conv2d_transpose3x3_fwd(
        self.handle, ctypes.byref(x), arr, bias.data_ptr() if bias is not None else None, flags, ctypes.byref(y),
        ctypes.byref(y_relu) if y_relu is not None else None, self._stream()))

  # -- pooling
  def maxpool_s2(self, x, ksize, y):
    self._check(self.lib.dd_maxpool_s2_fwd(self.handle, ctypes.byref(x), ksize, ctypes.byref(y), self._stream()))

  def avgpool(self, x, factor, y):
    self._check(self.lib.dd_avgpool_fwd(self.handle, ctypes.byref(x), factor, ctypes.byref(y), self._stream()))

  # -- encoder
  def standardize_variance_batch(self, jobs, table):
    """jobs: list of (src desc, dd_standardize_params, std_out desc | None, var_out desc | None) with identical [n,h,w];
    table: uint8 device tensor of at least len(jobs) * dd_standardize_variance_job_bytes() bytes."""
    count = len(jobs)
    ptr_t = ctypes.POINTER(dd_tensor)
    src = (ptr_t * count)(*[ctypes.pointer(j[0]) for j in jobs])
    prm = (dd_standardize_params * count)(*[j[1] for j in jobs])
    so = (ptr_t * count)(*[ctypes.pointer(j[2]) if j[2] is not None else ptr_t() for j in jobs])
    vo = (ptr_t * count)(*[ctypes.pointer(j[3]) if j[3] is not None else ptr_t() for j in jobs])
    self._check(self.lib.dd_standardize_variance_batch(self.handle, count, src, prm, so, vo, table.data_ptr(),
                                                       table.numel(), self._stream()))

  def standardize_variance(self, src, params, std_out, var_out):
    self._check(self.lib.dd_standardize_variance(
        self.handle, ctypes.byref(src), ctypes.byref(params), ctypes.byref(std_out) if std_out is not None else None,
        ctypes.byref(var_out) if var_out is not None else None, self._stream()))

  def assemble_input(self, table_dev, tuples, n, out):
    self._check(self.lib.dd_assemble_input(self.handle, table_dev.data_ptr(), tuples, n, ctypes.byref(out),
                                           self._stream()))

  # -- kernel prediction / multi-scale
  def kernel_predict(self, src, logits, ksize, features, images_per_tuple, out):
    self._check(self.lib.dd_kernel_predict_fwd(self.handle, ctypes.byref(src), ctypes.byref(logits), ksize, features,
                                               images_per_tuple, ctypes.byref(out), self._stream()))

  def compose_head(self, small, large, w_host, b_host, c_mid, y):
    self._check(self.lib.dd_compose_head_fwd(self.handle, ctypes.byref(small), ctypes.byref(large),
                                             w_host.data_ptr(), b_host.data_ptr(), c_mid, ctypes.byref(y),
                                             self._stream()))

  def compose_tail(self, t, w_host, b_host, c_mid, small, large, inv, out):
    self._check(self.lib.dd_compose_tail_fwd(self.handle, ctypes.byref(t), w_host.data_ptr(), b_host.data_ptr(), c_mid,
                                             ctypes.byref(small), ctypes.byref(large),
                                             ctypes.byref(inv) if inv is not None else None, ctypes.byref(out),
                                             self._stream()))

  def compose_scales(self, small, large, packed, inv, out):
    """Fused compose_scales; `packed` = (device weight blob, host float parameters, dtype code) of pack_compose_weights()."""
    blob_dev, params_host, code = packed
    self._check(self.lib.dd_compose_scales_fwd(self.handle, ctypes.byref(small), ctypes.byref(large), blob_dev.data_ptr(),
                                               params_host.ctypes.data, code,
                                               ctypes.byref(inv) if inv is not None else None, ctypes.byref(out),
                                               self._stream()))

  def post_kp(self, x, blob_dev, src, ksize, features, images_per_tuple, out):
    """Fused 1x1 post-processing + kernel-prediction apply of one scale (dd_post_kp_fwd)."""
    self._check(self.lib.dd_post_kp_fwd(self.handle, ctypes.byref(x), blob_dev.data_ptr(), ctypes.byref(src), ksize, features,
                                        images_per_tuple, ctypes.byref(out), self._stream()))

  def invert_standardization(self, x, inv, y):
    self._check(self.lib.dd_invert_standardization(self.handle, ctypes.byref(x), ctypes.byref(inv), ctypes.byref(y),
                                                   self._stream()))

  # -- training (every call is a thin, checked pass-through; see include/dd_b200.h)
  def call(self, name, *args):
    """Generic checked call: ctx.call('dd_relu_bwd', byref(...), ...) appends the current stream."""
    self._check(getattr(self.lib, name)(self.handle, *args, self._stream()))

  def cast_copy(self, x, y):
    self._check(self.lib.dd_cast_copy(self.handle, ctypes.byref(x), ctypes.byref(y), self._stream()))

  def l2_flush(self, scratch):
    self._check(self.lib.dd_l2_flush(self.handle, scratch.data_ptr(), scratch.numel() * scratch.element_size(),
                                     self._stream()))


class Communicator:
  """dd_comm over NCCL: the flat-gradient all-reduce of data-parallel training through the C ABI.  The unique id travels
  over torch.distributed (any initialised backend) - plumbing only; the reduction itself is ncclAllReduce on the caller's
  stream."""

  def __init__(self, ctx, rank, world):
    import torch.distributed as dist
    self.ctx, self.rank, self.world = ctx, rank, world
    ident = ctypes.create_string_buffer(128)
    if rank == 0:
      ctx._check(ctx.lib.dd_comm_unique_id(ident))
    box = [bytes(ident.raw)]
    dist.broadcast_object_list(box, src=0)
    handle = ctypes.c_void_p()
    ctx._check(ctx.lib.dd_comm_init(ctx.handle, box[0], rank, world, ctypes.byref(handle)))
    self.handle = handle

  def all_reduce_sum(self, tensor):
    assert tensor.dtype == torch.float32 and tensor.is_cuda and tensor.is_contiguous()
    self.ctx._check(self.ctx.lib.dd_comm_allreduce_sum_f32(self.handle, tensor.data_ptr(), tensor.numel(), self.ctx._stream()))

  def close(self):
    if self.handle:
      self.ctx.lib.dd_comm_destroy(self.handle)
      self.handle = None


def pack_post_kp_weights(w1, b1, w2, b2, ksize, features):
  """Host blob of dd_post_kp_fwd from the TF tensors of AdjustNumberOfChannels: w1 [1,1,cin,O], b1 [O], w2 [1,1,O,O], b2 [O]."""
  import numpy as np
  lib = load_library()
  w1 = np.ascontiguousarray(np.asarray(w1, dtype=np.float32).reshape(-1, np.asarray(w1).shape[-1]))
  w2 = np.ascontiguousarray(np.asarray(w2, dtype=np.float32).reshape(-1, np.asarray(w2).shape[-1]))
  b1 = np.ascontiguousarray(b1, dtype=np.float32)
  b2 = np.ascontiguousarray(b2, dtype=np.float32)
  cin = w1.shape[0]
  blob = np.zeros(lib.dd_post_kp_weights_bytes(cin, ksize, features), dtype=np.uint8)
  rc = lib.dd_post_kp_pack_weights(w1.ctypes.data, b1.ctypes.data, w2.ctypes.data, b2.ctypes.data, cin, ksize, features,
                                   blob.ctypes.data)
  if rc != 0:
    raise DDError("dd_post_kp_pack_weights failed: %s" % lib.dd_last_error().decode())
  return blob


def pack_compose_weights(head_w, head_b, conv_w, conv_b, tail_w, tail_b, dtype=DD_F16):
  """Host data of dd_compose_scales_fwd: (weight blob [uint8], float parameters [292 float32], dtype code).  head_w [1,1,6,24] /
  [6,24], head_b [24], conv_w 4 x TF kernels [3,3,24,24] (kh, kw, cin, cout), conv_b 4 x [24], tail_w [1,1,24,1] / [24],
  tail_b [1]."""
  import numpy as np
  lib = load_library()
  c = 24
  blob = np.zeros(lib.dd_compose_weights_bytes(), dtype=np.uint8)
  kernels = [np.ascontiguousarray(np.asarray(w, dtype=np.float32).reshape(3, 3, c, c)) for w in conv_w]
  biases = [np.ascontiguousarray(np.asarray(b, dtype=np.float32).reshape(c)) for b in conv_b]
  ptrs = (ctypes.c_void_p * 4)(*[k.ctypes.data for k in kernels])
  bptrs = (ctypes.c_void_p * 4)(*[b.ctypes.data for b in biases])
  rc = lib.dd_compose_pack_weights(ptrs, bptrs, int(dtype), blob.ctypes.data)
  if rc != 0:
    raise DDError("dd_compose_pack_weights failed: %s" % lib.dd_last_error().decode())
  fl = np.zeros(lib.dd_compose_params_floats(), dtype=np.float32)
  fl[0:144] = np.asarray(head_w, dtype=np.float32).reshape(6, c).reshape(-1)
  fl[144:168] = np.asarray(head_b, dtype=np.float32).reshape(c)
  for i, b in enumerate(conv_b):
    fl[168 + i * c:168 + (i + 1) * c] = np.asarray(b, dtype=np.float32).reshape(c)
  fl[264:288] = np.asarray(tail_w, dtype=np.float32).reshape(c)
  fl[288] = float(np.asarray(tail_b, dtype=np.float32).reshape(-1)[0])
  return blob, fl, int(dtype)
