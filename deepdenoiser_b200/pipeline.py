"""Frame pipeline: Architecture.predict over a sequence of host frames with the host->device upload of frame
i+1 and the device->host download of frame i-1 overlapped with the kernels of frame i (three CUDA streams,
double-buffered device inputs and pinned host outputs).

The reference feeds one tile per session.run and copies every result back synchronously
(Prediction.py:363-382); at 1080p a frame moves 514 MB up and 406 MB down, ~18 ms of PCIe time that would
otherwise serialise with ~75 ms of compute.  This is plumbing around the same public call
(Architecture.predict); it changes no arithmetic.
"""
import torch


class FramePipeline:

  def __init__(self, architecture, scale_index=0):
    self.arch = architecture
    self.scale_index = scale_index
    architecture._ensure_device()
    self.dev = architecture.ctx.device
    self.s_in = torch.cuda.Stream(device=self.dev)
    self.s_out = torch.cuda.Stream(device=self.dev)
    self._dev_in = [None, None]       # double-buffered device copies of the inputs
    self._host_out = [None, None]     # double-buffered pinned outputs
    self._in_free = [None, None]      # event: compute that read input set k has finished
    self._out_free = [None, None]     # event: download from output set k has finished (host may reuse after sync)

  def _upload(self, frame, k):
    """frame: dict of PINNED host tensors.  Enqueues the copies on the input stream; returns the ready event."""
    with torch.cuda.stream(self.s_in):
      if self._in_free[k] is not None:
        self.s_in.wait_event(self._in_free[k])
      if self._dev_in[k] is None:
        self._dev_in[k] = {name: torch.empty(t.shape, dtype=torch.float32, device=self.dev) for name, t in frame.items()}
      for name, t in frame.items():
        self._dev_in[k][name].copy_(t, non_blocking=True)
      ev = torch.cuda.Event()
      ev.record(self.s_in)
    return ev

  def run(self, frames, on_result=None):
    """frames: iterable of {'source_image/0/<Pass>': pinned float32 [N,H,W,C]}.  For every frame, `on_result(index,
    {'prediction/<Pass>': pinned host tensor})` is called once its download has completed (the tensors are reused two
    frames later).  Returns the number of frames."""
    compute = torch.cuda.current_stream(self.dev)
    it = iter(frames)
    pending = []                      # (index, set k, event)
    try:
      nxt = next(it)
    except StopIteration:
      return 0
    ready = self._upload(nxt, 0)
    i = 0
    while nxt is not None:
      k = i & 1
      cur_ready = ready
      try:
        nxt = next(it)
        ready = self._upload(nxt, (i + 1) & 1)      # overlaps with the kernels of frame i
      except StopIteration:
        nxt = None
      compute.wait_event(cur_ready)
      result = self.arch.predict(self._dev_in[k])[self.scale_index]
      done = torch.cuda.Event()
      done.record(compute)
      self._in_free[k] = done
      # download on the output stream
      with torch.cuda.stream(self.s_out):
        self.s_out.wait_event(done)
        if self._host_out[k] is None:
          self._host_out[k] = {name: torch.empty(t.shape, dtype=t.dtype).pin_memory() for name, t in result.items()}
        elif self._out_free[k] is not None:
          self._out_free[k].synchronize()          # the consumer of two frames ago must be done with these buffers
        for name, t in result.items():
          t.record_stream(self.s_out)
          self._host_out[k][name].copy_(t, non_blocking=True)
        out_ev = torch.cuda.Event()
        out_ev.record(self.s_out)
      self._out_free[k] = out_ev
      pending.append((i, k, out_ev))
      # hand finished frames to the consumer (keeps at most one download in flight behind the compute)
      while len(pending) > 1:
        j, kk, ev = pending.pop(0)
        ev.synchronize()
        if on_result is not None:
          on_result(j, self._host_out[kk])
      i += 1
    for j, kk, ev in pending:
      ev.synchronize()
      if on_result is not None:
        on_result(j, self._host_out[kk])
    return i
