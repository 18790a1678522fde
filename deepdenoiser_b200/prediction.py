"""Inference driver - the reference's Prediction.py (TensorFlow/Prediction.py:188-520) without TensorFlow:
EXR directory -> features -> overlapping tiles -> Architecture.predict (all tiles batched along N) -> stitch ->
combine passes -> .npy files.

Kept from the reference: the CLI flags, the tile grid (tile 128, overlap 14, first / last tile clamped to the
border, Prediction.py:259-310), which part of every tile is kept (:396-427), the pass combination
color * (direct + indirect) and the 18 output files (:443-511), the .npy float32 HWC output format read by
Blender/NPYImporter.py.  Not kept: ./tmp.tfrecords, the Estimator, one session.run per tile.
Multi-GPU: tiles are independent, so ranks take interleaved tiles and rank 0 gathers (no data-path collective).
"""
import math
import os

import numpy as np
import torch

from .Naming import Naming
from .RenderPasses import RenderPasses


# ------------------------------------------------------------------------------------------------ checkpoints
def latest_checkpoint(model_directory, base_directory=None):
  """The newest `ckpt-<step>.npz` written by Training.py in the architecture's model_directory (tf.estimator.Estimator restores
  the latest checkpoint of model_dir for predict(), Prediction.py:327-331).  The directory is tried relative to the JSON's
  directory first (where Training.py writes), then relative to the working directory (the reference's behaviour)."""
  import glob
  for root in ([base_directory] if base_directory else []) + [os.getcwd()]:
    found = glob.glob(os.path.join(root, model_directory, "ckpt-*.npz"))
    if found:
      return max(found, key=lambda path: int(os.path.basename(path)[len("ckpt-"):-len(".npz")]))
  return None


def load_checkpoint_weights(path):
  """{TF variable name: array} of a Training.py checkpoint (Adam moments and the step counter are dropped)."""
  z = np.load(path)
  return {k: z[k] for k in z.files if k != "step" and not k.startswith("adam_m/") and not k.startswith("adam_v/")}


# ------------------------------------------------------------------------------------------------ EXR input
def exr_files(directory):
  """OpenEXRDirectory._exr_files (OpenEXRDirectory.py:118-124)."""
  return sorted(os.path.join(directory, f) for f in os.listdir(directory) if f.endswith(".exr"))


def load_exr(exr_path):
  """OpenEXRDirectory._load_exr (OpenEXRDirectory.py:126-152): decode from memory (utf-8 paths), BGR -> RGB, float32."""
  os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
  import cv2
  with open(exr_path, "rb") as stream:
    data = np.frombuffer(stream.read(), dtype=np.uint8)
  image = cv2.imdecode(data, cv2.IMREAD_UNCHANGED)
  if image is None:
    raise IOError("could not decode '%s'" % exr_path)
  if image.ndim == 2:
    image = np.repeat(image[..., None], 3, axis=2)
  image = image[..., :3][..., ::-1]           # BGR(A) -> RGB
  return np.ascontiguousarray(image, dtype=np.float32)


def save_exr(exr_path, image_rgb):
  """Writes a float32 RGB EXR (used by tests / synthetic fixtures; Blender writes the real ones)."""
  os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
  import cv2
  bgr = np.ascontiguousarray(image_rgb[..., ::-1], dtype=np.float32)
  if not cv2.imwrite(exr_path, bgr, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT]):
    raise IOError("could not write '%s'" % exr_path)


def find_pass_file(files, pass_name):
  """The dataset code matches '_<Pass>_' (OpenEXRDirectory.py:36,63); Prediction.py:230-231 matches any path that
  merely CONTAINS the pass name ('Normal' also hits 'Screen Space Normal').  Strict rule first, reference rule as
  the fallback."""
  token = "_" + pass_name + "_"
  for f in files:
    if token in os.path.basename(f):
      return f
  for f in files:
    if pass_name in f:
      return f
  return None


def load_features(architecture, directory):
  """{'source_image/0/<Pass>': float32 [H,W,3]} for every pass the architecture needs (Prediction.py:220-256);
  passes with load_data == False are the constants the reference feeds."""
  files = exr_files(directory)
  features, height, width = {}, None, None
  pending = []
  for fp in architecture.required_features():
    key = Naming.source_feature_name(fp.name, index=0)
    if not fp.load_data:
      pending.append((key, fp))
      continue
    path = find_pass_file(files, fp.name)
    if path is None:
      raise IOError("Image for '%s' could not be loaded or does not exist." % fp.name)
    image = load_exr(path)
    if height is None:
      height, width = image.shape[0], image.shape[1]
    elif image.shape[:2] != (height, width):
      raise ValueError("'%s' is %dx%d, expected %dx%d" % (path, image.shape[0], image.shape[1], height, width))
    features[key] = image
  for key, fp in pending:
    features[key] = fp.synthetic_source(1, height, width)[0].numpy()
  return features, height, width


# ------------------------------------------------------------------------------------------------ tiling
class Tile:
  __slots__ = ("y", "x", "size", "crop", "dest")

  def __init__(self, y, x, size, crop, dest):
    self.y, self.x, self.size = y, x, size
    self.crop = crop      # (y0, y1, x0, x1) inside the tile: the part that is kept
    self.dest = dest      # (y0, y1, x0, x1) in the image


def _axis_tiles(n, tile, overlap):
  """One axis of the reference's grid (Prediction.py:269-310, 396-427): stride tile - 2*overlap, first tile at 0,
  last tile at n - tile; kept = interior minus `overlap` on shared sides, the last tile keeps the remainder."""
  delta = tile - 2 * overlap
  count = int(math.ceil((n - 2 * overlap - 2 * delta) / delta)) + 2
  count = max(count, 1)
  out = []
  for i in range(count):
    if i == 0:
      lo = 0
    elif i == count - 1:
      lo = n - tile
    else:
      lo = i * delta
    c0, c1 = 0, tile
    if count == 1:
      pass
    elif i == 0:
      c1 = tile - overlap
    elif i == count - 1:
      existing = overlap + (count - 1) * delta
      c0 = tile - (n - existing)
    else:
      c0, c1 = overlap, tile - overlap
    out.append((lo, c0, c1))
  return out


def tile_grid(height, width, tile_size=128, tile_overlap_size=14):
  """All tiles of an image, row-major.  Small images shrink the tile (Prediction.py:259-266)."""
  smaller = min(height, width)
  if smaller < 16:
    raise ValueError("The image needs to have at least a side length of 16 pixels.")
  if smaller < tile_size:
    ratio = tile_overlap_size / tile_size
    tile_size = smaller
    tile_overlap_size = int(tile_size * ratio)
  tiles = []
  for (y, cy0, cy1) in _axis_tiles(height, tile_size, tile_overlap_size):
    for (x, cx0, cx1) in _axis_tiles(width, tile_size, tile_overlap_size):
      tiles.append(Tile(y, x, tile_size, (cy0, cy1, cx0, cx1), (y + cy0, y + cy1, x + cx0, x + cx1)))
  return tiles, tile_size, tile_overlap_size


def tile_table(tiles):
  """int32 [T][6] = {y, x, crop_y0, crop_y1, crop_x0, crop_x1}: the device table of dd_tiles_gather / dd_tiles_scatter."""
  return torch.tensor([[t.y, t.x, t.crop[0], t.crop[1], t.crop[2], t.crop[3]] for t in tiles], dtype=torch.int32).reshape(-1, 6)


def cut_tiles(image, tiles, ctx=None):
  """[H,W,C] tensor -> [T,size,size,C] batch of tiles.  CUDA images are cut by one dd_tiles_gather launch."""
  if ctx is not None and image.is_cuda:
    import ctypes
    from . import _lib
    image = image.contiguous()
    out = torch.empty((len(tiles), tiles[0].size, tiles[0].size, image.shape[2]), dtype=image.dtype, device=image.device)
    table = tile_table(tiles).to(image.device)
    ctx.call("dd_tiles_gather", ctypes.byref(_lib.desc(image.unsqueeze(0))), ctypes.c_void_p(table.data_ptr()),
             ctypes.byref(_lib.desc(out)))
    return out
  return torch.stack([image[t.y:t.y + t.size, t.x:t.x + t.size] for t in tiles], dim=0)


def stitch_tiles(batch, tiles, height, width, ctx=None, out=None):
  """[T,size,size,C] predictions -> [H,W,C] image, keeping each tile's crop (Prediction.py:384-441).  CUDA batches are
  pasted by one dd_tiles_scatter launch.  `out`: paste into this image instead of a new one (a rank's share of the tiles)."""
  if out is None:
    out = torch.empty((height, width, batch.shape[3]), dtype=batch.dtype, device=batch.device)
  if ctx is not None and batch.is_cuda:
    import ctypes
    from . import _lib
    batch = batch.contiguous()
    table = tile_table(tiles).to(batch.device)
    ctx.call("dd_tiles_scatter", ctypes.byref(_lib.desc(batch)), ctypes.c_void_p(table.data_ptr()),
             ctypes.byref(_lib.desc(out.unsqueeze(0))))
    return out
  for i, t in enumerate(tiles):
    cy0, cy1, cx0, cx1 = t.crop
    dy0, dy1, dx0, dx1 = t.dest
    out[dy0:dy1, dx0:dx1] = batch[i, cy0:cy1, cx0:cx1]
  return out


# ------------------------------------------------------------------------------------------------ prediction
def predict_image(architecture, features, height, width, tile_size=128, tile_overlap_size=14, full_frame=False,
                  tiles_per_batch=64, rank=0, world_size=1, predict_fn=None):
  """Denoises one image.  features: {'source_image/0/<Pass>': [H,W,C] array / tensor}.  Returns
  {'prediction/<Pass>': float32 [H,W,C] torch tensor} (largest scale only, like Prediction.model_fn :180-185).

  full_frame=True skips the tiling (height and width must be divisible by 2^sampling steps): the reference tiles only
  to bound TensorFlow's memory; the results differ at tile borders where the tiles truncate the receptive field."""
  dev = None
  ctx = None
  if predict_fn is None:
    # product path: frame resident on the device, tiles cut / pasted by dd_tiles_gather / dd_tiles_scatter
    architecture._ensure_device()
    ctx = architecture.ctx
    predict_fn = lambda f: architecture.predict(f)[0]   # noqa: E731
  tensors = {}
  for k, v in features.items():
    t = torch.as_tensor(v)
    if t.dim() == 2:
      t = t.unsqueeze(-1)
    if ctx is not None:
      t = t.to(ctx.device, torch.float32)      # the frame is uploaded ONCE; tiles are cut and pasted on the device
    tensors[k] = t
  if full_frame:
    out = predict_fn({k: v.unsqueeze(0) for k, v in tensors.items()})
    return {k: v[0] for k, v in out.items()}
  tiles, tile_size, tile_overlap_size = tile_grid(height, width, tile_size, tile_overlap_size)
  mine = list(range(rank, len(tiles), world_size))
  results = {}
  for b0 in range(0, len(mine), tiles_per_batch):
    idx = mine[b0:b0 + tiles_per_batch]
    batch_tiles = [tiles[i] for i in idx]
    batch = {k: cut_tiles(v, batch_tiles, ctx) for k, v in tensors.items()}
    out = predict_fn(batch)
    for k, v in out.items():
      results.setdefault(k, []).append(v)
      dev = v.device
  results = {k: torch.cat(v, dim=0) for k, v in results.items()}
  if world_size > 1:
    import torch.distributed as dist
    on_device = ctx is not None and dist.get_backend() == "nccl"
    if on_device:
      # every rank pastes ITS tiles into a zero frame; the kept crops partition the image, so a sum over the ranks (one NCCL
      # reduce per pass over NVLink) is the stitched frame, bit for bit.  (The host-side gather below pickles ~3 GB at 4K.)
      layout = [[(k, int(v.shape[3])) for k, v in results.items()]] if rank == 0 else [None]
      dist.broadcast_object_list(layout, src=0)                  # a rank may own no tile at all
      my_tiles = [tiles[i] for i in mine]
      out = {}
      for k, channels in layout[0]:
        frame = torch.zeros((height, width, channels), dtype=torch.float32, device=ctx.device)
        if my_tiles:
          stitch_tiles(results[k].float(), my_tiles, height, width, ctx, out=frame)
        dist.reduce(frame, dst=0)
        out[k] = frame
      return out if rank == 0 else None
    results = _gather_tiles(results, mine, len(tiles), rank, world_size, dev)
    if rank != 0:
      return None
  else:
    mine = list(range(len(tiles)))
  return {k: stitch_tiles(v, tiles, height, width, ctx if v.is_cuda else None) for k, v in results.items()}


def _gather_tiles(results, mine, n_tiles, rank, world_size, dev):
  """Rank 0 collects every rank's tiles (host-side gather: the outputs leave the GPU anyway)."""
  import torch.distributed as dist
  gathered = [None] * world_size
  dist.gather_object({k: v.cpu() for k, v in results.items()}, gathered if rank == 0 else None, dst=0)
  if rank != 0:
    return None
  merged = {}
  for k in results:
    first = gathered[0][k]
    full = torch.empty((n_tiles,) + tuple(first.shape[1:]), dtype=first.dtype)
    for r in range(world_size):
      full[list(range(r, n_tiles, world_size))] = gathered[r][k]
    merged[k] = full
  return merged


def combine_passes(predictions, ctx=None):
  """Prediction.py:443-481: lighting = color * (direct + indirect); image = sum of the lighting passes + volume + env +
  emission (alpha is ignored, as in the reference).  Missing passes contribute nothing.  With a Context and CUDA tensors the
  arithmetic runs in libdd_b200 (dd_muladd_fwd / dd_axpy) so the stitched passes never leave the device before the .npy write."""
  def get(name):
    return predictions.get(Naming.feature_prediction_name(name))

  on_device = ctx is not None and all(v.is_cuda for v in predictions.values())
  if on_device:
    import ctypes
    from . import _lib
    d = lambda t: ctypes.byref(_lib.desc(t.unsqueeze(0) if t.dim() == 3 else t))   # noqa: E731

  def lighting(color, direct, indirect):
    if not on_device:
      return color * (direct + indirect)
    color, direct, indirect = (t.float().contiguous() for t in (color, direct, indirect))
    out = torch.empty_like(color)
    ctx.call("dd_muladd_fwd", d(color), d(direct), d(indirect), d(out))
    return out

  def add(image, value):
    if image is None:
      return value.clone() if on_device else value
    if not on_device:
      return image + value
    ctx.call("dd_axpy", ctypes.c_float(1.0), d(value.float().contiguous()), d(image))
    return image

  image = None
  combined = {}
  for light in ("Diffuse", "Glossy", "Subsurface", "Transmission"):
    color, direct, indirect = get(light + " Color"), get(light + " Direct"), get(light + " Indirect")
    if color is None or direct is None or indirect is None:
      continue
    value = lighting(color, direct, indirect)
    combined[light] = value
    image = add(image, value)
  for name in (RenderPasses.VOLUME_DIRECT, RenderPasses.VOLUME_INDIRECT, RenderPasses.ENVIRONMENT, RenderPasses.EMISSION):
    value = get(name)
    if value is not None:
      image = add(image, value)
  return image, combined


def save_predictions(directory, predictions, image):
  """np.save('<input>/<Pass>.npy') for the combined image and every predicted pass (Prediction.py:487-511)."""
  written = []
  if image is not None:
    path = os.path.join(directory, RenderPasses.COMBINED + ".npy")
    np.save(path, image.detach().float().cpu().numpy())
    written.append(path)
  for key, value in predictions.items():
    name = key[len("prediction/"):]
    path = os.path.join(directory, name + ".npy")
    np.save(path, value.detach().float().cpu().numpy())
    written.append(path)
  return written
