#!/usr/bin/env python
"""cfg4 (BASELINE.json configs[3]): MultiScalePrediction U-Net, 3840x2160 tiled inference, 32-channel render-pass stack.

  python tools/bench_prediction.py [--height 2160 --width 3840] [--tiles 128,256,512,0] [--repeats 2]
  python -m torch.distributed.run --nproc-per-node N tools/bench_prediction.py ...      # tiles sharded over the ranks

Times deepdenoiser_b200.prediction.predict_image (the body of Prediction.py) end to end from HOST features: one upload of the
frame, tiles cut / pasted on the device, all 17 tuple passes per tile batch, pass combination, download of the 18 outputs.
Tile size 0 = full frame (no tiling).  The reference's own setting is 128 with overlap 14 (Prediction.py:34-41)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdenoiser_b200 import prediction, synthetic  # noqa: E402
from deepdenoiser_b200.Architecture import Architecture  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--height", type=int, default=2160)
  ap.add_argument("--width", type=int, default=3840)
  ap.add_argument("--tiles", default="128,256,512,0")
  ap.add_argument("--repeats", type=int, default=2)
  ap.add_argument("--stream-frames", type=int, default=4, help="full frames streamed through FramePipeline per rank (0 = skip)")
  ap.add_argument("--out", default=None)
  args = ap.parse_args()
  rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
  torch.cuda.set_device(local)
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  j = synthetic.baseline_architecture_json("unet32")
  j["b200"] = {"dtype": "float16"}
  weights = synthetic.randomize_biases(Architecture(j).weights)
  arch = Architecture(j, weights=weights, device=local)
  feats = synthetic.synthetic_features(arch, 1, args.height, args.width, seed=99)
  feats = {k: torch.from_numpy(v[0]).pin_memory() for k, v in feats.items()}
  mp = args.height * args.width / 1e6
  results, full, stream_rec = [], None, None
  pinned_out = {}

  def download(name, t):
    buf = pinned_out.get(name)
    if buf is None or buf.shape != t.shape:
      buf = torch.empty(t.shape, dtype=torch.float32).pin_memory()
      pinned_out[name] = buf
    buf.copy_(t, non_blocking=True)
    return buf
  # full frames as a STREAM (Prediction.py walks a directory of frames): FramePipeline overlaps the upload of frame i+1 and the
  # download of frame i-1 with the kernels of frame i; every rank streams its own frames (no exchange)
  if args.stream_frames > 0:
    from deepdenoiser_b200.pipeline import FramePipeline
    pipe = FramePipeline(arch)
    frame = {k: v.unsqueeze(0).contiguous().pin_memory() for k, v in feats.items()}
    pipe.run([frame] * 2)                                        # warm-up: buffers, weights, caches
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    t0 = time.perf_counter()
    done = pipe.run([frame] * args.stream_frames)
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    dt = time.perf_counter() - t0
    rec = {"tile": 0, "mode": "stream of %d full frames per rank through FramePipeline (copies overlapped)" % done,
           "seconds_per_frame": dt / done, "megapixels_per_s": world * done * mp / dt}
    if rank == 0:
      stream_rec = rec
      print(json.dumps(rec), flush=True)
    arch.network.release_buffers()
    torch.cuda.empty_cache()
  for tile in [int(t) for t in args.tiles.split(",") if t != ""]:
    overlap = max(2, int(round(tile * 14 / 128))) if tile else 0
    times = []
    out = None
    for _ in range(args.repeats + 1):
      torch.cuda.synchronize()
      if world > 1:
        dist.barrier()
      t0 = time.perf_counter()
      out = prediction.predict_image(arch, feats, args.height, args.width, tile_size=tile or 128, tile_overlap_size=overlap,
                                     full_frame=(tile == 0), tiles_per_batch=max(1, (1 << 22) // max(1, (tile or 128) ** 2)),
                                     rank=rank, world_size=world)
      if out is not None:
        image, _ = prediction.combine_passes(out, arch.ctx)
        host = {k: download(k, v) for k, v in out.items()}
        host["Combined"] = download("Combined", image)
      torch.cuda.synchronize()
      if world > 1:
        dist.barrier()
      times.append(time.perf_counter() - t0)
    best = min(times[1:])
    rec = {"tile": tile, "overlap": overlap, "seconds": best, "megapixels_per_s": mp / best}
    if rank == 0:
      if tile == 0:
        full = out
      results.append((rec, out))
      print(json.dumps(rec), flush=True)
    arch.network.release_buffers()
    torch.cuda.empty_cache()
  if rank == 0:
    summary = {"metric": "tiled inference, %dx%d, U-Net KPCN 32-ch (cfg4)" % (args.width, args.height), "n_gpus": world,
               "results": [r for r, _ in results], "stream": stream_rec}
    if full is not None:
      for rec, out in results:
        if rec["tile"]:
          rec["max_abs_diff_vs_full_frame"] = max(float((out[k].cpu() - full[k].cpu()).abs().max()) for k in out)
    print(json.dumps(summary))
    if args.out:
      with open(args.out, "w") as f:
        f.write(json.dumps(summary) + "\n")
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
