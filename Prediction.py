#!/usr/bin/env python
"""python Prediction.py <architecture.json> --input DIR [--tile_size 128 --tile_overlap_size 14 --threads N
--data_format channels_first|channels_last] [--full_frame] [--dtype float16|float32]

Same command line as the reference's TensorFlow/Prediction.py:23-53 (extra flags are additions).  Denoises the
render passes found as EXR files in DIR with the B200 path and writes <Pass>.npy + Combined.npy next to them.
Multi-GPU: `python -m torch.distributed.run --nproc-per-node N Prediction.py ...` shards the tiles over N GPUs.
"""
import argparse
import json
import multiprocessing
import os
import sys

import torch

from deepdenoiser_b200 import prediction
from deepdenoiser_b200.Architecture import Architecture

parser = argparse.ArgumentParser(description="Prediction for the DeepDenoiser (B200-native path).")
parser.add_argument("json_filename", help="The json specifying all the relevant details.")
parser.add_argument("--input", type=str, help="Make a prediction for the files in this directory.")
parser.add_argument("--tile_size", default=128,
                    help="Width and heights of the tiles into which the image is split before denoising.")
parser.add_argument("--tile_overlap_size", default=14,
                    help="Border size of the tiles that is overlapping to avoid artifacts.")
parser.add_argument("--threads", default=multiprocessing.cpu_count() + 1, help="Number of threads to use.")
parser.add_argument("--data_format", type=str, default="channels_first", choices=["channels_first", "channels_last"],
                    help="Accepted for compatibility; the device layout is always NHWC.")
parser.add_argument("--full_frame", action="store_true",
                    help="Denoise the whole frame at once instead of 128x128 tiles (no 1.65x overlap overhead).")
parser.add_argument("--dtype", default=None, choices=["float16", "bfloat16", "float32"], help="Override the JSON's b200.dtype.")
parser.add_argument("--weights", default=None,
                    help=".npz with the variables (TF names); default: the latest Training.py checkpoint in the JSON's model_directory.")


def main(parsed_arguments):
  try:
    with open(parsed_arguments.json_filename, "r") as f:
      parsed_architecture_json = json.load(f)
  except Exception:  # noqa: BLE001
    print("Expected a valid architecture json file.")
    return 1
  assert os.path.isdir(parsed_arguments.input)
  if parsed_arguments.dtype:
    parsed_architecture_json.setdefault("b200", {})["dtype"] = parsed_arguments.dtype
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if torch.cuda.is_available():
    torch.cuda.set_device(local)        # every launch uses torch.cuda.current_stream() of THIS device
  if world > 1:
    import torch.distributed as dist
    if torch.cuda.is_available():
      dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
      dist.init_process_group("gloo")
  weights = None
  checkpoint = parsed_arguments.weights or prediction.latest_checkpoint(
      parsed_architecture_json.get("model_directory", ""), os.path.dirname(os.path.abspath(parsed_arguments.json_filename)))
  if checkpoint:
    weights = prediction.load_checkpoint_weights(checkpoint)
    if rank == 0:
      print("weights:", checkpoint)
  elif rank == 0:
    print("weights: seeded initialisation (no checkpoint in the model directory, no --weights)")
  architecture = Architecture(parsed_architecture_json, source_data_format="channels_last",
                              data_format=parsed_arguments.data_format, device=local, weights=weights)
  features, height, width = prediction.load_features(architecture, parsed_arguments.input)
  predictions = prediction.predict_image(
      architecture, features, height, width, int(parsed_arguments.tile_size), int(parsed_arguments.tile_overlap_size),
      full_frame=parsed_arguments.full_frame, rank=rank, world_size=world)
  if rank == 0:
    loaded = {fp.name for fp in architecture.feature_predictions if fp.load_data}
    predictions = {k: v for k, v in predictions.items() if k[len("prediction/"):] in loaded}
    image, _ = prediction.combine_passes(predictions, getattr(architecture, "ctx", None))
    for path in prediction.save_predictions(parsed_arguments.input, predictions, image):
      print(path)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  args, unparsed = parser.parse_known_args()
  sys.exit(main(args))
