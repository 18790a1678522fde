#!/usr/bin/env python
"""python Training.py <training.json> [--validate] [--threads N] [--train_epochs E] [--validation_interval V]
[--data_format channels_first|channels_last] [--synthetic_tiles N --synthetic_tile_size S --steps_per_epoch K]

Same command line as the reference's TensorFlow/Training.py:33-61 (extra flags are additions).  The Estimator loop
(Training.py:1252-1284) becomes a plain loop: forward + loss + explicit backward kernels + TF-form Adam
(deepdenoiser_b200/training.py), one all-reduce of the flat gradient buffer per step when launched with
`python -m torch.distributed.run --nproc-per-node N Training.py ...` (tiles are sharded over the ranks).
Checkpoints (.npz: weights under their TF variable names + Adam moments + step) are written to the JSON's
model_directory every 500 steps (Training.py:1214-1218) and the latest one is resumed; scalars go to
<model_directory>/training_log.jsonl.
Data: like the reference, `<base_tfrecords_directory>/training/*.tfrecords.gz` + `training.json` written by
TFRecordsCreator.py (read without TensorFlow by deepdenoiser_b200/tfrecords.py; data augmentation on the GPU,
deepdenoiser_b200/augmentation.py).  When that directory does not exist, or with --synthetic_tiles, the script trains on
seeded synthetic render-pass tiles; --write_synthetic_tfrecords N first writes N such tiles in the reference's format.
"""
import argparse
import glob
import json
import multiprocessing
import os
import sys
import time

import torch

import numpy as np

from deepdenoiser_b200 import augmentation, synthetic, tfrecords
from deepdenoiser_b200.Architecture import Architecture
from deepdenoiser_b200.training import Trainer, TrainingSettings

parser = argparse.ArgumentParser(description="Training for the DeepDenoiser (B200-native path).")
parser.add_argument("json_filename", help="The json specifying all the relevant details.")
parser.add_argument("--validate", action="store_true", help="Perform a validation step.")
parser.add_argument("--threads", default=multiprocessing.cpu_count() + 1, help="Number of threads to use")
parser.add_argument("--train_epochs", type=int, default=10000, help="Number of epochs to train.")
parser.add_argument("--validation_interval", type=int, default=1, help="Number of epochs after which a validation is made.")
parser.add_argument("--data_format", type=str, default="channels_first", choices=["channels_first", "channels_last"],
                    help="Accepted for compatibility; the device layout is always NHWC.")
parser.add_argument("--synthetic_tiles", type=int, default=None, help="Global batch of synthetic tiles per step.")
parser.add_argument("--synthetic_tile_size", type=int, default=64)
parser.add_argument("--steps_per_epoch", type=int, default=10)
parser.add_argument("--checkpoint_steps", type=int, default=500)
parser.add_argument("--write_synthetic_tfrecords", type=int, default=0,
                    help="Write this many synthetic tiles as <base_tfrecords_directory>/training (reference format) and use them.")
parser.add_argument("--micro_batch", type=int, default=None,
                    help="Tiles per forward/backward pass on each rank (gradient accumulation); default: the whole per-rank batch.")
parser.add_argument("--precision", default=None, choices=["float32", "float16", "bfloat16"],
                    help="float16 (default): tensor-core path (fp16 activations, fp32 master weights); float32: exact path.")


def synthetic_batch(architecture, tiles, size, seed):
  noisy = synthetic.synthetic_features(architecture, tiles, size, size, seed=seed)
  clean = synthetic.synthetic_features(architecture, tiles, size, size, seed=seed + 7919)
  features = {k: torch.from_numpy(v) for k, v in noisy.items()}
  targets = {"target_image/" + fp.name: torch.from_numpy(clean["source_image/0/" + fp.name])
             for fp in architecture.feature_predictions if fp.load_data}
  return features, targets


def write_synthetic_dataset(architecture, directory, tiles, size, seed=4242):
  """Synthetic stand-in for TFRecordsCreator.py: `tiles` examples, one source per example, 16 samples per pixel."""
  noisy = synthetic.synthetic_features(architecture, tiles, size, size, seed=seed)
  clean = synthetic.synthetic_features(architecture, tiles, size, size, seed=seed + 7919)
  every = list(architecture.feature_predictions) + list(architecture.auxiliary_features)

  def examples():
    for e in range(tiles):
      feats = {}
      for fp in every:
        if not fp.load_data:
          continue
        feats["source_image/16/0/" + fp.name] = noisy["source_image/0/" + fp.name][e]
        if fp.is_target:
          feats["target_image/" + fp.name] = clean["source_image/0/" + fp.name][e]
      yield feats

  settings = {"tiles_height_width": size, "number_of_sources_per_example": 1, "source_samples_per_pixel_list": [16]}
  return tfrecords.write_tile_dataset(directory, "training", examples(), settings)


_DATASETS = {}


def tile_dataset(architecture, training_json, base, mode, settings_name=None):
  """One TileDataset per (mode, sidecar json), cached: its per-file record counts are computed once."""
  directory = os.path.join(base, training_json["base_tfrecords_directory"])
  key = (mode, settings_name)
  if key not in _DATASETS:
    _DATASETS[key] = tfrecords.TileDataset(
        os.path.join(directory, mode), os.path.join(directory, (settings_name or mode) + ".json"), architecture,
        number_of_source_index_tuples=int(training_json.get("number_of_source_index_tuples", 1)))
  return _DATASETS[key]


def validation_sets(architecture, training_json, base):
  """evaluation_jsons + extract_evaluation_json_information (Training.py:916-941, 1236-1250): one evaluation per
  `validation*.json` sidecar (statistics files excluded); with group_by_samples_per_pixel the records of `validation_<spp>.json`
  live under `validation/<spp>`, otherwise directly under `validation`."""
  directory = os.path.join(base, training_json["base_tfrecords_directory"])
  out = []
  for name in sorted(os.listdir(directory)):
    stem, ext = os.path.splitext(name)
    if not (stem.startswith("validation") and ext == ".json" and "statistics" not in stem and
            os.path.isfile(os.path.join(directory, name))):
      continue
    with open(os.path.join(directory, name), "r", encoding="utf-8") as f:
      spp = json.load(f)["source_samples_per_pixel_list"][0]
    sub = os.path.join("validation", str(spp))
    mode = sub if os.path.isdir(os.path.join(directory, sub)) else "validation"
    if os.path.isdir(os.path.join(directory, mode)):
      out.append((stem, tile_dataset(architecture, training_json, base, mode, stem)))
  return out


def validate(trainer, datasets, per_rank, rank, world, threads=0):
  """estimator.evaluate (Training.py:866-877): the mean loss over the validation records, no augmentation, no shuffle, every
  rank evaluates its share and the (sum, count) pair is all-reduced.  Returns {name: mean loss}."""
  results = {}
  for name, dataset in datasets:
    total = torch.zeros(2, dtype=torch.float64, device=trainer.dev)
    for sources, targets in dataset.batches(per_rank, shuffle_seed=None, rank=rank, world=world, drop_remainder=False,
                                            threads=threads):
      n = next(iter(targets.values())).shape[0]
      trainer.forward({k: torch.from_numpy(v) for k, v in sources.items()})
      loss = trainer.loss_and_gradient({k: torch.from_numpy(v) for k, v in targets.items()})
      total[0] += loss.double().squeeze() * n
      total[1] += n
    if world > 1:
      import torch.distributed as dist
      dist.all_reduce(total)
    results[name] = float((total[0] / total[1].clamp(min=1)).item())
  return results


def tfrecord_batches(architecture, training_json, base, per_rank, rank, world, trainer, epoch_seed, threads=0):
  """input_fn_tfrecords (Training.py:728-850): records -> (sources, targets) examples -> device augmentation -> batches."""
  dataset = tile_dataset(architecture, training_json, base, "training")
  usage = augmentation.DataAugmentationUsage.from_json(training_json)
  augment = augmentation.DeviceAugmenter(trainer.ctx, usage)
  rng = np.random.default_rng(epoch_seed * 7919 + rank)
  for sources, targets in dataset.batches(per_rank, shuffle_seed=epoch_seed, rank=rank, world=world, threads=threads):
    draws = augmentation.draw(usage, per_rank, rng)
    yield augment(sources, targets, draws)


def main(parsed_arguments):
  with open(parsed_arguments.json_filename, "r") as f:
    training_json = json.load(f)
  base = os.path.dirname(os.path.abspath(parsed_arguments.json_filename))
  with open(os.path.join(base, training_json["architecture"]), "r") as f:      # relative to the training json (:953-961)
    architecture_json = json.load(f)
  architecture_json.setdefault("b200", {})["dtype"] = "float32"               # the Trainer owns the training arithmetic mode
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)          # every launch below uses torch.cuda.current_stream() of THIS device
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  architecture = Architecture(architecture_json, source_data_format="channels_last",
                              data_format=parsed_arguments.data_format, device=local)
  settings = TrainingSettings(training_json)
  precision = parsed_arguments.precision or "float16"
  trainer = Trainer(architecture, settings, precision=precision)
  comm = None
  if world > 1:
    # the gradient exchange runs through libdd_b200's own NCCL communicator; torch.distributed carries the unique id only
    from deepdenoiser_b200 import _lib
    comm = _lib.Communicator(trainer.ctx, rank, world)
  model_dir = os.path.join(base, architecture.model_directory)
  os.makedirs(model_dir, exist_ok=True)
  checkpoints = sorted(glob.glob(os.path.join(model_dir, "ckpt-*.npz")), key=lambda p: int(p.split("-")[-1][:-4]))
  if checkpoints:
    trainer.load_checkpoint(checkpoints[-1])
    if rank == 0:
      print("resumed from", checkpoints[-1], "at step", trainer.step_count)
  global_tiles = parsed_arguments.synthetic_tiles or settings.batch_size
  if global_tiles % world:
    raise ValueError("the global batch (%d tiles) must be divisible by the number of ranks (%d)" % (global_tiles, world))
  per_rank = global_tiles // world
  log = open(os.path.join(model_dir, "training_log.jsonl"), "a") if rank == 0 else None
  size = parsed_arguments.synthetic_tile_size
  records_dir = os.path.join(base, training_json.get("base_tfrecords_directory", ""), "training")
  if parsed_arguments.write_synthetic_tfrecords and rank == 0:
    write_synthetic_dataset(architecture, os.path.dirname(records_dir), parsed_arguments.write_synthetic_tfrecords, size)
  if world > 1:
    dist.barrier()
  use_records = parsed_arguments.synthetic_tiles is None and os.path.isdir(records_dir)
  if rank == 0:
    print("data:", ("TFRecords under " + records_dir) if use_records else "seeded synthetic tiles")

  def epoch_batches(epoch):
    if use_records:
      for features, targets in tfrecord_batches(architecture, training_json, base, per_rank, rank, world, trainer, epoch + 1,
                                                threads=min(8, max(1, int(parsed_arguments.threads) // max(1, world)))):
        yield features, targets
    else:
      for _ in range(parsed_arguments.steps_per_epoch):
        yield synthetic_batch(architecture, per_rank, size, seed=1000003 * trainer.step_count + rank)

  for epoch in range(parsed_arguments.train_epochs):
    for features, targets in epoch_batches(epoch):
      if use_records:
        size = next(iter(features.values())).shape[1]
      t0 = time.perf_counter()
      loss = float(trainer.train_step(features, targets, world_size=world, comm=comm,
                                      micro_batch=parsed_arguments.micro_batch).item())
      dt = time.perf_counter() - t0
      skipped, scale_factor = trainer.update_loss_scale()      # the .item() above synchronised anyway
      if rank == 0:
        rec = {"step": trainer.step_count, "epoch": epoch, "loss": loss, "learning_rate": settings.learning_rate,
               "batch_size": global_tiles, "tile": size, "ranks": world, "seconds": dt, "precision": precision,
               "megapixels_per_second": global_tiles * size * size / 1e6 / dt}
        if precision == "float16":
          rec.update({"skipped_steps": skipped, "loss_scale_factor": scale_factor})
        log.write(json.dumps(rec) + "\n")
        log.flush()
        print(json.dumps(rec))
      if rank == 0 and trainer.step_count % parsed_arguments.checkpoint_steps == 0:
        trainer.save_checkpoint(os.path.join(model_dir, "ckpt-%d.npz" % trainer.step_count))
    if parsed_arguments.validate and (epoch + 1) % parsed_arguments.validation_interval == 0:
      if use_records:
        # the reference evaluates on <base_tfrecords_directory>/validation (Training.py:1236-1250, 1268-1282)
        results = validate(trainer, validation_sets(architecture, training_json, base), per_rank, rank, world,
                           threads=min(8, max(1, int(parsed_arguments.threads) // max(1, world))))
        if rank == 0:
          rec = {"validation_loss": results, "epoch": epoch, "step": trainer.step_count}
          log.write(json.dumps(rec) + "\n")
          log.flush()
          print(json.dumps(rec))
      else:
        features, targets = synthetic_batch(architecture, per_rank, size, seed=424243 + rank)
        trainer.forward(features)
        val = trainer.loss_and_gradient(targets).clone()
        if world > 1:
          dist.all_reduce(val)
          val /= world
        if rank == 0:
          print(json.dumps({"validation_loss": {"synthetic": float(val.item())}, "epoch": epoch, "step": trainer.step_count}))
  if rank == 0:
    trainer.save_checkpoint(os.path.join(model_dir, "ckpt-%d.npz" % trainer.step_count))
  if world > 1:
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  args, unparsed = parser.parse_known_args()
  sys.exit(main(args))
