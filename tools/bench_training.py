#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4]: tiles of 256x256x32-ch, U-Net [64,96,128]x4 KPCN K=5, SMAPE loss of
TrainingExample.json) on the tensor-core (fp16) or exact (fp32) path.

  python tools/bench_training.py [--tiles T] [--size S] [--steps K] [--warmup W] [--precision float16|bfloat16|float32] [--arch unet32]
  python -m torch.distributed.run --nproc-per-node N tools/bench_training.py ...     # T tiles PER RANK (weak scaling)

One step = forward (all tuple passes) + loss + backward + gradient all-reduce (N > 1) + Adam + weight repack, inputs resident
in HBM.  Prints one JSON line: tiles/s, trained megapixels/s, achieved TFLOP/s (3 x forward MACs of the conv stack) and the
split of the step into phases (CUDA events)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepdenoiser_b200 import synthetic  # noqa: E402
from deepdenoiser_b200.Architecture import Architecture  # noqa: E402
from deepdenoiser_b200.training import Trainer, TrainingSettings  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--tiles", type=int, default=4)
  ap.add_argument("--size", type=int, default=256)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=2)
  ap.add_argument("--precision", default="float16")
  ap.add_argument("--arch", default="unet32")
  ap.add_argument("--out", default=None)
  args = ap.parse_args()
  rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
  torch.cuda.set_device(local)
  dist = None
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  j = synthetic.baseline_architecture_json(args.arch)
  j["b200"] = {"dtype": "float32"}
  arch = Architecture(j, device=local)
  trainer = Trainer(arch, TrainingSettings({"learning_rate": 1e-4}), precision=args.precision)
  noisy = synthetic.synthetic_features(arch, args.tiles, args.size, args.size, seed=77 + rank)
  clean = synthetic.synthetic_features(arch, args.tiles, args.size, args.size, seed=7996 + rank)
  feats = {k: torch.from_numpy(v).cuda() for k, v in noisy.items()}
  targets = {"target_image/" + fp.name: torch.from_numpy(clean["source_image/0/" + fp.name]).cuda()
             for fp in arch.feature_predictions if fp.load_data}
  ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
  phases = {"forward": 0.0, "loss": 0.0, "backward": 0.0, "optimizer": 0.0}
  losses = []

  def step(timed):
    e = [ev() for _ in range(5)] if timed else None
    if timed: e[0].record()
    trainer.forward(feats)
    if timed: e[1].record()
    loss = trainer.loss_and_gradient(targets)
    if timed: e[2].record()
    trainer.backward()
    if timed: e[3].record()
    scale = 1.0
    if dist is not None:
      dist.all_reduce(trainer.grad)
      scale = 1.0 / world
    trainer.apply_gradients(scale)
    if timed:
      e[4].record()
      torch.cuda.synchronize()
      for i, k in enumerate(phases):
        phases[k] += e[i].elapsed_time(e[i + 1])
    return loss

  for _ in range(max(args.warmup, 1)):
    first = step(False)
  losses.append(float(first.item()))
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  l0 = trainer.ctx.launch_count()
  t0, t1 = ev(), ev()
  t0.record()
  for _ in range(args.steps):            # the timed region: K steps back to back, no host synchronisation inside
    last = step(False)
  t1.record()
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  losses.append(float(last.item()))
  launches = (trainer.ctx.launch_count() - l0) // args.steps
  ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device="cuda")
  if dist is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms = float(ms.item())
  step(True)                              # one extra instrumented step for the phase split
  phases = {k: v for k, v in phases.items()}
  if rank == 0:
    tuples = len(arch.feature_prediction_tuples)
    px = args.tiles * args.size * args.size
    mac = arch.spec.mac_per_pixel(arch.features_per_tuple)
    flops = 3.0 * 2.0 * mac * px * tuples * world
    line = {"metric": "training step (tiles of %dx%dx%d-ch, %s)" % (args.size, args.size, arch.number_of_input_channels, args.arch),
            "precision": args.precision, "n_gpus": world, "tiles_per_gpu": args.tiles, "tuple_passes": tuples,
            "ms_per_step": ms, "tiles_per_s": world * args.tiles / ms * 1e3, "megapixels_per_s": world * px / 1e6 / ms * 1e3,
            "tflops": flops / ms / 1e9, "phases_ms": phases,
            "gpu_launches_per_step": int(launches), "loss_first_last": [losses[0], losses[-1]],
            "max_memory_gb": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(line))
    if args.out:
      with open(args.out, "w") as f:
        f.write(json.dumps(line) + "\n")
  if dist is not None:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
