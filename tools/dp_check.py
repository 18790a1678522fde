"""Data-parallel correctness on 2+ GPUs (run under torchrun): every rank trains on its shard of one global batch;
after the all-reduce the averaged gradient / the updated weights must equal the single-process result on the whole
batch (up to fp32 summation order: the first Adam step is lr * g / |g|, so 2e-5 = 2 % of one step of lr 1e-3).  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdenoiser_b200 import synthetic  # noqa: E402
from deepdenoiser_b200.Architecture import Architecture  # noqa: E402
from deepdenoiser_b200.training import Trainer, TrainingSettings  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
j = synthetic.example_architecture_json()
j["architecture"]["core_architecture"]["number_of_filters_for_convolution_blocks"] = [16, 24, 32]
j["architecture"]["core_architecture"]["number_of_convolutions_per_block"] = 2
j["b200"] = {"dtype": "float32"}
host = Architecture(j)
weights = synthetic.randomize_biases(host.weights)
tiles, size = 2 * world, 32
noisy = synthetic.synthetic_features(host, tiles, size, size, seed=5)
clean = synthetic.synthetic_features(host, tiles, size, size, seed=6)
feats = {k: torch.from_numpy(v) for k, v in noisy.items()}
targs = {"target_image/" + fp.name: torch.from_numpy(clean["source_image/0/" + fp.name]) for fp in host.feature_predictions}
per = tiles // world
shard = slice(rank * per, (rank + 1) * per)
trainer = Trainer(Architecture(j, weights=weights, device=local), TrainingSettings())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# the exchange goes through libdd_b200's own NCCL communicator (dd_comm_*); torch.distributed only carries the unique id
from deepdenoiser_b200 import _lib  # noqa: E402
comm = _lib.Communicator(trainer.ctx, rank, world)
loss = trainer.train_step({k: v[shard] for k, v in feats.items()}, {k: v[shard] for k, v in targs.items()}, world_size=world,
                          comm=comm)
torch.cuda.synchronize()
e0.record()
comm.all_reduce_sum(trainer.grad)
e1.record()
torch.cuda.synchronize()
if rank == 0:
  single = Trainer(Architecture(j, weights=weights, device=local), TrainingSettings())
  full_loss = single.train_step(feats, targs, world_size=1)
  dw = float((single.theta - trainer.theta).abs().max())
  print(json.dumps({"ranks": world, "dp_loss": float(loss), "single_process_loss": float(full_loss),
                    "max_weight_difference_after_one_step": dw, "allreduce_ms_flat_grad": e0.elapsed_time(e1),
                    "grad_elements": trainer.count, "exchange": "dd_comm_allreduce_sum_f32 (NCCL through the C ABI)", "ok": bool(dw < 2e-5 and abs(float(loss) - float(full_loss)) < 1e-4)}))
comm.close()
dist.barrier()
dist.destroy_process_group()
