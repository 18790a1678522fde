"""Debug probe: per-tensor gradient error of the fp16 training path vs the float64 oracle for several loss scales / sizes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_training as T  # noqa: E402
from deepdenoiser_b200.Architecture import Architecture  # noqa: E402
from deepdenoiser_b200.training import Trainer, TrainingSettings  # noqa: E402


def run(kind, h, w, precision, loss_scale=None, tuple_type="SINGLE"):
  j = T.small_example(filters=(16, 24, 32), n_convs=2, k=3, tuple_type=tuple_type)
  host, weights, features, targets = T.make_problem(j, n=2, h=h, w=w)
  trainer = Trainer(Architecture(j, weights=weights), TrainingSettings({"loss_difference": kind}), precision=precision,
                    loss_scale=loss_scale)
  trainer.forward({k: torch.from_numpy(v) for k, v in features.items()})
  loss = float(trainer.loss_and_gradient({k: torch.from_numpy(v) for k, v in targets.items()}).item())
  trainer.backward()
  want_loss, want, _ = T.oracle_loss_and_grads(j, weights, features, targets, kind=kind)
  got = trainer.gradients()
  rows = []
  for k, g in want.items():
    a, b = got[k].reshape(-1).astype(np.float64), g.reshape(-1)
    cos = float(a @ b / max(1e-30, np.linalg.norm(a) * np.linalg.norm(b)))
    rows.append((k.replace("reused_core_architecture/", "").replace("reused_compose_scales/", "cmp/"),
                 float(np.abs(a - b).max()) / max(1e-9, float(np.abs(b).max())), cos, float(np.abs(b).max())))
  print("== %s %dx%d %s scale=%s loss %.5f vs %.5f" % (kind, h, w, precision, loss_scale, loss, want_loss))
  for r in rows:
    print("   %-32s err %.3e cos %.5f |g| %.2e" % r)


if __name__ == "__main__":
  run("SQUARED", 16, 24, "float32")
  run("SQUARED", 16, 24, "float16")
  run("SQUARED", 16, 24, "float16", loss_scale=1.0)
  run("SQUARED", 16, 24, "float16", loss_scale=4096.0)
  run("SQUARED", 32, 48, "float16")
  run("ABSOLUTE", 16, 24, "float16")
