"""Generates tests/golden/<case>.npz: outputs of the float64 numpy oracle (oracle/reference_model.py) on the
seeded synthetic inputs / weights of tests/cases.py.

These are regression pins of the restated oracle; the outputs of the reference's own code for the same cases are
tests/golden/refshim_<case>.npz (tests/golden/make_reference_golden.py), which the oracle reproduces to 1e-10.
Re-run after an intentional oracle change:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
from oracle import np_ops, reference_model  # noqa: E402


def digest(arrays):
  h = hashlib.sha256()
  for k in sorted(arrays):
    h.update(k.encode())
    h.update(np.ascontiguousarray(arrays[k]).tobytes())
  return h.hexdigest()


def main():
  for name in cases.GOLDEN_CASES:
    j, arch, weights, features = cases.build(name)
    oracle = reference_model.Architecture(j, ops=np_ops, dtype=np.float64, weights=weights)
    out = oracle.predict_numpy(features)
    payload = {"inputs_sha256": np.array(digest(features)), "weights_sha256": np.array(digest(weights))}
    for s, d in enumerate(out):
      for k, v in d.items():
        payload["%d|%s" % (s, k)] = v.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
    print(name, len(out), "scales", len(out[0]), "passes")


if __name__ == "__main__":
  main()
