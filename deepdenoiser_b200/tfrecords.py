"""TensorFlow-free reader / writer of the reference's training data.

What TFRecordsCreator.py writes (TFRecordsCreator.py:143-158, 221-252) and Training.input_fn_tfrecords reads
(Training.py:502-524, 728-792):
  * files `<name>_<i>.tfrecords.gz`: a GZIP stream of TFRecord frames
        uint64 length | uint32 masked_crc32c(length) | payload | uint32 masked_crc32c(payload)        (little endian)
    masked_crc(x) = rotr15(crc32c(x)) + 0xa282ead8  [TensorFlow core/lib/hash/crc32c.h, core/lib/io/record_writer.cc];
  * payload = a serialized tf.train.Example protobuf: Example{1: Features{1: map<string, Feature>}},
    Feature{1: BytesList{1: repeated bytes} | 2: FloatList{1: packed float} | 3: Int64List{1: packed varint}};
    every feature here is a single bytes value = the raw float32 HWC tile (`image[x1:x2, y1:y2].tostring()`), keyed
    `source_image/<spp>/<index>/<Pass>` and `target_image/<Pass>` (Naming.py:57-81);
  * sidecar `<mode>.json`: {tiles_height_width, number_of_sources_per_example, source_samples_per_pixel_list}
    (TFRecordsCreator.py:164-178).
TensorFlow itself is not available here, so the framing and the protobuf wire format follow their published specifications;
tests/test_tfrecords.py pins them with known-answer vectors (CRC-32C check value, hand-encoded Example bytes).
The CRC runs in C (libdd_b200.so::dd_crc32c, host code); everything else is numpy / stdlib."""
import ctypes
import glob
import gzip
import json
import os
import random
import struct

import numpy as np

from .Architecture import FeaturePredictionType
from .Naming import Naming

_MASK_DELTA = 0xA282EAD8


def _crc32c(data):
  from . import _lib
  lib = _lib.load_library()
  buf = bytes(data)
  return int(lib.dd_crc32c(buf, len(buf))) & 0xFFFFFFFF


def masked_crc32c(data):
  crc = _crc32c(data)
  return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ protobuf wire format
def _read_varint(buf, pos):
  result, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _write_varint(value):
  out = bytearray()
  while True:
    b = value & 0x7F
    value >>= 7
    if value:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _fields(buf):
  """Yields (field_number, wire_type, value) of one message; length-delimited values are memoryviews (no copy)."""
  view = memoryview(buf)
  pos, end = 0, len(view)
  while pos < end:
    key, pos = _read_varint(view, pos)
    number, wire = key >> 3, key & 7
    if wire == 0:
      value, pos = _read_varint(view, pos)
    elif wire == 1:
      value, pos = view[pos:pos + 8], pos + 8
    elif wire == 2:
      size, pos = _read_varint(view, pos)
      value, pos = view[pos:pos + size], pos + size
    elif wire == 5:
      value, pos = view[pos:pos + 4], pos + 4
    else:
      raise ValueError("unsupported protobuf wire type %d" % wire)
    yield number, wire, value


def _parse_feature(buf):
  for number, wire, value in _fields(buf):
    if wire != 2:
      continue
    if number == 1:      # BytesList
      return [v for n, w, v in _fields(value) if n == 1]
    if number == 2:      # FloatList (packed or repeated fixed32)
      out = []
      for n, w, v in _fields(value):
        if n == 1:
          out.append(np.frombuffer(v, dtype="<f4"))
      return np.concatenate(out) if out else np.zeros(0, np.float32)
    if number == 3:      # Int64List
      out = []
      for n, w, v in _fields(value):
        if n != 1:
          continue
        if w == 0:
          out.append(v)
        else:
          pos = 0
          while pos < len(v):
            x, pos = _read_varint(v, pos)
            out.append(x)
      return np.array([x - (1 << 64) if x >= (1 << 63) else x for x in out], dtype=np.int64)
  return []


def parse_example(payload):
  """tf.parse_single_example without a schema: {feature name: list of bytes-like | float32 array | int64 array}."""
  result = {}
  for number, wire, features in _fields(payload):
    if number != 1 or wire != 2:
      continue
    for n2, w2, entry in _fields(features):
      if n2 != 1 or w2 != 2:
        continue
      key, value = None, None
      for n3, w3, v3 in _fields(entry):
        if n3 == 1:
          key = bytes(v3).decode("utf-8")
        elif n3 == 2:
          value = _parse_feature(v3)
      if key is not None:
        result[key] = value
  return result


def _ld(number, payload):
  return _write_varint((number << 3) | 2) + _write_varint(len(payload)) + payload


def serialize_example(features):
  """{name: bytes | float array | int array} -> serialized tf.train.Example (keys sorted: deterministic map order)."""
  entries = b""
  for key in sorted(features):
    value = features[key]
    if isinstance(value, (bytes, bytearray, memoryview)):
      feature = _ld(1, _ld(1, bytes(value)))
    else:
      arr = np.asarray(value)
      if arr.dtype.kind == "f":
        feature = _ld(2, _ld(1, arr.astype("<f4").tobytes()))
      else:
        feature = _ld(3, _ld(1, b"".join(_write_varint(int(x) & ((1 << 64) - 1)) for x in arr.reshape(-1))))
    entries += _ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feature))
  return _ld(1, entries)


# ------------------------------------------------------------------------------------------------ TFRecord framing
def read_records(path, verify_crc=True):
  """Yields the payloads of a .tfrecords or .tfrecords.gz file (tf.data.TFRecordDataset(compression_type='GZIP'),
  Training.py:830)."""
  opener = gzip.open if path.endswith(".gz") else open
  with opener(path, "rb") as f:
    while True:
      header = f.read(12)
      if not header:
        return
      if len(header) != 12:
        raise IOError("%s: truncated record header" % path)
      length, length_crc = struct.unpack("<QI", header)
      if verify_crc and masked_crc32c(header[:8]) != length_crc:
        raise IOError("%s: corrupted record length" % path)
      payload = f.read(length)
      footer = f.read(4)
      if len(payload) != length or len(footer) != 4:
        raise IOError("%s: truncated record" % path)
      if verify_crc and masked_crc32c(payload) != struct.unpack("<I", footer)[0]:
        raise IOError("%s: corrupted record payload" % path)
      yield payload


def write_records(path, payloads):
  """tf.python_io.TFRecordWriter + TFRecordsWriter._compress (TFRecordsCreator.py:221-252): .gz suffix => GZIP."""
  opener = gzip.open if path.endswith(".gz") else open
  with opener(path, "wb") as f:
    for payload in payloads:
      header = struct.pack("<Q", len(payload))
      f.write(header)
      f.write(struct.pack("<I", masked_crc32c(header)))
      f.write(payload)
      f.write(struct.pack("<I", masked_crc32c(payload)))


# ------------------------------------------------------------------------------------------------ index tuples
def source_index_tuples(number_of_sources_per_example, number_of_source_index_tuples, number_of_sources_per_target, rng=random):
  """Training.source_index_tuples (Training.py:879-913)."""
  if number_of_sources_per_example < number_of_sources_per_target:
    raise Exception("The source index tuples contain unique indices. That is not possible if there are fewer source examples "
                    "than indices per tuple.")
  index_tuples = []
  if number_of_sources_per_target == 1:
    complete, remaining = divmod(number_of_source_index_tuples, number_of_sources_per_example)
    for _ in range(complete):
      for index in range(number_of_sources_per_example):
        index_tuples.append([index])
    for _ in range(remaining):
      index_tuples.append([rng.randint(0, number_of_sources_per_example - 1)])
  else:
    if number_of_sources_per_target > 2:
      raise Exception("More than two source inputs are currently not supported!")
    for _ in range(number_of_source_index_tuples):
      t = []
      while len(t) < number_of_sources_per_target:
        index = rng.randint(0, number_of_sources_per_example - 1)
        if index not in t:
          t.append(index)
      index_tuples.append(t)
  required = sorted({i for t in index_tuples for i in t})
  return index_tuples, required


# ------------------------------------------------------------------------------------------------ dataset
class TileDataset:
  """The (sources, targets) examples of input_fn_tfrecords (Training.py:728-792) as numpy dictionaries.

  directory: `<base_tfrecords_directory>/<mode>` holding `*.tfrecords[.gz]`; settings: the sidecar json (path or dict)."""

  def __init__(self, directory, settings, architecture, index_tuples=None, required_indices=None,
               number_of_source_index_tuples=1, verify_crc=True):
    if isinstance(settings, str):
      with open(settings, "r", encoding="utf-8") as f:
        settings = json.load(f)
    self.tiles_height_width = int(settings["tiles_height_width"])
    self.number_of_sources_per_example = int(settings["number_of_sources_per_example"])
    self.source_samples_per_pixel_list = list(settings["source_samples_per_pixel_list"])
    self.architecture = architecture
    self.files = sorted(glob.glob(os.path.join(directory, "**", "*.tfrecords*"), recursive=True))
    if not self.files:
      raise FileNotFoundError("no .tfrecords files under %s" % directory)
    if index_tuples is None:
      index_tuples, required_indices = source_index_tuples(self.number_of_sources_per_example, number_of_source_index_tuples,
                                                           architecture.number_of_sources_per_target)
    self.index_tuples, self.required_indices = index_tuples, required_indices
    self.verify_crc = verify_crc
    self.feature_predictions = list(architecture.feature_predictions) + list(architecture.auxiliary_features)

  def _decode(self, parsed, key, channels):
    value = parsed.get(key)
    if not value:
      raise KeyError("feature '%s' is missing from the example" % key)
    s = self.tiles_height_width
    return np.frombuffer(value[0], dtype="<f4").reshape(s, s, channels)      # tf.decode_raw + reshape (:519-524)

  def examples_of_record(self, payload):
    """feature_parser (Training.py:763-792): one (sources, targets) pair per (samples per pixel, index tuple)."""
    parsed = parse_example(payload)
    s = self.tiles_height_width
    out = []
    for spp in self.source_samples_per_pixel_list:
      for index_tuple in self.index_tuples:
        sources, targets = {}, {}
        for fp in self.feature_predictions:
          for i, index in enumerate(index_tuple):
            key = Naming.source_feature_name(fp.name, index=i)
            if fp.load_data:
              sources[key] = self._decode(parsed, Naming.source_feature_name(fp.name, samples_per_pixel=spp, index=index),
                                          fp.number_of_channels)
            else:    # synthesised passes: ones for colours, 0.5 for direct / indirect (:531-537)
              assert fp.feature_prediction_type != FeaturePredictionType.AUXILIARY
              value = 1.0 if fp.feature_prediction_type == FeaturePredictionType.COLOR else 0.5
              sources[key] = np.full((s, s, fp.number_of_channels), value, dtype=np.float32)
          if fp.is_target:
            key = Naming.target_feature_name(fp.name)
            if fp.load_data:
              targets[key] = self._decode(parsed, key, fp.number_of_channels)
            else:
              value = 1.0 if fp.feature_prediction_type == FeaturePredictionType.COLOR else 0.5
              targets[key] = np.full((s, s, fp.number_of_channels), value, dtype=np.float32)
        out.append((sources, targets))
    return out

  def _examples_of_file(self, path):
    out = []
    for payload in read_records(path, verify_crc=self.verify_crc):
      out.extend(self.examples_of_record(payload))
    return out

  def examples(self, files=None, threads=0, with_file=False):
    """All (sources, targets) pairs of `files`, in file order.  threads > 0: files are read, gunzipped, CRC-checked and parsed
    by a pool of worker threads `threads` files ahead of the consumer (tf.data's num_parallel_reads / num_parallel_calls,
    Training.py:830-834; zlib and the CRC run outside the GIL), so the GPU step is not throttled by the input pipeline."""
    files = list(self.files if files is None else files)
    if threads <= 0:
      for path in files:
        index = 0
        for payload in read_records(path, verify_crc=self.verify_crc):
          for pair in self.examples_of_record(payload):
            yield (path, index, pair) if with_file else pair
            index += 1
      return
    import collections
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads) as pool:
      pending = collections.deque()
      it = iter(files)
      for path in it:
        pending.append((path, pool.submit(self._examples_of_file, path)))
        if len(pending) >= threads:
          break
      while pending:
        path, future = pending.popleft()
        done = future.result()
        nxt = next(it, None)
        if nxt is not None:
          pending.append((nxt, pool.submit(self._examples_of_file, nxt)))
        for index, pair in enumerate(done):
          yield (path, index, pair) if with_file else pair

  def record_counts(self, threads=0):
    """Records per file (framing only: gunzip + length fields, no CRC / parse), cached; aligned with self.files."""
    if getattr(self, "_record_counts", None) is None:
      def count(path):
        return sum(1 for _ in read_records(path, verify_crc=False))
      if threads > 0:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as pool:
          counts = list(pool.map(count, self.files))
      else:
        counts = [count(path) for path in self.files]
      self._record_counts = dict(zip(self.files, counts))
    return [self._record_counts[f] for f in self.files]

  def examples_per_record(self):
    return len(self.source_samples_per_pixel_list) * len(self.index_tuples)

  def rank_share(self, files, rank, world):
    """Example-granular sharding: the examples of `files` (in that order) form one global sequence, rank r owns the
    contiguous range [r*per, (r+1)*per) with per = total // world, so EVERY rank sees the same number of examples (the
    remainder is dropped) and therefore runs the same number of steps - each step holds a blocking all-reduce, unequal
    counts would deadlock the job at the end of the epoch.  Returns [(file, first example, one past last example)]."""
    self.record_counts()
    per_record = self.examples_per_record()
    counts = [self._record_counts[f] * per_record for f in files]
    total = sum(counts)
    per = total // world
    lo, hi = rank * per, (rank + 1) * per
    out, start = [], 0
    for f, c in zip(files, counts):
      a, b = max(lo, start), min(hi, start + c)
      if a < b:
        out.append((f, a - start, b - start))
      start += c
    return out, per

  def batches(self, batch_size, epochs=1, shuffle_seed=None, rank=0, world=1, drop_remainder=True, threads=0):
    """dataset.shuffle(20 * batch).batch(batch) (Training.py:836-839), files shuffled per epoch (:825-826, with a seed every
    rank shares) and sharded over `world` ranks at EXAMPLE granularity (rank_share): every rank yields the same number of
    batches.  Yields (sources, targets) dictionaries of [B,S,S,C] float32 arrays."""
    rng = random.Random(shuffle_seed)
    for _ in range(epochs):
      files = list(self.files)
      if shuffle_seed is not None:
        rng.shuffle(files)
      if world > 1:
        share, _ = self.rank_share(files, rank, world)
      else:
        share = [(f, 0, None) for f in files]
      bounds = {f: (a, b) for f, a, b in share}

      def stream_of(share=share, bounds=bounds):
        it = self.examples([f for f, _, _ in share], threads=threads, with_file=True)
        for path, index, pair in it:
          a, b = bounds[path]
          if index >= a and (b is None or index < b):
            yield pair

      pool, limit = [], (20 * batch_size if shuffle_seed is not None else batch_size)
      stream = stream_of()
      exhausted = False
      while True:
        while not exhausted and len(pool) < limit:
          try:
            pool.append(next(stream))
          except StopIteration:
            exhausted = True
        if len(pool) < batch_size and (drop_remainder or not pool):
          break
        take = []
        for _ in range(min(batch_size, len(pool))):
          take.append(pool.pop(rng.randrange(len(pool)) if shuffle_seed is not None else 0))
        sources = {k: np.stack([t[0][k] for t in take]) for k in take[0][0]}
        targets = {k: np.stack([t[1][k] for t in take]) for k in take[0][1]}
        yield sources, targets


def write_tile_dataset(directory, name, examples, settings, examples_per_tfrecords=16, compress=True):
  """Writes what TFRecordsCreator.py produces for one mode: `<directory>/<name>/<name>_<i>.tfrecords[.gz]` and the sidecar
  `<directory>/<name>.json`.  examples: iterable of {feature name: float32 HWC array}."""
  out_dir = os.path.join(directory, name)
  os.makedirs(out_dir, exist_ok=True)
  index, pending, files = 0, [], []

  def flush():
    nonlocal index, pending
    if pending:
      path = os.path.join(out_dir, "%s_%d.tfrecords%s" % (name, index, ".gz" if compress else ""))
      write_records(path, pending)
      files.append(path)
      index += 1
      pending = []

  for features in examples:
    pending.append(serialize_example({k: np.ascontiguousarray(v, dtype="<f4").tobytes() for k, v in features.items()}))
    if len(pending) >= examples_per_tfrecords:
      flush()
  flush()
  with open(os.path.join(directory, name + ".json"), "w", encoding="utf-8") as f:
    json.dump(settings, f, sort_keys=True, indent=2)
  return files
