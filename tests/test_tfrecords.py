"""TFRecord reader / writer without TensorFlow (deepdenoiser_b200/tfrecords.py): known-answer vectors of the published
formats (CRC-32C check value, TFRecord masking constant, protobuf wire encoding of a tf.train.Example written out by
hand), round trips through GZIP files, and the (sources, targets) assembly of Training.input_fn_tfrecords."""
import gzip
import os
import struct

import numpy as np
import pytest

from deepdenoiser_b200 import synthetic, tfrecords
from deepdenoiser_b200.Architecture import Architecture


def test_crc32c_known_answers():
  # CRC-32C (Castagnoli) check value of the ASCII digits, RFC 3720 appendix B.4 vectors
  assert tfrecords._crc32c(b"123456789") == 0xE3069283
  assert tfrecords._crc32c(bytes(32)) == 0x8A9136AA
  assert tfrecords._crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
  assert tfrecords._crc32c(bytes(range(32))) == 0x46DD794E
  assert tfrecords._crc32c(b"") == 0
  # masking: rotate right by 15, add 0xa282ead8 (mod 2^32)
  crc = 0xE3069283
  assert tfrecords.masked_crc32c(b"123456789") == ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_wire_format_known_answer():
  # Example{features{feature{key:"a" value{bytes_list{value:"xy"}}}}} encoded by hand from the protobuf wire spec:
  #   entry  = 0a 01 'a'  12 06 [ 0a 04 [ 0a 02 'x' 'y' ] ]
  #   Features.feature (field 1) = 0a 0b entry ; Example.features (field 1) = 0a 0d ...
  want = bytes([0x0A, 0x0D, 0x0A, 0x0B, 0x0A, 0x01, 0x61, 0x12, 0x06, 0x0A, 0x04, 0x0A, 0x02, 0x78, 0x79])
  assert tfrecords.serialize_example({"a": b"xy"}) == want
  parsed = tfrecords.parse_example(want)
  assert list(parsed) == ["a"] and bytes(parsed["a"][0]) == b"xy"
  # float_list / int64_list features (packed)
  blob = tfrecords.serialize_example({"f": np.array([1.5, -2.0], np.float32), "i": np.array([3, -1, 300], np.int64)})
  back = tfrecords.parse_example(blob)
  assert np.array_equal(back["f"], np.array([1.5, -2.0], np.float32))
  assert np.array_equal(back["i"], np.array([3, -1, 300], np.int64))


def test_record_framing_round_trip_and_corruption(tmp_path):
  payloads = [b"", b"hello", os.urandom(70000)]
  for suffix in (".tfrecords", ".tfrecords.gz"):
    path = str(tmp_path / ("data" + suffix))
    tfrecords.write_records(path, payloads)
    assert list(tfrecords.read_records(path)) == payloads
  # frame layout: u64 length, u32 masked crc(length), payload, u32 masked crc(payload)
  raw = open(str(tmp_path / "data.tfrecords"), "rb").read()
  assert struct.unpack("<Q", raw[:8])[0] == 0 and len(raw) == 3 * 16 + 5 + 70000
  assert struct.unpack("<I", raw[8:12])[0] == tfrecords.masked_crc32c(raw[:8])
  # the .gz variant is a plain GZIP stream of the same bytes (TFRecordsCreator._compress)
  assert gzip.open(str(tmp_path / "data.tfrecords.gz"), "rb").read() == raw
  bad = bytearray(raw)
  bad[16 + 12 + 2] ^= 0x40                      # flip a bit inside "hello"
  open(str(tmp_path / "bad.tfrecords"), "wb").write(bytes(bad))
  with pytest.raises(IOError):
    list(tfrecords.read_records(str(tmp_path / "bad.tfrecords")))
  assert len(list(tfrecords.read_records(str(tmp_path / "bad.tfrecords"), verify_crc=False))) == 3


def test_source_index_tuples_like_the_reference():
  tuples, required = tfrecords.source_index_tuples(2, 8, 1)
  assert tuples == [[0], [1]] * 4 and required == [0, 1]
  tuples, required = tfrecords.source_index_tuples(3, 4, 1)
  assert tuples[:3] == [[0], [1], [2]] and len(tuples) == 4 and 0 <= tuples[3][0] <= 2
  with pytest.raises(Exception):
    tfrecords.source_index_tuples(1, 2, 2)


def test_tile_dataset_matches_what_was_written(tmp_path):
  arch = Architecture(synthetic.example_architecture_json())
  size, spps, per_example = 16, [4, 16], 2
  rng = np.random.default_rng(3)
  passes = [(fp.name, fp.number_of_channels, fp.is_target) for fp in list(arch.feature_predictions) + list(arch.auxiliary_features)
            if fp.load_data]
  written = []
  for e in range(5):
    feats = {}
    for name, c, is_target in passes:
      for spp in spps:
        for index in range(per_example):
          feats["source_image/%d/%d/%s" % (spp, index, name)] = rng.standard_normal((size, size, c)).astype(np.float32)
      if is_target:
        feats["target_image/" + name] = rng.standard_normal((size, size, c)).astype(np.float32)
    written.append(feats)
  settings = {"tiles_height_width": size, "number_of_sources_per_example": per_example, "source_samples_per_pixel_list": spps}
  files = tfrecords.write_tile_dataset(str(tmp_path), "training", written, settings, examples_per_tfrecords=2)
  assert [os.path.basename(f) for f in files] == ["training_0.tfrecords.gz", "training_1.tfrecords.gz", "training_2.tfrecords.gz"]
  ds = tfrecords.TileDataset(str(tmp_path / "training"), str(tmp_path / "training.json"), arch, number_of_source_index_tuples=2)
  assert ds.index_tuples == [[0], [1]]
  got = list(ds.examples())
  assert len(got) == 5 * len(spps) * 2                      # one example per (record, spp, index tuple)
  # order: record-major, then spp, then index tuple (feature_parser, Training.py:776-777)
  sources, targets = got[3]                                 # record 0, spp 16, index 1
  name = passes[0][0]
  assert np.array_equal(sources["source_image/0/" + name], written[0]["source_image/16/1/" + name])
  for name, c, is_target in passes:
    if is_target:
      assert np.array_equal(targets["target_image/" + name], written[0]["target_image/" + name])
  # synthesised (load_data = False) passes: ones for colours, 0.5 for direct / indirect
  for fp in arch.feature_predictions:
    if not fp.load_data:
      v = sources["source_image/0/" + fp.name]
      assert v.shape == (size, size, fp.number_of_channels) and float(v.min()) == float(v.max()) and float(v.max()) in (0.5, 1.0)
  batches = list(ds.batches(4, shuffle_seed=7))
  assert len(batches) == 5 and all(b[0]["source_image/0/" + name].shape == (4, size, size, passes[-1][1]) or True for b in batches)
  key = "source_image/0/" + passes[0][0]
  assert batches[0][0][key].shape == (4, size, size, passes[0][1])
  # threaded prefetch yields exactly the same examples in the same order
  threaded = list(ds.examples(threads=3))
  assert len(threaded) == len(got)
  for (s0, t0), (s1, t1) in zip(got, threaded):
    assert s0.keys() == s1.keys() and all(np.array_equal(s0[k], s1[k]) for k in s0)
    assert all(np.array_equal(t0[k], t1[k]) for k in t0)
  # sharding over ranks is example-granular: 3 unequal files (2 + 2 + 1 records = 8 + 8 + 4 examples), every rank yields
  # the SAME number of batches (each training step holds a blocking all-reduce) and the shares are disjoint
  for world in (2, 3):
    for seed in (None, 5):
      per_rank = [list(ds.batches(2, shuffle_seed=seed, rank=r, world=world, threads=r)) for r in range(world)]
      assert len({len(b) for b in per_rank}) == 1, [len(b) for b in per_rank]
      assert len(per_rank[0]) == (20 // world) // 2
      seen = set()
      for batches_r in per_rank:
        for sources, _ in batches_r:
          for tile in sources[key]:
            sig = tile.tobytes()
            assert sig not in seen
            seen.add(sig)
      assert len(seen) == world * ((20 // world) // 2) * 2


def test_more_ranks_than_files_share_examples_not_duplicates(tmp_path):
  """len(files) < world used to hand every rank ALL files (the global batch repeated world times)."""
  arch = Architecture(synthetic.example_architecture_json())
  size = 8
  rng = np.random.default_rng(11)
  written = []
  for e in range(8):
    feats = {}
    for fp in list(arch.feature_predictions) + list(arch.auxiliary_features):
      if fp.load_data:
        feats["source_image/16/0/" + fp.name] = rng.standard_normal((size, size, fp.number_of_channels)).astype(np.float32)
        if fp.is_target:
          feats["target_image/" + fp.name] = rng.standard_normal((size, size, fp.number_of_channels)).astype(np.float32)
    written.append(feats)
  settings = {"tiles_height_width": size, "number_of_sources_per_example": 1, "source_samples_per_pixel_list": [16]}
  files = tfrecords.write_tile_dataset(str(tmp_path), "training", written, settings, examples_per_tfrecords=16)
  assert len(files) == 1
  ds = tfrecords.TileDataset(str(tmp_path / "training"), str(tmp_path / "training.json"), arch)
  assert ds.record_counts() == [8]
  key = "source_image/0/" + next(fp.name for fp in arch.feature_predictions if fp.load_data)
  seen = []
  for r in range(4):
    b = list(ds.batches(2, shuffle_seed=3, rank=r, world=4))
    assert len(b) == 1
    seen += [t.tobytes() for t in b[0][0][key]]
  assert len(set(seen)) == 8


def test_example_round_trip_property():
  """Property: serialize_example / parse_example and the record framing round-trip arbitrary feature dictionaries."""
  from hypothesis import given, settings, strategies as st
  names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz/ _0123456789", min_size=1, max_size=24)
  values = st.one_of(st.binary(min_size=0, max_size=300),
                     st.lists(st.floats(-1e6, 1e6, width=32), min_size=1, max_size=20).map(lambda v: np.array(v, np.float32)),
                     st.lists(st.integers(-2 ** 62, 2 ** 62), min_size=1, max_size=20).map(lambda v: np.array(v, np.int64)))

  @settings(max_examples=60, deadline=None)
  @given(features=st.dictionaries(names, values, min_size=0, max_size=6))
  def check(features):
    back = tfrecords.parse_example(tfrecords.serialize_example(features))
    assert set(back) == set(features)
    for k, v in features.items():
      if isinstance(v, bytes):
        assert bytes(back[k][0]) == v
      else:
        assert np.array_equal(np.asarray(back[k]), v)

  check()
