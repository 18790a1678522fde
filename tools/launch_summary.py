#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total ms, share."""
import collections
import csv
import sys


def main():
  rows = list(csv.reader(open(sys.argv[1])))
  hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
  H = rows[hdr]
  ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
  skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
  agg = collections.defaultdict(lambda: [0, 0.0])
  for r in rows[hdr + 1 + skip:]:
    if len(r) <= vi:
      continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
    name = r[ki].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += ms
  tot = sum(v[1] for v in agg.values())
  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%5d %10.3f ms %5.1f%%" % (k[:60], v[0], v[1], 100 * v[1] / tot))
  print("total %.3f ms" % tot)


if __name__ == "__main__":
  main()
