#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libdd_b200.so (tensor-core / TMA / TMEM / async-copy instructions): profiles/rNN_sass_evidence.txt.
usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "LDSM", "LDGSTS", "SYNCS", "REDG", "RED"]


def main():
  so = os.path.join(ROOT, "deepdenoiser_b200", "libdd_b200.so")
  sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
  demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
  counts, order, cur = {}, [], None
  for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      cur = m.group(1)
      counts[cur] = collections.Counter()
      order.append(cur)
      continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
      counts[cur]["instr"] += 1
      op = m.group(1)
      for k in KEYS:
        if op == k or op.startswith(k + "."):
          counts[cur][k] += 1
          break
  print("SASS evidence (cuobjdump -sass deepdenoiser_b200/libdd_b200.so, sm_100a): tensor-core / TMA / TMEM / async-copy mnemonics per kernel")
  print("UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,")
  print("HMMA + LDSM = mma.sync + ldmatrix, LDGSTS = cp.async, SYNCS = mbarrier ops, RED/REDG = global reductions\n")
  for fn in sorted(order, key=lambda f: -sum(v for k, v in counts[f].items() if k != "instr")):
    c = counts[fn]
    tags = "  ".join("%s:%d" % (k, c[k]) for k in KEYS if c[k])
    if tags:
      name = demangle(fn)
      name = name[:name.rfind(")(") + 1] if ")(" in name else name.split("(dd::")[0].split("(const")[0]
      print("%-62s %5d instr  %s" % (name[:62], c["instr"], tags))


if __name__ == "__main__":
  main()
