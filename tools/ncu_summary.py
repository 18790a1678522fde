#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py report.ncu-rep out.csv"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "Block Size", "Grid Size",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.max.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def main():
  rep, out = sys.argv[1], sys.argv[2]
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  header, units, data = rows[0], rows[1], rows[2:]
  with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch %d" % (i + 1) for i in range(len(data))])
    for key in KEEP:
      cols = [i for i, h in enumerate(header) if h == key or h.endswith("." + key)]
      if not cols:
        continue
      i = cols[0]
      w.writerow([key, units[i]] + [d[i] for d in data])
  # roofline.traffic of bench.py: dram bytes of the longest captured conv_rows_kernel launch (profiles/ncu_traffic.json)
  def col(name):
    c = [i for i, h in enumerate(header) if h == name]
    return c[0] if c else None
  kn, rd, wr, dur = col("Kernel Name"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
  convs = [d for d in data if kn is not None and "conv_rows_kernel" in d[kn]]
  if convs and rd is not None and wr is not None:
    import json
    import os
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    d = max(convs, key=lambda r: float(r[dur].replace(",", "")) if dur is not None else 0.0)    # the longest captured launch
    total = float(d[rd].replace(",", "")) * scale.get(units[rd], 1.0) + float(d[wr].replace(",", "")) * scale.get(units[wr], 1.0)
    path = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json")
    known = json.load(open(path)) if os.path.exists(path) else {}
    known["conv_rows_kernel"] = {"bytes": total, "launch": "launch %d of %s (%s %s under ncu --set full)" %
                                 (convs.index(d) + 1, os.path.basename(rep), d[dur] if dur is not None else "?", units[dur] if dur is not None else "")}
    json.dump(known, open(path, "w"), indent=1)


if __name__ == "__main__":
  main()
