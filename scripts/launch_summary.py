"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (developer tool).
usage: python scripts/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/xyz.csv"""
import collections
import csv
import re
import sys


def main():
  path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
  rows = [r for r in csv.reader(open(path)) if len(r) > 5]
  hdr = rows[0]
  ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
  agg = collections.OrderedDict()
  for r in rows[1:]:
    try:
      v = float(r[vi].replace(",", ""))
    except ValueError:
      continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("gpp::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
  tot = sum(v for _, v in agg.values())
  print(f"# ncu launch list (own kernels, -k regex:k_) of `{cmd}`")
  print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
  print(f"# total {tot:.2f} ms over {sum(n for n, _ in agg.values())} launches")
  print("kernel,launches,total_ms,share,avg_us")
  for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{v:.3f},{v / tot:.4f},{1e3 * v / n:.1f}")


if __name__ == "__main__":
  main()
