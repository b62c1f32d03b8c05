"""Summarise an ncu launch list (gpu__time_duration.sum per launch) into per-kernel shares."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    name = r[ki]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    short = name if name.startswith("pm::") else "torch:" + name.split("<")[0][:60]
    m = re.match(r"pm::(\w+)<(.*)>", name)
    if m:
        short = f"pm::{m.group(1)}<{m.group(2)}>"
    t = float(r[vi].replace(",", ""))
    tot[short] += t; cnt[short] += 1
all_t = sum(tot.values())
print(f"# ncu launch list summary: {path}\n# {sum(cnt.values())} launches, {all_t / 1e6:.3f} ms total (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| `{k}` | {cnt[k]} | {t / 1e6:.3f} | {100 * t / all_t:.1f}% | {t / cnt[k] / 1e3:.1f} |")
