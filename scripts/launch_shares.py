"""Per-kernel totals of the LAST iteration in an ncu launch list (csv of `--metrics gpu__time_duration.sum`): everything from
the last launch whose name contains argv[2] (default: flow_noise_mix = the start of a training step)."""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
start = sys.argv[2] if len(sys.argv) > 2 else "flow_noise_mix"
rows = []
for x in csv.DictReader(lines):
    if x["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")[:64]
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    rows.append((k, v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v, x["Grid Size"]))
idx = [i for i, r in enumerate(rows) if start in r[0]]
sel = rows[idx[-1]:] if idx else rows
agg, tot = {}, 0.0
for k, v, g in sel:
    a = agg.setdefault(k, [0, 0.0, g])
    a[0] += 1
    a[1] += v
    tot += v
print(f"{len(sel)} launches, {tot / 1e3:.2f} ms\n")
print("| kernel | launches | total us | share | avg us | grid (last) |\n|---|---:|---:|---:|---:|---|")
for k, (c, v, g) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"| `{k}` | {c} | {v:.0f} | {100 * v / tot:.1f}% | {v / c:.1f} | {g} |")
