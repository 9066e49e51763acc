"""Turns an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into the per-kernel share table used in profiles/*.md.
usage: python profiles/summarize_launches.py <launches.csv> "<title>" "<command note>" > <summary.md>"""
import csv, re, sys
from collections import defaultdict

path, title, note = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
    name = r["Kernel Name"]
    short = re.sub(r"^void\s+", "", name).replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    short = re.sub(r"\((SgItemKind|SgSerialKind|bool)\)", "", short)
    m = re.match(r"([A-Za-z0-9_]+(?:<[^>]*>)?)", short)
    short = m.group(1) if m else short
    rows.append((short, ms))
agg = defaultdict(list)
for k, ms in rows:
    agg[k].append(ms)
total = sum(ms for _, ms in rows)
print(f"# {title}\n\n{note}\n\ntotal {total:.1f} ms over {len(rows)} launches\n")
print("| kernel | launches | total ms | share | avg us | max ms |\n|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if sum(v) / total < 0.001:
        continue
    print(f"| `{k}` | {len(v)} | {sum(v):.3f} | {100 * sum(v) / total:.1f}% | {1e3 * sum(v) / len(v):.1f} | {max(v):.3f} |")
