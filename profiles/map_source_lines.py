"""Maps ncu's per-instruction counts and stall samples (`ncu -i X.ncu-rep --page source --csv`) to source lines through
`nvdisasm -g` of a cubin built from the SAME sources with -lineinfo.  usage:
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -x cu -cubin -o k.cubin seqfrost_b200/csrc/sigma_kernels.cu
    python profiles/map_source_lines.py X.ncu-rep k.cubin <kernel name substring in the report> <mangled substring in the cubin> [top]"""
import collections, csv, re, subprocess, sys, os

rep, cubin, kname, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l and l.rstrip().endswith(":"))
ins, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("//---") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
blocks, cur = [], None
for l in out:
    if l.startswith('"Kernel Name"'):
        cur = [l]; blocks.append(cur)
    elif cur is not None:
        cur.append(l)
blk = next(b for b in blocks if kname in b[0])
rows = list(csv.reader(blk[1:])); hdr, rows = rows[0], rows[1:]
ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")


def op(t):
    a = t.split()
    return ((a[1] if a[0].startswith("@") else a[0]) if a else "").split(".")[0]


n = min(len(rows), len(ins))
mism = sum(op(rows[k][cs].strip()) != op(ins[k][0]) for k in range(n))
print(f"{len(ins)} instructions in the cubin, {len(rows)} in the report, opcode mismatches {mism}")
agg, aggs = collections.Counter(), collections.Counter()
for k in range(n):
    agg[ins[k][1]] += int(rows[k][ci] or 0); aggs[ins[k][1]] += int(rows[k][cw] or 0)
tot, tots = sum(agg.values()), max(sum(aggs.values()), 1)
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seqfrost_b200", "csrc")
src = {f: open(os.path.join(root, f)).read().splitlines() for f in ("sigma_items.cuh", "sigma_kernels.cu")}
print("| instructions | stall samples | where | source |\n|---|---|---|---|")
for key, c in sorted(agg.items(), key=lambda kv: -(kv[1] / tot + aggs[kv[0]] / tots))[:top]:
    line = src[key[0]][key[1] - 1].strip()[:110].replace("|", "\\|") if key and key[0] in src else ""
    print(f"| {100 * c / tot:.1f}% | {100 * aggs[key] / tots:.1f}% | {key[0]}:{key[1]} | `{line}` |" if key else f"| {100 * c / tot:.1f}% | {100 * aggs[key] / tots:.1f}% | - | |")
