"""Aggregate the SASS source page of a .ncu-rep into contiguous hot regions:
python scripts/ncu_hot.py rep.ncu-rep [min_share]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["Instructions Executed"]) for r in rows)
samples = sum(int(r["# Samples"]) for r in rows)
print("total warp instr", tot, "samples", samples)
# regions: split where executed count changes by more than 2x
regions = []
cur = None
for i, r in enumerate(rows):
    n = int(r["Instructions Executed"])
    if cur is None or not (0.6 * cur["n"] <= n <= 1.6 * cur["n"]):
        cur = dict(i0=i, n=max(n, 1), instr=0, execd=0, samples=0, mufu=0, ops={})
        regions.append(cur)
    cur["instr"] += 1
    cur["execd"] += n
    cur["samples"] += int(r["# Samples"])
    op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
    cur["ops"][op.split(".")[0]] = cur["ops"].get(op.split(".")[0], 0) + 1
    cur["i1"] = i
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
for g in regions:
    if g["execd"] / tot >= thr:
        top = sorted(g["ops"].items(), key=lambda kv: -kv[1])[:8]
        print(f"rows {g['i0']:5d}-{g['i1']:5d} static {g['instr']:5d} exec/instr {g['execd']/g['instr']:.3g} "
              f"share {100*g['execd']/tot:5.1f}% samples {100*g['samples']/samples:5.1f}%  {top}")
