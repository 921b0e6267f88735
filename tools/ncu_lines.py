"""Not a test: per-source-line instruction counts and stall samples of one kernel from an ncu report.

    python tools/ncu_lines.py <report.ncu-rep> <object.o> [items]

Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` of the same object file (needs -lineinfo)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj = sys.argv[1], sys.argv[2]
items = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur, seq = None, []
for l in dis.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, data = rows[1], rows[2:]
isrc, iex, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
n = min(len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0])
for i in range(n):
    agg[seq[i][0]][0] += int(data[i][iex])
    agg[seq[i][0]][1] += int(data[i][isamp])
tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"{len(seq)} SASS instructions in the object, {len(data)} in the report; {tot / items:.0f} warp instructions per item")
srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clip-ppo_b200", "csrc")
cache = {}
for k, v in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if v[0] > tot * 0.004 or v[1] > ts * 0.004:
        f, ln = k if k else ("?", 0)
        path = os.path.join(srcdir, f)
        if f not in cache:
            cache[f] = open(path).read().split("\n") if os.path.exists(path) else None
        text = cache[f][ln - 1].strip()[:100] if cache[f] and 0 < ln <= len(cache[f]) else ""
        print(f"{f}:{ln:4d}  {v[0] / items:9.1f} instr ({100 * v[0] / tot:4.1f} %)  samples {100 * v[1] / ts:5.1f} %  {text}")
