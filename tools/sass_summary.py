"""Not a test: per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel, from the shipped library.

    python tools/sass_summary.py > profiles/sass_summary.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce,
HMMA = legacy mma.sync, LDGSTS = cp.async (guide: /opt/skills/guides/B200_PROFILING.md)."""
import collections
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clip-ppo_b200", "libclipppo_b200.so")
cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
sass = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True).stdout
cols = ("UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "LDGSTS", "FFMA", "MUFU", "instr")
rows, cur, counts = [], None, None


def classify(op):
    if re.match(r"UTC\w*MMA", op): return "UTC*MMA"
    for k in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "LDGSTS", "FFMA", "MUFU"):
        if op.startswith(k): return k
    return None


for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts))
        cur, counts = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        counts["instr"] += 1
        k = classify(m.group(1))
        if k:
            counts[k] += 1
if cur:
    rows.append((cur, counts))


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (sm_100a)\n# " + "  ".join(f"{c:>8s}" for c in cols) + "  kernel")
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: -r[1]["instr"]):
    tot.update(c)
    d = demangle(name)
    d = d.replace("clipppo::(anonymous namespace)::", "").replace("clipppo::<unnamed>::", "").replace("clipppo::", "").replace("void ", "")
    d = re.sub(r"\((int|bool)\)", "", d)                     # template-argument casts
    d = re.sub(r">\(.*$", ">", d) if ">(" in d else re.sub(r"\(.*$", "", d)     # drop the parameter list
    print("  " + "  ".join(f"{c[k]:8d}" for k in cols) + "  " + d[:100])
print("  " + "  ".join(f"{tot[k]:8d}" for k in cols) + "  TOTAL")
