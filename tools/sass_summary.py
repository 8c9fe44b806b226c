"""Per-kernel SASS mnemonic counts of the in-tree library (cuobjdump -sass): what proves the sm_100a code paths
(FFMA2 / FMNMX3 packed math, UBLKCP + SYNCS = bulk-TMA staging, USETMAXREG = warp-specialised register split) and that no
tensor-core / TMEM ops are used on this path (north_star). Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rustpotter_b200", "librustpotter_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        counts[fn][m.group(2).split(".")[0]] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
keys = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMNMX3", "MUFU", "LDS", "STS", "LDG", "STG", "LDCU", "SHFL", "BAR", "UBLKCP",
        "SYNCS", "USETMAXREG", "LDTM", "STTM", "UTCHMMA", "HMMA"]
print(f"SASS mnemonic counts per kernel of `{os.path.relpath(lib, ROOT)}` (static instruction counts, sm_100a)\n")
print("| kernel | total | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for (f, h), name in sorted(zip(counts.items(), names), key=lambda kv: -sum(kv[0][1].values())):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("rp::", "")
    name = re.sub(r"\((rp::)?(Dtw|Judge|Filter|const|float|unsigned|int|long|DtwPairsArgs).*", "", name)
    print("| `" + name[:80] + "` | " + str(sum(h.values())) + " | " + " | ".join(str(h.get(k, 0)) for k in keys) + " |")
