"""SASS instruction count per source line of one kernel (instruction-cache footprint audit).
usage: python profiles/sass_footprint.py <kernel substring> [top]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "torchdrivesim_b200", "_build", "libtds_b200.so")
pat = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    counts = collections.Counter()
    for f in os.listdir(d):
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, f)], capture_output=True, text=True).stdout
        fn, line = None, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", l)
            if m:
                fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                line = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            if fn and pat in fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
                counts[line] += 1
print("total SASS instructions:", sum(counts.values()))
srcs = {}
for (f, ln), c in sorted(counts.items(), key=lambda kv: -kv[1])[:top]:
    path = os.path.join(ROOT, "torchdrivesim_b200", "csrc", f)
    if f not in srcs and os.path.exists(path):
        srcs[f] = open(path).read().splitlines()
    text = srcs[f][ln - 1].strip()[:100] if f in srcs and ln - 1 < len(srcs[f]) else ""
    print(f"{c:5d}  {f}:{ln:<4d} {text}")
