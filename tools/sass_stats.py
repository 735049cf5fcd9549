"""Instruction mix per kernel from `cuobjdump -sass` (no GPU needed).
    python tools/sass_stats.py <obj-or-so> [name-filter-regex]
"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1]
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name, stats = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        stats[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        stats[name][m.group(2)] += 1
for name, c in stats.items():
    if flt and not flt.search(name):
        continue
    total = sum(c.values())
    print("%5d  %s" % (total, name))
    print("       " + "  ".join("%s:%d" % kv for kv in c.most_common(14)))
