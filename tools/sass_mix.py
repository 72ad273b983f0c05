#!/usr/bin/env python3
"""Static SASS instruction mix per kernel of the built library (cuobjdump -sass).
Usage: python tools/sass_mix.py [substring ...]   -- prints total instructions and the top opcodes per matching kernel."""
import collections
import re
import subprocess
import sys

so = "radar_signal_process_b200/libradar_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name = None
mix = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and name:
        mix[name][m.group(1)] += 1
pats = sys.argv[1:]
for k in sorted(mix):
    if pats and not any(p in k for p in pats):
        continue
    c = mix[k]
    print("%6d  %s" % (sum(c.values()), k))
    print("        " + " ".join("%s:%d" % kv for kv in c.most_common(14)))
