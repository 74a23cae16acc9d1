#!/usr/bin/env python
"""SASS instruction counts per kernel and per out-of-line device function of the built library (cuobjdump -xelf all + nvdisasm -c).
usage: python tools/sass_sizes.py [lib.so] [name filter regex]"""
import os, re, subprocess, sys, tempfile, collections
so = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "opus_codec_b200", "libopus_b200.so"))
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, stdout=subprocess.DEVNULL)
for cub in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):
    txt = subprocess.run(["nvdisasm", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    counts, order, cur = collections.Counter(), [], None
    for l in txt.splitlines():
        m = re.match(r"\s*\.type\s+(\S+),@function", l)
        if m:
            cur = m.group(1)
            if cur not in counts: order.append(cur)
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): counts[cur] += 1
    if not counts: continue
    names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    print("## " + cub)
    for n, d in zip(order, names):
        d = re.sub(r"\(.*", "", d)
        if flt and not flt.search(d): continue
        print("%8d  %s" % (counts[n], d))
