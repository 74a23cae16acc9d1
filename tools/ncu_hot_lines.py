#!/usr/bin/env python
"""Per-source-line instruction counts for a kernel: joins `ncu --page source --print-source sass --csv` (executed
instruction counts per SASS instruction) with `nvdisasm -g` line info of the cubin in the .so (same instruction order).
Usage: ncu_hot_lines.py <report.ncu-rep> <kernel-regex (mangled section name)> <lib.so> [top] [ncu kernel-name regex, default = kernel-regex]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, kern, so = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
ncu_kern = sys.argv[5] if len(sys.argv) > 5 else kern
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, stdout=subprocess.DEVNULL)
dis = []
for cub in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):      # one cubin per translation unit: take the one holding the kernel
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    if re.search(r"\.section\s+\.text\.\S*" + kern, txt):
        dis = txt.splitlines(); break
# instruction -> (file,line) in order, per function (functions called by the kernel are separate .text sections inlined or not)
sections = collections.OrderedDict(); cur = None; loc = ("?", 0)
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m: cur = m.group(1); sections[cur] = []; continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: loc = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m and cur: sections[cur].append((m.group(1).strip(), loc))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + ncu_kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = rows[1]; ia = h.index("Source"); ie = h.index("Instructions Executed"); it = h.index("Thread Instructions Executed"); iss = h.index("# Samples")
insts = [(r[ia].strip(), int(r[ie] or 0), int(r[it] or 0), int(r[iss] or 0)) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
# the ncu listing is the kernel function followed by its callees, each in nvdisasm order: match greedily by opcode sequence
def op(s): return s.split()[0] if not s.startswith("@") else s.split()[1]
flat = []
name = [k for k in sections if re.search(kern, k) and "_Z" in k]
order = name + [k for k in sections if k not in name]
agg = collections.Counter(); aggt = collections.Counter(); aggs = collections.Counter()
pos = 0
# try: concatenate sections in an order that matches instruction count
for k in order:
    flat += sections[k]
if len(flat) < len(insts):
    print("warning: nvdisasm has fewer instructions (%d) than ncu (%d)" % (len(flat), len(insts)))
# align by searching each ncu chunk start within sections
j = 0
secidx = {k: 0 for k in sections}
# simple approach: ncu lists functions in address order == nvdisasm section order of the final link? fall back to opcode matching per section
import itertools
allsecs = list(sections.items())
i = 0
used = set()
while i < len(insts):
    best = None
    for k, lst in allsecs:
        if k in used or not lst: continue
        n = min(len(lst), len(insts) - i, 40)
        if all(op(lst[t][0]) == op(insts[i + t][0]) for t in range(n)):
            best = k; break
    if best is None:
        i += 1; continue
    lst = sections[best]; used.add(best)
    for t in range(min(len(lst), len(insts) - i)):
        loc = lst[t][1]
        agg[loc] += insts[i + t][1]; aggt[loc] += insts[i + t][2]; aggs[loc] += insts[i + t][3]
    i += len(lst)
tot = sum(agg.values()) or 1; tots = sum(aggs.values()) or 1
print("kernel %s: %d warp-instructions attributed, %d stall samples" % (kern, tot, tots))
print("%-22s %6s %7s %7s %7s" % ("file:line", "inst%", "thr/in", "smpl%", ""))
by_samples = os.environ.get("HOT_BY", "inst") == "samples"
for loc, v in (sorted(agg.items(), key=lambda kv: -aggs[kv[0]])[:top] if by_samples else agg.most_common(top)):
    print("%-22s %6.2f %7.1f %7.2f" % ("%s:%d" % loc, 100.0 * v / tot, aggt[loc] / max(1, v), 100.0 * aggs[loc] / tots))
