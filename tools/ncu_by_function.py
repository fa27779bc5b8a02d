"""Aggregate an `ncu --page source --csv` dump (per SASS instruction) by SOURCE FUNCTION.

usage: tools/ncu_by_function.py <source.csv> <nvdisasm --print-line-info output> <kernel-name-substring>
The SASS offsets of ncu's rows are matched to nvdisasm's `//## File ..., line N` annotations of the same cubin, and lines
are mapped to the enclosing device function of the .cu/.cuh files under hyper-ray-tracer_b200/csrc.
"""
import bisect
import collections
import csv
import os
import re
import sys

src_csv, disasm, kern = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hyper-ray-tracer_b200", "csrc")

# offset -> (file, line) from nvdisasm
off2line = {}
cur = None
infunc = False
for line in open(disasm):
    m = re.search(r"\.section\s+\.text\.(\S+),", line)
    if m:
        infunc = kern in m.group(1)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", line)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur


def functions(path):
    res = []
    for i, l in enumerate(open(path), 1):
        if l.startswith((" ", "\t", "//", "#", "}")):
            continue
        m = re.search(r"\b([A-Za-z_][A-Za-z_0-9]*)\s*\(", l)
        if m and ("__device__" in l or "__global__" in l):
            res.append((i, m.group(1) if m.group(1) != "__launch_bounds__" else re.findall(r"\b([A-Za-z_0-9]+)\s*\(", l)[-1]))
    return res


fmap = {f: functions(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))}


def func_of(fl):
    f, ln = fl
    if f not in fmap:
        return f
    lst = fmap[f]
    i = bisect.bisect_right([x[0] for x in lst], ln) - 1
    return lst[i][1] if i >= 0 else f


rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    off = int(r[0], 16) - base
    fn = func_of(off2line.get(off, ("?", 0)))
    def num(name):
        try:
            return float(r[col[name]])
        except (ValueError, KeyError):
            return 0.0
    d = agg[fn]
    for name, key in (("Instructions Executed", "warp_inst"), ("Thread Instructions Executed", "thread_inst"), ("# Samples", "samples")):
        d[key] += num(name)
        tot[key] += num(name)
    for s in stall_cols:
        d[s] += num(s)
        tot[s] += num(s)
print(f"total warp_inst {tot['warp_inst']:.3e} thread_inst {tot['thread_inst']:.3e} samples {tot['samples']:.0f}  lanes/inst {tot['thread_inst'] / tot['warp_inst']:.2f}")
print(f"{'function':28s} {'warp_inst%':>10s} {'lanes':>6s} {'samples%':>9s}  top stalls")
for fn, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:28]:
    st = sorted(((s, d[s]) for s in stall_cols), key=lambda kv: -kv[1])[:3]
    print(f"{fn:28s} {100 * d['warp_inst'] / tot['warp_inst']:10.2f} {d['thread_inst'] / max(1, d['warp_inst']):6.2f} {100 * d['samples'] / tot['samples']:9.2f}  " +
          ", ".join(f"{s[6:]} {100 * v / max(1, d['samples']):.0f}%" for s, v in st))
