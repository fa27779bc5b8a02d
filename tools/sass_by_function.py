"""Static SASS size of one kernel by SOURCE FUNCTION: nvdisasm --print-line-info output -> instructions per function.
usage: tools/sass_by_function.py <nvdisasm --print-line-info output> <kernel-name-substring>"""
import bisect
import collections
import os
import re
import sys

disasm, kern = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hyper-ray-tracer_b200", "csrc")


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


def func_of(f, ln):
    if f not in fmap:
        return f
    lst = fmap[f]
    i = bisect.bisect_right([x[0] for x in lst], ln) - 1
    return lst[i][1] if i >= 0 else f


count = collections.Counter()
cur, infunc, sect = None, False, None
for line in open(disasm):
    m = re.search(r"\.section\s+\.text\.(\S+),", line)
    if m:
        infunc = kern in m.group(1)
        sect = m.group(1)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line) and cur:
        count[func_of(*cur)] += 1
tot = sum(count.values())
print(f"{kern}: {tot} instructions = {tot * 16 / 1024:.1f} KB")
for fn, n in count.most_common(40):
    print(f"  {fn:28s} {n:6d}  {100 * n / tot:5.1f} %")
