"""Attribute an ncu report's per-instruction counters to source lines: joins `ncu --page source --print-source sass`
with the line info of `nvdisasm -g` on the library's cubin (innermost inlined location).
usage: python scripts/ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "flame_ros_b200", "lib", "libflame_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
asm = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(asm) if ".section" in l and ".text." in l and kern in l][0]
end = [i for i, l in enumerate(asm) if i > start + 5 and l.startswith("//---------------------")][0]
cur, addr2line = None, {}
for l in asm[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
inst, samp, tot_i, tot_s = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    a, n, sm = int(r[ia], 16) - base, int(r[ie]), int(r[isamp])
    k = addr2line.get(a, (None, ""))[0]
    inst[k] += n
    samp[k] += sm
    tot_i += n
    tot_s += sm
print("kernel %s: %d warp instructions, %d samples" % (kern, tot_i, tot_s))
print("| source line | samples | % | warp instructions | % |\n|---|---|---|---|---|")
for k, sm in samp.most_common(top):
    print("| %s:%s | %d | %.1f | %d | %.1f |" % (k[0] if k else "?", k[1] if k else "?", sm, 100.0 * sm / max(tot_s, 1), inst[k], 100.0 * inst[k] / max(tot_i, 1)))
