"""Development aid: where (source line) a kernel's local-memory spill instructions (STL/LDL) sit.
usage: python scripts/spill_map.py <lib.so> <cubin-name-substring> <kernel-name-substring>"""
import re, sys, subprocess, collections, os, tempfile
lib, cub, kern = sys.argv[1:4]
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null 2>&1", shell=True)
f = [x for x in os.listdir(d) if cub in x][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout
m = re.search(r"\.text\.[^\n]*" + re.escape(kern) + r".*?(?=\n\s*//-+ \.text\.|\Z)", txt, re.S)
cur = None; tot = 0; cnt = collections.Counter(); ops = collections.Counter()
for line in m.group(0).split("\n"):
    mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if mm: cur = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    m2 = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m2:
        tot += 1; ops[m2.group(1).split(".")[0]] += 1
        if " STL" in line or " LDL" in line: cnt[(cur, "STL" if " STL" in line else "LDL")] += 1
print("static instructions", tot, dict(ops.most_common(12)))
agg = collections.Counter()
for (c, k), v in cnt.items(): agg[c] += v
for c, v in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])): print(c, v)
