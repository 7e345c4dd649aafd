"""Development aid: static SASS instruction count of a kernel attributed to source functions (inlined code included).
usage: python scripts/static_by_func.py <lib.so> <cubin-substr> <kernel-substr>"""
import re, sys, subprocess, collections, os, tempfile
lib, cub, kern = sys.argv[1:4]
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null 2>&1", shell=True)
f = [x for x in os.listdir(d) if cub in x][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout
m = re.search(r"\.text\.[^\n]*" + re.escape(kern) + r".*?(?=\n\s*//-+ \.text\.|\Z)", txt, re.S)
srcs = {}
def func_of(fn, line):
    if fn not in srcs:
        try: srcs[fn] = open(os.path.join("/root/repo/rl_rocket_b200/csrc", fn)).read().split("\n")
        except OSError: srcs[fn] = None
    s = srcs[fn]
    if s is None: return fn
    for i in range(min(line, len(s)) - 1, -1, -1):
        if re.match(r"^(RLR_DEV|__device__|template|__global__|struct|int launch)", s[i]) and "(" in s[i]:
            return re.sub(r"\(.*", "", s[i])[-40:]
    return "?"
cur = None; cnt = collections.Counter(); tot = 0
for line in m.group(0).split("\n"):
    mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if mm: cur = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        tot += 1; cnt[func_of(*cur)] += 1
print("static", tot)
for k, v in cnt.most_common(30): print(f"{v:6d}  {k}")
