"""Development aid: stall samples of an ncu report (--set full --import-source on) per source function / source line, with the
stall reasons that make them up.  usage: python scripts/ncu_lines.py <rep.ncu-rep> <kernel-mangled-substring> [lib.so] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, ksub = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "/root/repo/rl_rocket_b200/_C/librlrocket.so"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
d = tempfile.mkdtemp()
subprocess.run(f"cd {d} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null 2>&1", shell=True)
txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, "kernels_fast.sm_100a.cubin")], capture_output=True, text=True).stdout
m = re.search(r"\.text\.[^\n]*" + re.escape(ksub) + r".*?(?=\n\s*//-+ \.text\.|\Z)", txt, re.S)
off2line, cur = {}, None
for line in m.group(0).split("\n"):
    mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if mm:
        cur = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/", line)
    if m2 and cur: off2line[int(m2.group(1), 16)] = cur
srcl = open("/root/repo/rl_rocket_b200/csrc/rocket_kernels.cuh").read().split("\n")
def func_of(line):
    for i in range(min(line, len(srcl)) - 1, -1, -1):
        if re.match(r"^(RLR_DEV|__device__|template|__global__|struct|int launch)", srcl[i]) and "(" in srcl[i]:
            return re.sub(r"\(.*", "", srcl[i])[-36:]
    return "?"
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ix["Address"]], 16)
byfunc = collections.defaultdict(collections.Counter); byline = collections.defaultdict(collections.Counter)
ex_line = collections.Counter(); tot = 0
for r in data:
    f, l = off2line.get(int(r[ix["Address"]], 16) - base, ("?", 0))
    key = func_of(l) if f == "rocket_kernels.cuh" else f
    lk = (f, l)
    n = int(r[ix["# Samples"]]); tot += n
    ex_line[lk] += int(r[ix["Instructions Executed"]])
    for h in reasons:
        v = int(r[ix[h]] or 0)
        if v: byfunc[key][h] += v; byline[lk][h] += v
print("total samples", tot)
print("--- by function")
for k, c in sorted(byfunc.items(), key=lambda kv: -sum(kv[1].values()))[:22]:
    s = sum(c.values())
    print(f"{s/tot*100:5.1f}%  {k:40s} " + " ".join(f"{h[6:]}={v/s*100:.0f}%" for h, v in c.most_common(4)))
print("--- by line")
for k, c in sorted(byline.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(c.values())
    text = srcl[k[1] - 1].strip()[:70] if k[0] == "rocket_kernels.cuh" and 0 < k[1] <= len(srcl) else k[0]
    print(f"{s/tot*100:5.2f}%  L{k[1]:<5d} x{ex_line[k]:<8d} " + " ".join(f"{h[6:]}={v/s*100:.0f}%" for h, v in c.most_common(3)) + "   | " + text)
