"""Summarise an ncu report: key raw metrics, stall reasons, and executed instructions attributed to source functions.
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep <kernel-mangled-substring> <warp_steps>"""
import collections, csv, io, re, subprocess, sys
rep, ksub, warp_steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; d = dict(zip(hdr, rows[2])); unit = dict(zip(hdr, rows[1]))
print(d["Kernel Name"][:60], d["Block Size"], d["Grid Size"])
for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
          "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
    print("  ", k, d.get(k), unit.get(k, ""))
st = [(k, float(v.replace(",", ""))) for k, v in d.items() if "smsp__average_warp" in k and "issue_stalled" in k and v not in ("", "n/a")]
for k, v in sorted(st, key=lambda x: -x[1])[:9]:
    print("     ", k.replace("smsp__average_warps_issue_stalled_", ""), v)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
subprocess.run("cd /tmp && rm -rf cub2 && mkdir cub2 && cd cub2 && cuobjdump -xelf all /root/repo/rl_rocket_b200/_C/librlrocket.so >/dev/null 2>&1", shell=True)
txt = subprocess.run(["nvdisasm", "-g", "/tmp/cub2/kernels_fast.sm_100a.cubin"], capture_output=True, text=True).stdout
m = re.search(r"\.text\.[^\n]*" + re.escape(ksub) + r".*?(?=\n\s*//-+ \.text\.|\Z)", txt, re.S)
off2line, cur = {}, None
for line in m.group(0).split("\n"):
    mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if mm:
        cur = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/", line)
    if m2 and cur: off2line[int(m2.group(1), 16)] = cur
srcl = open("/root/repo/rl_rocket_b200/csrc/rocket_kernels.cuh").read().split("\n")
if m is None:
    sys.exit(0)
def func_of(line):
    for i in range(line - 1, -1, -1):
        if re.match(r"^(RLR_DEV|__device__|template|__global__|struct|int launch)", srcl[i]) and "(" in srcl[i]:
            return re.sub(r"\(.*", "", srcl[i])[-36:]
    return "?"
base = int(data[0][ix["Address"]], 16)
ex, sm, byline = collections.Counter(), collections.Counter(), collections.Counter()
mix = collections.Counter()
for r in data:
    f, l = off2line.get(int(r[ix["Address"]], 16) - base, ("?", 0))
    key = func_of(l) if f == "rocket_kernels.cuh" else f
    n = int(r[ix["Instructions Executed"]]); ex[key] += n; sm[key] += int(r[ix["# Samples"]]); byline[l] += n
    tok = r[ix["Source"]].split(); op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]; mix[op] += n
tot, ts = sum(ex.values()), sum(sm.values())
print("warp-instr per warp-step:", tot / warp_steps, " static instrs:", len(data))
print("mix:", [(k, round(v / tot * 100, 1)) for k, v in mix.most_common(14)])
for k, v in ex.most_common(16):
    print(f"{v/tot*100:5.1f}% instr {sm[k]/ts*100:5.1f}% samples {v/warp_steps:7.0f}/warp-step  {k}")
print("hot lines:", [(l, round(v / warp_steps)) for l, v in byline.most_common(10)])
