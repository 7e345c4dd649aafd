"""Development aid: which warps make the tail of one step_kernel<Sys6> launch (from a -DRLR_WARP_TRACE=1 tuning build).
usage: RLR_LIB=rl_rocket_b200/_C/librlrocket_WT.so python scripts/warp_trace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rl_rocket_b200 import _native as N
from rl_rocket_b200.config import EnvSpec, env_config_6dof
from rl_rocket_b200.engine import RocketBatch

lib = N.load()
n = int(os.environ.get("QB_N", "65536"))
spec = EnvSpec(6, max_episode_steps=800, **env_config_6dof())
env = RocketBatch(spec, n, seed=1, math_mode="fast"); env.reset()
a = torch.rand((200, n, 3), device="cuda") * 2 - 1
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
for t in range(160): env.step_tensor(a[t])
buf = np.zeros((8192, 16), np.uint64)
for rep in range(3):
    flush.sum(); torch.cuda.synchronize()
    env.step_tensor(a[160 + rep]); torch.cuda.synchronize()
    lib.rlr_debug_warp_trace(buf.ctypes.data_as(C.c_void_p))
    w = buf[: n // 32].astype(np.int64)
    sm, t0, t1, fl = w[:, 0], w[:, 1], w[:, 2], w[:, 3]
    base = t0.min()
    s, e = (t0 - base) / 1e3, (t1 - base) / 1e3
    att, ev, fin = fl & 255, (fl >> 8) & 255, (fl >> 16) & 255
    d = e - s
    print(f"--- launch {rep}: {n // 32} warps; first start 0, last start {s.max():.2f} us, last end {e.max():.2f} us")
    print("    warp duration us: mean %.2f  p50 %.2f  p90 %.2f  p99 %.2f  max %.2f" % (d.mean(), *np.percentile(d, [50, 90, 99]), d.max()))
    print("    end time us:      mean %.2f  p50 %.2f  p90 %.2f  p99 %.2f  max %.2f" % (e.mean(), *np.percentile(e, [50, 90, 99]), e.max()))
    for name, m in (("2 attempts, no event, no reset", (att == 2) & (ev == 0) & (fin == 0)), ("3+ attempts", att >= 3), ("event", ev > 0),
                    ("reset", fin > 0), ("1 attempt", att == 1), ("3+ att & event", (att >= 3) & (ev > 0)), ("event & reset", (ev > 0) & (fin > 0))):
        if m.any():
            print(f"    {name:32s} {m.mean()*100:5.1f}% of warps: duration mean {d[m].mean():6.2f} max {d[m].max():6.2f}  end mean {e[m].mean():6.2f} max {e[m].max():6.2f}")
    per_sm_end = np.array([e[sm == k].max() for k in np.unique(sm)])
    per_sm_cnt = np.array([(sm == k).sum() for k in np.unique(sm)])
    print("    per-SM last end us: mean %.2f min %.2f max %.2f; warps per SM min %d max %d; SMs used %d" %
          (per_sm_end.mean(), per_sm_end.min(), per_sm_end.max(), per_sm_cnt.min(), per_sm_cnt.max(), len(per_sm_end)))
    late = np.argsort(e)[-8:]
    print("    latest warps (end us, dur us, att, ev, fin, sm):", [(round(e[i], 2), round(d[i], 2), int(att[i]), int(ev[i]), int(fin[i]), int(sm[i])) for i in late])
    # per-warp arrival times (us after the warp's start) at the phase marks, in program order
    order = [(0, "loads"), (1, "init"), (2, "RK"), (3, "event"), (4, "reward"), (5, "out-st"), (9, "fin-st"), (6, "state-st"), (8, "obs+reset"), (7, "stats")]
    ph = (w[:, 4:14] - t0[:, None]) / 1e3
    def row(sel):
        return "  ".join(f"{nm} {ph[sel][:, k].mean():5.2f}" for k, nm in order)
    common = (att == 2) & (ev == 0) & (fin == 0)
    print("    marks, common warps (mean):   ", row(common))
    print("    marks, 3-attempt, no event/reset:", row((att == 3) & (ev == 0) & (fin == 0)))
    print("    marks, event warps (mean):    ", row(ev > 0))
    print("    marks, reset-only warps (mean):", row((ev == 0) & (fin > 0)))
    for i in late:
        print(f"    marks, late warp att {int(att[i])} ev {int(ev[i])} fin {int(fin[i])} start {s[i]:.2f}: ", "  ".join(f"{nm} {ph[i, k]:5.2f}" for k, nm in order))
    hist, edges = np.histogram(e, bins=12)
    print("    end-time histogram:", list(zip(np.round(edges[:-1], 1), hist)))
