#!/usr/bin/env python
"""Step time (CUDA events, warm L2) of the benchmark workload at a few episode phases -- a quick probe for kernel tuning.
    python tools/late_phase.py [config] [phases...]          SAG_B200_LIB=<variant .so> selects a build variant."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from safe_adaptation_gym_b200.benchmark import TASKS  # noqa: E402
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "point_gtg"
phases = [int(x) for x in sys.argv[2:]] or [60, 100, 300, 600, 900]
cfg = bench.CONFIGS[cfgname]
n = int(os.environ.get("ENVS", cfg["envs"]))
dev = torch.device("cuda:0")
env = BatchedSafeAdaptationGym("xmls/%s.xml" % cfg["robot"], num_envs=n, device=dev)
env.seed(666)
tasks = [TASKS[t]() for t in cfg["tasks"]]
env.set_task([tasks[e % len(tasks)] for e in range(n)])
L, h = env._lib, env._h
p = BatchedSafeAdaptationGym._p
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1234)
act = torch.empty((n, 2), dtype=torch.float32, device=dev)
t = 0
out = []
for ph in phases:
    while t < ph:
        act.uniform_(-1, 1, generator=g)
        L.check(L.L.sag_step(h, p(act), p(env._obs), p(env._reward), None, p(env._cost), p(env._done), sp)); t += 1
    ts = []
    for _ in range(10):
        act.uniform_(-1, 1, generator=g)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        L.check(L.L.sag_step(h, p(act), p(env._obs), p(env._reward), None, p(env._cost), p(env._done), sp)); t += 1
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    out.append(f"{ph}: med {ts[5]:.0f} min {ts[0]:.0f} max {ts[-1]:.0f}")
    if os.environ.get("SAG_TIMING_READ"):
        buf = (C.c_ulonglong * 18)()
        L.L.sag_debug_read.argtypes = [C.c_void_p, C.c_void_p]
        L.check(L.L.sag_debug_read(h, buf))
        v = list(buf)
        ne = max(1, v[15])
        names = ["prologue", "substep-pre", "det-broad1", "setup-rows", "pgs", "post", "robot-int", "eos-passA", "setup-table", "eos-rest", "epilogue", "det-narrow1", "det-broad2", "det-narrow2", "env_step total", "env-steps"]
        out.append("\n   cycles/env-step: " + ", ".join(f"{names[i]} {v[i] / ne:.0f}" for i in range(16) if v[i] and i != 15) + f" | env-steps {v[15]}\n"
                   + f"   slowest env-step of the interval: total {v[16] >> 40} cycles, pgs {((v[16] >> 20) & 0xfffff) << 10}, detect {(v[16] & 0xfffff) << 10}, setup {(v[17] & 0xffffffffff) << 10}\n")
print(os.environ.get("SAG_B200_LIB", "default").split("/")[-1], cfgname, "step us @phase |", " | ".join(out))
