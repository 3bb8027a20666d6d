#!/usr/bin/env python
"""Reset latency probe (CUDA events): whole-batch reset at a few batch sizes, and the sparse auto-reset of the
environments that expire in a step.     python tools/probe_reset.py [robot] [task]"""
import ctypes as C
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safe_adaptation_gym_b200.benchmark import TASKS  # noqa: E402
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym  # noqa: E402

robot = sys.argv[1] if len(sys.argv) > 1 else "point"
task = sys.argv[2] if len(sys.argv) > 2 else "go_to_goal"
dev = torch.device("cuda:0")
p = BatchedSafeAdaptationGym._p
out = {"robot": robot, "task": task, "whole_batch_reset_us": {}, "with_first_observation_us": {}, "sparse_reset_us": {}}
for n in (256, 4096, 65536):
    env = BatchedSafeAdaptationGym("xmls/%s.xml" % robot, num_envs=n, device=dev)
    env.seed(666)
    env.set_task(TASKS[task]())
    L, h = env._lib, env._h
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    wr = torch.zeros(n, dtype=torch.uint8, device=dev)
    for key, fn in (("whole_batch_reset_us", lambda: L.L.sag_reset(h, None, 0, 0, sp)),
                    ("with_first_observation_us", lambda: L.L.sag_reset_obs(h, None, 0, 0, p(env._obs), p(wr), sp))):
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); L.check(fn()); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        out[key][str(n)] = {"median": round(ts[3], 1), "min": round(ts[0], 1), "max": round(ts[-1], 1)}
    # sparse: k environments (spread over the batch) carry the NEEDS_RESET flag
    for k in (1, 64, 256):
        if k > n:
            continue
        ts = []
        for _ in range(7):
            fl = env.get_field("flags")
            idx = (torch.arange(k, device=dev) * (n // k)).long()
            fl[idx] |= 4
            env.set_field("flags", fl)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); L.check(L.L.sag_reset_obs(h, None, 1, 0, p(env._obs), p(wr), sp)); b.record(stream)
            torch.cuda.synchronize()
            assert int(wr.sum()) == k, (int(wr.sum()), k)
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        out["sparse_reset_us"]["%d of %d" % (k, n)] = {"median": round(ts[3], 1), "min": round(ts[0], 1), "max": round(ts[-1], 1)}
    env.close()
print(json.dumps(out))
