#!/usr/bin/env python
"""Python-API overhead: env.step() (tensors in, tensors out, device resident) against the raw sag_step call it wraps, at
steps 300-500 of the benchmark episode."""
import ctypes as C
import os
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_adaptation_gym_b200 as sag  # noqa: E402
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym  # noqa: E402

n = 65536
for copy in (None, None, False, True):
    env = sag.make("point", "go_to_goal", seed=666, num_envs=n, device="cuda:0")
    if copy is not None:
        env.copy_outputs = copy
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    act = torch.empty((n, 2), device="cuda")
    env.rollout(300)
    torch.cuda.synchronize()
    L, h, p = env._lib, env._h, BatchedSafeAdaptationGym._p
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    for _ in range(200):
        act.uniform_(-1, 1, generator=g)
        if copy is None:
            L.check(L.L.sag_step(h, p(act), p(env._obs), p(env._reward), None, p(env._cost), p(env._done), sp))
        else:
            obs, rew, done, info = env.step(act)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-28s %.3f ms/step  %.3e env-steps/s" % ("raw sag_step" if copy is None else "env.step copy_outputs=%s" % copy, 1e3 * dt / 200, n * 200 / dt))
    env.close()
