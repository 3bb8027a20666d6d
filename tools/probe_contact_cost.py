#!/usr/bin/env python
"""Probe: how long is a step when exactly K environments are in contact and all others are quiet?
Puts the robot of K environments (env ids 0, 32, 64, ...) next to their first vase, driving into it, and times
sag_step with CUDA events.  PROBE_K=0,256,2048 selects the K values."""
import ctypes as C
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safe_adaptation_gym_b200 import tasks
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym

n = 65536
dev = torch.device("cuda:0")


def time_steps(env, act, steps=12):
    L, h = env._lib, env._h
    p = BatchedSafeAdaptationGym._p
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    ts = []
    for i in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        L.check(L.L.sag_step(h, p(act), p(env._obs), p(env._reward), None, p(env._cost), p(env._done), sp))
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return ts


for K, stride in [(int(k), 32) for k in os.environ.get("PROBE_K", "0,1,64,2048").split(",")]:
    env = BatchedSafeAdaptationGym("xmls/point.xml", num_envs=n, device=dev, config={"action_noise": 0.0})
    env.seed(1); env.set_task(tasks.GoToGoal())
    _ = env.observation
    robot = env.get_field("robot"); objs = env.get_field("objects")
    idx = torch.arange(K, device=dev) * stride
    if K:
        vx, vy = objs[0, 9, idx], objs[1, 9, idx]          # first vase (slot 9)
        robot[0, idx] = vx - 0.205; robot[1, idx] = vy; robot[2, idx] = 0.0   # sphere 5 mm from the vase face, arrow overlapping
        robot[3, idx] = 0.3; robot[4, idx] = 0.0; robot[5, idx] = 0.0
        objs[2, 9, idx] = 0.0
        env.set_field("robot", robot); env.set_field("objects", objs)
    _ = env.observation
    act = torch.zeros((n, 2), device=dev); act[:, 0] = 1.0
    ts = time_steps(env, act)
    moved = float((env.get_field("objects")[0, 9, idx] - (vx if K else 0)).abs().max()) if K else 0.0
    print(f"K={K:5d} stride={stride:2d}  step us: " + " ".join(f"{t:7.1f}" for t in ts) + f"   vase moved {moved:.3f}")
    env.close()
