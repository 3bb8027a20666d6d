#!/usr/bin/env python
"""Wall-clock of env.reset() through the host API (sag_reset_host / sag_observe_host) after 30 steps, for three
BASELINE configs -- the probe that exposed the local-memory re-size stall."""
import ctypes as C, os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
from safe_adaptation_gym_b200.benchmark import TASKS
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym
for cfgname in ("point_gtg", "haul_push_car", "car_gtg_pb"):
    cfg = bench.CONFIGS[cfgname]; n = cfg["envs"]
    env = BatchedSafeAdaptationGym("xmls/%s.xml" % cfg["robot"], num_envs=n, device="cuda:0")
    env.seed(666); tasks = [TASKS[t]() for t in cfg["tasks"]]
    env.set_task([tasks[e % len(tasks)] for e in range(n)])
    L, h = env._lib, env._h; p = BatchedSafeAdaptationGym._p
    obs_h = torch.empty((n, env.obs_dim), dtype=torch.float32).pin_memory()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    act = torch.zeros((n, 2), device="cuda")
    out = []
    for rep in range(4):
        for _ in range(30):
            L.check(L.L.sag_step(h, p(act.uniform_(-1, 1)), p(env._obs), p(env._reward), None, p(env._cost), p(env._done), sp))
        torch.cuda.synchronize()
        t0 = time.perf_counter(); L.check(L.L.sag_reset_host(h, None, 0, 0, None)); t1 = time.perf_counter()
        L.check(L.L.sag_observe_host(h, p(obs_h))); t2 = time.perf_counter()
        L.check(L.L.sag_reset_host(h, None, 0, 0, p(obs_h))); t3 = time.perf_counter()
        out.append("reset %.2f ms, observe_host %.2f ms, reset+obs %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
    print(cfgname, " | ".join(out))
    env.close()
