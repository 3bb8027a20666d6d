#!/usr/bin/env python
"""Step times of BASELINE.json's other configurations (configs[2..3]: car go_to_goal + press_buttons at 32,768 envs,
point / car HaulBox + PushBox at 16,384 envs).  Not bench lines -- parity-test cases timed for DESIGN.md 7."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safe_adaptation_gym_b200 import tasks
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym


def run(robot, names, n, steps=200):
    env = BatchedSafeAdaptationGym("xmls/%s.xml" % robot, num_envs=n, device=torch.device("cuda:0"))
    env.seed(666)
    env.set_task([getattr(tasks, t)() for t in names])
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(steps + 10):
        a = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(a); e1.record(); torch.cuda.synchronize()
        if i >= 10:
            ts.append(e0.elapsed_time(e1))
    print(robot, names[:2], n, "env-steps/s %.3e" % (n * steps / (sum(ts) * 1e-3)),
          "ms/step first10 %.3f last10 %.3f" % (sum(ts[:10]) / 10, sum(ts[-10:]) / 10), flush=True)
    env.close()


if __name__ == "__main__":
    steps = int(os.environ.get("STEPS", "200"))
    which = os.environ.get("WHICH", "0,1,2").split(",")
    if "0" in which: run("car", ["GoToGoal", "PressButtons"] * 16384, 32768, steps)
    if "1" in which: run("point", ["HaulBox", "PushBox"] * 8192, 16384, steps)
    if "2" in which: run("car", ["HaulBox", "PushBox"] * 8192, 16384, steps)
