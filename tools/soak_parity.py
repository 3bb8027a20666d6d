#!/usr/bin/env python
"""One-off soak: the CUDA path against the CPU oracle, bit for bit, over long episodes of every task and both robots
(tests/common.run_parity with more environments and steps than the test-suite uses).  Test infrastructure, not product."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import run_parity  # noqa: E402

TASKS = ["go_to_goal", "go_to_goal_scarce", "go_to_goal_damping", "go_to_goal_motor", "catch_goal", "unsupervised", "press_buttons",
         "press_buttons_scarce", "collect", "push_box", "push_box_scarce", "haul_box", "roll_rod", "dribble_ball"]

if __name__ == "__main__":
    n = int(os.environ.get("SOAK_N", "48"))
    steps = int(os.environ.get("SOAK_STEPS", "1000"))
    seed0 = int(os.environ.get("SOAK_SEED", "1000"))
    cfg = {"action_noise": 0.01}
    if os.environ.get("SOAK_KNOBS"):   # adaptation knobs on: Cauchy-scaled ctrl ranges (incl. inverted ones), random bound
        cfg.update({"robot_ctrl_range_scale": 0.5, "random_bound": True})
    if os.environ.get("SOAK_GREMLINS"):   # a user-defined task with gremlins on top of every registry task (world.py:157-165)
        cfg["num_gremlins"] = int(os.environ["SOAK_GREMLINS"])
    for robot in ("point", "car"):
        for i, task in enumerate(TASKS):
            t0 = time.time()
            for policy in ("drive", "random"):
                s = run_parity("cuda", task, n=n, steps=steps if policy == "drive" else steps // 4, seed=seed0 + i, policy=policy,
                               check_every=5, robot=robot, config=cfg)
                print(robot, task, policy, {k: (round(float(v), 3) if not isinstance(v, int) else v) for k, v in s.items()},
                      "%.0fs" % (time.time() - t0), flush=True)
    print("soak ok")
