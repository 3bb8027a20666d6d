#!/usr/bin/env python
"""Workload statistics of the benchmark episode, on the CPU (test-infrastructure build of the kernel body with
-DSAG_PROFILE): per episode phase, how many environments are not quiet, how many run the contact solver, and how much
solver work a contact step is (passes, contacts, rows, sweeps).  Drives the design of the busy kernels.

    python tools/workload_stats.py [robot] [task] [n_envs] [steps]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from safe_adaptation_gym_b200 import _abi  # noqa: E402
from safe_adaptation_gym_b200.benchmark import TASKS  # noqa: E402
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym  # noqa: E402

NC = 16


def main():
    robot = sys.argv[1] if len(sys.argv) > 1 else "point"
    task = sys.argv[2] if len(sys.argv) > 2 else "go_to_goal"
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
    lib = os.path.join(ROOT, "tests", "hostemu", "libsag_hostemu_prof.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-DSAG_PROFILE", "-o", lib,
                           os.path.join(ROOT, "tests", "hostemu", "sag_hostemu.cpp")])
    L = _abi.SagLib(lib, host_api=False)
    env = BatchedSafeAdaptationGym("xmls/%s.xml" % robot, num_envs=n, _test_lib=L)
    env.seed(666)
    env.set_task([TASKS[task]() for _ in range(n)])
    prof = np.zeros((n, NC), dtype=np.int64)
    L.L.sag_prof_set.argtypes = [C.c_void_p]
    L.L.sag_prof_set(prof.ctypes.data_as(C.c_void_p))
    g = torch.Generator(); g.manual_seed(1234)
    names = ["passes", "collide", "ncon", "need_sub", "rowvis", "floorvis", "pairchk", "notquiet", "sweeps", "nrow", "nb", "anyrow"]
    print("step  notquiet%  contact%(need>0)  passes/cenv  need_sub/cenv ncon/pass nrow/pass nb/pass sweeps/pass rowvis/pass floorvis/pass anyrow/pass")
    acc = np.zeros(NC)
    hist = {}
    accn = np.zeros(3)
    for t in range(steps):
        prof[:] = 0
        act = torch.empty((n, 2), dtype=torch.float32).uniform_(-1, 1, generator=g)
        env.step(act)
        nq = (prof[:, 7] > 0).sum()
        cen = (prof[:, 0] > 0)
        nc = cen.sum()
        acc += prof.sum(0); accn += [n, nq, nc]
        if t >= steps - 200:
            for e in np.nonzero(cen)[0]:
                key = (int(prof[e, 12]), int(prof[e, 14]), int(prof[e, 13]), int(prof[e, 15]))
                hist[key] = hist.get(key, 0) + 1
        if (t + 1) % int(os.environ.get('EVERY','50')) == 0:
            p = acc
            ps = max(p[0], 1)
            print(f"{t+1:4d}  {100*accn[1]/accn[0]:6.2f}  {100*accn[2]/accn[0]:6.2f}   {p[0]/max(accn[2],1):5.2f} {p[3]/max(accn[2],1):5.2f}  {p[2]/ps:5.2f} {p[9]/ps:5.2f} {p[10]/ps:5.2f} "
                  f"{p[8]/ps:5.2f} {p[4]/ps:6.2f} {p[5]/ps:6.2f} {p[11]/ps:5.2f}", flush=True)
            acc[:] = 0; accn[:] = 0
    tot = sum(hist.values())
    print("contact env-steps of the last 200 steps by (max ncon, max nrow, max nb, any object-object contact): share")
    for k, v in sorted(hist.items(), key=lambda kv: -kv[1]):
        print(k, f"{100.0 * v / tot:6.2f} %")


if __name__ == "__main__":
    main()
