#!/usr/bin/env python
"""Stand-alone lidar / cost kernel micro-benchmark (573 B / env and 90 B / env algorithmic).

    python bench_kernels.py [--envs 65536] [--iters 50]

Times sag_lidar and sag_cost on caller-owned SoA buffers with CUDA events, L2 flushed between launches,
and reports achieved algorithmic GB/s against the measured HBM peak.  One JSON line per kernel."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from safe_adaptation_gym_b200 import _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    L = _abi.load()
    n, nslots, nh = args.envs, 21, 9
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(0)
    robot = torch.rand((3, n), dtype=torch.float64, device=dev, generator=g) * 4 - 2
    obj = torch.rand((2, nslots, n), dtype=torch.float64, device=dev, generator=g) * 4 - 2
    group = torch.ones((nslots, n), dtype=torch.uint8, device=dev)
    group[20] = 2
    out = torch.empty((n, 48), dtype=torch.float32, device=dev)
    rxy = robot[:2].contiguous()
    hz = (torch.rand((2, nh, n), device=dev, generator=g) * 4 - 2).contiguous()
    contact = torch.zeros(n, dtype=torch.uint8, device=dev)
    cost = torch.empty(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peak = 6538.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def run(name, fn, bytes_per_env):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            if not args.no_flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        med = ts[len(ts) // 2] * 1e-3
        ach = bytes_per_env * n / med / 1e9
        print(json.dumps({"kernel": name, "envs": n, "us_median": med * 1e6, "us_min": ts[0] * 1e3, "b_alg_per_env": bytes_per_env,
                          "achieved_gbs": ach, "peak_gbs": peak, "frac": ach / peak, "l2": "warm" if args.no_flush else "flushed"}), flush=True)

    run("k_lidar", lambda: L.check(L.L.sag_lidar(robot.data_ptr(), obj.data_ptr(), group.data_ptr(), n, nslots, out.data_ptr(), sp)), 573)
    run("k_cost", lambda: L.check(L.L.sag_cost(rxy.data_ptr(), hz.data_ptr(), contact.data_ptr(), n, nh, C.c_double(0.2), cost.data_ptr(), sp)), 90)


if __name__ == "__main__":
    main()
