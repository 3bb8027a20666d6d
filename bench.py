#!/usr/bin/env python
"""bench.py -- env-steps/sec of the point go_to_goal hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on all host cores

Workload (N=1): BASELINE.json configs[1] -- point go_to_goal, 65,536 environments per GPU, pseudo-lidar
observation + hazard/vase/pillar cost, i.i.d. U(-1,1) actions, action_noise 0.01 (reference default).
A "step" = one env.step over the whole batch (5 physics substeps per env, reward, cost, 60-float obs).

value        : env-steps/s with actions resident in HBM, timed per step with CUDA events on the launching
               stream, L2 flushed (256 MiB memset) between steps.
e2e          : same metric through sag_step_host (C ABI, pinned HOST buffers): H2D of the actions and D2H of
               obs/reward/cost/done inside the timed region, every step (the bulk D2H overlaps the busy kernel).
roofline     : algorithmic bytes per env-step (2582 B for the fused point go_to_goal step)
               x envs / average k_step duration vs the measured HBM copy bandwidth.
cpu_baseline : the oracle port (oracle/sag_oracle.c, pthreads) on the box's host cores, bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 65536
B_ALG_STEP = 2582  # fused step, point go_to_goal, vases simulated [bytes / env-step]
METRIC = "env_steps_per_sec_point_go_to_goal"
UNIT = "env-steps/s"
WORKLOAD = "point go_to_goal, 65536 envs/GPU, lidar obs + hazard/vase/pillar cost, U(-1,1) actions"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(n_envs, steps, threads, seed=666):
    """env-steps/s of the oracle port on `threads` host threads (bounded sample of the same workload)."""
    import oracle as O
    envs = [O.OracleEnv("point", "go_to_goal", seed=seed, env_gid=i) for i in range(n_envs)]
    for e in envs:
        e.reset(0)
    t0 = time.perf_counter()
    count, sr, sc = O.batch_rollout(envs, steps, threads)
    dt = time.perf_counter() - t0
    return count / dt, count, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = 64 * cores
    rates = []
    for _ in range(args.warmup):
        cpu_oracle_rate(n_envs, 20, cores)
    t_total, c_total = 0.0, 0
    sample_steps = 250
    for _ in range(args.steps):
        r, c, dt = cpu_oracle_rate(n_envs, sample_steps, cores)
        rates.append(r); t_total += dt; c_total += c
        if t_total > 120:
            break
    value = c_total / t_total
    sample = f"{n_envs} envs x {sample_steps} steps per bench step, {len(rates)} bench steps, oracle port (C, pthreads)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, len(rates)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's own stack (dm_control/MuJoCo) is not installable here; this arm times the CPU oracle port"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps (capped at 2000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from safe_adaptation_gym_b200 import _abi, tasks
    from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs_per_gpu
    K, W = args.steps, max(3, args.warmup)
    env = BatchedSafeAdaptationGym("xmls/point.xml", num_envs=n, device=dev, env_id_base=rank * n, max_episode_steps=0)
    env.seed(666)
    env.set_task(tasks.GoToGoal())
    L, h = env._lib, env._h
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)
    # synthetic actions, resident in HBM: i.i.d. U(-1,1) per env per step (env.action_space.sample()), drawn on the device
    # outside the timed event pair.  (A short cyclic ring of action batches would give every env a periodic action
    # sequence, i.e. a systematic drift into obstacles -- not what a random policy does.)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    act_buf = torch.empty((n, 2), dtype=torch.float32, device=dev)

    def new_actions():
        act_buf.uniform_(-1.0, 1.0, generator=g)
    obs, rew, cost, done = env._obs, env._reward, env._cost, env._done
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    p = BatchedSafeAdaptationGym._p

    def launch(i):
        L.check(L.L.sag_step(h, p(act_buf), p(obs), p(rew), None, p(cost), p(done), sp))

    nstep = 0
    for i in range(W):
        new_actions(); launch(i); nstep += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches = 0
    t_wall0 = time.perf_counter()
    for i in range(K):
        new_actions()
        flush.zero_()  # L2 flush (not timed: outside the event pair)
        evs[i][0].record(stream)
        launch(W + i); launches += 2; nstep += 1  # k_step_quiet + k_step_coop
        if nstep % 1000 == 0:  # episode length used by the reference's tooling (tests/test_safety_gym.py:78)
            L.check(L.L.sag_reset(h, None, 0, 0, sp)); launches += 1
        evs[i][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(ms))
    # warm-L2 back-to-back variant (state stays resident in the 126 MB L2, as in a real RL loop)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        new_actions(); launch(i)
    e1.record(stream)
    torch.cuda.synchronize()
    warm_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms, warm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, warm_ms = float(t[0]), float(t[1])
    value = world * n * K / (total_ms * 1e-3)
    value_warm = world * n * K / (warm_ms * 1e-3)

    # ---- end-to-end through the C ABI with pinned host buffers, on a second handle reset to the same episode so that
    # it covers the same episode phase (steps W .. W+Ke after reset) as `value`
    Ke = args.e2e_steps or min(K, 2000)
    env2 = BatchedSafeAdaptationGym("xmls/point.xml", num_envs=n, device=dev, env_id_base=rank * n, max_episode_steps=0)
    env2.seed(666)
    env2.set_task(tasks.GoToGoal())
    h2 = env2._h
    act_h = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    obs_h = torch.empty((n, env.obs_dim), dtype=torch.float32).pin_memory()
    rew_h = torch.empty((n,), dtype=torch.float64).pin_memory()
    cost_h = torch.empty((n,), dtype=torch.uint8).pin_memory()
    done_h = torch.empty((n,), dtype=torch.uint8).pin_memory()
    cpu_gen = torch.Generator(); cpu_gen.manual_seed(99 + rank)
    act_h.uniform_(-1, 1, generator=cpu_gen)
    torch.cuda.synchronize()
    for i in range(W):
        L.check(L.L.sag_step_host(h2, p(act_h), p(obs_h), p(rew_h), p(cost_h), p(done_h)))
    if world > 1:
        dist.barrier()
    te = 0.0
    for i in range(Ke):
        act_h.uniform_(-1, 1, generator=cpu_gen)   # the host policy's work is not timed
        t0 = time.perf_counter()
        L.check(L.L.sag_step_host(h2, p(act_h), p(obs_h), p(rew_h), p(cost_h), p(done_h)))  # H2D + kernels + D2H + sync
        te += time.perf_counter() - t0
    torch.cuda.synchronize()
    t = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    te = float(t[0])
    e2e_value = world * n * Ke / te
    h2d = n * 2 * 4
    d2h = n * (env.obs_dim * 4 + 8 + 1 + 1)
    env2.close()

    # ---- per-task statistics: the only collective on this path (NCCL all-reduce of a [14,3] fp64 buffer)
    L.check(L.L.sag_reset(h, None, 0, 0, sp))
    stats = env.task_stats(reset=True)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    stats = stats.cpu().numpy()

    if rank == 0:
        peak, peak_src = peaks()
        avg_s = total_ms * 1e-3 / K
        achieved = B_ALG_STEP * n / avg_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("step_dram_bytes_per_launch")
            except Exception:
                traffic = None
        first10_s = float(sum(ms[:10]) / max(1, len(ms[:10]))) * 1e-3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "global_envs": world * n, "parallelism": f"env-shard x{world}",
                       "l2": "flushed between steps (256 MiB memset outside the event pair)", "action_noise": 0.01,
                       "episode_phase": f"steps {W}..{W + K} after reset (1000-step episodes)"},
            "value_l2_warm": value_warm,
            "value_l2_warm_note": f"back-to-back steps without the flush, LATER episode phase (steps {W + K}..{W + 2 * K}): more contacts than `value`'s phase",
            "ms_per_step_first10": float(sum(ms[:10]) / max(1, len(ms[:10]))), "ms_per_step_last10": float(sum(ms[-10:]) / max(1, len(ms[-10:]))),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "api": "sag_step_host (C ABI, pinned host buffers)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "fused step = k_step_quiet + k_step_coop", "b_alg_per_env_step": B_ALG_STEP,
                         "peak_source": peak_src,
                         "frac_quiet_phase": B_ALG_STEP * n / first10_s / 1e9 / peak},
            "clocks": clocks,
            "episode_stats": {"go_to_goal": {"sum_return": float(stats[3, 0]), "sum_cost": float(stats[3, 1]),
                                             "episodes": float(stats[3, 2])}},
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            ne, ns = 64 * cores, 1000
            r, c, dt = cpu_oracle_rate(ne, ns, cores)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{ne} envs x {ns} steps = {c} env-steps in {dt:.1f}s, oracle port (C, pthreads); "
                                              "the reference's MuJoCo stack is not installable here"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
