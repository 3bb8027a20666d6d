"""N>1 path on CPU: two gloo ranks, each owning a shard of the global env-id range; the only collective is the
all-reduce of the per-task statistics buffer."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import env_state, make_env
    from safe_adaptation_gym_b200.stats import reduce_task_stats
    n = 32
    names = ["go_to_goal", "press_buttons"] * (n // 2)
    env = make_env("hostemu", n, names, seed=7, env_id_base=rank * n, max_episode_steps=20)
    act = torch.zeros((n, 2))
    act[:, 0] = 0.5
    for _ in range(45):
        env.step(act)
    local = env.task_stats().clone()
    total = reduce_task_stats(env)
    robot, _ = env_state(env)
    q.put((rank, local.numpy(), total.numpy(), robot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, l0, t0, rob0), (r1, l1, t1, rob1) = res
    np.testing.assert_array_equal(t0, t1)
    np.testing.assert_array_equal(t0, l0 + l1)
    assert t0[3, 2] == 2 * 16 * 2 and t0[8, 2] == 2 * 16 * 2  # 2 finished episodes per env, 16 envs per task per rank
    # shards are different environments (global ids differ) ...
    assert not np.array_equal(rob0, rob1)
    # ... and identical to the same global ids in a single 64-env handle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import env_state, make_env
    env = make_env("hostemu", 64, ["go_to_goal", "press_buttons"] * 32, seed=7, max_episode_steps=20)
    act = torch.zeros((64, 2))
    act[:, 0] = 0.5
    for _ in range(45):
        env.step(act)
    rob, _ = env_state(env)
    np.testing.assert_array_equal(rob[:32], rob0)
    np.testing.assert_array_equal(rob[32:], rob1)
