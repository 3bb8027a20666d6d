"""Shared helpers for the parity tests: build the test-only host emulation, drive the batched env and the
oracle side by side."""
import os
import subprocess

import numpy as np
import torch

import oracle as O
from safe_adaptation_gym_b200 import _abi, tasks
from safe_adaptation_gym_b200.benchmark import TASKS
from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTEMU_SRC = os.path.join(ROOT, "tests", "hostemu", "sag_hostemu.cpp")
HOSTEMU_LIB = os.path.join(ROOT, "tests", "hostemu", "libsag_hostemu.so")
_hostemu = None


def hostemu_lib():
    """g++ build of the kernel body (tests/hostemu) -- test infrastructure, see the file header."""
    global _hostemu
    if _hostemu is None:
        deps = [HOSTEMU_SRC, os.path.join(ROOT, "safe_adaptation_gym_b200", "csrc", "sag_core.cuh"),
                os.path.join(ROOT, "safe_adaptation_gym_b200", "csrc", "sag_layout.h"), os.path.join(ROOT, "include", "sag_b200.h"),
                os.path.join(ROOT, "include", "sag_detmath.h")]
        if not os.path.exists(HOSTEMU_LIB) or os.path.getmtime(HOSTEMU_LIB) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off"] + (["-mfma"] if "fma" in open("/proc/cpuinfo").read().split() else []) + ["-o", HOSTEMU_LIB, HOSTEMU_SRC])
        _hostemu = _abi.SagLib(HOSTEMU_LIB, host_api=False)
    return _hostemu


def make_env(backend, n, task_names, seed, config=None, robot="point", **kw):
    """backend: 'hostemu' (CPU, kernel body compiled by g++) or 'cuda' (the product library)."""
    lib = hostemu_lib() if backend == "hostemu" else None
    env = BatchedSafeAdaptationGym("xmls/%s.xml" % robot, config=config, num_envs=n, _test_lib=lib, **kw)
    env.seed(seed)
    if isinstance(task_names, str):
        task_names = [task_names] * n
    env.set_task([TASKS[t]() for t in task_names])
    return env


def make_oracles(n, task_names, seed, config=None, gid_base=0, robot="point"):
    if isinstance(task_names, str):
        task_names = [task_names] * n
    cfg = {k: v for k, v in (config or {}).items()}
    envs = [O.OracleEnv(robot, task_names[e], config=cfg, seed=seed, env_gid=gid_base + e) for e in range(n)]
    for e in envs:
        assert e.reset(0) == 0
    return envs


def env_state(env):
    """(robot [n,6] (+6 car extras), objects [n,32,6]) as numpy"""
    n = env.num_envs
    robot = env.get_field("robot").cpu().numpy()[:, :n].T.copy()
    if env.robot_name == "car":
        robot = np.concatenate([robot, env.get_field("robot_ext").cpu().numpy()[:, :n].T], axis=1)
    objs = env.get_field("objects").cpu().numpy()[:, :, :n].transpose(2, 1, 0).copy()
    return robot, objs


def oracle_state(envs):
    robot = np.stack([np.concatenate([e.robot_state, e.robot_ext]) if e.robot == O.CAR else e.robot_state for e in envs])
    objs = np.zeros((len(envs), 32, 6))
    for i, e in enumerate(envs):
        o = e.objects()
        objs[i, :len(o)] = o[:, 2:8]
    return robot, objs


def _car_action(err, rng):
    """differential drive; the car's front is body -y (car.xml:19-20)"""
    fwd = np.clip(1.0 - abs(err), 0.0, 1.0) * 0.02
    a = np.array([fwd + 0.01 * np.clip(err, -1, 1), fwd - 0.01 * np.clip(err, -1, 1)])
    return np.clip(a + 0.003 * rng.normal(size=2), -1, 1)


def drive_action(oenv, rng, p_random=0.15):
    """scripted drive-to-target policy from the oracle state, so that goals / hazards / vases are visited"""
    s = oenv.robot_state
    objs = oenv.objects()
    ts = oenv.task_state
    kinds = objs[:, 0].astype(int)
    target = None
    if (kinds == O.BUTTON).any():
        btn = np.where(kinds == O.BUTTON)[0]
        groups = objs[btn, 1].astype(int)
        goal = btn[groups == 2]
        target = objs[goal[0], 2:4] if len(goal) else objs[btn[int(ts[2]) % len(btn)], 2:4]
    elif (kinds >= O.BOX).any() and oenv.task != O.TASK_ID["haul_box"]:  # push box, rod or ball
        box = objs[kinds >= O.BOX][0, 2:4]
        goal = objs[kinds == O.GOAL][0, 2:4]
        d = goal - box
        target = box - 0.35 * d / (np.linalg.norm(d) + 1e-9)
        if np.linalg.norm(target - s[:2]) < 0.15:
            target = goal
    elif (kinds == O.GOAL).any():
        target = objs[kinds == O.GOAL][0, 2:4]
    if target is None or rng.uniform() < p_random:
        return rng.uniform(-1, 1, 2)
    d = target - s[:2]
    if oenv.robot == O.CAR:
        return _car_action((np.arctan2(d[1], d[0]) - (s[2] - np.pi / 2) + np.pi) % (2 * np.pi) - np.pi, rng)
    err = (np.arctan2(d[1], d[0]) - s[2] + np.pi) % (2 * np.pi) - np.pi
    a = np.array([np.clip(1.0 - abs(err), 0.02, 1.0), np.clip(2.0 * err, -1, 1)])
    return np.clip(a + 0.2 * rng.normal(size=2), -1, 1)


def run_parity(backend, task_names, n, steps, seed, config=None, policy="drive", check_every=1, robot="point"):
    """Step the batched env and n oracle envs on identical actions and require BIT-EXACT agreement every step:
    robot / object state (float64), reward (float64), cost, done, and the float32 observation (== the oracle's
    float64 observation rounded to float32).  Exactness holds because both sides use only correctly rounded
    IEEE operations in the same order (nvcc -fmad=false, gcc -ffp-contract=off) plus include/sag_detmath.h.
    Returns summary statistics (so tests can assert that interesting events actually happened)."""
    cfg = dict(config or {})
    cfg.setdefault("action_noise", 0.0)
    env = make_env(backend, n, task_names, seed, cfg, robot=robot)
    orc = make_oracles(n, task_names, seed, cfg, robot=robot)
    r0, o0 = env_state(env)
    ro, oo = oracle_state(orc)
    np.testing.assert_array_equal(r0, ro)
    np.testing.assert_array_equal(o0, oo)
    obs = env.observation.cpu().numpy()
    for e in range(n):
        np.testing.assert_array_equal(obs[e], orc[e].observation().astype(np.float32))
    rng = np.random.RandomState(seed + 17)
    stats = {"cost": 0.0, "reward": 0.0, "goals": 0, "contacts": 0, "moved_objects": 0.0}
    for t in range(steps):
        acts = np.zeros((n, 2), dtype=np.float32)
        for e in range(n):
            a = drive_action(orc[e], rng) if policy == "drive" else rng.uniform(-1, 1, 2)
            acts[e] = a.astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(acts))
        obs = obs.cpu().numpy(); rew = rew.cpu().numpy(); cost = info["cost"].cpu().numpy(); done = done.cpu().numpy()
        np.testing.assert_array_equal(np.broadcast_to(torch.as_tensor(info["bound"]).cpu().numpy(), (n,)), [o.bound for o in orc])
        for e in range(n):
            oobs, orew, ocost, odone, rc = orc[e].step(acts[e].astype(np.float64))
            assert rc == 0
            msg = f"env {e} step {t} task {orc[e].task}"
            assert cost[e] == ocost, msg
            assert bool(done[e]) == odone, msg
            r = np.atleast_1d(rew[e])
            np.testing.assert_array_equal(r, orew[:r.size], err_msg=msg)
            np.testing.assert_array_equal(obs[e], oobs.astype(np.float32), err_msg=msg)
            stats["cost"] += ocost
            stats["reward"] += orew[r.size - 1]
            stats["goals"] += int(orew[r.size - 1] > 0.5)
            stats["contacts"] += len(orc[e].contacts())
        if t % check_every == 0 or t == steps - 1:
            r1, o1 = env_state(env)
            ro, oo = oracle_state(orc)
            np.testing.assert_array_equal(r1, ro, err_msg=f"robot state step {t}")
            np.testing.assert_array_equal(o1, oo, err_msg=f"object state step {t}")
            ng = int(cfg.get("num_gremlins", 0))
            if ng:  # mocap position of the last kinematics pass + weld anchors
                g = env.get_field("gremlins").cpu().numpy()[:, :n].T
                for e in range(n):
                    np.testing.assert_array_equal(g[e, :2 + 3 * ng], orc[e].gremlin_state[2:], err_msg=f"gremlin state env {e} step {t}")
    stats["moved_objects"] = float(np.abs(oo[:, :, :2] - o0[:, :, :2]).max())
    env.close()
    return stats


def run_overflow_case(backend, robot="point"):
    """Capacity overflow = the reference's PhysicsError path (safe_adaptation_gym.py:73-75: reward -10, done, cost 0):
    all ten vases of go_to_goal are stacked into one heap that the robot drives into, so the contact list (16) and the
    constrained-body table (8) overflow.  Same injected state in the batched env and in the oracle; exact agreement."""
    n = 4
    cfg = {"action_noise": 0.0}
    env = make_env(backend, n, "go_to_goal", 77, cfg, robot=robot)
    orc = make_oracles(n, "go_to_goal", 77, cfg, robot=robot)
    objs = env.get_field("objects")
    rob = env.get_field("robot")
    for e in range(n):
        rx, ry = float(rob[0, e]), float(rob[1, e])
        for k, s in enumerate(range(9, 19)):          # vase slots
            x, y = rx + 0.16 + 0.03 * (k % 4) + 0.01 * e, ry + 0.03 * (k // 4) - 0.03
            objs[0, s, e], objs[1, s, e], objs[2, s, e] = x, y, 0.1 * k
            orc[e].set_obj(s, x=x, y=y, yaw=0.1 * k)
        st = orc[e].robot_state; st[2] = 0.0; st[3] = 0.5
        orc[e].robot_state = st
        rob[2, e], rob[3, e] = 0.0, 0.5
    env.set_field("objects", objs); env.set_field("robot", rob)
    _ = env.observation
    for o in orc:
        o.forward()
    dones = 0
    for t in range(6):
        acts = np.tile(np.array([[1.0, 0.0]], dtype=np.float32), (n, 1))
        if robot == "car":
            acts[:] = 0.02
        obs, rew, done, info = env.step(torch.from_numpy(acts))
        for e in range(n):
            oobs, orew, ocost, odone, rc = orc[e].step(acts[e].astype(np.float64))
            assert bool(done[e]) == odone and float(rew[e]) == orew[0] and float(info["cost"][e]) == ocost, (t, e)
            np.testing.assert_array_equal(obs[e].cpu().numpy(), oobs.astype(np.float32))
            dones += int(odone)
        r1, o1 = env_state(env)
        ro, oo = oracle_state(orc)
        np.testing.assert_array_equal(r1, ro); np.testing.assert_array_equal(o1, oo)
    env.close()
    return dones


def run_chain_case(backend, robot="point", task="go_to_goal", steps=60):
    """Several movable bodies in one solve, below the capacity limits: the vases are lined up in front of the robot with
    1 cm gaps (every second one rotated), the robot pushes the first into the second into the third ...  Exercises
    multi-body Gauss-Seidel, the object-pair phase, floor rows of many bodies and sleeping, in every kernel variant."""
    n = 6
    cfg = {"action_noise": 0.0}
    env = make_env(backend, n, task, 78, cfg, robot=robot)
    orc = make_oracles(n, task, 78, cfg, robot=robot)
    objs = env.get_field("objects")
    rob = env.get_field("robot")
    kinds = orc[0].objects()[:, 0].astype(int)
    vases = [s for s in range(len(kinds)) if kinds[s] == O.VASE]
    for e in range(n):
        rx, ry = float(rob[0, e]), float(rob[1, e])
        heading = 0.0 if robot == "point" else np.pi / 2   # the car drives along body -y: yaw pi/2 -> world +x
        for k, s in enumerate(vases):
            if k < 2 + e:   # 2 .. 7 vases in the chain
                x, y, yaw = rx + 0.33 + 0.21 * k, ry + 0.004 * k, (0.3 if k % 2 else 0.0)
            else:           # the rest far away
                x, y, yaw = rx + 3.0 + 0.5 * k, ry + 3.0, 0.0
            objs[0, s, e], objs[1, s, e], objs[2, s, e] = x, y, yaw
            orc[e].set_obj(s, x=x, y=y, yaw=yaw)
        boxes = [s for s in range(len(kinds)) if kinds[s] in (O.BOX, O.ROD, O.BALL)]
        if boxes:   # the task's movable body goes first in the chain, the vases behind it
            s = boxes[0]
            x, y = rx + 0.45, ry + 0.01
            objs[0, s, e], objs[1, s, e], objs[2, s, e] = x, y, 0.2 * e
            orc[e].set_obj(s, x=x, y=y, yaw=0.2 * e)
            for k, sv in enumerate(vases):
                if k < 2 + e // 2:
                    x, y, yaw = rx + 0.45 + 0.42 + 0.21 * k, ry + 0.004 * k, (0.3 if k % 2 else 0.0)
                    objs[0, sv, e], objs[1, sv, e], objs[2, sv, e] = x, y, yaw
                    orc[e].set_obj(sv, x=x, y=y, yaw=yaw)
        for s in range(len(kinds)):   # everything else out of the way
            if kinds[s] in (O.PILLAR, O.HAZARD, O.BUTTON):
                x, y = rx - 3.0 - 0.3 * s, ry - 3.0
                objs[0, s, e], objs[1, s, e] = x, y
                orc[e].set_obj(s, x=x, y=y)
        st = orc[e].robot_state; st[2] = heading
        orc[e].robot_state = st
        rob[2, e] = heading
    env.set_field("objects", objs); env.set_field("robot", rob)
    _ = env.observation
    for o in orc:
        o.forward()
    moved, maxcon = 0.0, 0
    o0 = oracle_state(orc)[1]
    for t in range(steps):
        acts = np.tile(np.array([[1.0, 0.0]], dtype=np.float32), (n, 1))
        if robot == "car":
            acts[:] = np.array([[0.02, 0.02]], dtype=np.float32)   # both wheels forward: the car drives along body -y
        obs, rew, done, info = env.step(torch.from_numpy(acts))
        for e in range(n):
            oobs, orew, ocost, odone, rc = orc[e].step(acts[e].astype(np.float64))
            assert bool(done[e]) == odone and float(info["cost"][e]) == ocost, (t, e)
            np.testing.assert_array_equal(np.atleast_1d(rew[e].cpu().numpy()), orew[:1], err_msg=f"step {t} env {e}")
            np.testing.assert_array_equal(obs[e].cpu().numpy(), oobs.astype(np.float32), err_msg=f"step {t} env {e}")
            maxcon = max(maxcon, len(orc[e].contacts()))
        r1, o1 = env_state(env)
        ro, oo = oracle_state(orc)
        np.testing.assert_array_equal(r1, ro, err_msg=f"step {t}"); np.testing.assert_array_equal(o1, oo, err_msg=f"step {t}")
    moved = float(np.abs(oo[:, :, :2] - o0[:, :, :2]).max())
    nmoved = int((np.abs(oo[:, :, :2] - o0[:, :, :2]).max(axis=2) > 1e-6).sum(axis=1).max())
    env.close()
    return moved, nmoved, maxcon


def run_fullsize_parity(robot, task_cycle, n, steps, n_sample, seed=666, action_noise=0.01, use_host_api=False):
    """Oracle parity at a BASELINE batch size: the CUDA env steps all n environments (device-generated U(-1,1) actions,
    action noise on, so the in-step Philox stream is exercised), and n_sample of them -- half drawn at random, half the
    ones with the smallest initial clearance, i.e. the likeliest to touch something -- are mirrored by oracle
    environments created with the same seed / global env id / task.  Every step the sampled rows of obs / reward / cost /
    done must be bit-exact, every 50 steps the full state of the sampled environments too.  With this many environments
    the work list of every step carries hundreds to thousands of entries (dynamic fetch, bails from the free kernel)."""
    cfg = {"action_noise": action_noise}
    names = [task_cycle[e % len(task_cycle)] for e in range(n)]
    env = make_env("cuda", n, names, seed, cfg, robot=robot)
    clear = env.get_field("task_f64")[7, :n].cpu().numpy()
    rs = np.random.RandomState(seed)
    ids = set(np.argsort(clear)[:n_sample // 2].tolist())
    while len(ids) < n_sample:
        ids.add(int(rs.randint(n)))
    ids = np.array(sorted(ids))
    orc = [O.OracleEnv(robot, names[e], config=cfg, seed=seed, env_gid=int(e)) for e in ids]
    for o in orc:
        assert o.reset(0) == 0
    idt = torch.as_tensor(ids, device="cuda")
    r0, o0 = env_state(env)
    ro, oo = oracle_state(orc)
    np.testing.assert_array_equal(r0[ids], ro)
    np.testing.assert_array_equal(o0[ids], oo)
    g = torch.Generator(device="cuda"); g.manual_seed(seed + 1)
    L = env._lib
    od = env.obs_dim
    if use_host_api == "packed":
        # the output block of sag_host_alloc_outputs (one pinned block laid out like the device staging area): sag_step_host
        # then moves all four outputs with a single device-to-host copy per step
        import ctypes as C
        ptrs = [C.c_void_p() for _ in range(4)]
        L.check(L.L.sag_host_alloc_outputs(env._h, *[C.byref(x) for x in ptrs]))

        class _View:
            def __init__(self, ptr, ctype, shape):
                self.ptr = ptr
                self.arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=shape)

            def data_ptr(self):
                return self.ptr

            def __getitem__(self, i):
                return torch.from_numpy(self.arr[i].copy())
        act_h = torch.empty((n, 2), dtype=torch.float32).pin_memory()
        obs_h = _View(ptrs[0].value, C.c_float, (n, od)); rew_h = _View(ptrs[1].value, C.c_double, (n,))
        cost_h = _View(ptrs[2].value, C.c_uint8, (n,)); done_h = _View(ptrs[3].value, C.c_uint8, (n,))
    elif use_host_api:
        act_h = torch.empty((n, 2), dtype=torch.float32).pin_memory()
        obs_h = torch.empty((n, od), dtype=torch.float32).pin_memory()
        rew_h = torch.empty((n,), dtype=torch.float64).pin_memory()
        cost_h = torch.empty((n,), dtype=torch.uint8).pin_memory()
        done_h = torch.empty((n,), dtype=torch.uint8).pin_memory()
    stats = {"cost": 0.0, "contacts": 0, "max_worklist": 0}
    for t in range(steps):
        act = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
        if use_host_api:
            act_h.copy_(act); torch.cuda.synchronize()
            L.check(L.L.sag_step_host(env._h, act_h.data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), cost_h.data_ptr(), done_h.data_ptr()))
            obs_s, rew_s, cost_s, done_s = obs_h[ids].numpy(), rew_h[ids].numpy(), cost_h[ids].numpy(), done_h[ids].numpy()
        else:
            obs, rew, done, info = env.step(act)
            obs_s, rew_s = obs[idt].cpu().numpy(), rew[idt].cpu().numpy()
            cost_s, done_s = info["cost"][idt].cpu().numpy(), done[idt].cpu().numpy()
        acts = act[idt].cpu().numpy()
        for k, o in enumerate(orc):
            oobs, orew, ocost, odone, rc = o.step(acts[k].astype(np.float64))
            msg = f"global env {ids[k]} step {t} task {names[ids[k]]}"
            assert rc == 0, msg
            assert float(cost_s[k]) == ocost and bool(done_s[k]) == odone, msg
            assert float(rew_s[k]) == orew[0], msg
            np.testing.assert_array_equal(obs_s[k], oobs.astype(np.float32), err_msg=msg)
            stats["cost"] += ocost
            stats["contacts"] += len(o.contacts())
        if t % 50 == 49 or t == steps - 1:
            r1, o1 = env_state(env)
            ro, oo = oracle_state(orc)
            np.testing.assert_array_equal(r1[ids], ro, err_msg=f"robot state step {t}")
            np.testing.assert_array_equal(o1[ids], oo, err_msg=f"object state step {t}")
            mm = env.get_field("task_i32")[9, :n]
            stats["max_worklist"] = max(stats["max_worklist"], int((mm != 0).sum()))
    if use_host_api == "packed":
        L.L.sag_host_free(ptrs[0])
    env.close()
    return stats


def check_autoreset_and_stats(backend):
    """VERDICT r01 / ADVICE: (1) auto-reset hands back the FIRST observation of the new episode, done and info['truncated']
    for exactly the environments that were reset, also when they are out of phase; (2) per-task statistics stay with the
    task the episode ran under across set_task, and a manual reset of an unfinished episode is not an episode;
    (3) outputs are fresh tensors unless copy_outputs=False."""
    n, T = 6, 7
    env = make_env(backend, n, ["go_to_goal", "press_buttons"] * 3, 21, {"action_noise": 0.0}, max_episode_steps=T)
    dev = env.device
    act = torch.zeros((n, 2), device=dev)
    # put env 2 out of phase: it has already done 4 steps of its episode
    ti = env.get_field("task_i32"); ti[6, 2] = 4; env.set_field("task_i32", ti)
    kept = []
    for t in range(1, 2 * T + 1):
        obs, rew, done, info = env.step(act)
        kept.append(obs)
        expect = np.zeros(n, dtype=bool)
        expect[[0, 1, 3, 4, 5]] = (t % T == 0)
        expect[2] = ((t + 4) % T == 0)
        np.testing.assert_array_equal(done.cpu().numpy(), expect, err_msg=f"step {t}")
        np.testing.assert_array_equal(info["truncated"].cpu().numpy(), expect, err_msg=f"step {t}")
        if expect.any():   # the rows of the reset envs are the first observation of the new episode
            fresh = env.observation
            idx = np.nonzero(expect)[0]
            assert torch.equal(obs[idx], fresh[idx]), t
            ns = env.get_field("task_i32")[6, :n].cpu().numpy()
            assert (ns[idx] == 0).all()
    assert kept[0].data_ptr() != kept[1].data_ptr() and not torch.equal(kept[0], kept[-1])   # fresh tensors
    st = env.task_stats().cpu().numpy()
    # env ids 0, 2, 4 run go_to_goal (task 3), 1, 3, 5 press_buttons (task 8): two finished episodes each (env 2 at steps 3 and 10)
    assert st[3, 2] == 3 * 2 and st[8, 2] == 3 * 2, st[:, 2]
    env.close()


def check_stats_follow_task(backend):
    n, T = 4, 5
    env = make_env(backend, n, "go_to_goal", 22, {"action_noise": 0.0}, max_episode_steps=T)
    act = torch.zeros((n, 2), device=env.device)
    for _ in range(T + 2):           # one finished episode per env + 2 steps of the next
        env.step(act)
    from safe_adaptation_gym_b200 import tasks
    env.set_task(tasks.PushBox())    # cuts the running episodes short (not counted); the finished ones stay with go_to_goal
    for _ in range(T):
        env.step(act)
    env.reset()                      # manual reset right after an auto-reset: nothing unfinished is counted
    st = env.task_stats(reset=True).cpu().numpy()
    assert st[tasks.GoToGoal.task_id, 2] == n and st[tasks.PushBox.task_id, 2] == n, st[:, 2]
    assert env.task_stats().cpu().numpy().sum() == 0.0
    env.close()
