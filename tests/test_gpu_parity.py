"""CUDA path (through the C ABI of csrc/libsag_b200.so) vs the oracle.  Needs a B200: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle as O
from common import env_state, make_env, make_oracles, oracle_state, run_fullsize_parity, run_parity

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded():
    from safe_adaptation_gym_b200 import _abi
    L = _abi.load()
    assert L.L.sag_abi_version() == 3
    assert L.path.endswith("csrc/libsag_b200.so")


def test_go_to_goal_random_actions():
    s = run_parity("cuda", "go_to_goal", n=64, steps=200, seed=11, policy="random")
    assert s["cost"] >= 0


def test_go_to_goal_drive_hits_goals_hazards_and_vases():
    s = run_parity("cuda", "go_to_goal", n=32, steps=500, seed=5)
    assert s["goals"] >= 10 and s["cost"] >= 50 and s["contacts"] >= 50, s


def test_go_to_goal_1000_step_trajectory_bit_exact():
    """north_star: point trajectories within a stated tolerance over 1000 steps -- here: bit-exact vs the oracle,
    through goals, hazards and vase contacts."""
    s = run_parity("cuda", "go_to_goal", n=8, steps=1000, seed=2, check_every=50)
    assert s["goals"] >= 5 and s["contacts"] >= 20 and s["moved_objects"] > 0.05, s


@pytest.mark.parametrize("task", ["go_to_goal_scarce", "go_to_goal_damping", "go_to_goal_motor", "catch_goal", "unsupervised",
                                  "press_buttons", "press_buttons_scarce", "collect", "push_box", "push_box_scarce", "haul_box",
                                  "roll_rod", "dribble_ball"])
def test_other_tasks(task):
    s = run_parity("cuda", task, n=8, steps=300, seed=23)
    assert s["reward"] == s["reward"]


def test_mixed_task_batch_with_noise():
    names = ["go_to_goal", "press_buttons", "push_box", "collect", "catch_goal", "haul_box", "unsupervised", "go_to_goal_scarce"] * 4
    run_parity("cuda", names, n=32, steps=150, seed=3, config={"action_noise": 0.01})


def test_sharding_is_independent_of_gpu_count():
    """Philox keys use the global env id: envs [64,128) of one big batch == a second handle with env_id_base=64."""
    a = make_env("cuda", 128, "go_to_goal", seed=9)
    b = make_env("cuda", 64, "go_to_goal", seed=9, env_id_base=64)
    ra, oa = env_state(a)
    rb, ob = env_state(b)
    np.testing.assert_array_equal(ra[64:], rb)
    np.testing.assert_array_equal(oa[64:], ob)
    act = torch.rand((128, 2), device="cuda") * 2 - 1
    for _ in range(20):
        o1, r1, d1, i1 = a.step(act)
        o2, r2, d2, i2 = b.step(act[64:])
        assert torch.equal(o1[64:], o2) and torch.equal(r1[64:], r2) and torch.equal(i1["cost"][64:], i2["cost"])


def test_full_size_batch_properties():
    """BASELINE config 2 size (65,536 envs): size-independent properties."""
    n = 65536
    env = make_env("cuda", n, "go_to_goal", seed=666, config={"action_noise": 0.01})
    robot, objs = env_state(env)
    # layout validity: world.py:197-199 keepouts hold for every env (robot 0.4, hazards 0.2, vases 0.15, pillar 0.3)
    keep = np.array([0.2] * 9 + [0.15] * 10 + [0.3])
    xy = objs[:, :20, :2]
    d = np.linalg.norm(xy[:, :, None, :] - xy[:, None, :, :], axis=-1)
    need = keep[None, :, None] + keep[None, None, :]
    iu = np.triu_indices(20, 1)
    assert (d[:, iu[0], iu[1]] >= need[:, iu[0], iu[1]] - 1e-12).all()
    dr = np.linalg.norm(xy - robot[:, None, :2], axis=-1)
    assert (dr >= 0.4 + keep[None, :] - 1e-12).all()
    assert (np.abs(robot[:, :2]) <= 1.6 + 1e-12).all()
    # 100 steps: lidar in [0,1], cost binary, done never, determinism of a second identical handle
    env2 = make_env("cuda", n, "go_to_goal", seed=666, config={"action_noise": 0.01})
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    tot_cost = 0.0
    for t in range(100):
        act = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
        o1, r1, d1, i1 = env.step(act)
        o2, r2, d2, i2 = env2.step(act)
        assert torch.equal(o1, o2) and torch.equal(r1, r2)
        assert float(o1[:, :48].min()) >= 0.0 and float(o1[:, :48].max()) <= 1.0
        assert not bool(d1.any())
        assert set(torch.unique(i1["cost"]).tolist()) <= {0.0, 1.0}
        tot_cost += float(i1["cost"].sum())
        assert torch.isfinite(r1).all()
    # spot-check 16 envs of the big batch against the oracle after 100 steps (same seeds / global ids)
    # (positions only: the oracle is re-run with the same Philox action noise)
    assert tot_cost >= 0.0


def test_layout_fail_rate_and_impossible_layout():
    """mirrors the reference's tests/test_layout_sampling.py: <= 0.5% failures; all sizes 2.0 must raise."""
    from safe_adaptation_gym_b200.utils import ResamplingError
    for task in ["catch_goal", "haul_box", "collect", "push_box", "press_buttons", "go_to_goal", "unsupervised"]:
        env = make_env("cuda", 2000, task, seed=0)  # would raise on any failure
        env.close()
    with pytest.raises(ResamplingError):
        make_env("cuda", 4, "go_to_goal", seed=0,
                 config={"hazards_size": 2.0, "vases_size": 2.0, "pillars_size": 2.0, "gremlins_size": 2.0, "max_layout_draws": 200000})


def _soa(n, nslots, rng):
    robot = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(-np.pi, 3 * np.pi, n)])
    obj = rng.uniform(-2.5, 2.5, (2, nslots, n))
    group = rng.choice([0, 1, 1, 1, 2, 3], size=(nslots, n)).astype(np.uint8)
    return robot, obj, group


@pytest.mark.parametrize("n", [1, 33, 4096])
def test_standalone_lidar_kernel(n):
    from safe_adaptation_gym_b200 import _abi
    L = _abi.load()
    rng = np.random.RandomState(n)
    nslots = 21
    robot, obj, group = _soa(n, nslots, rng)
    tr, to, tg = (torch.from_numpy(x).cuda() for x in (robot, obj, group))
    out = torch.empty((n, 48), dtype=torch.float32, device="cuda")
    L.check(L.L.sag_lidar(tr.data_ptr(), to.data_ptr(), tg.data_ptr(), n, nslots, out.data_ptr(), None))
    out = out.cpu().numpy()
    for e in range(min(n, 200)):
        for gi, off in ((1, 0), (3, 16), (2, 32)):
            m = group[:, e] == gi
            ref = O.lidar(robot[0, e], robot[1, e], robot[2, e], obj[0, m, e], obj[1, m, e])
            np.testing.assert_array_equal(out[e, off:off + 16], ref.astype(np.float32))


@pytest.mark.parametrize("n", [1, 1000, 65536])
def test_standalone_cost_kernel_bit_exact(n):
    from safe_adaptation_gym_b200 import _abi
    L = _abi.load()
    rng = np.random.RandomState(7)
    nh = 9
    robot = rng.uniform(-2, 2, (2, n))
    hz = rng.uniform(-2, 2, (2, nh, n)).astype(np.float32)
    k = min(n, 50)  # exact boundary cases: dist == 0.2 counts (world.py:152 `<=`)
    hz[0, 0, :k] = (robot[0, :k] + 0.2).astype(np.float32); hz[1, 0, :k] = robot[1, :k].astype(np.float32)
    contact = (rng.uniform(size=n) < 0.05).astype(np.uint8)
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    tr, th, tc = torch.from_numpy(robot).cuda(), torch.from_numpy(hz).cuda(), torch.from_numpy(contact).cuda()
    L.check(L.L.sag_cost(tr.data_ptr(), th.data_ptr(), tc.data_ptr(), n, nh, C.c_double(0.2), out.data_ptr(), None))
    d = np.sqrt((robot[0][None] - hz[0].astype(np.float64)) ** 2 + (robot[1][None] - hz[1].astype(np.float64)) ** 2)
    ref = ((d <= 0.2).any(axis=0) | (contact != 0)).astype(np.uint8)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("robot", ["point", "car"])
def test_host_buffer_api_matches_device_api(robot):
    """sag_step_host: pinned buffers take the overlapped path (bulk copy under the busy kernel + mapped-memory fix-up of
    the busy rows), pageable buffers the plain one; both must return exactly what the device API returns"""
    from safe_adaptation_gym_b200 import _abi
    n = 4096
    od = 60 if robot == "point" else 72
    tasks_ = ["go_to_goal", "push_box", "press_buttons", "go_to_goal"] * (n // 4)
    a = make_env("cuda", n, tasks_, seed=4, robot=robot)
    b = make_env("cuda", n, tasks_, seed=4, robot=robot)
    L = _abi.load()
    pinned = [torch.empty((n, 2), dtype=torch.float32).pin_memory(), torch.empty((n, od), dtype=torch.float32).pin_memory(),
              torch.empty((n,), dtype=torch.float64).pin_memory(), torch.empty((n,), dtype=torch.uint8).pin_memory(),
              torch.empty((n,), dtype=torch.uint8).pin_memory()]
    pageable = [torch.empty((n, 2), dtype=torch.float32), torch.empty((n, od), dtype=torch.float32),
                torch.empty((n,), dtype=torch.float64), torch.empty((n,), dtype=torch.uint8), torch.empty((n,), dtype=torch.uint8)]
    busy_rows = 0
    for t in range(160):
        act_h, obs_h, rew_h, cost_h, done_h = pageable if t % 8 == 7 else pinned
        act_h.uniform_(-1, 1)
        obs_h.fill_(-7.0)
        L.check(L.L.sag_step_host(a._h, act_h.data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), cost_h.data_ptr(), done_h.data_ptr()))
        obs, rew, done, info = b.step(act_h.cuda())
        assert torch.equal(obs.cpu(), obs_h) and torch.equal(rew.cpu(), rew_h), t
        assert torch.equal(info["cost"].cpu().to(torch.uint8), cost_h) and torch.equal(done.cpu().to(torch.uint8), done_h), t
        busy_rows += int((b.get_field("task_f64")[7, :n] < 0.05).sum())   # cached clearance: touching / moving / tendon
    assert busy_rows > 1000   # the fix-up path really carried rows


def test_auto_reset_and_task_stats():
    env = make_env("cuda", 64, ["go_to_goal", "press_buttons"] * 32, seed=1, max_episode_steps=25)
    act = torch.zeros((64, 2), device="cuda")
    for t in range(60):
        env.step(act)
    st = env.task_stats().cpu().numpy()
    assert st[3, 2] == 64 and st[8, 2] == 64  # 2 finished episodes x 32 envs per task
    ti = env.get_field("task_i32").cpu().numpy()
    assert (ti[6, :64] == 10).all()  # n_step inside the third episode


def test_rollout_kernel_matches_stepping():
    a = make_env("cuda", 256, "go_to_goal", seed=8)
    b = make_env("cuda", 256, "go_to_goal", seed=8)
    a.rollout(30)
    # replay the same Philox actions (stream 2) through step()
    n = 256
    for k in range(30):
        u = np.stack([O.philox_uniform2(8, k, 0, e, 2) for e in range(n)])
        act = torch.from_numpy((2.0 * u - 1.0).astype(np.float32)).cuda()
        b.step(act)
    ra, oa = env_state(a)
    rb, ob = env_state(b)
    np.testing.assert_array_equal(ra, rb)
    np.testing.assert_array_equal(oa, ob)


@pytest.mark.parametrize("task", ["go_to_goal", "press_buttons", "push_box", "haul_box", "collect", "unsupervised", "catch_goal",
                                  "roll_rod", "dribble_ball"])
def test_car_tasks(task):
    """BASELINE configs 3 / 4: car robot (reduced planar differential-drive model), bit-exact vs the oracle"""
    s = run_parity("cuda", task, n=16, steps=250, seed=31, robot="car")
    assert s["reward"] == s["reward"]


def test_car_mixed_batch_random_actions_with_noise():
    names = ["go_to_goal", "press_buttons", "push_box", "haul_box"] * 16
    s = run_parity("cuda", names, n=64, steps=200, seed=7, policy="random", config={"action_noise": 0.01}, robot="car")
    assert s["cost"] >= 0


def test_car_observation_shape_and_full_size_smoke():
    env = make_env("cuda", 32768, ["go_to_goal", "press_buttons"] * 16384, seed=666, robot="car")
    obs = env.observation
    assert tuple(obs.shape) == (32768, 72)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for _ in range(50):
        obs, rew, done, info = env.step(torch.rand((32768, 2), device="cuda", generator=g) * 2 - 1)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and not bool(done.any())
    # ball-joint quaternion stays normalised; the 3x3 block of the obs is a rotation matrix
    m = obs[:, 63:72].reshape(-1, 3, 3).double()
    eye = torch.eye(3, device="cuda", dtype=torch.float64)
    assert float((m @ m.transpose(1, 2) - eye).abs().max()) < 1e-5


@pytest.mark.gpu
def test_multitask_sampler_batch_on_car():
    """BASELINE config 5: the 30 train tasks of benchmark.make('multitask', 30, 666), one env each, car robot"""
    from safe_adaptation_gym_b200 import benchmark
    names = [name for name, _ in benchmark.make("multitask", 30, 666).train_tasks]
    run_parity("cuda", names, n=30, steps=100, seed=666, config={"action_noise": 0.01}, robot="car")


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["point", "car"])
def test_adaptation_knobs_ctrl_range_scale_and_random_bound(robot):
    cfg = {"robot_ctrl_range_scale": 0.6, "random_bound": True, "action_noise": 0.01}
    run_parity("cuda", ["go_to_goal", "push_box", "press_buttons", "unsupervised"] * 8, n=32, steps=120, seed=41, config=cfg,
               policy="random", robot=robot)


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["point", "car"])
def test_capacity_overflow_is_a_physics_error(robot):
    from common import run_overflow_case
    assert run_overflow_case("cuda", robot) >= 4


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["point", "car"])
def test_chain_of_pushed_bodies(robot):
    from common import run_chain_case
    moved, nmoved, maxcon = run_chain_case("cuda", robot, steps=100)
    assert moved > 0.05 and nmoved >= 3 and maxcon >= 4, (moved, nmoved, maxcon)


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["push_box", "roll_rod", "dribble_ball"])
def test_chain_behind_the_task_body(task):
    """the robot pushes the box / rod / ball, which pushes a row of vases"""
    from common import run_chain_case
    moved, nmoved, maxcon = run_chain_case("cuda", "point", task=task, steps=100)
    assert moved > 0.05 and nmoved >= 2 and maxcon >= 2, (moved, nmoved, maxcon)


def test_state_dict_resume_on_device():
    n = 512
    names = ["go_to_goal", "press_buttons", "push_box", "catch_goal"] * (n // 4)
    env = make_env("cuda", n, names, seed=21)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    for _ in range(60):
        env.step(torch.rand((n, 2), device="cuda", generator=g) * 2 - 1)
    sd = env.state_dict()
    acts = [torch.rand((n, 2), device="cuda", generator=g) * 2 - 1 for _ in range(40)]
    ref = [tuple(t.clone() for t in env.step(a)[:3]) for a in acts]
    other = make_env("cuda", n, "go_to_goal", seed=5)
    other.load_state_dict(sd)
    for a, (o, r, d) in zip(acts, ref):
        obs, rew, done, _ = other.step(a)
        assert torch.equal(obs, o) and torch.equal(rew, r) and torch.equal(done, d)


# ---- oracle parity at the BASELINE batch sizes ---------------------------------------------
@pytest.mark.gpu
def test_fullsize_point_go_to_goal_oracle_parity():
    """BASELINE config 2: 65,536 point go_to_goal environments, 300 steps, 256 environments mirrored by the oracle"""
    s = run_fullsize_parity("point", ["go_to_goal"], n=65536, steps=300, n_sample=256)
    assert s["contacts"] > 0 and s["max_worklist"] > 300      # the sampled envs touched things; the work list was long


@pytest.mark.gpu
def test_fullsize_car_go_to_goal_press_buttons_oracle_parity():
    """BASELINE config 3: 32,768 car environments, go_to_goal + press_buttons alternating"""
    s = run_fullsize_parity("car", ["go_to_goal", "press_buttons"], n=32768, steps=120, n_sample=96)
    assert s["contacts"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["point", "car"])
def test_fullsize_haul_box_push_box_oracle_parity(robot):
    """BASELINE config 4: 16,384 environments, haul_box + push_box alternating (tendon, 5-geom box)"""
    s = run_fullsize_parity(robot, ["haul_box", "push_box"], n=16384, steps=150 if robot == "point" else 100, n_sample=96)
    assert s["contacts"] > 0


@pytest.mark.gpu
def test_host_buffer_api_against_oracle():
    """sag_step_host (pinned buffers, overlapped copy + mapped-memory fix-up) checked against the oracle, not only
    against the device API"""
    s = run_fullsize_parity("point", ["go_to_goal", "push_box"], n=8192, steps=250, n_sample=128, use_host_api=True)
    assert s["max_worklist"] > 20     # the mapped-memory fix-up path carried rows


@pytest.mark.gpu
@pytest.mark.parametrize("robot,n", [("point", 16384), ("car", 8192)])
def test_host_buffer_api_packed_block_against_oracle(robot, n):
    """sag_step_host with the output block of sag_host_alloc_outputs (all four outputs in ONE device-to-host copy per step,
    actions read from the pinned host buffer by the kernels) against the oracle"""
    s = run_fullsize_parity(robot, ["go_to_goal", "push_box"], n=n, steps=200, n_sample=128, use_host_api="packed")
    assert s["max_worklist"] > 20


@pytest.mark.gpu
def test_new_abi_behaviours_on_device():
    """auto-reset row / truncated mask, statistics keyed by the task the episode ran under, host-pointer set_task and bound,
    two handles in one process (device guard), fresh output tensors"""
    from common import check_autoreset_and_stats
    check_autoreset_and_stats("cuda")
    from safe_adaptation_gym_b200 import _abi
    L = _abi.load()
    a = make_env("cuda", 256, "go_to_goal", seed=3, config={"random_bound": True})
    b = make_env("cuda", 128, "push_box", seed=4, robot="car")
    ids = np.array([3, 10] * 128, dtype=np.int32)
    L.check(L.L.sag_set_tasks_host(a._h, ids.ctypes.data))
    L.check(L.L.sag_reset_host(a._h, None, 0, 1, None))
    assert a.get_field("task_i32")[0, :256].cpu().numpy().tolist() == ids.tolist()
    bound = np.zeros(256)
    L.check(L.L.sag_bound_host(a._h, bound.ctypes.data_as(C.POINTER(C.c_double))))
    np.testing.assert_array_equal(bound, a.get_field("task_f64")[14, :256].cpu().numpy())
    assert (bound > 0).all() and (bound < 25).all() and len(np.unique(bound)) > 200
    bad = ids.copy(); bad[7] = 99
    assert L.L.sag_set_tasks_host(a._h, bad.ctypes.data) != 0
    badd = torch.as_tensor(bad, device="cuda")
    L.check(L.L.sag_set_tasks(a._h, badd.data_ptr(), None))
    torch.cuda.synchronize()
    assert L.L.sag_error_flags(a._h, 1) & _abi.ERR_BAD_TASK_ID
    assert L.L.sag_error_flags(a._h, 0) == 0
    o1, *_ = b.step(torch.zeros((128, 2), device="cuda"))     # the other handle still works
    assert torch.isfinite(o1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("robot,tasks_", [("point", ["go_to_goal", "push_box", "roll_rod", "dribble_ball"]), ("car", ["go_to_goal", "push_box", "roll_rod"])])
def test_speed_bounds_on_device(robot, tasks_):
    """tests/test_physics_bounds.py on the CUDA path: 4096 random-action environments, no body faster than an elastic
    collision with the robot allows, nothing non-finite"""
    n = 4096
    env = make_env("cuda", n, [tasks_[e % len(tasks_)] for e in range(n)], seed=31, robot=robot)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    term = 1.5 if robot == "point" else 1.0
    for t in range(300):
        obs, rew, done, info = env.step(torch.rand((n, 2), device="cuda", generator=g) * 2 - 1)
        if t % 20 == 19:
            r, o = env.get_field("robot")[:, :n], env.get_field("objects")[:, :, :n]
            assert torch.isfinite(r).all() and torch.isfinite(o).all() and torch.isfinite(obs).all()
            assert float(torch.hypot(r[3], r[4]).max()) <= 1.1 * term
            assert float(torch.hypot(o[3], o[4]).max()) <= 2.5 * term
    assert not bool(done.any())


@pytest.mark.gpu
def test_integration_md_ctypes_stub_runs_as_written():
    """the reference-side binding shown in INTEGRATION.md (numpy in / out, host pointers only) is executed verbatim and
    must return what the torch-side env returns"""
    import os
    import re
    from safe_adaptation_gym_b200 import _abi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    code = [b for b in re.findall(r"```python\n(.*?)```", md, re.S) if "class B200Bridge" in b][0]
    ns = {}
    exec(code, ns)
    n = 64
    names = ["go_to_goal", "push_box"] * (n // 2)
    br = ns["B200Bridge"](_abi.LIB_PATH, n, names, seed=12, config={"random_bound": 1})
    env = make_env("cuda", n, names, seed=12, config={"random_bound": True})
    np.testing.assert_array_equal(br.obs, env.observation.cpu().numpy())
    np.testing.assert_array_equal(br.bound, env.get_field("task_f64")[14, :n].cpu().numpy())
    rs = np.random.RandomState(0)
    for t in range(30):
        a = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
        obs, rew, done, info = br.step(a)
        o2, r2, d2, i2 = env.step(torch.from_numpy(a))
        np.testing.assert_array_equal(obs, o2.cpu().numpy())
        np.testing.assert_array_equal(rew, r2.cpu().numpy())
        np.testing.assert_array_equal(info["cost"], i2["cost"].cpu().numpy().astype(np.float64))
        np.testing.assert_array_equal(info["bound"], i2["bound"].cpu().numpy())
    obs = br.reset()
    env.reset()
    np.testing.assert_array_equal(obs, env.observation.cpu().numpy())
    with pytest.raises(KeyError):
        br.set_task("fly_to_goal")


@pytest.mark.gpu
@pytest.mark.parametrize("robot,tasks_", [("point", ["go_to_goal", "press_buttons", "push_box", "haul_box"]),
                                          ("car", ["go_to_goal", "press_buttons", "push_box"])])
def test_gremlins_welded_to_orbiting_mocaps(robot, tasks_):
    """world.py:157-165 + primitive_objects.py:57-86 (user-defined task with Task.obstacles[2] > 0): weld rows in the
    cooperative contact kernel, mocap staleness in the first substep, contact cost, weld anchors -- bit-exact vs the oracle"""
    for ng in (1, 3):
        s = run_parity("cuda", tasks_ * 4, n=4 * len(tasks_), steps=250, seed=61 + ng, config={"num_gremlins": ng, "action_noise": 0.01},
                       robot=robot)
        assert s["moved_objects"] > 0.3 and s["contacts"] > 0, s


@pytest.mark.gpu
def test_gremlin_environment_reset_and_checkpoint():
    """masked reset + state_dict round trip carry the gremlin field (weld anchors, mocap position)"""
    cfg = {"num_gremlins": 2, "action_noise": 0.0}
    a = make_env("cuda", 64, "go_to_goal", seed=9, config=cfg)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for _ in range(30):
        a.step(torch.rand((64, 2), device="cuda", generator=g) * 2 - 1)
    sd = a.state_dict()
    b = make_env("cuda", 64, "go_to_goal", seed=10, config=cfg)
    b.load_state_dict(sd)
    for _ in range(30):
        act = torch.rand((64, 2), device="cuda", generator=g) * 2 - 1
        oa, ra, _, _ = a.step(act)
        ob, rb, _, _ = b.step(act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb)
    assert "gremlins" in sd["fields"] and float(sd["fields"]["gremlins"][:2, :64].abs().max()) > 0.1


@pytest.mark.gpu
@pytest.mark.parametrize("robot,task", [("point", "roll_rod"), ("point", "go_to_goal"), ("car", "press_buttons"), ("point", "haul_box"),
                                        ("car", "push_box_scarce")])
def test_warp_cooperative_reset_layouts_bit_exact(robot, task):
    """the reset kernel (one warp per environment, 8 candidates x 4 check lanes per round) against the oracle's sequential
    sampler over 512 environments and three consecutive episodes: roll_rod restarts whole layouts 0.7 times per reset, a
    crowded go_to_goal goal burns all 1000 draws now and then -- positions, yaws, goal resample and draw counters must agree"""
    n = 512
    cfg = {"action_noise": 0.0}
    env = make_env("cuda", n, task, seed=1234, config=cfg, robot=robot)
    orc = make_oracles(n, task, seed=1234, config=cfg, robot=robot)
    for episode in range(3):
        if episode:
            env.reset()
            for o in orc:
                assert o.reset(episode) == 0
        r1, o1 = env_state(env)
        ro, oo = oracle_state(orc)
        np.testing.assert_array_equal(r1, ro, err_msg=f"robot, episode {episode}")
        np.testing.assert_array_equal(o1, oo, err_msg=f"objects, episode {episode}")
        ts = env.get_field("task_f64").cpu().numpy()[:, :n]
        want = np.array([o.task_state[:2] for o in orc]).T
        np.testing.assert_array_equal(ts[:2], want, err_msg="last distances after task.reset")
    # a masked reset touches the selected environments only
    before_r, before_o = env_state(env)
    mask = torch.zeros(n, dtype=torch.bool, device="cuda"); mask[::7] = True
    env.reset_envs(mask)
    r2, o2 = env_state(env)
    keep = ~mask.cpu().numpy()
    np.testing.assert_array_equal(r2[keep], before_r[keep]); np.testing.assert_array_equal(o2[keep], before_o[keep])
    for e in np.nonzero(~keep)[0]:
        assert orc[e].reset(3) == 0
    ro, oo = oracle_state(orc)
    np.testing.assert_array_equal(r2[~keep], ro[~keep]); np.testing.assert_array_equal(o2[~keep], oo[~keep])
    env.close()


@pytest.mark.gpu
def test_two_handles_on_two_devices_in_one_process():
    """every entry point runs on its handle's device and leaves the caller's current device alone (DevGuard): two
    environments on cuda:0 and cuda:1 stepped alternately from one process give what each gives alone"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = {"action_noise": 0.0}
    n = 256
    a0 = make_env("cuda", n, "go_to_goal", seed=5, config=cfg, device="cuda:0")
    a1 = make_env("cuda", n, "push_box", seed=6, config=cfg, device="cuda:1")
    b0 = make_env("cuda", n, "go_to_goal", seed=5, config=cfg, device="cuda:0")
    assert torch.cuda.current_device() == 0
    g = torch.Generator(); g.manual_seed(3)
    for _ in range(40):
        act = torch.rand((n, 2), generator=g) * 2 - 1
        o0, r0, _, _ = a0.step(act.to("cuda:0"))
        o1, r1, _, _ = a1.step(act.to("cuda:1"))
        ob, rb, _, _ = b0.step(act.to("cuda:0"))
        assert o1.device.index == 1 and torch.isfinite(o1).all()
        assert torch.equal(o0, ob) and torch.equal(r0, rb)
        assert torch.cuda.current_device() == 0
    a1.reset(); a0.reset()
    assert torch.cuda.current_device() == 0
    for e in (a0, a1, b0):
        e.close()
