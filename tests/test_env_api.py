"""Host-side logic of the batched env (reference API semantics), exercised on the host-emulated kernels (CPU)."""
import numpy as np
import pytest
import torch

import oracle as O
import safe_adaptation_gym_b200 as sag
from common import check_autoreset_and_stats, check_stats_follow_task, env_state, hostemu_lib, make_env
from safe_adaptation_gym_b200 import benchmark, tasks
from safe_adaptation_gym_b200.utils import ResamplingError


def _make(robot="point", task="go_to_goal", n=8, **kw):
    return sag.make(robot, task, num_envs=n, _test_lib=hostemu_lib(), **kw)


def test_make_reset_step_shapes_and_info():
    env = _make(n=5, seed=666, config={"obstacles_size_noise_scale": 1.0})  # dead key accepted (world.py:29)
    obs = env.reset()
    assert tuple(obs.shape) == (5, 60) and obs.dtype == torch.float32
    assert env.action_space.shape == (2,) and env.observation_space.shape == (60,)
    obs, reward, done, info = env.step(np.stack([env.action_space.sample() for _ in range(5)]))
    assert tuple(reward.shape) == (5,) and reward.dtype == torch.float64
    assert done.dtype == torch.bool and not bool(done.any())          # go_to_goal.py:45: never terminal
    assert set(info) == {"cost", "bound"} and float(info["bound"][0]) == 25.0   # safe_adaptation_gym.py:79, world.py:32
    assert env.lidar_observations.shape == (5, 48)
    with pytest.raises(NotImplementedError):
        env.render()


def test_reset_before_task_asserts_and_unknown_names_raise():
    env = sag.make("point", None, num_envs=2, _test_lib=hostemu_lib())
    with pytest.raises(AssertionError):   # safe_adaptation_gym.py:93-96
        env.reset()
    with pytest.raises(KeyError):
        sag.make("point", "fly_to_goal", num_envs=2, _test_lib=hostemu_lib())
    with pytest.raises(NotImplementedError):
        sag.make("doggo", "go_to_goal", num_envs=2, _test_lib=hostemu_lib())
    with pytest.raises(NotImplementedError):
        sag.make("point", "go_to_goal", num_envs=2, rgb_observation=True, _test_lib=hostemu_lib())
    with pytest.raises(KeyError):
        sag.make("point", "go_to_goal", num_envs=2, config={"no_such_key": 1}, _test_lib=hostemu_lib())


def test_seed_and_reset_semantics():
    """seed(s) restarts the stream; reset() without seed moves to the next episode (safe_adaptation_gym.py:97-101,113-118)"""
    env = _make(n=4, seed=7)
    a, _ = env_state(env)
    env.reset()
    b, _ = env_state(env)
    assert not np.array_equal(a, b)                  # next episode, new layout
    env.reset(seed=7)                                # same key, first episode again ...
    env2 = _make(n=4, seed=7)
    c, _ = env_state(env)
    d, _ = env_state(env2)
    # ... make() = seed + set_task builds episode 0; reset(seed=7) builds episode 0 of the same key
    np.testing.assert_array_equal(c, d)
    # and it equals the oracle's episode 0 for the same seed / global ids
    orc = [O.OracleEnv("point", "go_to_goal", seed=7, env_gid=e) for e in range(4)]
    for o in orc:
        assert o.reset(0) == 0
    np.testing.assert_array_equal(c, np.stack([o.robot_state for o in orc]))


def test_reset_with_task_options_single_and_per_env_list():
    env = _make(n=6, seed=3)
    obs = env.reset(options={"task": tasks.PressButtons()})           # README.md:62-66
    ti = env.get_field("task_i32").numpy()
    assert (ti[0, :6] == tasks.PressButtons.task_id).all()
    names, ids = [], []
    for name, task in benchmark.make("multitask", 6, 666).train_tasks:
        names.append(name); ids.append(task)
    env.reset(options={"task": ids})
    ti = env.get_field("task_i32").numpy()
    assert ti[0, :6].tolist() == [t.task_id for t in ids]
    with pytest.raises(ValueError):
        env.set_task([tasks.GoToGoal()] * 5)


def test_impossible_layout_raises_resampling_error():
    """tests/test_layout_sampling.py:40-50,71-78 of the reference: all obstacle sizes 2.0 cannot be placed"""
    with pytest.raises(ResamplingError):
        _make(n=2, config={"hazards_size": 2.0, "vases_size": 2.0, "pillars_size": 2.0, "gremlins_size": 2.0, "max_layout_draws": 50000})
    assert issubclass(ResamplingError, AssertionError)   # utils.py:6


def test_layout_fail_rate_like_reference_test():
    """tests/test_layout_sampling.py:53-68: 200 layouts per task, at most 1 failure (here: none may fail, make() would raise)"""
    for task in ["catch_goal", "haul_box", "collect", "push_box", "press_buttons", "go_to_goal", "unsupervised"]:
        env = _make(task=task, n=200, seed=0)
        robot, objs = env_state(env)
        assert np.isfinite(robot).all()


def test_auto_reset_and_episode_statistics():
    env = _make(n=8, seed=1, max_episode_steps=10)
    act = torch.zeros((8, 2))
    for _ in range(25):
        env.step(act)
    st = env.task_stats().numpy()
    assert st[tasks.GoToGoal.task_id, 2] == 16            # 2 finished episodes per env
    ti = env.get_field("task_i32").numpy()
    assert (ti[6, :8] == 5).all()                         # 5 steps into the third episode
    assert (ti[8, :8] == 2).all()                         # episode counter


def test_car_observation_layout():
    env = _make(robot="car", n=3, seed=5)
    obs = env.reset().numpy()
    assert obs.shape == (3, 72)
    np.testing.assert_allclose(obs[:, 50], 9.81, rtol=1e-6)                      # accelerometer z
    np.testing.assert_array_equal(obs[:, 63:72].reshape(3, 3, 3), np.broadcast_to(np.eye(3, dtype=np.float32), (3, 3, 3)))  # ballquat -> I


def test_random_bound_and_ctrl_scale_are_per_task_instance():
    """world.py:72-78 + task.py:85-94: drawn once per Task instance -- kept by reset(), redrawn by set_task /
    reset(options={'task': ...}); bound ~ U(0, max_bound); ctrl-range scale = Cauchy * scale + 1"""
    n = 512
    env = _make(n=n, seed=9, config={"random_bound": True, "robot_ctrl_range_scale": 0.1, "max_bound": 10})
    b0 = env.step(np.zeros((n, 2), np.float32))[3]["bound"].clone()
    assert tuple(b0.shape) == (n,) and float(b0.min()) >= 0.0 and float(b0.max()) <= 10.0
    assert 4.0 < float(b0.mean()) < 6.0 and float(b0.std()) > 2.0
    sc0 = env.get_field("task_f64")[12:14, :n].clone()
    q = np.percentile(sc0.numpy().ravel(), [25, 50, 75])
    np.testing.assert_allclose(q, [0.9, 1.0, 1.1], atol=0.03)     # quartiles of a Cauchy(1, 0.1)
    env.reset()
    assert torch.equal(env.step(np.zeros((n, 2), np.float32))[3]["bound"], b0)
    assert torch.equal(env.get_field("task_f64")[12:14, :n], sc0)
    env.reset(options={"task": tasks.GoToGoal()})
    b1 = env.step(np.zeros((n, 2), np.float32))[3]["bound"]
    assert not torch.equal(b1, b0) and not torch.equal(env.get_field("task_f64")[12:14, :n], sc0)
    # defaults: fixed bound, unit scale
    env = _make(n=4, seed=9)
    assert env.get_field("task_f64")[12:15, :4].T.tolist() == [[1.0, 1.0, 25.0]] * 4


def test_state_dict_resume_is_bit_exact():
    """checkpoint / resume: a restored env continues exactly like the original (mixed tasks, goal resamples, button
    cooldowns, moving bodies and the Philox draw counters all live in the saved fields)"""
    n = 24
    names = ["go_to_goal", "press_buttons", "push_box", "catch_goal", "collect", "haul_box"] * 4
    env = _make(n=n, seed=21)
    by_name = {c.name: c for c in tasks.TASK_CLASSES}
    env.set_task([by_name[t]() for t in names])
    g = torch.Generator(); g.manual_seed(3)
    for _ in range(40):
        env.step(torch.rand((n, 2), generator=g) * 2 - 1)
    sd = env.state_dict()
    acts = [torch.rand((n, 2), generator=g) * 2 - 1 for _ in range(30)]
    ref = []
    for a in acts:
        obs, rew, done, info = env.step(a)
        ref.append((obs.clone(), rew.clone(), done.clone(), info["cost"].clone()))
    other = _make(n=n, seed=99)           # a different env object, different seed and tasks
    other.load_state_dict(sd)
    assert [t.name for t in other._tasks] == names
    for a, (o, r, d, c) in zip(acts, ref):
        obs, rew, done, info = other.step(a)
        assert torch.equal(obs, o) and torch.equal(rew, r) and torch.equal(done, d) and torch.equal(info["cost"], c)
    with pytest.raises(ValueError):
        _make(n=n + 1).load_state_dict(sd)


def test_autoreset_rows_truncated_mask_and_fresh_outputs():
    check_autoreset_and_stats("hostemu")


def test_statistics_stay_with_the_task_the_episode_ran_under():
    check_stats_follow_task("hostemu")


def test_state_dict_carries_config_and_rejects_a_different_one():
    a = _make(n=3, seed=5, config={"action_noise": 0.0})
    sd = a.state_dict()
    assert sd["config"]["action_noise"] == 0.0 and sd["env_id_base"] == 0 and sd["max_episode_steps"] == 0
    b = _make(n=3, seed=5)                      # default action_noise 0.01
    with pytest.raises(ValueError):
        b.load_state_dict(sd)
    c = _make(n=3, seed=9, config={"action_noise": 0.0})
    c.load_state_dict(sd)                       # same configuration: accepted


def test_bad_task_id_is_reported_not_indexed():
    from safe_adaptation_gym_b200 import _abi
    env = _make(n=4, seed=5)
    ids = torch.tensor([3, 3, 77, 3], dtype=torch.int32)
    env._lib.check(env._lib.L.sag_set_tasks(env._h, ids.data_ptr(), None))
    assert env._lib.L.sag_error_flags(env._h, 1) & _abi.ERR_BAD_TASK_ID
    assert env.get_field("task_i32")[0, :4].tolist() == [3, 3, 3, 3]
