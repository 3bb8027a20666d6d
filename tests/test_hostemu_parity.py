"""Kernel body (sag_core.cuh compiled by g++, tests/hostemu) vs the oracle -- runs on the GPU-less
container.  The same checks run against the CUDA library in tests/test_gpu_parity.py."""
import pytest

from common import run_parity


def test_go_to_goal_random_actions():
    s = run_parity("hostemu", "go_to_goal", n=6, steps=150, seed=11, policy="random")
    assert s["cost"] >= 0


def test_go_to_goal_drive_hits_goals_hazards_and_vases():
    s = run_parity("hostemu", "go_to_goal", n=6, steps=400, seed=5)
    assert s["goals"] >= 3 and s["cost"] >= 10 and s["contacts"] >= 10, s


@pytest.mark.parametrize("task", ["go_to_goal_scarce", "go_to_goal_damping", "go_to_goal_motor", "catch_goal", "unsupervised",
                                  "press_buttons", "press_buttons_scarce", "collect", "push_box", "push_box_scarce", "haul_box",
                                  "roll_rod", "dribble_ball"])
def test_other_tasks(task):
    s = run_parity("hostemu", task, n=3, steps=250, seed=23)
    assert s["reward"] == s["reward"]


def test_mixed_task_batch_with_noise():
    names = ["go_to_goal", "press_buttons", "push_box", "collect", "catch_goal", "haul_box", "unsupervised", "go_to_goal_scarce"]
    run_parity("hostemu", names, n=8, steps=120, seed=3, config={"action_noise": 0.01})


@pytest.mark.parametrize("task", ["go_to_goal", "press_buttons", "push_box", "haul_box", "collect", "unsupervised", "catch_goal",
                                  "roll_rod", "dribble_ball"])
def test_car_tasks(task):
    s = run_parity("hostemu", task, n=3, steps=150, seed=31, robot="car")
    assert s["reward"] == s["reward"]


def test_car_random_actions_with_noise():
    run_parity("hostemu", ["go_to_goal", "press_buttons", "push_box", "haul_box"], n=4, steps=120, seed=7, policy="random",
               config={"action_noise": 0.01}, robot="car")


def test_multitask_sampler_batch_on_car():
    """BASELINE config 5: one environment per task of benchmark.make('multitask', 30, 666).train_tasks, car robot --
    covers all 14 registry tasks' device paths in one batch (README.md:59-63)"""
    from safe_adaptation_gym_b200 import benchmark
    names = [name for name, _ in benchmark.make("multitask", 30, 666).train_tasks]
    assert len(names) == 30 and {"roll_rod", "dribble_ball"} & set(names)
    run_parity("hostemu", names, n=30, steps=60, seed=666, config={"action_noise": 0.01}, robot="car")


@pytest.mark.parametrize("robot", ["point", "car"])
def test_adaptation_knobs_ctrl_range_scale_and_random_bound(robot):
    """world.py:72-78: Cauchy-scaled ctrlrange (incl. inverted ranges) and U(0, max_bound) constraint bound per Task instance"""
    cfg = {"robot_ctrl_range_scale": 0.6, "random_bound": True, "action_noise": 0.01}
    run_parity("hostemu", ["go_to_goal", "push_box", "press_buttons", "unsupervised"] * 4, n=16, steps=80, seed=41, config=cfg,
               policy="random", robot=robot)


@pytest.mark.parametrize("robot", ["point", "car"])
def test_capacity_overflow_is_a_physics_error(robot):
    from common import run_overflow_case
    assert run_overflow_case("hostemu", robot) >= 4   # every env hit the PhysicsError path (sticky until reset)


@pytest.mark.parametrize("robot", ["point", "car"])
def test_chain_of_pushed_bodies(robot):
    from common import run_chain_case
    moved, nmoved, maxcon = run_chain_case("hostemu", robot)
    assert moved > 0.05 and nmoved >= 3 and maxcon >= 4, (moved, nmoved, maxcon)


@pytest.mark.parametrize("task", ["push_box", "roll_rod", "dribble_ball"])
def test_chain_behind_the_task_body(task):
    """the robot pushes the box / rod / ball, which pushes a row of vases"""
    from common import run_chain_case
    moved, nmoved, maxcon = run_chain_case("hostemu", "point", task=task, steps=60)
    assert moved > 0.05 and nmoved >= 2 and maxcon >= 2, (moved, nmoved, maxcon)


@pytest.mark.parametrize("robot,task", [("point", "go_to_goal"), ("point", "push_box"), ("car", "press_buttons")])
def test_gremlins_welded_to_orbiting_mocaps(robot, task):
    """world.py:157-165 + primitive_objects.py:57-86 for a user-defined task with Task.obstacles[2] = 2 (no registry task has
    gremlins): weld rows, mocap staleness in the first substep, contact cost; plus the weld anchors / mocap state"""
    s = run_parity("hostemu", task, n=4, steps=200, seed=61, config={"num_gremlins": 2, "action_noise": 0.01}, robot=robot)
    assert s["moved_objects"] > 0.3, s   # the gremlins orbit their spawn position at radius gremlins_travel = 0.35
