"""safe_adaptation_gym_b200 -- B200-native batched implementation of safe-adaptation-gym's per-step
environment loop behind the reference's own API (``make / env.reset(options={'task'}) / env.step /
env.set_task``, ``benchmark.make``).  See DESIGN.md."""
from typing import Dict, Optional

from safe_adaptation_gym_b200 import benchmark, tasks  # noqa: F401
from safe_adaptation_gym_b200.benchmark import ROBOTS_BASENAMES, TASKS
from safe_adaptation_gym_b200.utils import ResamplingError  # noqa: F401

__all__ = ['make', 'benchmark', 'tasks', 'TASKS', 'ResamplingError']


def make(robot_name: str,
         task_name: Optional[str] = None,
         seed: int = 666,
         config: Optional[Dict] = None,
         rgb_observation: bool = False,
         render_options: Optional[Dict] = None,
         render_lidar_and_collision=True,
         num_envs: int = 1,
         device=None,
         **kwargs):
    """Reference signature (safe_adaptation_gym/__init__.py:6-24) + `num_envs` / `device`.

    `render_lidar_and_collision` only adds visual sites in the reference (render.py) and is ignored here.
    """
    from safe_adaptation_gym_b200.env import BatchedSafeAdaptationGym
    env = BatchedSafeAdaptationGym(
        ROBOTS_BASENAMES[robot_name.lower()],
        config=config,
        rgb_observation=rgb_observation,
        render_lidars_and_collision=False,
        render_options=render_options,
        num_envs=num_envs,
        device=device,
        **kwargs)
    env.seed(seed)
    if task_name is not None:
        task = TASKS[task_name.lower()]()
        env.set_task(task)
    return env
