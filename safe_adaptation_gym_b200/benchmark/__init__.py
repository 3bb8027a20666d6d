"""Task registry + benchmark samplers: the `benchmark.make` surface of the reference
(benchmark/__init__.py:11-84), host-side Python.

Registry keys are the snake_case class names in alphabetical CamelCase order, which is what the
reference's ``inspect.getmembers`` comprehension yields (benchmark/__init__.py:16-20); the order matters
because the sampler permutes ``list(TASKS.items())``.
"""
import re
from typing import Dict, Iterator, Tuple, Type

import numpy as np

from safe_adaptation_gym_b200 import tasks as _tasks
from safe_adaptation_gym_b200.benchmark.task_sampler import TaskSampler

BENCHMARKS = {"multitask", "task_adaptation"}
ROBOTS = {"point", "car", "doggo"}
ROBOTS_BASENAMES = {r: "xmls/%s.xml" % r for r in ("point", "car", "doggo")}

_CAMEL = re.compile(r"(?<!^)(?=[A-Z])")


def _registry() -> Dict[str, Type[_tasks.Task]]:
    classes = sorted(_tasks.__all__)  # == inspect.getmembers order (sorted by attribute name)
    return {_CAMEL.sub("_", c).lower(): getattr(_tasks, c) for c in classes}


TASKS = _registry()
TASK_IDS = {name: cls.task_id for name, cls in TASKS.items()}


class Benchmark:
    """`train_tasks` / `test_tasks` are generator properties of (task_name, Task), as in the reference."""

    def __init__(self, train_sampler: TaskSampler, test_sampler: TaskSampler, batch_size: int):
        self._samplers = {"train": train_sampler, "test": test_sampler}
        self._batch_size = batch_size

    def _draw(self, which: str) -> Iterator[Tuple[str, _tasks.Task]]:
        for _ in range(self._batch_size):
            item = self._samplers[which].sample()
            if item is None:
                return
            yield item

    @property
    def train_tasks(self) -> Iterator[Tuple[str, _tasks.Task]]:
        return self._draw("train")

    @property
    def test_tasks(self) -> Iterator[Tuple[str, _tasks.Task]]:
        return self._draw("test")

    @property
    def batch_size(self) -> int:
        return self._batch_size


def make(benchmark_name: str, batch_size: int = 16, seed: int = 666) -> Benchmark:
    """One shared RandomState feeds both samplers (benchmark/__init__.py:70-84)."""
    assert benchmark_name in BENCHMARKS, "Supplied a wrong benchmark name."
    rs = np.random.RandomState(seed)
    if benchmark_name == "multitask":
        train, heldout = TASKS, TASKS
    else:  # 'task_adaptation': permuted names, first 5 train, remaining 9 held out
        order = rs.permutation(list(TASKS.keys()))
        train = {k: TASKS[k] for k in order[:5]}
        heldout = {k: TASKS[k] for k in order[5:]}
    return Benchmark(TaskSampler(rs, train), TaskSampler(rs, heldout), batch_size)
