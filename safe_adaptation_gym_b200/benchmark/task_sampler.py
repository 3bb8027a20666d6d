"""benchmark/task_sampler.py:10-19 of the reference, same draw (`rs.permutation(items)[0]`)."""
from dataclasses import dataclass
from typing import Mapping, Optional, Tuple, Type

import numpy as np

from safe_adaptation_gym_b200.tasks import Task


@dataclass
class TaskSampler:
    rs: np.random.RandomState
    tasks: Mapping[str, Type[Task]]

    def sample(self) -> Optional[Tuple[str, Task]]:
        if len(self.tasks) == 0:
            return
        task_name, task = self.rs.permutation(list(self.tasks.items()))[0]
        return task_name, task()
