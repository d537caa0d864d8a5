"""Task registry (`isaacgymenvs/tasks/__init__.py:88-114`), restricted to the fused tasks."""
from .cartpole import Cartpole
from .locomotion import Ant, Humanoid

isaacgym_task_map = {
    "Ant": Ant,
    "Cartpole": Cartpole,
    "Humanoid": Humanoid,
}
