"""Task registry (`isaacgymenvs/tasks/__init__.py:88-114`), restricted to the fused tasks."""
from .anymal_terrain import AnymalTerrain
from .cartpole import Cartpole
from .locomotion import Ant, Humanoid
from .shadow_hand import ShadowHand

isaacgym_task_map = {
    "Ant": Ant,
    "AnymalTerrain": AnymalTerrain,
    "Cartpole": Cartpole,
    "Humanoid": Humanoid,
    "ShadowHand": ShadowHand,
}
