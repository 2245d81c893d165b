"""navbot_ppo_b200 — B200-native LiDAR-navigation environment and PPO trainer behind the reference's
`Env.reset()/step()` and `PPO` surface (project_ppo/src/environment_new.py, ppo.py).

    from navbot_ppo_b200 import Env, VecEnv, PPO, NetActor, NetCritic

Importing the package loads navbot_ppo_b200/libnavbot_b200.so (there is no CPU fallback)."""
from .env import Env, VecEnv  # noqa: F401
from .nets import NetActor, NetCritic  # noqa: F401
from .ppo import PPO  # noqa: F401

__all__ = ["Env", "VecEnv", "PPO", "NetActor", "NetCritic"]
