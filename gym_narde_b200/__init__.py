"""gym_narde_b200 -- B200-native batched Narde environment (drop-in for gym_narde's env path).

Public surface (mirrors gym_narde/__init__.py + gym_narde/envs):
  gym_narde_b200.envs.NardeEnv / Narde   single-game facade, same names as the reference
  gym_narde_b200.VecNardeEnv             N lock-step games on the GPU
  gym_narde_b200.AfterstateMLP / AfterstateActor   DecomposedDQN(198) scorer and the greedy batched actor
  gym_narde_b200.NardeGameManager        interactive turn manager (my_game/narde_game_manager.py surface)
  gym_narde_b200.make(id)                'narde-v0' (reference rules) / 'Narde-v0' (README rules)
When gymnasium is importable both ids are also registered with it (max_episode_steps=1000, as
gym_narde/__init__.py:3-7 does).
"""
__version__ = "0.1.0"


def _lazy(name):
    if name == "VecNardeEnv":
        from .vec_env import VecNardeEnv
        return VecNardeEnv
    if name == "NardeEnv":
        from .envs.narde_env import NardeEnv
        return NardeEnv
    if name == "Narde":
        from .envs.narde import Narde
        return Narde
    if name == "AfterstateMLP":
        from .mlp import AfterstateMLP
        return AfterstateMLP
    if name == "AfterstateActor":
        from .actor import AfterstateActor
        return AfterstateActor
    if name == "NardeGameManager":
        from .narde_game_manager import NardeGameManager
        return NardeGameManager
    raise AttributeError(name)


def __getattr__(name):
    return _lazy(name)


def make(env_id="narde-v0", **kwargs):
    from .envs.narde_env import NardeEnv

    base = env_id.split(":")[-1]
    if base == "narde-v0":
        return NardeEnv(rules="reference", **kwargs)
    if base == "Narde-v0":
        return NardeEnv(rules="full", **kwargs)
    raise ValueError("unknown env id %r" % env_id)


try:  # optional gymnasium registration (gym_narde/__init__.py:3-7)
    from gymnasium.envs.registration import register as _register

    _register(id="narde-v0", entry_point="gym_narde_b200.envs:NardeEnv", max_episode_steps=1000)
    _register(id="Narde-v0", entry_point="gym_narde_b200.envs:NardeEnv", max_episode_steps=1000,
              kwargs={"rules": "full"})
except Exception:
    pass
