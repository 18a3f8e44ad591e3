"""A minimal stand-in for `gymnasium` (not installed in this image, no network) -- TEST INFRASTRUCTURE.

Only the names the reference touches: Env, Wrapper, spaces.{Box, Discrete, Tuple},
envs.registration.{register, registry}, register, make (with the 'module:id' form and the TimeLimit wrapper
that `max_episode_steps` implies).  install() puts it into sys.modules when the real package is absent, so that
the reference (`import gymnasium as gym`, gym.make('gym_narde:narde-v0')) and this repo's registration branch
(gym_narde_b200/__init__.py) can run; with the real gymnasium installed it does nothing.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np


def _build():
    gym = types.ModuleType("gymnasium")

    class Env:
        metadata = {}
        render_mode = None

        def __init__(self, *a, **k):
            pass

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, **kw):
            return self.env.reset(**kw)

        def step(self, action):
            return self.env.step(action)

        def render(self):
            return self.env.render()

        def close(self):
            return self.env.close()

    class TimeLimit(Wrapper):
        """gymnasium.wrappers.TimeLimit: truncated=True once max_episode_steps steps were taken."""

        def __init__(self, env, max_episode_steps):
            super().__init__(env)
            self._max_episode_steps = int(max_episode_steps)
            self._elapsed_steps = 0

        def reset(self, **kw):
            self._elapsed_steps = 0
            return self.env.reset(**kw)

        def step(self, action):
            obs, reward, terminated, truncated, info = self.env.step(action)
            self._elapsed_steps += 1
            if self._elapsed_steps >= self._max_episode_steps:
                truncated = True
            return obs, reward, terminated, truncated, info

    class _Space:
        def sample(self):
            raise NotImplementedError

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return np.random.randint(self.low, self.high + 1, size=self.shape).astype(self.dtype)
            return np.random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    class Discrete(_Space):
        def __init__(self, n):
            self.n = n

        def sample(self):
            return int(np.random.randint(0, self.n))

    class Tuple(_Space):
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box, spaces.Discrete, spaces.Tuple = Box, Discrete, Tuple
    wrappers = types.ModuleType("gymnasium.wrappers")
    wrappers.TimeLimit = TimeLimit
    envs = types.ModuleType("gymnasium.envs")
    registration = types.ModuleType("gymnasium.envs.registration")
    registry = {}

    def register(id, entry_point=None, max_episode_steps=None, kwargs=None, **kw):
        registry[id] = {"entry_point": entry_point, "max_episode_steps": max_episode_steps, "kwargs": dict(kwargs or {})}

    def make(id, **kwargs):
        if ":" in id:
            mod, id = id.split(":", 1)
            importlib.import_module(mod)
        spec = registry[id]
        ep = spec["entry_point"]
        if isinstance(ep, str):
            mod, cls = ep.split(":")
            ep = getattr(importlib.import_module(mod), cls)
        env = ep(**{**spec["kwargs"], **kwargs})
        if spec["max_episode_steps"]:
            env = TimeLimit(env, spec["max_episode_steps"])
        return env

    registration.register, registration.registry = register, registry
    envs.registration = registration
    gym.Env, gym.Wrapper, gym.spaces, gym.envs, gym.wrappers = Env, Wrapper, spaces, envs, wrappers
    gym.register, gym.make = register, make
    gym.__narde_stub__ = True
    return {"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.envs": envs,
            "gymnasium.envs.registration": registration, "gymnasium.wrappers": wrappers}


def install():
    """Returns the gymnasium module in sys.modules (the real one if installed, else the stub)."""
    try:
        import gymnasium
        return gymnasium
    except ImportError:
        pass
    mods = _build()
    sys.modules.update(mods)
    return mods["gymnasium"]
