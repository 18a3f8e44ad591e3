"""CPU: gym registration of the drop-in ids (gym_narde/__init__.py:3-7) and the make() fallback's TimeLimit."""
import sys

import gymnasium_stub


def test_registration_branch_runs_with_gymnasium_present():
    gym = gymnasium_stub.install()
    import gym_narde_b200
    assert gym_narde_b200._register_with_gymnasium() is True
    reg = gym.envs.registration.registry
    for env_id in ("narde-v0", "Narde-v0"):
        spec = reg[env_id]
        ep = spec["entry_point"] if isinstance(spec, dict) else spec.entry_point
        mes = spec["max_episode_steps"] if isinstance(spec, dict) else spec.max_episode_steps
        assert ep == "gym_narde_b200.envs:NardeEnv" and mes == 1000          # gym_narde/__init__.py:6


def test_make_fallback_applies_the_time_limit():
    import gym_narde_b200

    class FakeEnv:
        unwrapped = property(lambda self: self)
        current_player = 1

        def reset(self, **kw):
            return "obs", {}

        def step(self, a):
            return "obs", 0, False, False, {}

    env = gym_narde_b200.TimeLimit(FakeEnv(), 3)
    env.reset()
    assert [env.step(0)[3] for _ in range(4)] == [False, False, True, True]
    env.reset()
    assert env.step(0)[3] is False and env.unwrapped.current_player == 1 and env.current_player == 1
    assert gym_narde_b200.MAX_EPISODE_STEPS == 1000
