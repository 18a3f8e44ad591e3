"""NardeEnv -- drop-in for gym_narde/envs/narde_env.py:NardeEnv, computed on the GPU.

rules="reference" (default, id `narde-v0`): identical call surface AND identical consumption of the
process-global numpy RNG (two np.random.randint(1,7) per step, the roll-off loop in reset), so a
script seeded with np.random.seed(s) sees the same dice, observations and rewards as with the
reference.  rules="full" (id `Narde-v0`): the README contract -- get_valid_actions(roll), Box(198)
observation, reward +1 iff WHITE wins.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi
from .. import state as S
from .narde import Narde, _Dev, download_game, upload_game

try:  # gymnasium is optional (not installed in the build image)
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _Base = _gym.Env
except Exception:  # minimal stand-ins with the attributes callers use
    _gym = None

    class _Base:  # noqa: D401
        metadata = {}

        @property
        def unwrapped(self):
            return self

    class _Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return np.random.randint(self.low, self.high + 1, size=self.shape).astype(self.dtype)
            return np.random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    class _Discrete:
        def __init__(self, n):
            self.n = n

        def sample(self):
            return int(np.random.randint(0, self.n))

    class _Tuple:
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

    class _spaces:  # noqa: N801
        Box, Discrete, Tuple = _Box, _Discrete, _Tuple


class NardeEnv(_Base):
    metadata = {"render_modes": ["human"], "render_fps": 4}

    def __init__(self, render_mode=None, rules="reference", reward=None):
        super().__init__()
        if rules not in ("reference", "full"):
            raise ValueError("rules must be 'reference' or 'full'")
        _cabi.require_cuda()
        _cabi.load()
        self.rules = rules
        self.reward_mode = reward or ("mover12" if rules == "reference" else "white01")
        self.game = Narde()
        self.current_player = 1
        self.render_mode = render_mode
        self.last_roll = None
        if rules == "reference":
            # gym_narde/envs/narde_env.py:18-22
            self.observation_space = _spaces.Box(low=-15, high=15, shape=(24,), dtype=np.int32)
        else:
            # README.md:44-102, bounds tests/test_observation_space.py:5-202
            low = np.zeros(198, dtype=np.float32)
            high = np.ones(198, dtype=np.float32)
            high[3:96:4] = 6.0
            high[101:194:4] = 6.0
            high[96] = 7.5
            high[194] = 7.5
            self.observation_space = _spaces.Box(low=low, high=high, shape=(198,), dtype=np.float32)
        self.action_space = _spaces.Tuple((_spaces.Discrete(24 * 24), _spaces.Discrete(24 * 24)))

    # ---- observations --------------------------------------------------------------------
    def _get_obs(self):
        dev = _Dev.get()
        upload_game(dev, self.game, self.current_player)
        if self.rules == "reference":
            _cabi.obs24(dev.lo, dev.hi, dev.obs24)  # narde_env.py:24-25
            dev.sync()
            return dev.obs24.numpy()[0].astype(np.int32)
        _cabi.obs198(dev.lo, dev.hi, dev.obs198)
        dev.sync()
        return dev.obs198.numpy()[0].copy()

    # ---- reset (narde_env.py:105-120) ----------------------------------------------------
    def reset(self, *, seed=None, options=None):
        if seed is not None:
            np.random.seed(seed)
        self.game = Narde()
        while True:
            white_roll = np.random.randint(1, 7)
            black_roll = np.random.randint(1, 7)
            if white_roll != black_roll:
                break
        self.current_player = 1 if white_roll > black_roll else -1
        self.last_roll = None
        return self._get_obs(), {}

    # ---- step ----------------------------------------------------------------------------
    def step(self, action):
        if self.rules == "reference":
            return self._step_reference(action)
        return self._step_full(action)

    def _step_reference(self, action):
        """narde_env.py:27-103 through the narde_step_ref kernel (dice from the global numpy RNG)."""
        dice = [int(np.random.randint(1, 7)), int(np.random.randint(1, 7))]
        dev = _Dev.get()
        t = dev.torch
        upload_game(dev, self.game, self.current_player)
        dev.np["dice2"][0, :] = dice
        dev.np["codes"][0, :] = (int(action[0]), int(action[1]))
        p = dev.p
        rc = dev.lib.narde_step_ref(p["lo"], p["hi"], p["dice2"], p["codes"], 1, 0, p["obs24"], p["rew_i"], p["done"], None,
                                    dev.stream())
        if rc != 0:
            raise _cabi.NardeCudaError("narde_step_ref failed: %d" % rc)
        u = download_game(dev, self.game)          # (synchronises: every output of the kernel is in host memory)
        self.current_player = int(u["turn"][0])
        obs = dev.np["obs24"][0].astype(np.int32)
        reward = int(dev.np["rew_i"][0])
        done = bool(int(dev.np["done"][0]))
        self.last_roll = tuple(dice)
        return obs, reward, done, False, {}

    def roll_dice(self):
        """Two dice from the global numpy RNG (same call the reference makes, narde_env.py:29)."""
        self.last_roll = (int(np.random.randint(1, 7)), int(np.random.randint(1, 7)))
        return self.last_roll

    def get_valid_actions(self, roll=None):
        """README.md:156-165: legal full-turn actions of the player to move for `roll`."""
        if roll is None:
            roll = self.last_roll or self.roll_dice()
        self.last_roll = (abs(int(roll[0])), abs(int(roll[1])))
        return self.game.get_valid_actions(self.last_roll, self.current_player)

    def _step_full(self, action):
        """Full-rules turn: `action` is an element of get_valid_actions(roll) (or its index);
        an action that is not legal for the last roll forfeits the turn, like the reference's
        silent handling of invalid codes (narde_env.py:63)."""
        if self.last_roll is None:
            self.roll_dice()
        legal = self.game.get_valid_actions(self.last_roll, self.current_player)
        if isinstance(action, (int, np.integer)):
            chosen = legal[int(action)] if 0 <= int(action) < len(legal) else None
        else:
            chosen = tuple(tuple(m) for m in action) if action is not None else None
            if chosen not in legal:
                chosen = None
        if chosen is None and legal and action is None:
            chosen = None
        dev = _Dev.get()
        t = dev.torch
        upload_game(dev, self.game, self.current_player)
        a = S.encode_action(list(chosen)) if chosen else 0xFFFFFFFFFFFFFFFF
        dev.act.copy_(t.tensor([a - (1 << 64) if a >= (1 << 63) else a], dtype=t.int64))
        flags = _cabi.REWARD_MOVER12 if self.reward_mode == "mover12" else 0
        _cabi.apply_actions(dev.lo, dev.hi, dev.act, reward=dev.rew_f, done=dev.done, flags=flags)
        _cabi.obs198(dev.lo, dev.hi, dev.obs198)
        u = download_game(dev, self.game)
        self.current_player = int(u["turn"][0])
        self.last_roll = None
        obs = dev.obs198.numpy()[0].copy()
        return obs, float(dev.rew_f.numpy()[0]), bool(int(dev.done.numpy()[0])), False, {}

    # ---- misc (narde_env.py:122-141) -----------------------------------------------------
    def render(self):
        if self.render_mode == "human":
            board_str = ""
            for i in range(24):
                board_str += f"{self.game.board[i]:>3} "
                if (i + 1) % 6 == 0:
                    board_str += "\n"
            print(board_str)

    def close(self):
        pass

    def _check_game_ended(self):
        if self.current_player == 1 and self.game.borne_off_white == 15:
            return True, (1 if self.game.borne_off_black > 0 else 2)
        elif self.current_player == -1 and self.game.borne_off_black == 15:
            return True, (1 if self.game.borne_off_white > 0 else 2)
        return False, 0
