"""Single-game rules object with the reference's attribute surface, computed on the GPU.

Mirror of gym_narde/envs/narde.py:Narde (same field and method names, same return values) so
that callers such as train_deepq_pytorch.py (`env.unwrapped.game.get_valid_moves`) or the
reference's own tests run unchanged.  The fields live on the host exactly as in the reference
(callers mutate `game.board[...]` directly); every rules computation packs them into the
32-byte state record, runs the CUDA kernel through the C ABI for n = 1 and reads the answer back.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _cabi
from .. import state as S


def rotate_board(board):
    """gym_narde/envs/narde.py:16-17."""
    board = np.asarray(board)
    return np.concatenate((-board[12:], -board[:12])).astype(np.int32)


class _Dev:
    """Lazily allocated n=1 buffers shared by all Narde facades of the process.

    They live in PINNED HOST memory, which CUDA maps into the device address space: the kernels read the packed state /
    dice / codes and write their results straight over PCIe (the zero-copy I/O VecNardeEnv.step_host uses), so a facade
    call is one kernel launch and one stream synchronisation -- no host<->device copy operations (a call used to be 3
    H2D copies, a launch and 2 synchronous D2H copies)."""
    _inst = None

    def __init__(self):
        torch = _cabi.require_cuda()
        _cabi.load()
        self.torch = torch
        z = lambda shape, dtype: torch.zeros(shape, dtype=dtype).pin_memory()
        self.lo = z((1, 16), torch.uint8)
        self.hi = z((1, 16), torch.uint8)
        self.dice4 = z((1, 4), torch.uint8)
        self.dice2 = z((1, 2), torch.uint8)
        self.moves = z((1, _cabi.MAX_HALF_MOVES, 2), torch.uint8)
        self.counts = z(1, torch.int32)
        self.codes = z((1, 2), torch.int32)
        self.obs24 = z((1, 24), torch.int32)
        self.obs198 = z((1, 198), torch.float32)
        self.rew_i = z(1, torch.int32)
        self.rew_f = z(1, torch.float32)
        self.done = z(1, torch.uint8)
        self.act = z(1, torch.int64)
        self.actions = z((1, 4096), torch.int64)
        self.overflow = z(1, torch.uint8)
        self.board8 = z((1, 24), torch.int8)
        self.flag = z(1, torch.uint8)
        # numpy views and raw pointers of the hot single-env calls (packing through generic [n,16] array code and
        # re-validating every tensor per call cost more than the kernel: 98 -> see tools/facade_latency.py)
        self.lo_i8 = self.lo.numpy().view(np.int8)[0]
        self.hi_i8 = self.hi.numpy().view(np.int8)[0]
        self.hi_u8 = self.hi.numpy()[0]
        self.np = {k: getattr(self, k).numpy() for k in ("dice4", "dice2", "moves", "counts", "codes", "obs24", "rew_i", "done")}
        self.p = {k: C.c_void_p(getattr(self, k).data_ptr()) for k in
                  ("lo", "hi", "dice4", "dice2", "moves", "counts", "codes", "obs24", "rew_i", "done")}
        self.lib = _cabi.load()

    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def sync(self):
        """Wait for the kernels enqueued so far: their results are then in the (host) buffers."""
        self.torch.cuda.current_stream().synchronize()

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst


def upload_game(dev, game, turn):
    """The state record (gym_narde_b200/state.py layout) of one game, written straight into the pinned planes."""
    board = np.asarray(game.board)
    dev.lo_i8[:] = board[:16]
    dev.hi_i8[:8] = board[16:]
    h = dev.hi_u8
    h[8] = game.borne_off_white
    h[9] = game.borne_off_black
    dev.hi_i8[10] = turn
    h[11] = (S.FLAG_FIRST_W if game.first_turn_white else 0) | (S.FLAG_FIRST_B if game.first_turn_black else 0)
    h[12:16] = 0


def download_game(dev, game):
    dev.sync()
    try:
        game.board[:16] = dev.lo_i8  # keep the caller's array object alive when possible
        game.board[16:] = dev.hi_i8[:8]
    except Exception:
        game.board = np.concatenate([dev.lo_i8, dev.hi_i8[:8]]).astype(np.int32)
    h = dev.hi_u8
    flags = int(h[11])
    game.borne_off_white = int(h[8])
    game.borne_off_black = int(h[9])
    game.first_turn_white = bool(flags & S.FLAG_FIRST_W)
    game.first_turn_black = bool(flags & S.FLAG_FIRST_B)
    return {"turn": [int(dev.hi_i8[10])], "done": [bool(flags & S.FLAG_DONE)]}


class Narde:
    def __init__(self):
        # gym_narde/envs/narde.py:21-29
        self.board = np.zeros(24, dtype=np.int32)
        self.board[23] = 15
        self.board[11] = -15
        self.borne_off_white = 0
        self.borne_off_black = 0
        self.first_turn_white = True
        self.first_turn_black = True

    def get_perspective_board(self, current_player):
        """gym_narde/envs/narde.py:31-34 (narde_obs24 kernel)."""
        dev = _Dev.get()
        upload_game(dev, self, 1 if current_player == 1 else -1)
        _cabi.obs24(dev.lo, dev.hi, dev.obs24)
        dev.sync()
        return dev.obs24.numpy()[0].astype(np.int32)

    def get_valid_moves(self, roll, current_player=1):
        """gym_narde/envs/narde.py:58-92 (narde_half_moves kernel): ordered list of (from, to|'off')."""
        roll = [int(d) for d in roll]
        if len(roll) > 4 or any(d < 1 or d > 6 for d in roll):
            raise ValueError("roll must hold 1..4 dice in 1..6")
        dev = _Dev.get()
        upload_game(dev, self, 1 if current_player == 1 else -1)
        d4 = dev.np["dice4"][0]
        d4[:] = 0
        d4[:len(roll)] = roll
        p = dev.p
        rc = dev.lib.narde_half_moves(p["lo"], p["hi"], p["dice4"], 1, 0, p["moves"], p["counts"], dev.stream())
        if rc != 0:
            raise _cabi.NardeCudaError("narde_half_moves failed: %d" % rc)
        dev.sync()
        n = int(dev.np["counts"][0])
        return [(f, 'off' if t == S.OFF else t) for f, t in dev.np["moves"][0, :n].tolist()]

    def execute_rotated_move(self, move, current_player):
        """gym_narde/envs/narde.py:36-56 (narde_apply_actions kernel, half-move only)."""
        dev = _Dev.get()
        upload_game(dev, self, 1 if current_player == 1 else -1)
        a = S.encode_action([move])
        dev.act.copy_(dev.torch.tensor([a - (1 << 64) if a >= (1 << 63) else a], dtype=dev.torch.int64))
        _cabi.apply_actions(dev.lo, dev.hi, dev.act, flags=_cabi.HALF_MOVES_ONLY)
        download_game(dev, self)

    def _violates_block_rule(self, board):
        """gym_narde/envs/narde.py:139-184 (narde_violates_block_rule kernel)."""
        dev = _Dev.get()
        b = np.clip(np.asarray(board, dtype=np.int64), -127, 127).astype(np.int8).reshape(1, 24)
        dev.board8.copy_(dev.torch.from_numpy(b))
        _cabi.violates_block_rule(dev.board8, dev.flag)
        dev.sync()
        return bool(dev.flag.numpy()[0])

    def validate_move(self, move, roll, current_player=1):
        """gym_narde/envs/narde.py:186-192."""
        return tuple(move) in [tuple(m) for m in self.get_valid_moves(roll, current_player)]

    # ---- README contract (Tier N) --------------------------------------------------------
    def get_valid_actions(self, roll, current_player=1):
        """README.md:156-165: the set of legal full-turn actions ((src, dst), ...) for a 2-dice roll."""
        d1, d2 = int(roll[0]), int(roll[1])
        dev = _Dev.get()
        upload_game(dev, self, 1 if current_player == 1 else -1)
        dev.dice2.copy_(dev.torch.tensor([[abs(d1), abs(d2)]], dtype=dev.torch.uint8))
        _cabi.enumerate_actions(dev.lo, dev.hi, dev.dice2, dev.actions, dev.counts, dev.overflow)
        dev.sync()
        n = min(int(dev.counts.numpy()[0]), dev.actions.shape[1])
        acts = dev.actions.numpy()[0, :n].view(np.uint64).copy()
        return [tuple(S.decode_action(a)) for a in acts]
