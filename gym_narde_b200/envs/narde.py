"""Single-game rules object with the reference's attribute surface, computed on the GPU.

Mirror of gym_narde/envs/narde.py:Narde (same field and method names, same return values) so
that callers such as train_deepq_pytorch.py (`env.unwrapped.game.get_valid_moves`) or the
reference's own tests run unchanged.  The fields live on the host exactly as in the reference
(callers mutate `game.board[...]` directly); every rules computation packs them into the
32-byte state record, runs the CUDA kernel through the C ABI for n = 1 and reads the answer back.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi
from .. import state as S


def rotate_board(board):
    """gym_narde/envs/narde.py:16-17."""
    board = np.asarray(board)
    return np.concatenate((-board[12:], -board[:12])).astype(np.int32)


class _Dev:
    """Lazily allocated n=1 device buffers shared by all Narde facades of the process."""
    _inst = None

    def __init__(self):
        torch = _cabi.require_cuda()
        _cabi.load()
        dev = torch.device("cuda")
        self.torch = torch
        self.lo = torch.zeros((1, 16), dtype=torch.uint8, device=dev)
        self.hi = torch.zeros((1, 16), dtype=torch.uint8, device=dev)
        self.dice4 = torch.zeros((1, 4), dtype=torch.uint8, device=dev)
        self.dice2 = torch.zeros((1, 2), dtype=torch.uint8, device=dev)
        self.moves = torch.zeros((1, _cabi.MAX_HALF_MOVES, 2), dtype=torch.uint8, device=dev)
        self.counts = torch.zeros(1, dtype=torch.int32, device=dev)
        self.codes = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        self.obs24 = torch.zeros((1, 24), dtype=torch.int32, device=dev)
        self.obs198 = torch.zeros((1, 198), dtype=torch.float32, device=dev)
        self.rew_i = torch.zeros(1, dtype=torch.int32, device=dev)
        self.rew_f = torch.zeros(1, dtype=torch.float32, device=dev)
        self.done = torch.zeros(1, dtype=torch.uint8, device=dev)
        self.act = torch.zeros(1, dtype=torch.int64, device=dev)
        self.actions = torch.zeros((1, 4096), dtype=torch.int64, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.uint8, device=dev)
        self.board8 = torch.zeros((1, 24), dtype=torch.int8, device=dev)
        self.flag = torch.zeros(1, dtype=torch.uint8, device=dev)

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst


def upload_game(dev, game, turn):
    lo, hi = S.pack_states(np.asarray(game.board, dtype=np.int64), game.borne_off_white, game.borne_off_black,
                           turn, bool(game.first_turn_white), bool(game.first_turn_black))
    dev.lo.copy_(dev.torch.from_numpy(lo))
    dev.hi.copy_(dev.torch.from_numpy(hi))


def download_game(dev, game):
    u = S.unpack_states(dev.lo.cpu().numpy(), dev.hi.cpu().numpy())
    board = u["board"][0].astype(np.int32)
    try:
        game.board[:] = board  # keep the caller's array object alive when possible
    except Exception:
        game.board = board
    game.borne_off_white = int(u["off_w"][0])
    game.borne_off_black = int(u["off_b"][0])
    game.first_turn_white = bool(u["first_w"][0])
    game.first_turn_black = bool(u["first_b"][0])
    return u


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
        return dev.obs24.cpu().numpy()[0].astype(np.int32)

    def get_valid_moves(self, roll, current_player=1):
        """gym_narde/envs/narde.py:58-92 (narde_half_moves kernel): ordered list of (from, to|'off')."""
        roll = [int(d) for d in roll]
        if len(roll) > 4 or any(d < 1 or d > 6 for d in roll):
            raise ValueError("roll must hold 1..4 dice in 1..6")
        dev = _Dev.get()
        upload_game(dev, self, 1 if current_player == 1 else -1)
        d4 = np.zeros((1, 4), dtype=np.uint8)
        d4[0, :len(roll)] = roll
        dev.dice4.copy_(dev.torch.from_numpy(d4))
        _cabi.half_moves(dev.lo, dev.hi, dev.dice4, dev.moves, dev.counts)
        n = int(dev.counts.cpu()[0])
        mv = dev.moves.cpu().numpy()[0, :n]
        return [(int(f), 'off' if int(t) == S.OFF else int(t)) for f, t in mv]

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
        return bool(dev.flag.cpu()[0])

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
        n = min(int(dev.counts.cpu()[0]), dev.actions.shape[1])
        acts = dev.actions.cpu().numpy()[0, :n].view(np.uint64)
        return [tuple(S.decode_action(a)) for a in acts]
